#!/bin/bash
# GPU tests, then march variants on configs 4 / 3 / 5 (stage times, walks vs steps, hash check)
T=${1:-r02l}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
cp lighter_b200/liblighter_b200.so /tmp/default.so
for v in default $(ls lighter_b200/variants | sed 's/lib_\(.*\)\.so/\1/'); do
  if [ $v = default ]; then cp /tmp/default.so lighter_b200/liblighter_b200.so; else cp lighter_b200/variants/lib_$v.so lighter_b200/liblighter_b200.so; fi
  for w in config4 config3 config5; do
    timeout 300 python bench.py --workload $w --no-cpu-baseline --steps 3 --warmup 2 --e2e-steps 1 > gpurun_out/${T}_bench_${w}_$v.json 2> gpurun_out/${T}_bench_${w}_$v.err
    T=$T v=$v w=$w python - <<'PY'
import json, os
T, v, w = os.environ["T"], os.environ["v"], os.environ["w"]
try:
    src = [l for f in (f"gpurun_out/{T}_bench_{w}_{v}.json", f"gpurun_out/{T}_bench_{w}_{v}.err") for l in open(f) if l.startswith("{")]
    j = json.loads(src[-1])
    print("%-8s %-8s ms/step %.1f" % (v, w, j["ms_per_step"]), {a: round(b, 1) for a, b in j["stage_ms"].items()}, "parity", j["parity"]["match"], j["parity"]["lightmap_fnv1a64"], "wall %.3f" % j["bake_wall_s"])
except Exception as e:
    print(v, w, "failed", e, open(f"gpurun_out/{T}_bench_{w}_{v}.err").read()[-600:])
PY
  done
done
cp /tmp/default.so lighter_b200/liblighter_b200.so
