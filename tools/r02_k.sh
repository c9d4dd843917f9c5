#!/bin/bash
# GPU tests on the default build, then the visibility-walk variants (tools/build_variants.sh) on config 4: stage times + lightmap hash check
T=${1:-r02k}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
cp lighter_b200/liblighter_b200.so /tmp/default.so
for v in default $(ls lighter_b200/variants | sed 's/lib_\(.*\)\.so/\1/'); do
    if [ $v = default ]; then cp /tmp/default.so lighter_b200/liblighter_b200.so; else cp lighter_b200/variants/lib_$v.so lighter_b200/liblighter_b200.so; fi
    timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 2 --e2e-steps 1 > gpurun_out/${T}_bench_$v.json 2> gpurun_out/${T}_bench_$v.err
    T=$T v=$v python - <<'PY'
import json, os
T, v = os.environ["T"], os.environ["v"]
try:
    j = json.loads([l for l in open(f"gpurun_out/{T}_bench_{v}.json") if l.startswith("{")][-1])
    c = j["counters"]
    print("%-10s ms/step %.1f" % (v, j["ms_per_step"]), {a: round(b, 1) for a, b in j["stage_ms"].items()}, "parity", j["parity"]["match"],
          "links", c["n_rad_links"], "tri tests/seg %.2f" % (c["n_ray_tri_tests"] / max(c["n_rad_segments"] + c["n_ao_segments"], 1)), "clocks", j["clocks"])
except Exception as e:
    print(v, "failed", e, open(f"gpurun_out/{T}_bench_{v}.err").read()[-600:])
PY
done
cp /tmp/default.so lighter_b200/liblighter_b200.so
