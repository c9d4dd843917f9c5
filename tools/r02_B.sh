#!/bin/bash
# quick A/B of library variants on config 4 (stage times + hash)
T=${1:-r02B}
mkdir -p gpurun_out
cp lighter_b200/liblighter_b200.so /tmp/default.so
for v in default $(ls lighter_b200/variants | sed 's/lib_\(.*\)\.so/\1/'); do
    if [ $v = default ]; then cp /tmp/default.so lighter_b200/liblighter_b200.so; else cp lighter_b200/variants/lib_$v.so lighter_b200/liblighter_b200.so; fi
    timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 2 --e2e-steps 1 > gpurun_out/${T}_bench_$v.json 2> gpurun_out/${T}_bench_$v.err
    T=$T v=$v python - <<'PY'
import json, os
T, v = os.environ["T"], os.environ["v"]
try:
    src = [l for f in (f"gpurun_out/{T}_bench_{v}.json", f"gpurun_out/{T}_bench_{v}.err") for l in open(f) if l.startswith("{")]
    j = json.loads(src[-1])
    print("%-10s ms/step %.1f" % (v, j["ms_per_step"]), {a: round(b, 1) for a, b in j["stage_ms"].items()}, "parity", j["parity"]["match"], "wall %.4f" % j["bake_wall_s"])
except Exception as e:
    print(v, "failed", e, open(f"gpurun_out/{T}_bench_{v}.err").read()[-600:])
PY
done
cp /tmp/default.so lighter_b200/liblighter_b200.so
