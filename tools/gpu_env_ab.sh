#!/bin/bash
# bench the default library under a list of environment settings:  tools/gpu_env_ab.sh TAG "A=1" "B=2 C=3" ...
mkdir -p gpurun_out
T=$1; shift
i=0
for e in "$@"; do
    i=$((i+1))
    env $e timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 2 --e2e-steps 1 > gpurun_out/${T}_env$i.json 2> gpurun_out/${T}_env$i.err
    T=$T i=$i e="$e" python - <<'PY'
import json, os
T, i, e = os.environ["T"], os.environ["i"], os.environ["e"]
try:
    j = json.loads(open(f"gpurun_out/{T}_env{i}.json").read().strip().splitlines()[-1])
    print("%-24s ms/step %.1f" % (e, j["ms_per_step"]), {a: round(b, 1) for a, b in j["stage_ms"].items()}, "e2e wall %.3f" % j.get("bake_wall_s", 0))
except Exception as ex:
    print(e, "failed", ex)
PY
done
