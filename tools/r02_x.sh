#!/bin/bash
# A/B by environment variable on configs 4 / 3 / 5 / mesh2: stage times + hash check
T=${1:-r02x}; VAR=${2:-LTR_MARCH_ORDER}
mkdir -p gpurun_out
for v in 1 0; do
  for w in config4 config3 config5 mesh2; do
    env $VAR=$v timeout 300 python bench.py --workload $w --no-cpu-baseline --steps 3 --warmup 2 --e2e-steps 1 > gpurun_out/${T}_bench_${w}_$v.json 2> gpurun_out/${T}_bench_${w}_$v.err
    T=$T v=$v w=$w VAR=$VAR python - <<'PY'
import json, os
T, v, w = os.environ["T"], os.environ["v"], os.environ["w"]
try:
    src = [l for f in (f"gpurun_out/{T}_bench_{w}_{v}.json", f"gpurun_out/{T}_bench_{w}_{v}.err") for l in open(f) if l.startswith("{")]
    j = json.loads(src[-1])
    print("%s=%s %-8s ms/step %.1f" % (os.environ["VAR"], v, w, j["ms_per_step"]), {a: round(b, 2) for a, b in j["stage_ms"].items()}, "parity", j["parity"]["match"], "wall %.4f" % j["bake_wall_s"], "nodes", j["counters"]["n_node_visits"])
except Exception as e:
    print(v, w, "failed", e, open(f"gpurun_out/{T}_bench_{w}_{v}.err").read()[-600:])
PY
  done
done
