#!/bin/bash
T=${1:-r02u}
mkdir -p gpurun_out
bash tools/r02_t.sh $T
for w in config4 config3 config5; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline --steps 3 --warmup 2 --e2e-steps 1 > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err
  T=$T w=$w python - <<'PY'
import json, os
T, w = os.environ["T"], os.environ["w"]
src = [l for f in (f"gpurun_out/{T}_bench_{w}.json", f"gpurun_out/{T}_bench_{w}.err") for l in open(f) if l.startswith("{")]
j = json.loads(src[-1])
print("%-8s ms/step %.1f" % (w, j["ms_per_step"]), {a: round(b, 1) for a, b in j["stage_ms"].items()}, "parity", j["parity"]["match"], j["parity"]["lightmap_fnv1a64"], "wall %.3f" % j["bake_wall_s"], "nodes", j["counters"]["n_node_visits"], "tris", j["counters"]["n_tri_tests"])
PY
done
