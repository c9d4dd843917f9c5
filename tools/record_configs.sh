#!/bin/bash
# One bench line per BASELINE config on one GPU (configs 1, 2, 3, 5 + the headline config 4), kept under profiles/ each round,
# and the single-GPU lightmap hashes bench.py checks sharded runs against (tests/golden/bench_hashes.json).
#   bash tools/record_configs.sh r02      (on the GPU box; outputs in gpurun_out/, copy the ones to keep into profiles/)
T=${1:-r02}
mkdir -p gpurun_out
for W in mesh1 mesh2 config3 config5 config4; do
  timeout 900 python bench.py --workload $W --steps 3 --warmup 3 --record-hash > gpurun_out/${T}_bench_$W.json 2> gpurun_out/${T}_bench_$W.err || tail -3 gpurun_out/${T}_bench_$W.err
done
python - <<PY
import json
out = {}
for w in ("mesh1", "mesh2", "config3", "config5", "config4"):
    try:
        d = json.load(open(f"gpurun_out/${T}_bench_{w}.json"))
        out[w] = d["parity"]["lightmap_fnv1a64"]
        print(w, d["parity"]["lightmap_fnv1a64"], "ms/step %.2f" % d["ms_per_step"], "e2e wall %.4f" % d["bake_wall_s"], {k: round(v, 2) for k, v in d["stage_ms"].items()})
        if "same_config" in d: print("   same_config", d["same_config"])
        if "cpu_baseline" in d: print("   cpu", {k: v for k, v in d["cpu_baseline"].items() if k != "per_unit"})
    except Exception as e:
        print(w, "FAILED", e)
json.dump(out, open("gpurun_out/${T}_bench_hashes.json", "w"), indent=1)
PY
