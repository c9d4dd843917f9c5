#!/bin/bash
# One GPU call: parity tests, then A/B of the entry-set walks on the bench workload (bundle entry sets on / off).
mkdir -p gpurun_out
T=${1:-r01b}
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_pytest.log 2>&1
tail -5 gpurun_out/${T}_pytest.log
LTR_RAD_ENTRY=0 LTR_AO_ENTRY=0 timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --e2e-steps 1 > gpurun_out/${T}_bench_root.json 2> gpurun_out/${T}_bench_root.err
timeout 600 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --e2e-steps 1 > gpurun_out/${T}_bench_entry.json 2> gpurun_out/${T}_bench_entry.err
T=$T python - <<'PY'
import json, os
T = os.environ["T"]
for k in ("root", "entry"):
    try:
        j = json.loads(open(f"gpurun_out/{T}_bench_{k}.json").read().strip().splitlines()[-1])
        print(k, "ms/step %.1f" % j["ms_per_step"], {a: round(b, 1) for a, b in j["stage_ms"].items()}, "node visits", j["counters"]["n_ray_node_visits"],
              "e2e wall", j.get("bake_wall_s"))
    except Exception as e:
        print(k, "failed", e)
PY
