#!/bin/bash
# round 2, first GPU call: sanity tests, config-5 launch list (the 3 s per bake), host phase trace of one config-4 bake, baseline lines of configs 3/5
T=r02a
mkdir -p gpurun_out
nproc > gpurun_out/${T}_nproc.txt
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -2 gpurun_out/${T}_pytest.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_launches_step_config5.csv \
    python tools/profile_step.py config5 > gpurun_out/${T}_step_config5.log 2>&1
tail -1 gpurun_out/${T}_step_config5.log
LTR_TRACE=1 LTR_TRACE_BVH=1 timeout 300 python - > gpurun_out/${T}_trace_config4.log 2>&1 <<'PY'
from lighter_b200 import api, scenes
sc = scenes.workload("config4")
for i in range(2):
    out = api.bake(sc)
    print("wall", out["wall_s"], {k: v for k, v in out["stats"].items() if k.startswith("t_")})
PY
tail -3 gpurun_out/${T}_trace_config4.log
for W in config3 config5; do
  timeout 300 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_$W.json 2> gpurun_out/${T}_bench_$W.err
  tail -c 300 gpurun_out/${T}_bench_$W.json
done
