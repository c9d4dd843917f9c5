#!/bin/bash
# Round-1 profile run (one GPU): launch list of the bench command, launch list of one clean step,
# and a full capture of the three kernels that make up >95 % of the step.  Outputs in gpurun_out/.
set -x
W=${1:-config4}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r01_launches_bench_$W.csv \
    python bench.py --workload $W --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r01_bench_under_ncu_$W.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r01_launches_step_$W.csv \
    python tools/profile_step.py $W > gpurun_out/r01_step_$W.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"rad_candidates|rad_visibility|direct_march" -c 4 \
    -o gpurun_out/r01_full_$W python tools/profile_step.py $W > gpurun_out/r01_full_$W.log 2>&1
python bench.py --workload $W --no-cpu-baseline > gpurun_out/r01_bench_$W.json 2> gpurun_out/r01_bench_$W.err
tail -c 1500 gpurun_out/r01_bench_$W.json
