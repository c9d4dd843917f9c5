#!/bin/bash
# Round-1 profile run (one GPU): launch list of the bench command, launch list of one clean step, and a
# full capture of the kernels that make up >95 % of the step.  Outputs in gpurun_out/ (copy summaries to profiles/).
set -x
W=${1:-config4}
T=${2:-r01}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches_bench_$W.csv \
    python bench.py --workload $W --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${T}_bench_under_ncu_$W.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_launches_step_$W.csv \
    python tools/profile_step.py $W > gpurun_out/${T}_step_$W.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"rad_candidates|rad_visibility|direct_march|lumel_fix|ao_trace" -c 16 \
    -o gpurun_out/${T}_full_$W python tools/profile_step.py $W > gpurun_out/${T}_full_$W.log 2>&1
python bench.py --workload $W > gpurun_out/${T}_bench_$W.json 2> gpurun_out/${T}_bench_$W.err
tail -c 1500 gpurun_out/${T}_bench_$W.json
