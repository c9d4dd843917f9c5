#!/bin/bash
# Lean profile run (one GPU, ~3 min): launch list of one clean step, launch list of the bench command, the bench line itself.
# The full ncu captures are taken separately, a few kernels at a time (each replayed kernel costs a save/restore of the
# whole device state: ~20 s per kernel with the 12 GB candidate buffer resident).
W=${1:-config4}
T=${2:-r01}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_launches_step_$W.csv \
    python tools/profile_step.py $W > gpurun_out/${T}_step_$W.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches_bench_$W.csv \
    python bench.py --workload $W --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/${T}_bench_under_ncu_$W.log 2>&1
python bench.py --workload $W > gpurun_out/${T}_bench_$W.json 2> gpurun_out/${T}_bench_$W.err
tail -c 600 gpurun_out/${T}_bench_$W.json
