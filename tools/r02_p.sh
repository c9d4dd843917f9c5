#!/bin/bash
# Round-2 profile evidence on config 4 (one GPU): launch list of one clean step, launch list of the default bench command,
# ncu --set full (with source counters) of one launch of each of the three top kernels.
T=${1:-r02p}
W=config4
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_launches_step_$W.csv \
    python tools/profile_step.py $W > gpurun_out/${T}_step_$W.log 2>&1
tail -2 gpurun_out/${T}_step_$W.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${T}_launches_bench_$W.csv \
    python bench.py --no-cpu-baseline > gpurun_out/${T}_bench_under_ncu_$W.log 2>&1
wc -l gpurun_out/${T}_launches_bench_$W.csv
for K in rad_visibility:2 rad_candidates:2 direct_march:0; do
    k=${K%%:*}; s=${K##*:}
    timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:$k --launch-skip $s --launch-count 1 -o gpurun_out/${T}_$k -f \
        python tools/profile_step.py $W > gpurun_out/${T}_ncu_$k.log 2>&1
    ls -la gpurun_out/${T}_$k.ncu-rep
done
python bench.py > gpurun_out/${T}_bench_$W.json 2> gpurun_out/${T}_bench_$W.err
tail -c 400 gpurun_out/${T}_bench_$W.json
