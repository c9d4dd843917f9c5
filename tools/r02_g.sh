#!/bin/bash
# ncu --set full of the visibility kernel (2 launches of the bigger ones) on the quarter-size sibling of config 4, with source counters
T=r02g
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:rad_visibility --launch-skip 2 --launch-count 1 -o gpurun_out/${T}_vis -f \
    python tools/profile_step.py config4_quarter > gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_ncu.log; ls -la gpurun_out/${T}_vis.ncu-rep
