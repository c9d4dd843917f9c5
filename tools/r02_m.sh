#!/bin/bash
T=${1:-r02m}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -5 gpurun_out/${T}_pytest.log
timeout 600 python tools/sample_fn_timing.py config4 3 > gpurun_out/${T}_sample_fn_config4.json 2> gpurun_out/${T}_sample_fn_config4.err; tail -c 1500 gpurun_out/${T}_sample_fn_config4.json; tail -3 gpurun_out/${T}_sample_fn_config4.err
