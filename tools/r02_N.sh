#!/bin/bash
# packet form of the visibility kernel: parity tests with the default build, then the A/B of the variants on config 4
T=${1:-r02N}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
bash tools/r02_B.sh $T
