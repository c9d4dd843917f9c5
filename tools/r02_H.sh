#!/bin/bash
# version-2 entry sets: parity tests with the 16-entry shaft variant, then the A/B of all variants on config 4
T=${1:-r02H}
mkdir -p gpurun_out
cp lighter_b200/liblighter_b200.so /tmp/default.so
cp lighter_b200/variants/lib_e2_16.so lighter_b200/liblighter_b200.so
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/${T}_pytest_e2_16.log 2>&1; tail -3 gpurun_out/${T}_pytest_e2_16.log
cp /tmp/default.so lighter_b200/liblighter_b200.so
bash tools/r02_B.sh $T
