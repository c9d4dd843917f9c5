#!/bin/bash
T=${1:-r02z}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
bash tools/r02_x.sh $T LTR_BVH_LATE
