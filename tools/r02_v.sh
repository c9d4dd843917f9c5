#!/bin/bash
T=${1:-r02v}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -4 gpurun_out/${T}_pytest.log
bash tools/record_configs.sh $T
cat gpurun_out/${T}_bench_hashes.json
