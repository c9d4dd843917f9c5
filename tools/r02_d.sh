#!/bin/bash
T=r02d
mkdir -p gpurun_out
python tools/diag_vs_reference.py rad1 > gpurun_out/${T}_diag_rad1.log 2>&1
LTR_BVH_HOST=1 python tools/diag_vs_reference.py rad1 > gpurun_out/${T}_diag_rad1_host.log 2>&1
cat gpurun_out/${T}_diag_rad1.log; echo ====; cat gpurun_out/${T}_diag_rad1_host.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; tail -8 gpurun_out/${T}_pytest.log
