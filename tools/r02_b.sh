#!/bin/bash
# round 2, call b: device BVH builder tests + whole GPU suite + e2e trace of config 4 with the device-built tree
T=r02b
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bvh.py -x -q -s > gpurun_out/${T}_pytest_bvh.log 2>&1; tail -15 gpurun_out/${T}_pytest_bvh.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
LTR_TRACE=1 timeout 300 python - > gpurun_out/${T}_trace_config4.log 2>&1 <<'PY'
from lighter_b200 import api, scenes
sc = scenes.workload("config4")
for i in range(3):
    out = api.bake(sc)
    print("wall", out["wall_s"], {k: v for k, v in out["stats"].items() if k.startswith("t_") or k.startswith("gpu_ms")})
PY
grep -E "wall|scene BVH|accel|upload" gpurun_out/${T}_trace_config4.log | tail -24
