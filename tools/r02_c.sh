#!/bin/bash
# round 2, call c (2 GPUs): whole GPU suite incl. sharded == solo, lumel stage on config 5, 2-GPU bench of config 4
T=r02c
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -5 gpurun_out/${T}_pytest.log
LTR_TRACE=1 timeout 300 python - > gpurun_out/${T}_trace.log 2>&1 <<'PY'
from lighter_b200 import api, scenes
for name in ("config5", "config4", "config3"):
    sc = scenes.workload(name)
    for i in range(2):
        out = api.bake(sc)
    print(name, "wall", out["wall_s"], {k: round(v, 4) if isinstance(v, float) else v for k, v in out["stats"].items() if k.startswith("t_") or k.startswith("gpu_ms")})
PY
grep -E "wall|reference-order" gpurun_out/${T}_trace.log | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err
tail -c 1500 gpurun_out/${T}_bench_n2.json; tail -5 gpurun_out/${T}_bench_n2.err
