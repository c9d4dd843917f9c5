#!/bin/bash
# 8 GPUs: sharded == solo bit-identity on four scenes (log kept), then the bench line of config 4 (hash checked against the single-GPU hash)
T=${1:-r02i}
N=${2:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/multi_gpu_check.py config4_sibling rad1 mesh2 mesh2:sampled config4_quarter corner > gpurun_out/${T}_multi_gpu_check_n$N.log 2>&1
grep -E "bit-identical|Error|error" gpurun_out/${T}_multi_gpu_check_n$N.log | tail -8
LTR_TRACE_HOST=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${T}_bench_n$N.json").read().strip().splitlines()[-1])
    print("N=$N ms/step %.2f  e2e wall %.4f  parity %s" % (d["ms_per_step"], d["bake_wall_s"], d["parity"]["match"]))
    print(" stage_ms", {k: round(v, 2) for k, v in d["stage_ms"].items()})
    print(" host_s", {k: round(v, 4) for k, v in d["e2e"]["host_s"].items()})
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${T}_bench_n$N.err").read()[-3000:])
PY
