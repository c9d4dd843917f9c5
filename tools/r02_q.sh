#!/bin/bash
# N ranks: bench line (hash checked) + per-rank stage times of a resident bake
T=${1:-r02q}
N=${2:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${T}_bench_n$N.json") if l.startswith("{")][-1])
    print("N=$N ms/step %.2f  e2e wall %.4f  parity %s  clocks %s" % (d["ms_per_step"], d["bake_wall_s"], d["parity"]["match"], d["clocks"]))
    print(" stage_ms", {k: round(v, 2) for k, v in d["stage_ms"].items()})
    print(" host_s", {k: round(v, 4) for k, v in d["e2e"]["host_s"].items()})
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${T}_bench_n$N.err").read()[-3000:])
PY
LTR_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 tools/multi_gpu_stage_times.py config4 > gpurun_out/${T}_stage_times_n$N.log 2> gpurun_out/${T}_stage_times_n$N.err
grep "^rank" gpurun_out/${T}_stage_times_n$N.log
grep "rank 0\]" gpurun_out/${T}_stage_times_n$N.err | tail -40
