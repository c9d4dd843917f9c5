#!/bin/bash
# final multi-GPU evidence at N ranks: (N=2: the sharded==solo pytest) (N>=4: tools/multi_gpu_check.py on six scenes), bench line, per-rank stage times
T=${1:-r02U}
N=${2:-2}
mkdir -p gpurun_out
if [ "$N" = 2 ]; then
    timeout 600 python -m pytest tests/test_gpu_dropin.py -m gpu -x -q > gpurun_out/${T}_pytest_n2.log 2>&1; tail -2 gpurun_out/${T}_pytest_n2.log
else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/multi_gpu_check.py config4_sibling rad1 mesh2 mesh2:sampled config4_quarter corner > gpurun_out/${T}_multi_gpu_check_n$N.log 2>&1
    grep -E "bit-identical|Error|error|differ" gpurun_out/${T}_multi_gpu_check_n$N.log | tail -8
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${T}_bench_n$N.json").read().strip().splitlines()[-1])
    print("N=$N ms/step %.2f  e2e wall %.4f  parity %s" % (d["ms_per_step"], d["bake_wall_s"], d["parity"]["match"]), d.get("step_ms"))
    print(" stage_ms", {k: round(v, 2) for k, v in d["stage_ms"].items()})
    print(" host_s", {k: round(v, 4) for k, v in d["e2e"]["host_s"].items()})
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${T}_bench_n$N.err").read()[-3000:])
PY
LTR_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 tools/multi_gpu_stage_times.py config4 > gpurun_out/${T}_stage_times_n$N.log 2> gpurun_out/${T}_stage_times_n$N.err
grep -E "^rank" gpurun_out/${T}_stage_times_n$N.log | head -8
grep "ltr rank 0" gpurun_out/${T}_stage_times_n$N.err | tail -12 >> gpurun_out/${T}_stage_times_n$N.log
