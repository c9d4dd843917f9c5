#!/bin/bash
# e2e host-stage trace at N ranks
T=${1:-r02j}
N=${2:-8}
mkdir -p gpurun_out
nproc > gpurun_out/${T}_nproc.txt; free -g | head -2 >> gpurun_out/${T}_nproc.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/e2e_trace.py config4 4 > gpurun_out/${T}_e2e_trace_n$N.log 2> gpurun_out/${T}_e2e_trace_n$N.err
grep -E "^bake 3" gpurun_out/${T}_e2e_trace_n$N.log
grep "ltr host" gpurun_out/${T}_e2e_trace_n$N.err | tail -40
cat gpurun_out/${T}_nproc.txt
