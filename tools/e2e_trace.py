#!/usr/bin/env python
"""Per-rank host stage times of END-TO-END sharded bakes (ltr_Start -> ltr_GetStatus), torchrun, one process per GPU.
  LTR_TRACE=1 python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/e2e_trace.py [workload] [bakes]
Every rank prints its own host_s table of the last bake; with LTR_TRACE=1 the library's phase laps go to stderr."""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lighter_b200 import api, scenes  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
shard = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = (ctypes.c_char * 128)()
        assert api.lib().ltrx_NcclUniqueId(buf)
        idt = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda")
    dist.broadcast(idt, 0)
    shard = (rank, world, bytes(idt.cpu().tolist()))
sc = scenes.workload(sys.argv[1] if len(sys.argv) > 1 else "config4")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
keys = ("t_prexform", "t_accel", "t_upload", "t_samples", "t_direct", "t_radiosity", "t_ao", "t_finalize", "t_readback")
for it in range(n):
    hh = api.BakeHandle(sc, device=local, shard=shard, output_root_only=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if it == n - 1 and rank == int(os.environ.get("TRACE_RANK", "0")):
        os.environ["LTR_TRACE"] = "1"                      # the library reads it per bake: phase laps of this rank's last bake on stderr
        sys.stderr.write(f"---- rank {rank} last bake trace ----\n")
    w = hh.run()
    st = hh.stats()
    hh.close()
    line = f"bake {it} rank {rank}: wall {w * 1e3:7.1f} ms | " + " ".join(f"{k[2:]} {st[k] * 1e3:6.1f}" for k in keys)
    for r in range(world):
        if r == rank:
            print(line, flush=True)
        if world > 1:
            dist.barrier()
if world > 1:
    dist.destroy_process_group()
