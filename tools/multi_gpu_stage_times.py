#!/usr/bin/env python
"""Per-rank stage times of a sharded resident bake (torchrun, one process per GPU): shows load imbalance.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/multi_gpu_stage_times.py [workload]"""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lighter_b200 import api, scenes  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    buf = (ctypes.c_char * 128)()
    assert api.lib().ltrx_NcclUniqueId(buf)
    idt = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda")
dist.broadcast(idt, 0)
sc = scenes.workload(sys.argv[1] if len(sys.argv) > 1 else "config4")
h = api.BakeHandle(sc, device=local, shard=(rank, world, bytes(idt.cpu().tolist())) if world > 1 else None)
h.prepare()
for _ in range(2):
    ms = h.bake_resident()
st = h.stats()
line = f"rank {rank}: span {ms:7.1f} ms | " + " ".join(f"{k[7:]} {st[k]:6.1f}" for k in ("gpu_ms_samples", "gpu_ms_direct", "gpu_ms_march", "gpu_ms_radiosity", "gpu_ms_rad_pairs",
                                                                                      "gpu_ms_rad_vis", "gpu_ms_ao", "gpu_ms_finalize")) + f" | marches {st['n_marches']} queries {st['n_distance_queries']}"
for r in range(world):
    if r == rank:
        print(line, flush=True)
    dist.barrier()
h.close()
dist.destroy_process_group()
