#!/usr/bin/env python
"""Multi-GPU parity check, launched with torchrun (one process per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 \
      tools/multi_gpu_check.py [workload ...]

Every rank bakes its lumel shard (BVH replicated, per-bounce radiance all-gather over NCCL); rank 0
then bakes the same scene alone and the two results must be bit-identical: sharding changes who
computes a lumel, never the arithmetic or the summation order."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lighter_b200 import api, scenes  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
names = sys.argv[1:] or ["config4_sibling", "rad1", "mesh2", "mesh2:sampled"]
ok = True
for name in names:
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = (ctypes.c_char * 128)()
        assert api.lib().ltrx_NcclUniqueId(buf)
        idt = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda")
    dist.broadcast(idt, 0)
    nccl_id = bytes(idt.cpu().tolist())
    mode = 1 if name.endswith(":sampled") else 0            # "<scene>:sampled" = sampled soft-shadow extension mode
    base = name.split(":")[0]
    sc = scenes.NAMED[base]() if base in scenes.NAMED else scenes.workload(base)
    if mode:
        for lt in sc.lights:
            lt.shadow_sample_count = 8
    api.srand(1)
    out = api.bake(sc, device=local, shard=(rank, world, nccl_id), reset_rand=False, shadow_mode=mode)
    dist.barrier()
    if rank == 0:
        api.srand(1)
        solo = api.bake(sc, device=local, reset_rand=False, shadow_mode=mode)
        same = all(np.array_equal(a["rgb"].view(np.uint32), b["rgb"].view(np.uint32)) for a, b in zip(out["lightmaps"], solo["lightmaps"]))
        nsame = all((a["normals"] is None and b["normals"] is None) or np.array_equal(a["normals"].view(np.uint32), b["normals"].view(np.uint32))
                    for a, b in zip(out["lightmaps"], solo["lightmaps"]))
        st, s1 = out["stats"], solo["stats"]
        print(f"{name}: world={world} sharded==solo lightmaps bit-identical: {same}, normals: {nsame}; lumels {st['n_lumels_local']}/{st['n_lumels_total']} "
              f"local; wall sharded {out['wall_s']:.3f}s solo {solo['wall_s']:.3f}s; links local {st['n_rad_links']} solo {s1['n_rad_links']}", flush=True)
        ok = ok and same and nsame
    dist.barrier()
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
