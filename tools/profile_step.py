#!/usr/bin/env python
"""Run ONE profiled bake of a workload on a resident scene (after one warm-up bake), bracketed by
cudaProfilerStart/Stop so that `ncu --profile-from-start off ...` captures exactly one step.

  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py config4_quarter
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lighter_b200 import api, scenes  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "config4_quarter"
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cudart = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else None
if cudart is None:
    import torch  # noqa: F401  (pulls in a libcudart)
    cudart = ctypes.CDLL("libcudart.so.12")
sc = scenes.workload(name)
with api.BakeHandle(sc) as h:
    h.prepare()
    for _ in range(warm):
        h.bake_resident()
    cudart.cudaProfilerStart()
    ms = h.bake_resident()
    cudart.cudaProfilerStop()
    st = h.stats()
    print("profiled step: %.1f ms" % ms, {k: v for k, v in st.items() if k.startswith("gpu_ms") or k.startswith("t_")})
