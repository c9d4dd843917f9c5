#!/usr/bin/env python
"""Convert the reference's bundled test meshes (bin/*.data, OBJ-like text) into
tests/golden/meshes/*.npz so tests and bench.py can run where /root/reference does not exist
(the GPU box).  Run in the build container:  python tools/import_reference_meshes.py

The arrays are exactly what the reference's loader (lighter_test.cpp:28-108) hands to
ltr_MeshAddPart: de-duplicated float32 positions / normals / uvs (v flipped) and u32 indices.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lighter_b200.scenes import MESH_DIR, parse_data_mesh  # noqa: E402

REF_BIN = "/root/reference/bin"

if __name__ == "__main__":
    os.makedirs(MESH_DIR, exist_ok=True)
    for name in ("test-mesh", "test-set2-mesh1", "test-set2-mesh2"):
        p = parse_data_mesh(os.path.join(REF_BIN, name + ".data"))
        np.savez_compressed(os.path.join(MESH_DIR, name + ".npz"), pos=p.pos, nrm=p.nrm, uv=p.uv1, idx=p.idx)
        print(f"{name}: {len(p.pos)} verts, {len(p.idx) // 3} tris")
