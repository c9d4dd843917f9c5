import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def prims():
    return np.load(os.path.join(GOLD, "prims.npz"))


@pytest.fixture(scope="session")
def bakes():
    return np.load(os.path.join(GOLD, "bakes.npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def refprims():
    from oracle import REF_PRIMS_SO, RefPrims
    if not os.path.exists(REF_PRIMS_SO):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return RefPrims()


def bits_equal(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def scene_tris(scene):
    """World-space shadow triangles of a scenes.Scene whose instances all use identity-like float32
    matrices (v*M evaluated in float32 exactly as the reference does), degenerate ones dropped."""
    out = []
    for inst in scene.instances:
        m = np.asarray(inst.matrix, np.float32)
        for part in scene.meshes[inst.mesh].parts:
            if not part.shadow or not inst.shadow:
                continue
            p = part.pos.astype(np.float32)
            w = np.empty_like(p)
            for c in range(3):
                w[:, c] = ((p[:, 0] * m[0, c] + p[:, 1] * m[1, c]) + p[:, 2] * m[2, c]) + m[3, c] * np.float32(1.0)
            t = w[part.idx.reshape(-1, 3)].reshape(-1, 9)
            e1, e2 = t[:, 3:6] - t[:, 0:3], t[:, 6:9] - t[:, 0:3]
            cr = np.stack([e1[:, 1] * e2[:, 2] - e1[:, 2] * e2[:, 1], e1[:, 2] * e2[:, 0] - e1[:, 0] * e2[:, 2],
                           e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]], 1)
            keep = ~(np.abs(cr) < np.float32(0.001)).all(axis=1)
            out.append(t[keep])
    return np.concatenate(out) if out else np.zeros((0, 9), np.float32)
