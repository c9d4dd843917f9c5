"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the headers
declare, lays its structs out like the reference, and its host-side API behaves like the
reference's (return codes, defaults, status strings) -- all without touching a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import GOLD, ROOT
from lighter_b200 import api, scenes


def declared_symbols(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    return re.findall(r"LTRAPI[^;]*?\b(ltrx?_\w+)\s*\(", txt)


def test_library_exports_every_declared_symbol():
    L = api.lib()
    decl = declared_symbols("lighter.h") + declared_symbols("lighter_b200.h")
    assert len(decl) >= 17 + 19
    for name in decl:
        assert hasattr(L, name), f"liblighter_b200.so does not export {name}"
    # the 17 entry points of the reference header (lighter.h:191,237-258)
    assert sorted(declared_symbols("lighter.h")) == sorted(api.LTR_SYMBOLS)
    assert set(api.LTRX_SYMBOLS) <= set(declared_symbols("lighter_b200.h"))


def test_struct_layout_matches_reference_abi():
    """tests/golden/abi_layout.txt was printed by a probe compiled against /root/reference/lighter.h."""
    types = {"ltr_MeshPartInfo": api.MeshPartInfo, "ltr_MeshInstanceInfo": api.MeshInstanceInfo, "ltr_LightInfo": api.LightInfo,
             "ltr_SampleInfo": api.SampleInfo, "ltr_SampleRequest": api.SampleRequest, "ltr_Config": api.Config,
             "ltr_WorkOutputInfo": api.WorkOutputInfo, "ltr_WorkOutput": api.WorkOutput, "ltr_WorkStatus": api.WorkStatus}
    n = 0
    for line in open(os.path.join(GOLD, "abi_layout.txt")):
        m = re.match(r"SZ\((\w+), (\d+)\)", line)
        if m:
            assert C.sizeof(types[m.group(1)]) == int(m.group(2)), line
            n += 1
        m = re.match(r"OFF\((\w+), (\w+), (\d+)\)", line)
        if m:
            assert getattr(types[m.group(1)], m.group(2)).offset == int(m.group(3)), line
            n += 1
    assert n > 50
    assert api.lib().ltrx_abi_checked() == 1        # the same table, asserted at compile time in C++


def test_next_power_of_two():
    L = api.lib()
    for x, want in [(0, 0), (1, 1), (2, 2), (3, 4), (5, 8), (64, 64), (65, 128), (1000, 1024), (2 ** 31, 2 ** 31), (2 ** 31 + 1, 0)]:
        assert L.ltr_NextPowerOfTwo(x) == want, x


def test_default_config_and_roundtrip():
    L = api.lib()
    cfg = api.Config()
    L.ltr_GetConfig(C.byref(cfg), None)
    d = scenes.default_config()
    for k in ("max_lightmap_size", "default_width", "default_height", "global_size_factor", "max_correct_dist", "max_correct_angle",
              "bounce_count", "ao_distance", "ao_multiplier", "ao_falloff", "ao_effect", "ao_num_samples", "blur_size", "ds2x",
              "generate_normalmap_data"):
        assert np.float32(getattr(cfg, k)) == np.float32(d[k]), k
    assert cfg.max_num_threads == 0x7FFF and cfg.max_tree_memory == 128 * 1024 * 1024
    assert not cfg.sample_fn and cfg.size_fn
    s = L.ltr_CreateScene()
    cfg.ao_distance, cfg.bounce_count = 2.5, 3
    assert L.ltr_SetConfig(s, C.byref(cfg)) == 1      # the reference returns 1 here, not LTRC_SUCCESS
    back = api.Config()
    L.ltr_GetConfig(C.byref(back), s)
    assert back.ao_distance == 2.5 and back.bounce_count == 3
    L.ltr_DestroyScene(s)


def test_default_size_func():
    L = api.lib()
    cfg = api.Config()
    L.ltr_GetConfig(C.byref(cfg), None)
    out = (api.u32 * 2)(0, 0)
    # basic scenario sizes: quads of area 4, 0.16, 14.44 at importance 1 -> 8, 1, 16
    for area, imp, want in [(4.0, 1.0, 8), (0.16, 1.0, 1), (14.44, 1.0, 16), (1e-9, 1.0, 1), (100.0, 0.3, 16)]:
        assert L.ltr_DefaultSizeFunc(C.byref(cfg), None, 0, None, 0, area, imp, out) == 1
        assert out[0] == out[1] == want, (area, imp, out[0])
    assert L.ltr_DefaultSizeFunc(C.byref(cfg), None, 0, None, 0, 1e9, 1.0, out) == 0       # above max_lightmap_size


def test_scene_setup_return_codes_and_status():
    L = api.lib()
    s = L.ltr_CreateScene()
    st = api.WorkStatus()
    assert L.ltr_GetStatus(s, C.byref(st)) != 0 and st.stage == b"not started" and st.completion == 0.0
    m = L.ltr_CreateMesh(s, b"mesh1", 5)
    q = scenes._quad_part()
    pi = api.MeshPartInfo(q.pos.ctypes.data, q.nrm.ctypes.data, q.uv1.ctypes.data, q.uv2.ctypes.data, 12, 12, 8, 8, q.idx.ctypes.data, 4, 6, 1)
    assert L.ltr_MeshAddPart(m, C.byref(pi)) == 1
    pi.index_count = 2                                 # index_count < 3 && % 3 != 0 -> rejected (lighter.cpp:1237)
    assert L.ltr_MeshAddPart(m, C.byref(pi)) == 0
    ii = api.MeshInstanceInfo()
    C.memmove(ii.matrix, scenes.IDENTITY.ctypes.data, 64)
    ii.importance, ii.shadow = 1.0, 1
    assert L.ltr_MeshAddInstance(m, C.byref(ii)) == 1
    info = api.WorkOutputInfo()
    L.ltr_GetWorkOutputInfo(s, C.byref(info))
    assert info.lightmap_count == 0 and info.sample_count == 0
    wo = api.WorkOutput()
    assert L.ltr_GetWorkOutput(s, 0, C.byref(wo)) == 0  # out of range -> 0 (lighter.cpp:1341)
    si = api.SampleInfo(7, api.VEC3(0, 0, 1), api.VEC3(0, 0, 1), api.VEC3(9, 9, 9))
    L.ltr_SampleAdd(s, C.byref(si))
    L.ltr_GetWorkOutputInfo(s, C.byref(info))
    assert info.sample_count == 1 and info.samples[0].id == 7 and info.samples[0].out_color[0] == 0.0
    L.ltr_Abort(s)
    L.ltr_DestroyScene(s)


def test_light_add_consumes_one_rand():
    """The reference draws one randf() per ltr_LightAdd (lighter.cpp:1300); AO offsets depend on it."""
    L = api.lib()
    libc = api._libc
    api.srand(1)
    first = [libc.rand() for _ in range(3)]
    api.srand(1)
    s = L.ltr_CreateScene()
    li = api.LightInfo(1, api.VEC3(0, 0, 1), api.VEC3(0, 0, 0), api.VEC3(0, 0, 0), api.VEC3(1, 1, 1), 4.0, 1.0, 0.1, 5, 0, 0, 0)
    L.ltr_LightAdd(s, C.byref(li))
    assert libc.rand() == first[1]
    L.ltr_DestroyScene(s)


def test_ltr_sleep():
    import time
    t0 = time.perf_counter()
    api.lib().ltr_Sleep(20)
    assert 0.015 < time.perf_counter() - t0 < 0.5


def test_failed_bake_is_visible_through_the_plain_api():
    """Without a CUDA device (this suite runs on CPU) a bake must FAIL LOUDLY -- there is no CPU fallback -- and a caller that
    only knows lighter.h must be able to see it: ltr_GetStatus ends the polling loop with the stage "failed: ...", no outputs."""
    import ctypes as C
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a CUDA device is present: nothing fails")
    except ImportError:
        pass
    from lighter_b200 import scenes
    h = api.BakeHandle(scenes.scene_basic())
    L = api.lib()
    st = api.WorkStatus()
    L.ltr_Start(h.h)
    for _ in range(20000):
        if not L.ltr_GetStatus(h.h, C.byref(st)):
            break
        L.ltr_Sleep(1)
    assert not L.ltr_GetStatus(h.h, C.byref(st))
    assert st.stage.startswith(b"failed: "), st.stage
    assert b"CUDA" in st.stage or b"device" in st.stage
    info = api.WorkOutputInfo()
    L.ltr_GetWorkOutputInfo(h.h, C.byref(info))
    assert info.lightmap_count == 0
    assert L.ltrx_GetError(h.h).decode() in st.stage.decode()
    h.close()
