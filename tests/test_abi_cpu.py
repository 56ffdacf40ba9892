"""CPU suite: the C-ABI library builds for sm_100a, loads, and exports every symbol include/rf_abi.h declares
(no compute calls without a GPU); host-side argument checks that need no device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "rf_abi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(rf_lib):
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(rf_lib, n), f"{n} declared in include/rf_abi.h but not exported by librf_b200.so"


def test_version_and_error_string(rf_lib):
    assert rf_lib.rf_version() == 1
    rc = rf_lib.rf_tsdf_clear_global(C.c_void_p(0), C.c_int64(8), C.c_void_p(0))
    assert rc == -1 and b"NULL" in rf_lib.rf_last_error()
    rc = rf_lib.rf_tsdf_clear_local(C.c_void_p(8), C.c_void_p(8), C.c_void_p(8), C.c_int64(-1), C.c_void_p(0))
    assert rc == -2


def test_argument_validation_without_a_gpu(rf_lib):
    """Bad arguments are rejected before anything is launched (no device needed): NULL pointers, ranges, alignment."""
    import numpy as np
    f = lambda *v: np.asarray(v, np.float32).ctypes.data_as(C.POINTER(C.c_float))
    null, p16, p4 = C.c_void_p(0), C.c_void_p(64), C.c_void_p(68)
    K = np.eye(3, dtype=np.float32).reshape(-1); Kp = K.ctypes.data_as(C.POINTER(C.c_float))
    # tracker maps: NULL, image too small, misaligned vertex buffer
    assert rf_lib.rf_track_vertex_normal(null, 8, 8, Kp, C.c_float(6), C.c_float(.06), 1, C.c_float(3), p16, p16, p16, null) == -1
    assert rf_lib.rf_track_vertex_normal(p16, 2, 8, Kp, C.c_float(6), C.c_float(.06), 1, C.c_float(3), p16, p16, p16, null) == -2
    assert rf_lib.rf_track_vertex_normal(p16, 8, 8, Kp, C.c_float(6), C.c_float(.06), 1, C.c_float(3), p16, p4, p16, null) == -3
    assert b"16-byte" in rf_lib.rf_last_error()
    # candidate reduction: NULL / empty
    ss = np.ones(6, np.float32); ssp = ss.ctypes.data_as(C.POINTER(C.c_float))
    assert rf_lib.rf_track_cal_transform(null, p16, p16, 4, ssp, 3, p16, null) == -1
    assert rf_lib.rf_track_cal_transform(p16, p16, p16, 0, ssp, 3, p16, null) == -2
    # scratch sizes are pure host functions
    rf_lib.rf_track_fitness_scratch_floats.restype = C.c_int64
    assert rf_lib.rf_track_fitness_scratch_floats(0, 680, 1200, 8) == 0
    assert rf_lib.rf_track_fitness_scratch_floats(1024, 680, 1200, 8) % (2 * 1024) == 0


def test_grid_desc_init_matches_tcnn_table(rf_lib):
    import numpy as np
    from oracle.tcnn_standin import grid_levels
    from remixfusion_b200.encodings import make_grid_desc
    for hs, res in ((16, 400), (19, 512), (21, 1750), (19, 5000), (12, 250)):
        pls = np.exp2(np.log2(res / 16) / 15)                                   # model/encodings.py:36
        s, r, n, o = grid_levels(16, 2, True, hs, 16, pls)
        d = make_grid_desc(16, 2, True, hs, 16, pls)
        assert [d.scale[i] for i in range(16)] == [float(v) for v in s]
        assert list(d.resolution[:16]) == r and list(d.size[:16]) == n and list(d.offset[:17]) == o
    d = make_grid_desc(1, 4, False, 0, 200, 1)                                   # GBV: model/scene_rep.py:60-74
    assert d.scale[0] == 199.0 and d.resolution[0] == 200 and d.offset[1] == 8000000 and d.n_params == 32000000
    from remixfusion_b200 import abi
    bad = abi.GridDesc()
    assert rf_lib.rf_grid_desc_init(C.byref(bad), 17, 2, 1, 19, 16, C.c_double(1.5)) == -2
    assert rf_lib.rf_grid_desc_init(C.byref(bad), 16, 3, 1, 19, 16, C.c_double(1.5)) == -4


def test_product_never_imports_oracle():
    """The product path must not route through oracle/ (tier rule ③)."""
    pkg = os.path.join(ROOT, "remixfusion_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports oracle"
                assert "liboracle" not in txt, f


def test_missing_library_fails_loudly(monkeypatch):
    from remixfusion_b200 import abi
    monkeypatch.setattr(abi, "_lib", None)
    monkeypatch.setattr(abi, "LIB_PATH", "/nonexistent/librf_b200.so")
    with pytest.raises(abi.RfError):
        abi.lib()


def test_mapvolume_rejects_mismatched_parameter_sizes():
    """A z-slab object must own slab-sized parameter tensors and the unsharded one the full grids: a full-size model passed
    together with a z_slab would be written at the wrong offset (checked at construction, no device needed)."""
    import numpy as np
    import pytest
    import torch
    from remixfusion_b200 import abi
    from remixfusion_b200.global_volume import MapVolume
    R = 16
    cfg = {"globalV": {"base_resolution": R}, "mapping": {"bound": [[0, 1], [0, 1], [0, 1]]}, "training": {"c_trunc": 0.1}}

    def model(nvox):
        m = type("M", (), {})(); m.GBV = type("E", (), {})(); m.GBW = type("E", (), {})()
        m.GBV.params = torch.zeros(4 * nvox); m.GBW.params = torch.zeros(nvox)
        return m
    K = np.eye(3)
    MapVolume(cfg, model(R ** 3), K)                                     # full volume, exact size
    MapVolume(cfg, model(R ** 3 + 8), K)                                 # tcnn pads to 8 entries
    MapVolume(cfg, model(4 * R * R), K, z_slab=(4, 8))                   # slab-sized tensors with a slab
    with pytest.raises(abi.RfError, match="z-slab"):
        MapVolume(cfg, model(R ** 3), K, z_slab=(4, 8))                  # full-size model + slab: refused
    with pytest.raises(abi.RfError, match="full volume"):
        MapVolume(cfg, model(4 * R * R), K)                              # slab-sized tensors without a slab
    with pytest.raises(abi.RfError, match="outside"):
        MapVolume(cfg, model(4 * R * R), K, z_slab=(14, 18))


def test_graphed_step_needs_eager_iterations():
    import pytest
    import torch
    from remixfusion_b200 import abi
    from remixfusion_b200.graph import GraphedMappingStep
    from remixfusion_b200.optim import Adam
    p = torch.nn.Parameter(torch.zeros(4))
    opt = Adam([p], capturable=True)
    with pytest.raises(abi.RfError, match="eager_steps"):
        GraphedMappingStep(torch.nn.Linear(2, 2), opt, 16, lambda r: 0, eager_steps=0)


def test_ray_chunk_policy(monkeypatch):
    """scene_rep._chunk_rays: the whole batch unless the planes cannot fit; RF_RAY_CHUNK forces a multiple of 128 (one decoder tile)."""
    import torch
    from remixfusion_b200 import scene_rep
    cpu = torch.device("cpu")
    monkeypatch.delenv("RF_RAY_CHUNK", raising=False)
    assert scene_rep._chunk_rays(None, None, 1000, 59, cpu) == 1000
    monkeypatch.setenv("RF_RAY_CHUNK", "300")
    assert scene_rep._chunk_rays(None, None, 1000, 59, cpu) == 256
    monkeypatch.setenv("RF_RAY_CHUNK", "5")
    assert scene_rep._chunk_rays(None, None, 1000, 59, cpu) == 128
