"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/pp_b200.h declares
(no compute calls without a GPU).  Argument validation that needs no device is exercised too."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from objectdetection_3d_b200 import _lib, build
    build.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pp_b200.h")).read()
    return sorted(set(re.findall(r"PP_API [^;(]*?\b(pp_\w+)\(", text)))


def test_header_symbols_exported(lib):
    from objectdetection_3d_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.SIGNATURES) == names, "ctypes signature table out of sync with the header"


def test_sass_is_sm100a_only():
    from objectdetection_3d_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_argument_validation_without_gpu(lib):
    from objectdetection_3d_b200 import _lib
    assert lib.pp_version() >= 100
    cfg = _lib.VoxelCfg()
    for i, v in enumerate([0, -39.68, -3, 69.12, 39.68, 1]):
        cfg.range[i] = v
    for i, v in enumerate([0.16, 0.16, 4]):
        cfg.vsize[i] = v
    cfg.grid[0], cfg.grid[1], cfg.grid[2] = 432, 496, 1
    cfg.max_points, cfg.max_voxels, cfg.num_feats = 32, 12000, 4
    assert lib.pp_voxelize_max_rows(1_000_000, ctypes.byref(cfg)) == 12000
    assert lib.pp_voxelize_max_rows(500, ctypes.byref(cfg)) == 500
    assert lib.pp_voxelize_workspace_bytes(1_000_000, ctypes.byref(cfg), 1) > \
        lib.pp_voxelize_workspace_bytes(1_000_000, ctypes.byref(cfg), 0) > 4_000_000
    # two-level NMS: a 2048^2 bitmask for the best boxes + a (N-2048)^2 one for the filtered rest
    assert lib.pp_nms_workspace_bytes(20000) >= (2048 * 32 + 17952 * 281) * 8
    assert lib.pp_nms_workspace_bytes_mode(20000, 1) > lib.pp_nms_workspace_bytes_mode(20000, 0) == lib.pp_nms_workspace_bytes(20000)
    # invalid arguments are rejected before any CUDA call
    rc = lib.pp_voxelize(None, 10, ctypes.byref(cfg), 7, None, None, None, None, ctypes.c_void_p(8), None, None, 0, None)
    assert rc == _lib.PP_ERR_INVALID and b"order" in lib.pp_last_error()
    rc = lib.pp_bbox_iou2d(None, 4, None, 4, 9, 1e-6, None, None)
    assert rc == _lib.PP_ERR_INVALID
    with pytest.raises(ValueError):
        _lib.check(rc)
    with pytest.raises(AssertionError):
        _lib.check(rc, AssertionError)


def test_voxel_cfg_regimes():
    import numpy as np
    from objectdetection_3d_b200.ops_numba import voxel_cfg
    c = voxel_cfg(np.float32, [0.1, 0.1, 0.3], [0, 0, 0, 40.0, 40.0, 30.0], 50, 100, 4)      # lists -> f32 regime
    assert (c.range_is_f64, c.vsize_is_f64) == (0, 0) and list(c.grid) == [400, 400, 100]
    c = voxel_cfg(np.float32, np.array([0.1, 0.1, 0.3], np.float32), np.array([0, 0, 0, 40.0, 40.0, 30.0]), 50, 100, 4)
    assert (c.range_is_f64, c.vsize_is_f64) == (1, 0) and list(c.grid) == [400, 400, 100]
    assert c.vsize[0] == float(np.float32(0.1))
    c = voxel_cfg(np.float32, np.array([0.16, 0.16, 4]), np.array([0, -39.68, -3, 69.12, 39.68, 1], np.float32), 32, 10, 4)
    assert (c.range_is_f64, c.vsize_is_f64) == (0, 1) and list(c.grid) == [432, 496, 1]
    c = voxel_cfg(np.float32, [1, 1, 1], np.array([0, 0, 0, 4, 4, 4]), 5, 10, 4)               # int64 range -> f64 math
    assert c.range_is_f64 == 1
