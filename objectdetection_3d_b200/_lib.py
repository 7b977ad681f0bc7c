"""ctypes binding of libpp_b200.so (the C ABI declared in include/pp_b200.h).

There is no CPU fallback: if the library is missing, or a call fails, this raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PP_B200_LIB") or os.path.join(_HERE, "libpp_b200.so")      # (override: development builds)

PP_OK, PP_ERR_INVALID, PP_ERR_CUDA, PP_ERR_WORKSPACE = 0, -1, -2, -3
ORDER_GIVEN, ORDER_REFLECTANCE_DESC, ORDER_PERM = 0, 1, 2
NMS_AABB2D, NMS_ROT_BEV, NMS_BOX3D = 0, 1, 2
COORS_XYZ_I32, COORS_BZYX_I32, COORS_BZYX_I64 = 0, 1, 2
NUM_I32, NUM_I64 = 0, 1
IOU_MODES = {"iou": 0, "iof": 1, "giou": 2}


class VoxelCfg(ctypes.Structure):
    """pp_voxel_cfg"""
    _fields_ = [("range", ctypes.c_double * 6), ("vsize", ctypes.c_double * 3),
                ("range_is_f64", ctypes.c_int32), ("vsize_is_f64", ctypes.c_int32),
                ("grid", ctypes.c_int32 * 3), ("max_points", ctypes.c_int32),
                ("max_voxels", ctypes.c_int32), ("num_feats", ctypes.c_int32)]


_vp, _i64, _i32, _f32, _f64, _sz = (ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_float,
                                     ctypes.c_double, ctypes.c_size_t)
_cfgp = ctypes.POINTER(VoxelCfg)

# name -> (restype, argtypes); must list every function declared in include/pp_b200.h
SIGNATURES = {
    "pp_version": (ctypes.c_int, []),
    "pp_last_error": (ctypes.c_char_p, []),
    "pp_launch_count": (_i64, []),
    "pp_profile_enable": (ctypes.c_int, [ctypes.c_int]),
    "pp_profile_report": (ctypes.c_int, [ctypes.c_char_p, _sz]),
    "pp_voxelize_max_rows": (_i64, [_i64, _cfgp]),
    "pp_voxelize_workspace_bytes": (_sz, [_i64, _cfgp, ctypes.c_int]),
    "pp_voxelize": (ctypes.c_int, [_vp, _i64, _cfgp, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pp_voxelize_features": (ctypes.c_int, [_vp, _i64, _cfgp, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz,
                                            _vp]),
    "pp_voxelize_scatter": (ctypes.c_int, [_vp, _i64, _cfgp, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz,
                                           _vp]),
    "pp_decorate": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _vp, ctypes.c_int, _i64, _vp, ctypes.c_int, ctypes.c_int,
                                   _f32, _f32, _f32, _f32, _vp, _vp]),
    "pp_pfn_layer": (ctypes.c_int, [_vp, _i64, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, ctypes.c_int, ctypes.c_int,
                                    _vp, _vp]),
    "pp_pillar_features": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _vp, ctypes.c_int, _i64, _vp, ctypes.c_int,
                                          ctypes.c_int, _f32, _f32, _f32, _f32, _vp, _vp, _vp, ctypes.c_int, _vp, _vp]),
    "pp_scatter_workspace_bytes": (_sz, [ctypes.c_int] * 4),
    "pp_scatter_dense": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _i64, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp, _sz, _vp]),
    "pp_scatter_mapped": (ctypes.c_int, [_vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         _vp, _vp]),
    "pp_box_encode": (ctypes.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "pp_box_decode": (ctypes.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "pp_limit_period": (ctypes.c_int, [_vp, _i64, _f32, _f32, _vp, _vp]),
    "pp_grid_anchors": (ctypes.c_int, [ctypes.POINTER(_f32), ctypes.POINTER(_f32), ctypes.c_int, ctypes.POINTER(_f32),
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp]),
    "pp_box_corners3d": (ctypes.c_int, [_vp, _i64, _vp, _vp]),
    "pp_box_aabb2d": (ctypes.c_int, [_vp, _i64, _vp, _vp]),
    "pp_bbox_iou2d": (ctypes.c_int, [_vp, _i64, _vp, _i64, ctypes.c_int, _f32, _vp, _vp]),
    "pp_iou_rotated_bev": (ctypes.c_int, [_vp, _i64, _vp, _i64, _vp, _vp]),
    "pp_box3d_overlap": (ctypes.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _vp]),
    "pp_box3d_check": (ctypes.c_int, [_vp, _i64, _f32, _vp, _vp]),
    "pp_assign_overlaps": (ctypes.c_int, [_vp, _i64, _vp, _i64, ctypes.c_int, _f32, _vp, _vp, _vp, _vp, _vp]),
    "pp_head_max_scores": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp]),
    "pp_head_select_decode": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, ctypes.c_int, _vp, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp]),
    "pp_head_direction_fixup": (ctypes.c_int, [_vp, _vp, _i64, _f32, _vp]),
    "pp_head_topk_workspace_bytes": (_sz, [_i64, _i64]),
    "pp_head_topk": (ctypes.c_int, [_vp, _i64, _i64, _vp, _vp, _sz, _vp]),
    "pp_compact_workspace_bytes": (_sz, [_i64]),
    "pp_preprocess_workspace_bytes": (_sz, [_i64]),
    "pp_preprocess_points": (ctypes.c_int, [_vp, _i64, ctypes.c_int, ctypes.c_int, _vp, _vp, ctypes.c_int, _vp, _vp, _vp, _sz, _vp]),
    "pp_points_minmax": (ctypes.c_int, [_vp, _i64, ctypes.c_int, _vp, _vp, _sz, _vp]),
    "pp_voxel_centroids": (ctypes.c_int, [_vp, _vp, _i64, ctypes.c_int, ctypes.c_int, _vp, _vp]),
    "pp_dense_to_sparse": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pp_iou_jit": (ctypes.c_int, [_vp, _i64, _vp, _i64, _f64, _vp, _vp]),
    "pp_nms_workspace_bytes": (_sz, [_i64]),
    "pp_nms": (ctypes.c_int, [_vp, _vp, _i64, _i64, _f32, _f32, _vp, _vp, _vp, _sz, _vp]),
    "pp_nms_workspace_bytes_mode": (_sz, [_i64, ctypes.c_int]),
    "pp_nms_mode": (ctypes.c_int, [_vp, _vp, _i64, _i64, _f32, _f32, ctypes.c_int, _vp, _vp, _vp, _sz, _vp]),
    "pp_sort_workspace_bytes": (_sz, [_i64]),
    "pp_sort_pairs_u32": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _sz, _vp]),
}

_lib = None


class PfnFused(ctypes.Structure):
    """pp_pfn_fused of include/pp_b200.h"""
    _fields_ = [("weight", ctypes.c_void_p), ("scale", ctypes.c_void_p), ("shift", ctypes.c_void_p),
                ("units", ctypes.c_int32), ("vx", ctypes.c_float), ("vy", ctypes.c_float), ("x_off", ctypes.c_float),
                ("y_off", ctypes.c_float), ("feat", ctypes.c_void_p)]


class PPError(RuntimeError):
    pass


def load():
    """Load libpp_b200.so; raises (no fallback) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PPError("libpp_b200.so is missing: run `python -m objectdetection_3d_b200.build` "
                          "(there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc, invalid_exc=ValueError):
    """Map a PP_ERR_* code to the exception type the reference would raise."""
    if rc == PP_OK:
        return
    msg = load().pp_last_error().decode("utf-8", "replace")
    if rc == PP_ERR_INVALID:
        raise invalid_exc(msg)
    raise PPError("libpp_b200 error %d: %s" % (rc, msg))


def launch_count():
    return int(load().pp_launch_count())


def profile(on):
    load().pp_profile_enable(1 if on else 0)


def profile_report():
    """{kernel name: (launches, total_ms)} collected since profile(True)."""
    buf = ctypes.create_string_buffer(1 << 16)
    check(load().pp_profile_report(buf, len(buf)))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.rsplit(" ", 2)
        out[name] = (int(cnt), float(ms))
    return out
