"""Swap the reference's hot-path symbols for the libpp_b200-backed ones, in place.

    import objectdetection_3d_b200.dropin as dropin
    dropin.install()        # the reference checkout must be importable (ops/, model/ on sys.path)

model/PointPillars.py binds VoxelGenerator, BBoxCoder, multiclass_nms, ... by name at import time
(model/PointPillars.py:16-18) and late-imports the IoU ops inside functions (model/utils.py:368-374,
model/PointPillars.py:899-905), so both the defining modules and the importing module are patched.
"""
import importlib

from . import model_utils, ops_numba, ops_numpy, ops_torch, pointpillars

PATCHES = {
    "ops.ops_numba": (ops_numba, ["points_to_voxel", "VoxelGenerator", "CustomVoxelGenerator", "iou_jit"]),
    "ops.ops_torch": (ops_torch, ["bbox2rotated_corners2D", "bbox2corners3D", "bbox_iou2D", "box3d_overlap",
                                  "check_coplanar", "check_nonzero"]),
    "ops.ops_numpy": (ops_numpy, ["global_outlier_check"]),
    "model.utils": (model_utils, ["Anchor3DRangeGenerator", "BBoxCoder", "limit_period", "multiclass_nms",
                                  "get_paddings_indicator", "CustomVoxelizer"]),
    "model.PointPillars": (pointpillars, ["PointPillarsVoxelization", "PFNLayer", "PillarFeatureNet", "Anchor3DHead"]),
}
# names model/PointPillars.py imported from the modules above
REBIND_IN_POINTPILLARS = {"VoxelGenerator": ops_numba, "Anchor3DRangeGenerator": model_utils, "BBoxCoder": model_utils,
                          "limit_period": model_utils, "multiclass_nms": model_utils,
                          "get_paddings_indicator": model_utils}

_saved = {}


def install():
    """Returns the list of (module, name) pairs that were replaced."""
    done = []
    for modname, (src, names) in PATCHES.items():
        mod = importlib.import_module(modname)
        for n in names:
            _saved.setdefault((modname, n), getattr(mod, n))
            setattr(mod, n, getattr(src, n))
            done.append((modname, n))
    pp = importlib.import_module("model.PointPillars")
    for n, src in REBIND_IN_POINTPILLARS.items():
        if hasattr(pp, n):
            _saved.setdefault(("model.PointPillars", n), getattr(pp, n))
            setattr(pp, n, getattr(src, n))
            done.append(("model.PointPillars", n))
    return done


def uninstall():
    for (modname, n), obj in _saved.items():
        setattr(importlib.import_module(modname), n, obj)
    _saved.clear()
