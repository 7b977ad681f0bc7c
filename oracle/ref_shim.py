"""TEST INFRASTRUCTURE ONLY -- import the reference's own Python modules.

Used by ``oracle/make_golden.py`` and by the oracle-pinning tests that run in the
build container, where the read-only reference checkout is mounted at
``/root/reference``.  That directory does not exist on the GPU box, so nothing in the
``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call into this module.

The reference imports four third-party packages that are not installed here
(pytorch3d, spconv, open3d, xgboost) plus ``addict``.  None of them is touched by the
functions on the hot path that we pin, so empty stub modules
are enough to make ``ops.ops_numba``, ``ops.ops_torch``, ``model.utils`` and
``model.PointPillars`` import unmodified.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "ops", "ops_numba.py"))


_loaded = None


def load():
    """Return a namespace with the reference modules (ops_numba, ops_torch, utils, pp)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference checkout not found at %s" % REFERENCE_ROOT)
    import torch

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if "pytorch3d" not in sys.modules:
        p3d = types.ModuleType("pytorch3d")
        p3d._C = types.SimpleNamespace(iou_box3d=None)
        sys.modules["pytorch3d"] = p3d
    for name in ("open3d", "xgboost"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if "addict" not in sys.modules:
        ad = types.ModuleType("addict")
        ad.Dict = dict
        sys.modules["addict"] = ad
    if "spconv" not in sys.modules:
        sp = types.ModuleType("spconv")
        spp = types.ModuleType("spconv.pytorch")
        sp.pytorch = spp
        sys.modules["spconv"] = sp
        sys.modules["spconv.pytorch"] = spp
    # model/utils.py:403 hard-codes .cuda(); on a CPU-only host make it the identity
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self

    import ops.ops_numba as ops_numba
    import ops.ops_torch as ops_torch
    import model.utils as model_utils
    import model.PointPillars as pp

    _loaded = types.SimpleNamespace(
        ops_numba=ops_numba, ops_torch=ops_torch, utils=model_utils, pp=pp
    )
    return _loaded
