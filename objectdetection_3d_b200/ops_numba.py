"""Drop-in for the reference's ``ops/ops_numba.py`` (same names, argument order and returns),
backed by the sm_100a kernels of libpp_b200 (csrc/pp_voxelize.cu, csrc/pp_boxes.cu).

``points`` may be a numpy array (the reference's contract: host in, host out -- the upload and the
read-back happen here) or a CUDA tensor (device in, device out; the only synchronisation is the
read of the pillar count that bounds the returned slices, where the reference slices at
ops/ops_numba.py:164-166).
"""
import ctypes

import numpy as np
import torch

from . import _lib


def _dev():
    if not torch.cuda.is_available():
        raise _lib.PPError("objectdetection_3d_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def voxel_cfg(points_dtype, voxel_size, coors_range, max_points, max_voxels, num_feats):
    """ops/ops_numba.py:139-145: lists are cast to points.dtype, ndarrays keep theirs (numba then
    promotes per operation); grid = round((range[3:] - range[:3]) / voxel_size)."""
    np_dtype = np.float32 if points_dtype in (torch.float32, np.float32, np.dtype("float32")) else np.dtype(points_dtype)
    if isinstance(voxel_size, torch.Tensor):
        voxel_size = voxel_size.detach().cpu().numpy()
    if isinstance(coors_range, torch.Tensor):
        coors_range = coors_range.detach().cpu().numpy()
    if not isinstance(voxel_size, np.ndarray):
        voxel_size = np.array(voxel_size, dtype=np_dtype)
    if not isinstance(coors_range, np.ndarray):
        coors_range = np.array(coors_range, dtype=np_dtype)
    grid = np.round((coors_range[3:] - coors_range[:3]) / voxel_size).astype(np.int32)
    cfg = _lib.VoxelCfg()
    for i in range(6):
        cfg.range[i] = float(coors_range[i])
    for i in range(3):
        cfg.vsize[i] = float(voxel_size[i])
        cfg.grid[i] = int(grid[i])
    cfg.range_is_f64 = int(coors_range.dtype not in (np.float32, np.float16))
    cfg.vsize_is_f64 = int(voxel_size.dtype not in (np.float32, np.float16))
    cfg.max_points = int(max_points)
    cfg.max_voxels = int(min(int(max_voxels), 2 ** 31 - 1))
    cfg.num_feats = int(num_feats)
    return cfg


def voxelize_device(points, cfg, order, perm=None, want_map=False):
    """Device-side core: returns (voxels, coors xyz int32, num int32, voxel_num device scalar, map)."""
    lib = _lib.load()
    assert points.is_cuda and points.dtype == torch.float32 and points.dim() == 2
    points = points.contiguous()
    n, c = points.shape
    dev = points.device
    rows = int(lib.pp_voxelize_max_rows(n, ctypes.byref(cfg)))
    voxels = torch.empty((rows, cfg.max_points, c), dtype=torch.float32, device=dev)
    coors = torch.empty((rows, 3), dtype=torch.int32, device=dev)
    num = torch.empty((rows,), dtype=torch.int32, device=dev)
    voxel_num = torch.empty((1,), dtype=torch.int32, device=dev)
    ws_bytes = int(lib.pp_voxelize_workspace_bytes(n, ctypes.byref(cfg), order))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    pmap = None
    if want_map:
        pmap = torch.empty((cfg.grid[2], cfg.grid[1], cfg.grid[0]), dtype=torch.int32, device=dev)
    if perm is not None:
        perm = perm.to(device=dev, dtype=torch.int32).contiguous()
    rc = lib.pp_voxelize(_ptr(points), n, ctypes.byref(cfg), order, _ptr(perm), _ptr(voxels), _ptr(coors), _ptr(num),
                         _ptr(voxel_num), _ptr(pmap), _ptr(ws), ws_bytes, _stream())
    _lib.check(rc)
    return voxels, coors, num, voxel_num, pmap


_numba_order = None


def numba_reflectance_order(reflectance):
    """The reference's own pre-order, ``points[:, 3].argsort()[::-1]`` under numba.njit (ops/ops_numba.py:262), computed
    with numba itself on the host: among EQUAL reflectances numba's quicksort order is what the reference follows, and
    it is not reproducible in parallel.  Only for ``exact_ties=True``; needs numba (the reference's own dependency)."""
    global _numba_order
    if _numba_order is None:
        import numba

        @numba.njit(cache=False)
        def order(a):
            return a.argsort()[::-1]
        _numba_order = order
    refl = np.ascontiguousarray(reflectance.detach().cpu().numpy() if isinstance(reflectance, torch.Tensor) else reflectance)
    return np.ascontiguousarray(_numba_order(refl)).astype(np.int32)


def points_to_voxel(points, voxel_size, coors_range, max_points, max_voxels, reflectance_sampling, perm=None,
                    exact_ties=False):
    """ops/ops_numba.py:109-168.  Returns (voxels f32 [M,P,C], coors int32 [M,3] xyz, num int32 [M]).

    reflectance_sampling=True: points[:, 3] descending.  Among EQUAL reflectances (real LiDAR intensity is quantised,
    so ties are the rule) the GPU order is: lower index first, -0.0 == +0.0 -- deterministic, but not the order the
    reference's numba quicksort happens to produce, so with ties the kept subset of an over-full pillar, the slot order
    and the pillar ids can differ from the reference's.  ``exact_ties=True`` computes the reference's order with numba
    on the host (one CPU argsort per call) and replays it: bit-exact with the reference on any input.  Any order can
    be replayed through ``perm``.  reflectance_sampling=False: the reference shuffles the caller's array in place
    (:190) and then takes the given order; so does this function.
    """
    is_numpy = isinstance(points, np.ndarray)
    if exact_ties and reflectance_sampling and perm is None:
        perm = numba_reflectance_order(points[:, 3])
    if is_numpy:
        if not reflectance_sampling and perm is None:
            np.random.shuffle(points)                                   # same side effect as :190
        dpts = torch.from_numpy(np.ascontiguousarray(points, dtype=np.float32)).to(_dev(), non_blocking=True)
    else:
        dpts = points
        if dpts.dtype != torch.float32:
            dpts = dpts.float()
        if not reflectance_sampling and perm is None:
            shuffled = dpts[torch.randperm(dpts.shape[0], device=dpts.device)]
            points.copy_(shuffled)
            dpts = points if points.dtype == torch.float32 else shuffled
    cfg = voxel_cfg(np.float32, voxel_size, coors_range, max_points, max_voxels, dpts.shape[1])
    if perm is not None:
        order = _lib.ORDER_PERM
        perm = torch.as_tensor(np.asarray(perm) if not isinstance(perm, torch.Tensor) else perm)
    else:
        order = _lib.ORDER_REFLECTANCE_DESC if reflectance_sampling else _lib.ORDER_GIVEN
    voxels, coors, num, voxel_num, _ = voxelize_device(dpts, cfg, order, perm)
    m = int(voxel_num.item())                                            # the slice bound of :164-166
    voxels, coors, num = voxels[:m], coors[:m], num[:m]
    if is_numpy:
        return voxels.cpu().numpy(), coors.cpu().numpy(), num.cpu().numpy()
    return voxels, coors, num


class VoxelGenerator:
    """ops/ops_numba.py:40-81"""

    def __init__(self, voxel_size, point_cloud_range, max_voxel_points, max_voxels):
        point_cloud_range = np.array(point_cloud_range, dtype=np.float32)
        voxel_size = np.array(voxel_size, dtype=np.float32)
        grid_size = (point_cloud_range[3:] - point_cloud_range[:3]) / voxel_size
        self._grid_size = np.round(grid_size).astype(np.int64)
        self._voxel_size = voxel_size
        self._point_cloud_range = point_cloud_range
        self._max_voxel_points = max_voxel_points
        self._max_voxels = max_voxels

    def generate(self, points, max_voxels, cloud_range, reflectance_sampling, exact_ties=False):
        return points_to_voxel(points, self._voxel_size, cloud_range, self._max_voxel_points, max_voxels,
                               reflectance_sampling, exact_ties=exact_ties)

    @property
    def voxel_size(self):
        return self._voxel_size

    @property
    def max_num_points_per_voxel(self):
        return self._max_voxel_points

    @property
    def point_cloud_range(self):
        return self._point_cloud_range

    @property
    def grid_size(self):
        return self._grid_size


class CustomVoxelGenerator:
    """ops/ops_numba.py:83-106"""

    def __init__(self, voxel_size, max_voxel_points, reflectance_sampling):
        self._voxel_size = np.array(voxel_size, dtype=np.float32)
        self._max_voxel_points = max_voxel_points
        self._reflectance_sampling = reflectance_sampling

    def generate(self, points, point_cloud_range, max_voxels):
        return points_to_voxel(points, self._voxel_size, point_cloud_range, self._max_voxel_points, max_voxels,
                               self._reflectance_sampling)

    @property
    def voxel_size(self):
        return self._voxel_size

    @property
    def max_num_points_per_voxel(self):
        return self._max_voxel_points


def iou_jit(boxes, query_boxes, eps=0.0):
    """ops/ops_numba.py:7-36: (N,4),(K,4) -> (N,K) overlaps."""
    lib = _lib.load()
    is_numpy = isinstance(boxes, np.ndarray)
    b = torch.as_tensor(boxes, dtype=torch.float32).to(_dev()).contiguous()
    q = torch.as_tensor(query_boxes, dtype=torch.float32).to(_dev()).contiguous()
    out = torch.empty((b.shape[0], q.shape[0]), dtype=torch.float32, device=b.device)
    _lib.check(lib.pp_iou_jit(_ptr(b), b.shape[0], _ptr(q), q.shape[0], float(eps), _ptr(out), _stream()))
    return out.cpu().numpy() if is_numpy else out
