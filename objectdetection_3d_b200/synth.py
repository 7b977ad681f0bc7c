"""Seeded synthetic inputs for the parity tests and the benchmark.

numpy only.  Columns of a point tile are [x, y, z, reflectance]; reflectance is
``permutation(N) / N`` (tie-free and distinct in float32 for N <= 2**23) unless
``ties=True`` (quantised to 256 levels like real LiDAR intensity).
"""
import numpy as np

# Geometries: G_ref is config.yaml verbatim, G_kitti the BASELINE.json shapes.
G_REF = dict(point_cloud_range=[0, 0, 0, 40.0, 40.0, 30.0], voxel_size=[0.1, 0.1, 0.3],
             max_voxel_points=50, max_voxels=7500000)
G_REF_PILLAR = dict(point_cloud_range=[0, 0, 0, 40.0, 40.0, 30.0], voxel_size=[0.16, 0.16, 30.0],
                    max_voxel_points=50, max_voxels=7500000)
G_KITTI = dict(point_cloud_range=[0, -39.68, -3, 69.12, 39.68, 1], voxel_size=[0.16, 0.16, 4],
               max_voxel_points=32, max_voxels=12000)

ANCHOR_SIZES = [[0.75, 0.75, 12], [1.3, 1.3, 17], [1.0, 1.75, 20]]          # config.yaml:65
ANCHOR_ROTATIONS = [[0.0, 0.0, 0.0], [0.0, 0.0, 1.57], [0.3142, 0.0, 0.0], [-0.3142, 0.0, 0.0]]  # :66


def _reflectance(rng, n, ties):
    if ties:
        return (rng.integers(0, 256, size=n) / 255.0).astype(np.float32)
    return (rng.permutation(n) / float(n)).astype(np.float32)


def forest_tile(n=120_000, seed=1234, point_cloud_range=G_REF["point_cloud_range"], ties=False,
                n_trunks=150):
    """F120k: 30 % ground, 40 % trunk cylinders, 30 % canopy blobs, clipped to the range."""
    rng = np.random.default_rng(seed)
    x0, y0, z0, x1, y1, z1 = [float(v) for v in point_cloud_range]
    ng, nt = int(0.3 * n), int(0.4 * n)
    nc = n - ng - nt
    ground = np.stack([rng.uniform(x0, x1, ng), rng.uniform(y0, y1, ng),
                       z0 + np.abs(rng.normal(0, 0.15, ng))], 1)
    mx, my = 0.05 * (x1 - x0), 0.05 * (y1 - y0)
    cx = rng.uniform(x0 + mx, x1 - mx, n_trunks)
    cy = rng.uniform(y0 + my, y1 - my, n_trunks)
    rad = rng.uniform(0.1, 0.4, n_trunks)
    hgt = rng.uniform(0.4, 0.83, n_trunks) * (z1 - z0)
    t = rng.integers(0, n_trunks, nt)
    ang = rng.uniform(0, 2 * np.pi, nt)
    trunk = np.stack([cx[t] + rad[t] * np.cos(ang) + rng.normal(0, 0.02, nt),
                      cy[t] + rad[t] * np.sin(ang) + rng.normal(0, 0.02, nt),
                      z0 + rng.uniform(0, 1, nt) * hgt[t]], 1)
    t = rng.integers(0, n_trunks, nc)
    canopy = np.stack([cx[t], cy[t], z0 + hgt[t]], 1) + rng.normal(0, 1.5, (nc, 3))
    xyz = np.concatenate([ground, trunk, canopy], 0)
    eps = 1e-3
    xyz[:, 0] = np.clip(xyz[:, 0], x0 + eps, x1 - eps)
    xyz[:, 1] = np.clip(xyz[:, 1], y0 + eps, y1 - eps)
    xyz[:, 2] = np.clip(xyz[:, 2], z0 + eps, z1 - eps)
    xyz = xyz[rng.permutation(n)]
    pts = np.empty((n, 4), dtype=np.float32)
    pts[:, :3] = xyz.astype(np.float32)
    pts[:, 3] = _reflectance(rng, n, ties)
    return pts


def dense_tile(n=1_000_000, seed=2024, geom=G_KITTI, n_cells=11_500, n_clusters=600, ties=False):
    """D1M: n points inside n_cells occupied pillars drawn as disc clusters; per-cell counts
    follow a lognormal(0, 1) weight (about 70 % of the pillars saturate P = 32 at n = 1e6)."""
    rng = np.random.default_rng(seed)
    rg = np.asarray(geom["point_cloud_range"], dtype=np.float64)
    vs = np.asarray(geom["voxel_size"], dtype=np.float64)
    gx, gy = int(round((rg[3] - rg[0]) / vs[0])), int(round((rg[4] - rg[1]) / vs[1]))
    n_cells = min(n_cells, gx * gy)
    # grow disc clusters until n_cells distinct cells are occupied
    cells = set()
    ccx = rng.uniform(0, gx, n_clusters)
    ccy = rng.uniform(0, gy, n_clusters)
    rad = 1.0
    while len(cells) < n_cells:
        k = rng.integers(0, n_clusters, 4 * n_cells)
        r = rad * np.sqrt(rng.uniform(0, 1, k.size))
        a = rng.uniform(0, 2 * np.pi, k.size)
        ix = np.floor(ccx[k] + r * np.cos(a)).astype(np.int64)
        iy = np.floor(ccy[k] + r * np.sin(a)).astype(np.int64)
        ok = (ix >= 0) & (ix < gx) & (iy >= 0) & (iy < gy)
        for c in (ix[ok] * gy + iy[ok]).tolist():
            cells.add(c)
            if len(cells) >= n_cells:
                break
        rad += 1.0
    cells = np.fromiter(cells, dtype=np.int64, count=n_cells)
    cells.sort()
    cells = cells[rng.permutation(n_cells)]
    w = rng.lognormal(0.0, 1.0, n_cells)
    which = rng.choice(n_cells, size=n, p=w / w.sum())
    ix, iy = cells[which] // gy, cells[which] % gy
    u = rng.uniform(0.05, 0.95, (n, 2))
    pts = np.empty((n, 4), dtype=np.float32)
    pts[:, 0] = (rg[0] + (ix + u[:, 0]) * vs[0]).astype(np.float32)
    pts[:, 1] = (rg[1] + (iy + u[:, 1]) * vs[1]).astype(np.float32)
    pts[:, 2] = rng.uniform(rg[2] + 1e-3, rg[5] - 1e-3, n).astype(np.float32)
    pts[:, 3] = _reflectance(rng, n, ties)
    return pts


def uniform_tile(n=1_000_000, seed=7, geom=G_KITTI, ties=False, margin=0.0):
    """D1M-overflow: uniform over the range (>> max_voxels occupied cells -> exercises the break);
    margin > 0 also emits points outside the range."""
    rng = np.random.default_rng(seed)
    rg = np.asarray(geom["point_cloud_range"], dtype=np.float64)
    ext = (rg[3:] - rg[:3]) * margin
    pts = np.empty((n, 4), dtype=np.float32)
    for j in range(3):
        pts[:, j] = rng.uniform(rg[j] - ext[j], rg[3 + j] + ext[j], n).astype(np.float32)
    pts[:, 3] = _reflectance(rng, n, ties)
    return pts


def nms_boxes(n=20_000, seed=4, extent=40.0, tilt=0.3):
    """NMS20k: 9-parameter boxes [x,y,z,dx,dy,dz,rx,ry,rz] with unique scores."""
    rng = np.random.default_rng(seed)
    sizes = np.asarray(ANCHOR_SIZES, dtype=np.float64)
    b = np.zeros((n, 9), dtype=np.float64)
    b[:, 0:2] = rng.uniform(0, extent, (n, 2))
    b[:, 3:6] = sizes[rng.integers(0, len(sizes), n)] * np.exp(rng.normal(0, 0.1, (n, 3)))
    b[:, 6:8] = rng.uniform(-tilt, tilt, (n, 2)) if tilt > 0 else 0.0
    b[:, 8] = rng.uniform(0, np.pi, n)
    scores = ((rng.permutation(n) + 0.5) / n).astype(np.float32).reshape(n, 1)
    return b.astype(np.float32), scores


def pfn_params(c_in=9, units=63, seed=11):
    """Random PFNLayer weights + BatchNorm running statistics (eval mode)."""
    rng = np.random.default_rng(seed)
    return dict(weight=(rng.normal(0, 0.3, (units, c_in))).astype(np.float32),
                gamma=rng.uniform(0.5, 1.5, units).astype(np.float32),
                beta=rng.normal(0, 0.2, units).astype(np.float32),
                mean=rng.normal(0, 0.5, units).astype(np.float32),
                var=rng.uniform(0.3, 2.0, units).astype(np.float32))
