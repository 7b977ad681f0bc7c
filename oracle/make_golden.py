"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the REFERENCE itself.

Run in the build container (where /root/reference is mounted):

    python -m oracle.make_golden

Every fixture stores the seeded inputs next to the reference's outputs, so the CPU tests
(oracle vs golden) and the GPU tests (CUDA vs golden) need neither the reference checkout
nor numba at run time.  Versions used are recorded in tests/golden/VERSIONS.json.
"""
import json
import os

import numpy as np
import torch

from objectdetection_3d_b200 import synth
from oracle import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("%-28s %8.1f kB" % (name, os.path.getsize(path) / 1024))


def voxel_cases(R):
    p2v = R.ops_numba.points_to_voxel
    small = dict(point_cloud_range=[0, -8.0, -3, 16.0, 8.0, 1], voxel_size=[0.16, 0.16, 4])

    def run(name, pts, vs, rg, P, cap, refl, note=""):
        pts_in = pts.copy()
        v, c, n = p2v(pts, vs, rg, P, cap, refl)
        # reflectance_sampling=False shuffles the caller's array in place (ops_numba.py:190):
        # store the post-call order so the pass can be replayed in "given order" mode
        save(name, points=pts_in, points_after=pts, voxel_size=np.asarray(vs), coors_range=np.asarray(rg),
             vs_is_list=np.array(not isinstance(vs, np.ndarray)), rg_is_list=np.array(not isinstance(rg, np.ndarray)),
             max_points=np.array(P), max_voxels=np.array(cap), reflectance=np.array(refl),
             voxels=v, coors=c, num=n)

    f32vs = lambda g: np.array(g["voxel_size"], dtype=np.float32)
    f64rg = lambda g: np.array(g["point_cloud_range"], dtype=np.float64)

    # model path regime (voxel f32, range f64), clustered, no overflow
    pts = synth.dense_tile(n=6000, seed=1, geom=small, n_cells=400, n_clusters=12)
    run("vox_model_clustered", pts, f32vs(small), f64rg(small), 8, 1000, True)
    # cap overflow -> break
    pts = synth.uniform_tile(n=6000, seed=2, geom=small, margin=0.1)
    run("vox_model_overflow", pts, f32vs(small), f64rg(small), 4, 700, True)
    # reflectance ties (numba quicksort tie order)
    pts = synth.dense_tile(n=6000, seed=3, geom=small, n_cells=300, n_clusters=10, ties=True)
    run("vox_model_ties", pts, f32vs(small), f64rg(small), 8, 1000, True)
    # all-f32 regime (lists), 3-D grid, boundary points on multiples of 0.1 (SURVEY 8 V1 KAT)
    rng = np.random.default_rng(4)
    k = np.arange(1, 400, dtype=np.float64) * 0.1
    pts = np.zeros((399 * 3, 4), dtype=np.float32)
    for j in range(3):
        blk = pts[j * 399:(j + 1) * 399]
        blk[:, :3] = rng.uniform(1, 29, (399, 3)).astype(np.float32)
        blk[:, j] = (k * (0.75 if j == 2 else 1.0)).astype(np.float32)
    pts[:, 3] = (rng.permutation(len(pts)) / len(pts)).astype(np.float32)
    run("vox_f32_boundary", pts.copy(), [0.1, 0.1, 0.3], [0, 0, 0, 40.0, 40.0, 30.0], 5, 5000, True)
    run("vox_f64_boundary", pts.copy(), np.array([0.1, 0.1, 0.3], dtype=np.float32),
        np.array([0, 0, 0, 40.0, 40.0, 30.0]), 5, 5000, True)
    # shuffle variant (in-place shuffle, replay) incl. out-of-range points and C = 5
    pts = synth.uniform_tile(n=3000, seed=5, geom=small, margin=0.2)
    pts = np.concatenate([pts, rng.random((3000, 1)).astype(np.float32)], 1)
    run("vox_shuffle_c5", pts, f32vs(small), f64rg(small), 6, 5000, False)
    # forest tile on the verbatim config.yaml geometry (3-D voxels)
    pts = synth.forest_tile(n=8000, seed=6)
    g = synth.G_REF
    run("vox_gref_forest", pts, f32vs(g), f64rg(g), 50, 7500000 // 100, True)


def pfn_case(R):
    torch.manual_seed(0)
    g = dict(point_cloud_range=[0, -8.0, -3, 16.0, 8.0, 1], voxel_size=[0.16, 0.16, 4])
    pts = synth.dense_tile(n=5000, seed=8, geom=g, n_cells=300, n_clusters=10)
    vox = R.pp.PointPillarsVoxelization("cpu", g["voxel_size"], g["point_cloud_range"], 16, 1000)
    voxels, coords, num = vox(pts)
    coors = torch.nn.functional.pad(coords, (1, 0), value=0)
    for name, feat in (("pfn_single64", [64]), ("pfn_two_layer", [32, 20])):
        net = R.pp.PillarFeatureNet(4, feat, g["voxel_size"], g["point_cloud_range"]).eval()
        layers = {}
        for i, l in enumerate(net.pfn_layers):
            with torch.no_grad():
                l.norm.running_mean.normal_(0, 0.5)
                l.norm.running_var.uniform_(0.3, 2.0)
                l.norm.weight.uniform_(0.5, 1.5)
                l.norm.bias.normal_(0, 0.2)
            layers.update({"w%d" % i: l.linear.weight.detach().numpy(), "gamma%d" % i: l.norm.weight.detach().numpy(),
                           "beta%d" % i: l.norm.bias.detach().numpy(), "mean%d" % i: l.norm.running_mean.numpy(),
                           "var%d" % i: l.norm.running_var.numpy()})
        captured = {}
        h = net.pfn_layers[0].register_forward_pre_hook(lambda m, a: captured.setdefault("dec", a[0].clone()))
        with torch.no_grad():
            out = net(voxels, num, coors)
        h.remove()
        # dense scatter semantics (spconv .dense() is absent: pinned by definition, PointPillars.py:565-571)
        H, W = 100, 100
        canvas = torch.zeros(1, out.shape[1], 1, H, W)
        canvas[coors[:, 0], :, coors[:, 1], coors[:, 2], coors[:, 3]] = out
        save(name, voxels=voxels.numpy(), num=num.numpy(), coors=coors.numpy(), n_layers=np.array(len(feat)),
             voxel_size=np.array(g["voxel_size"]), point_cloud_range=np.array(g["point_cloud_range"]),
             decorated=captured["dec"].numpy(), out=out.numpy(), canvas_hw=np.array([H, W]),
             canvas=canvas.view(1, -1, H, W).numpy(), **layers)


def box_cases(R):
    boxes, scores = synth.nms_boxes(n=600, seed=9, extent=12.0)
    tb = torch.from_numpy(boxes)
    rect = R.ops_torch.bbox2rotated_corners2D(tb)
    corners = R.ops_torch.bbox2corners3D(tb)
    iou = R.ops_torch.bbox_iou2D(rect[:200], rect[200:500])
    iof = R.ops_torch.bbox_iou2D(rect[:50], rect[200:300], mode="iof")
    giou = R.ops_torch.bbox_iou2D(rect[:50], rect[200:300], mode="giou")
    ioujit = R.ops_numba.iou_jit(rect[:40].numpy(), rect[300:360].numpy(), 0.0)
    ioujit1 = R.ops_numba.iou_jit(rect[:40].numpy(), rect[300:360].numpy(), 1.0)
    save("boxes_iou", boxes=boxes, rect=rect.numpy(), corners=corners.numpy(), iou=iou.numpy(), iof=iof.numpy(),
         giou=giou.numpy(), iou_jit=ioujit, iou_jit_eps1=ioujit1)

    # multiclass NMS, 2 classes, several thresholds; keep lists stored sorted by descending score
    rng = np.random.default_rng(10)
    sc2 = np.stack([scores[:, 0], ((rng.permutation(600) + 0.25) / 600).astype(np.float32)], 1)
    out = {}
    for si, sthr in enumerate((0.05, 0.3, 0.7)):
        for ii, ithr in enumerate((1e-5, 0.1, 0.5)):
            keep = R.utils.multiclass_nms(tb, torch.from_numpy(sc2), sthr, ithr, 2)
            for c, k in enumerate(keep):
                k = k.numpy()
                k = k[np.argsort(-sc2[k, c], kind="stable")]
                out["keep_s%d_i%d_c%d" % (si, ii, c)] = k
    save("nms_multiclass", boxes=boxes, scores=sc2, score_thrs=np.array([0.05, 0.3, 0.7]),
         iou_thrs=np.array([1e-5, 0.1, 0.5]), **out)

    # codec + limit_period
    anchors, _ = synth.nms_boxes(n=500, seed=11, extent=30.0, tilt=0.3)
    gts, _ = synth.nms_boxes(n=500, seed=12, extent=30.0, tilt=0.3)
    enc = R.utils.BBoxCoder.encode(torch.from_numpy(anchors), torch.from_numpy(gts))
    deltas = (rng.normal(0, 0.5, (500, 9))).astype(np.float32)
    dec = R.utils.BBoxCoder.decode(torch.from_numpy(anchors), torch.from_numpy(deltas))
    val = rng.uniform(-10, 10, 1000).astype(np.float32)
    lp = R.utils.limit_period(torch.from_numpy(val), 1, np.pi)
    lp2 = R.utils.limit_period(torch.from_numpy(val), 0.5, 2 * np.pi)
    save("codec", anchors=anchors, gts=gts, encoded=enc.numpy(), deltas=deltas, decoded=dec.numpy(), val=val,
         limit_1_pi=lp.numpy(), limit_05_2pi=lp2.numpy())

    # anchors
    gen = R.utils.Anchor3DRangeGenerator([[0, 0, 0, 40.0, 40.0, 30.0]], synth.ANCHOR_SIZES, synth.ANCHOR_ROTATIONS, 9)
    a57 = gen.grid_anchors((5, 7), device="cpu")
    a3d = R.utils.Anchor3DRangeGenerator([[0, -39.68, -1.78, 69.12, 39.68, -1.78]], [[1.6, 3.9, 1.56]],
                                         [[0, 0, 0], [0, 0, 1.57]], 9).grid_anchors((31, 27), device="cpu")
    save("anchors", a57=a57.numpy(), a_kitti=a3d.numpy())


def head_case(R):
    """Anchor3DHead.get_bboxes_single + assign_bboxes with nms_dim = 2 (PointPillars.py:886-1094)."""
    torch.manual_seed(1)
    head = R.pp.Anchor3DHead(num_classes=1, in_channels=8, nms_dim=2, nms_pre=300, nms_thresh=0.1, score_thr=0.3,
                             ranges=[[0, 0, 0, 40.0, 40.0, 30.0]], sizes=synth.ANCHOR_SIZES,
                             rotations=synth.ANCHOR_ROTATIONS, iou_thr=[[0.08, 0.2]])
    H, W, A = 20, 24, 12
    rng = np.random.default_rng(13)
    cls = torch.from_numpy(rng.normal(-0.5, 1.5, (A * 1, H, W)).astype(np.float32))
    reg = torch.from_numpy(rng.normal(0, 0.2, (A * 9, H, W)).astype(np.float32))
    dirs = torch.from_numpy(rng.normal(0, 1, (A * 6, H, W)).astype(np.float32))
    with torch.no_grad():
        b, s, l = head.get_bboxes_single(cls, reg, dirs)
    order = np.argsort(-s.numpy(), kind="stable")
    gts, _ = synth.nms_boxes(n=25, seed=14, extent=38.0, tilt=0.2)
    gts[:, 0:2] += 1.0
    with torch.no_grad():
        ab, ti, pi, ni = head.assign_bboxes(reg.unsqueeze(0), [torch.from_numpy(gts)])
    save("head", cls=cls.numpy(), reg=reg.numpy(), dirs=dirs.numpy(), bboxes=b.numpy()[order], scores=s.numpy()[order],
         labels=l.numpy()[order], gts=gts, assigned=ab.numpy(), target_idx=ti.numpy(), pos_idx=pi.numpy(),
         neg_idx=ni.numpy())


def main():
    os.makedirs(OUT, exist_ok=True)
    R = ref_shim.load()
    voxel_cases(R)
    pfn_case(R)
    box_cases(R)
    head_case(R)
    import numba
    with open(os.path.join(OUT, "VERSIONS.json"), "w") as f:
        json.dump({"numpy": np.__version__, "numba": numba.__version__, "torch": torch.__version__,
                   "reference": "michalp0lak/ObjectDetection_3D (read-only mount)"}, f, indent=1)


if __name__ == "__main__":
    main()
