"""TEST / MEASUREMENT INFRASTRUCTURE (build container only: needs /root/reference).

Times the UNMODIFIED reference (imported through oracle/ref_shim.py) and the oracle port
(oracle/pp_oracle.c, what bench.py's CPU legs run on the GPU box, where the reference checkout does
not exist) on the SAME inputs -- the D1M tile and the NMS20k boxes of bench.py -- and writes
profiles/ref_vs_port_cpu.json.  bench.py reads that file and carries
``cpu_baseline.port_speedup_over_reference`` so the GPU/port ratio can be converted into a
GPU/reference ratio.

    python scripts/calibrate_cpu_baseline.py [--reps 5]
"""
import argparse
import json
import os
import platform
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def timed(fn, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
    return {"best_s": min(ts), "median_s": statistics.median(ts), "reps": reps}, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--nms-reps", type=int, default=3)
    args = ap.parse_args()
    from objectdetection_3d_b200 import synth
    from oracle import oracle as O
    from oracle import ref_shim
    import bench

    R = ref_shim.load()
    O.lib()
    torch.set_num_threads(os.cpu_count())
    geom, pfn = synth.G_KITTI, synth.pfn_params(9, 63, seed=5)
    pts, boxes, scores = bench.make_frame(3000)
    res = {"inputs": bench.WORKLOAD, "host": {"cpu_count": os.cpu_count(), "torch_threads": torch.get_num_threads(),
                                              "machine": platform.processor() or platform.machine()},
           "versions": {"numpy": np.__version__, "torch": torch.__version__,
                        "numba": __import__("numba").__version__, "python": platform.python_version()},
           "note": "reference = /root/reference imported unmodified (oracle/ref_shim.py), numba JIT warm-up excluded; "
                   "port = oracle/pp_oracle.c (gcc -O2), single thread like the reference's numba kernels; "
                   "scatter stand-in for spconv .dense(): torch index_put (SURVEY 8d)"}

    # ---- stage 1: PointPillarsVoxelization.forward (model/PointPillars.py:330-354)
    vox = R.pp.PointPillarsVoxelization("cpu", geom["voxel_size"], geom["point_cloud_range"], geom["max_voxel_points"],
                                        geom["max_voxels"])
    vox.forward(pts[:20000].copy())          # JIT warm-up
    r_vox, (rv, rc, rn) = timed(lambda: vox.forward(pts.copy()), args.reps)
    p_vox, (pv, pc, pn) = timed(lambda: O.pointpillars_voxelization(pts, geom["voxel_size"], geom["point_cloud_range"],
                                                                    geom["max_voxel_points"], geom["max_voxels"]), args.reps)
    assert np.array_equal(rv.numpy(), pv) and np.array_equal(rc.numpy(), pc) and np.array_equal(rn.numpy(), pn)

    # ---- stage 2: PillarFeatureNet.forward (:480-526) + dense scatter (:565-571)
    net = R.pp.PillarFeatureNet(4, [64], geom["voxel_size"], geom["point_cloud_range"]).eval()
    with torch.no_grad():
        l = net.pfn_layers[0]
        l.linear.weight.copy_(torch.from_numpy(pfn["weight"])); l.norm.weight.copy_(torch.from_numpy(pfn["gamma"]))
        l.norm.bias.copy_(torch.from_numpy(pfn["beta"])); l.norm.running_mean.copy_(torch.from_numpy(pfn["mean"]))
        l.norm.running_var.copy_(torch.from_numpy(pfn["var"]))
    coors4 = torch.cat([torch.zeros((len(rc), 1), dtype=torch.int64), rc], 1)

    def ref_enc():
        with torch.no_grad():
            f = net(rv, rn, coors4)
            canvas = torch.zeros((1, f.shape[1], 1, 496, 432))
            canvas[coors4[:, 0], :, coors4[:, 1], coors4[:, 2], coors4[:, 3]] = f
            return canvas.view(1, -1, 496, 432)

    r_enc, rcanvas = timed(ref_enc, args.reps)

    def port_enc():
        f = O.pillar_feature_net(pv, pn, coors4.numpy(), [pfn], geom["voxel_size"], geom["point_cloud_range"])
        return O.scatter_dense(f, coors4.numpy().astype(np.int32), 1, 1, 496, 432)

    p_enc, pcanvas = timed(port_enc, args.reps)
    assert np.allclose(rcanvas.numpy(), pcanvas, rtol=1e-5, atol=1e-3)

    # ---- stage 3: multiclass_nms (model/utils.py:353-426), nms_dim == 2, the bench's thresholds
    tb, ts_ = torch.from_numpy(boxes), torch.from_numpy(scores)
    r_nms, rkeep = timed(lambda: R.utils.multiclass_nms(tb, ts_, bench.NMS_SCORE_THR, bench.NMS_IOU_THR, 2), args.nms_reps)
    p_nms, pkeep = timed(lambda: O.multiclass_nms(boxes, scores, bench.NMS_SCORE_THR, bench.NMS_IOU_THR, 2), args.nms_reps)
    assert set(rkeep[0].tolist()) == set(pkeep[0].tolist())

    res["reference"] = {"voxelize": r_vox, "pfn_scatter": r_enc, "nms": r_nms}
    res["port"] = {"voxelize": p_vox, "pfn_scatter": p_enc, "nms": p_nms}
    for k in ("best_s", "median_s"):
        rt = sum(res["reference"][s][k] for s in ("voxelize", "pfn_scatter", "nms"))
        pt = sum(res["port"][s][k] for s in ("voxelize", "pfn_scatter", "nms"))
        res["frame_" + k] = {"reference": rt, "port": pt, "port_speedup_over_reference": rt / pt}
    res["port_speedup_over_reference"] = res["frame_median_s"]["port_speedup_over_reference"]
    res["per_stage_speedup_median"] = {s: res["reference"][s]["median_s"] / res["port"][s]["median_s"]
                                       for s in ("voxelize", "pfn_scatter", "nms")}
    res["reference_threads"] = {"voxelize": 1, "pfn_scatter": torch.get_num_threads(), "nms": torch.get_num_threads()}
    out = os.path.join(ROOT, "profiles", "ref_vs_port_cpu.json")
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
