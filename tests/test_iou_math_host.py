"""The float64 IoU routines of csrc/pp_boxes.cuh (oriented 3-D box volume as branch-free line integrals, rotated BEV
footprints as their 2-D case) compiled for the HOST and checked against the float64 oracle, which computes the same
quantities by polygon clipping: the two algorithms check each other on a machine without a GPU.  (The GPU tests call
the same header through the kernels: tests/test_box3d.py, tests/test_rotated_iou.py.)  Both are test infrastructure;
the product has no CPU path."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not found")
    exe = str(tmp_path_factory.mktemp("iou_host") / "iou_host")
    subprocess.run([NVCC, "-O2", "-std=c++17", "-I", os.path.join(ROOT, "objectdetection_3d_b200", "csrc"), "-o", exe,
                    os.path.join(ROOT, "tests", "host", "iou_host.cu")], check=True, capture_output=True)

    def run(mode, a, b):
        d = os.path.dirname(exe)
        with open(os.path.join(d, "in.bin"), "wb") as f:
            np.array([a.shape[0], b.shape[0]], np.int32).tofile(f)
            np.ascontiguousarray(a, np.float32).tofile(f)
            np.ascontiguousarray(b, np.float32).tofile(f)
        subprocess.run([exe, mode, os.path.join(d, "in.bin"), os.path.join(d, "out.bin")], check=True)
        out = np.fromfile(os.path.join(d, "out.bin"), np.float32)
        nm = a.shape[0] * b.shape[0]
        if mode == "bev":
            return out.reshape(a.shape[0], -1)
        return out[:nm].reshape(a.shape[0], -1), out[nm:].reshape(a.shape[0], -1)
    return run


def _boxes(rng, n, centre, size_lo, size_hi, tilt, offset=0.0):
    b = np.zeros((n, 9), np.float32)
    b[:, 0:3] = rng.uniform(0, centre, (n, 3)) + offset
    b[:, 3:6] = np.exp(rng.uniform(np.log(size_lo), np.log(size_hi), (n, 3)))
    b[:, 6:8] = rng.uniform(-tilt, tilt, (n, 2))
    b[:, 8] = rng.uniform(-np.pi, np.pi, n)
    return b


def _cases():
    from objectdetection_3d_b200 import synth
    rng = np.random.default_rng(5)
    a, _ = synth.nms_boxes(n=150, seed=3, extent=6.0)
    b, _ = synth.nms_boxes(n=170, seed=4, extent=6.0)
    a[:, 2] = rng.uniform(0, 2, len(a)); b[:, 2] = rng.uniform(0, 2, len(b))
    yield "random", a, b
    yield "identical", a, a
    up = a.copy(); up[:, 6:8] = 0
    yield "upright", up, b
    ax = up.copy(); ax[:, 8] = rng.integers(0, 4, len(ax)) * np.float32(np.pi / 2)
    yield "axis aligned", ax, ax
    g = np.zeros((27, 9), np.float32)
    g[:, :3] = np.stack(np.meshgrid(range(3), range(3), range(3), indexing="ij"), -1).reshape(-1, 3)
    g[:, 3:6] = 2
    yield "integer grid (shared faces)", g, g
    sh = a.copy(); sh[:, 0] += 0.25
    yield "shifted copies", a, sh
    ne = a.copy(); ne[:, 3:6] *= 0.5; ne[:, 2] += 0.1
    yield "nested, parallel faces", a, ne
    ti = a.copy(); ti[:, 6:9] += rng.normal(0, 1e-5, (len(a), 3)).astype(np.float32)
    yield "tilted by 1e-5", a, ti
    yield "aspect 1e3", _boxes(rng, 120, 5, 0.01, 10, 0.5), _boxes(rng, 120, 5, 0.01, 10, 0.5)
    yield "1e4 from the origin", _boxes(rng, 120, 5, 0.5, 3, 0.3, 1e4), _boxes(rng, 120, 5, 0.5, 3, 0.3, 1e4)
    yield "millimetre boxes", _boxes(rng, 120, 0.01, 1e-3, 5e-3, 0.3), _boxes(rng, 120, 0.01, 1e-3, 5e-3, 0.3)
    yield "huge against tiny", _boxes(rng, 60, 5, 20, 50, 0.3), _boxes(rng, 200, 5, 0.05, 0.2, 0.3)
    yield "tilt up to 1.5 rad", _boxes(rng, 120, 4, 0.5, 3, 1.5), _boxes(rng, 120, 4, 0.5, 3, 1.5)
    cop = np.array([[18.550724, 0.6779661, 0., 1., 1.75, 20., 0.3142, 0., 0.],
                    [17.971014, 0.6779661, 0., 1.3, 1.3, 17., 0.3142, 0., 0.]], np.float32)
    yield "same tilted ground plane", cop, cop


def test_box3d_volume_routine_against_oracle(harness, oracle):
    for name, a, b in _cases():
        ca, cb = oracle.bbox2corners3D(a), oracle.bbox2corners3D(b)
        ovol, oiou = oracle.box3d_overlap(ca, cb)
        vol, iou = harness("box3d", ca.reshape(-1, 24), cb.reshape(-1, 24))
        assert not np.isnan(iou).any(), name
        assert np.abs(iou - oiou).max() < 1e-6, (name, np.abs(iou - oiou).max())
        assert np.abs(vol - ovol).max() < 1e-6 * max(1.0, ovol.max()), name
        if a is b:
            assert np.array_equal(iou, iou.T), name                   # canonical argument order: symmetric bit for bit


def test_rotated_bev_routine_against_oracle(harness, oracle):
    for name, a, b in _cases():
        ref = oracle.bbox_iou_rotated_bev(a, b)
        got = harness("bev", a, b)
        assert not np.isnan(got).any(), name
        assert np.abs(got - ref).max() < 1e-6, (name, np.abs(got - ref).max())
        if a is b:
            assert np.array_equal(got, got.T), name
