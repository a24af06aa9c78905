"""N1 -- the tracking front-end: bilateral filter, depth pyramid, vertex/normal maps, ICP.
CPU: properties of the oracle restatement.  GPU: the CUDA kernels against the oracle (tolerances written
out: these stages contain expf and order-dependent float reductions)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib
from oracle_lib import SDF, Oracle
from supereight_b200 import synth

W, H, DIM, MU, SIZE = 320, 240, 4.8, 0.1, 256
K = tuple(v * W / 640.0 for v in synth.DEFAULT_K)
ITER = [10, 5, 4]


def assert_close_frac(a, b, atol, max_frac=2e-4):
    """allclose, except for a tiny fraction of elements: a projected pixel that lands within rounding of a pixel
    boundary picks the neighbouring reference pixel on one of the two sides (different sample, not an error)."""
    bad = np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)) > atol
    assert bad.mean() <= max_frac, (int(bad.sum()), bad.size)


def build_scene(make, frames=3, filt=True):
    p = make(SDF, SIZE, DIM, W, H)
    pose = None
    for f in range(frames):
        d, pose = synth.corner_view(f, DIM, W, H, K)
        p.preprocess(d); p.filter_depth(filt, 3); p.integrate(pose, K, MU, f)
    p.raycast(pose, K, MU)
    return p, pose


def test_se3_exp_and_solve_closed_forms():
    lib = oracle_lib.load()
    x = np.array([0.1, -0.2, 0.05, 0.0, 0.0, 0.0], np.float32); T = np.zeros(16, np.float32)
    lib.seo_se3_exp(x.ctypes.data, T.ctypes.data)
    assert np.allclose(T.reshape(4, 4), np.array([[1, 0, 0, .1], [0, 1, 0, -.2], [0, 0, 1, .05], [0, 0, 0, 1]]), atol=1e-7)
    x = np.array([0, 0, 0, 0, 0, np.pi / 2], np.float32)
    lib.seo_se3_exp(x.ctypes.data, T.ctypes.data)
    assert np.allclose(T.reshape(4, 4)[:3, :3], [[0, -1, 0], [1, 0, 0], [0, 0, 1]], atol=1e-6)
    rng = np.random.default_rng(0)
    A = rng.normal(size=(40, 6)); JTJ = (A.T @ A).astype(np.float32); b = rng.normal(size=6).astype(np.float32)
    vals = np.concatenate([b, JTJ[np.triu_indices(6)]]).astype(np.float32); sol = np.zeros(6, np.float32)
    assert lib.seo_solve6(vals.ctypes.data, sol.ctypes.data) == 1
    assert np.allclose(sol, np.linalg.solve(JTJ.astype(np.float64), b), rtol=1e-3, atol=1e-4)
    assert lib.seo_solve6(np.zeros(27, np.float32).ctypes.data, sol.ctypes.data) == 0        # not positive definite


def test_oracle_preprocessing_properties():
    o = Oracle(SDF, SIZE, DIM, W, H)
    d, pose = synth.corner_view(0, DIM, W, H, K, dropout=0.02)
    o.preprocess(d); o.filter_depth(True, 3)
    raw = o.depth()
    iters = np.array(ITER, np.int32)
    o.track(pose, pose, K, 1e-5, iters)                                     # builds the pyramid (the map is empty: no inliers)
    f0, v0, n0 = o.pyramid(0)
    assert np.all(f0[raw == 0] == 0)                                        # holes stay holes
    assert np.abs(f0 - raw)[raw > 0].max() < 0.06 and np.median(np.abs(f0 - raw)[raw > 0]) < 2e-3    # smoothing, bounded at depth edges (e_delta = 0.1)
    d1, v1, n1 = o.pyramid(1)
    assert d1.shape == (H // 2, W // 2) and np.abs(d1 - raw[::2, ::2])[raw[::2, ::2] > 0].max() < 0.31
    valid = f0 > 0
    assert np.allclose(v0[..., 2][valid], f0[valid])                        # z of the back-projection is the depth
    good = n0[..., 0] != -2
    assert good.mean() > 0.9 and np.allclose(np.linalg.norm(n0[good], axis=1), 1, atol=1e-5)


def test_oracle_tracking_recovers_a_perturbed_pose():
    o, pose = build_scene(lambda *a: Oracle(*a))
    d, gt = synth.corner_view(3, DIM, W, H, K)
    o.preprocess(d); o.filter_depth(True, 3)
    start = gt.copy(); start[:3, 3] += np.array([0.02, -0.015, 0.01], np.float32)     # 2.7 cm off
    est, ok = o.track(start, pose, K, 1e-5, ITER)
    assert ok
    assert np.abs(est[:3, 3] - gt[:3, 3]).max() < 0.012 < np.abs(start[:3, 3] - gt[:3, 3]).max()
    td, red = o.tracking_data()
    assert red[28] > 0.5 * W * H and np.sqrt(red[0] / red[28]) < 2e-2
    # a pose from which nothing matches is rejected and restored (checkPoseKernel)
    far = gt.copy(); far[:3, 3] += 1.5
    est2, ok2 = o.track(far, pose, K, 1e-5, ITER)
    assert not ok2 and np.array_equal(est2, far)


@pytest.mark.gpu
def test_cuda_front_end_matches_oracle():
    from supereight_b200 import Map
    g, pose = build_scene(lambda *a: Map(*a))
    o, _ = build_scene(lambda *a: Oracle(*a))
    d, gt = synth.corner_view(3, DIM, W, H, K, dropout=0.01)
    for p in (g, o):
        p.preprocess(d); p.filter_depth(True, 3)
    start = gt.copy(); start[:3, 3] += np.array([0.02, -0.015, 0.01], np.float32)
    # one iteration at the finest level first: compares the kernels before iteration amplifies rounding
    ge, gok = g.track(start, pose, K, 1e-5, [1, 0, 0])
    oe, ook = o.track(start, pose, K, 1e-5, [1, 0, 0])
    for lvl in range(3):
        gd, gv, gn = g.pyramid(lvl); od, ov, on = o.pyramid(lvl)
        np.testing.assert_allclose(gd, od, rtol=2e-6, atol=1e-7)            # expf: device vs libm
        np.testing.assert_allclose(gv, ov, rtol=2e-6, atol=1e-6)
        assert np.array_equal(gn[..., 0] == -2, on[..., 0] == -2)
        good = on[..., 0] != -2
        np.testing.assert_allclose(gn[good], on[good], rtol=0, atol=2e-4)  # normals of nearly flat cross products
    gtd, gred = g.tracking_data(); otd, ored = o.tracking_data()
    assert np.count_nonzero(gtd["result"] != otd["result"]) <= 0.001 * W * H    # a threshold can flip within rounding
    same = (gtd["result"] == 1) & (otd["result"] == 1)
    assert_close_frac(gtd["error"][same], otd["error"][same], atol=2e-5)
    assert_close_frac(gtd["J"][same], otd["J"][same], atol=2e-4)
    # order-dependent float sums of ~7e4 terms of magnitude <= 9 (|J| <= 3): the oracle adds serially per row block,
    # the GPU by a tree; allow 1e-6 of the sum of magnitudes
    np.testing.assert_allclose(gred[:28], ored[:28], rtol=1e-4, atol=1e-6 * 10 * float(ored[28]))
    assert abs(gred[28] - ored[28]) <= 0.001 * W * H
    np.testing.assert_allclose(ge, oe, rtol=0, atol=2e-5)
    # the full coarse-to-fine schedule: both converge to the same pose
    ge, gok = g.track(start, pose, K, 1e-5, ITER)
    oe, ook = o.track(start, pose, K, 1e-5, ITER)
    assert gok and ook
    np.testing.assert_allclose(ge, oe, rtol=0, atol=2e-4)
    assert np.abs(ge[:3, 3] - gt[:3, 3]).max() < 0.012
    # renderTrack of the device-resident result == the oracle's colour map of the same codes
    want = np.empty((H, W, 4), np.uint8)
    td, _ = g.tracking_data()
    codes = np.ascontiguousarray(td["result"])                                  # keep alive while the C call reads it
    oracle_lib.load().seo_render_track(want.ctypes.data, codes.ctypes.data, 1, W, H)
    assert np.array_equal(g.render_track_last(), want)
    far = gt.copy(); far[:3, 3] += 1.5
    est2, ok2 = g.track(far, pose, K, 1e-5, ITER)
    assert not ok2 and np.array_equal(est2, far)
    # updatePoseKernel returning true ends a level (DenseSLAMSystem.cpp:182-183): with a threshold every update passes, the
    # 10/5/4 schedule must do exactly one iteration per level -- on the device that is the `converged` flag in HBM
    e1, ok1 = g.track(start, pose, K, 1.0, ITER)
    e2, ok2 = g.track(start, pose, K, 1e-9, [1, 1, 1])
    assert ok1 == ok2 and np.array_equal(e1.view(np.uint32), e2.view(np.uint32))
    o1, _ = o.track(start, pose, K, 1.0, ITER)
    o2, _ = o.track(start, pose, K, 1e-9, [1, 1, 1])
    assert np.array_equal(o1.view(np.uint32), o2.view(np.uint32))
    np.testing.assert_allclose(e1, o1, rtol=0, atol=2e-4)


@pytest.mark.gpu
def test_cuda_tracking_follows_a_sequence():
    """Closed loop: track each frame against the previous raycast, integrate with the tracked pose."""
    from supereight_b200 import Map
    g = Map(SDF, SIZE, DIM, W, H)
    est = None
    for f in range(10):
        d, gt = synth.corner_view(f, DIM, W, H, K)
        g.preprocess(d); g.filter_depth(True, 3)
        if f < 3:
            est = gt.copy()
        else:
            est, ok = g.track(est, rp, K, 1e-5, ITER)
            assert ok
            assert np.abs(est[:3, 3] - gt[:3, 3]).max() < 0.03            # ~1.5 voxels of drift at 19 mm voxels
        g.integrate(est, K, MU, f); g.raycast(est, K, MU); rp = est.copy()


@pytest.mark.gpu
def test_cuda_pyramid_with_odd_level_widths():
    """100 x 76 with four levels: 100 / 50 / 25 / 12 -- level 3 halves a 25-wide parent, whose row stride is not twice the
    child's width (halfSampleRobustImageKernel indexes with in.width(), preprocessing.cpp:209,217).  Device == oracle (and
    the oracle == the reference build there: tests/test_reference_build.py)."""
    from supereight_b200 import Map
    w, h = 100, 76
    k = tuple(v * w / 640.0 for v in synth.DEFAULT_K)
    g, o = Map(SDF, 128, DIM, w, h), Oracle(SDF, 128, DIM, w, h)
    d, pose = synth.corner_view(2, DIM, w, h, k, noise_mm=2.0, dropout=0.01)
    for p in (g, o):
        p.preprocess(d); p.filter_depth(True, 4)
        p.track(pose, pose, k, 1e-5, [0, 0, 0, 0])
    for lvl in range(4):
        gd, gv, gn = g.pyramid(lvl); od, ov, on = o.pyramid(lvl)
        assert gd.shape == (h >> lvl, w >> lvl)
        np.testing.assert_allclose(gd, od, rtol=2e-6, atol=1e-7)            # expf: device vs libm
        np.testing.assert_allclose(gv, ov, rtol=2e-6, atol=1e-6)
        assert np.array_equal(gn[..., 0] == -2, on[..., 0] == -2)
    assert (od > 0).mean() > 0.5
