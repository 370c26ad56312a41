"""K10c/K11/K12 parity: CUDA triangulatePoints / extract_3Dpoints / solvePnPRansac(EPNP) vs the CPU oracle and the
committed cv2 fixture.  Index lists and inlier sets bit-exact; fp64 values to 1e-9 relative (the device libm's
hypot/sin/cos/acos differ from glibc in the last ulp, so hypothesis models are not bit-equal).
Reference: visual_odometry.h:631-648, VO_utility.cpp:188-237."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _scene(n, seed, outlier_frac=0.3, noise=0.3):
    rs = np.random.RandomState(seed)
    K = np.array([[1300.0, 0, 640], [0, 1300.0, 512], [0, 0, 1]])
    X = np.stack([rs.uniform(-4, 4, n), rs.uniform(-3, 3, n), rs.uniform(4, 9, n)], -1)
    rvec = np.array([0.01, -0.02, 0.015])
    tvec = np.array([0.3, 0.05, 0.1])
    from oracle import oracle as O
    R = O.rodrigues_vec2mat(rvec)
    x = O.project_points(X, R, tvec, K) + rs.randn(n, 2) * noise
    out = rs.rand(n) < outlier_frac
    x[out] = np.stack([rs.uniform(0, 1280, out.sum()), rs.uniform(0, 1024, out.sum())], -1)
    return K, X, x.astype(np.float32), rvec, tvec


def test_triangulate_golden_cv2(ctx):
    z = np.load(os.path.join(GOLD, "pose_600.npz"))
    X4 = ctx.triangulatePoints(z["P1"], z["P2"], z["x1"], z["x2"])
    a, b = X4[:3] / X4[3], z["X4"][:3] / z["X4"][3]
    assert np.abs(a - b).max() <= 1e-5 * np.abs(b).max()


def test_triangulate_vs_oracle(ctx, oracle):
    z = np.load(os.path.join(GOLD, "pose_600.npz"))
    X4 = ctx.triangulatePoints(z["P1"], z["P2"], z["x1"], z["x2"])
    X4o = oracle.triangulate_points(z["P1"], z["P2"], z["x1"], z["x2"])
    assert X4.shape == X4o.shape == (4, 600)
    # the homogeneous sign is arbitrary; compare sign-normalised vectors (f32, 2 ulp)
    s, so = np.sign(X4[3]), np.sign(X4o[3])
    assert np.allclose(X4 * s, X4o * so, rtol=3e-7, atol=1e-9)
    assert len(ctx.triangulatePoints(z["P1"], z["P2"], z["x1"][:0], z["x2"][:0]).T) == 0


@pytest.mark.parametrize("tol,min3d", [(3.0, 5), (0.1, 5), (1.0, 700)])
def test_extract_3dpoints_vs_oracle(ctx, oracle, tol, min3d):
    z = np.load(os.path.join(GOLD, "pose_600.npz"))
    K = z["K"]
    R2 = oracle.rodrigues_vec2mat(z["rvec"])
    X4 = oracle.triangulate_points(z["P1"], z["P2"], z["x1"], z["x2"])
    X4[:, 5] *= -1            # a point behind the camera ...
    X4[3, 9] = 0.0            # ... and one at infinity (w == 0 => scale 1)
    ctx.params.reprojection_tolerance, ctx.params.min_num_3dpoints = tol, min3d
    pts, idx = ctx.extract_3Dpoints(z["x1"], z["x2"], np.eye(3), np.zeros(3), R2, z["tvec"], K, K, X4)
    po, io = oracle.extract_3dpoints(z["x1"], z["x2"], np.eye(3), np.zeros(3), R2, z["tvec"], K, K, X4, tol, min3d)
    ctx.params.reprojection_tolerance, ctx.params.min_num_3dpoints = 3.0, 5
    assert np.array_equal(idx, io)
    assert np.array_equal(pts, po)
    if min3d == 700:
        assert len(idx) == 0
    else:
        assert 0 < len(idx) < 600


def test_pnp_ransac_golden_cv2(ctx):
    """cv2.solvePnPRansac fixture: same inlier set, pose within 1e-6 (LAPACK vs Jacobi null-space bases differ at the
    hypothesis level, the refit on the common inlier set agrees)"""
    z = np.load(os.path.join(GOLD, "pose_600.npz"))
    ok, rvec, tvec, inl, hyps = ctx.solvePnPRansac(z["X"], z["x"], z["K"], 1000, 1.0, 0.99)
    assert ok and np.array_equal(inl, z["pnp_inliers"])
    assert np.abs(rvec - z["pnp_rvec"]).max() < 1e-6 and np.abs(tvec - z["pnp_tvec"]).max() < 1e-6


@pytest.mark.parametrize("n,frac,iters,conf", [(600, 0.3, 1000, 0.99), (4000, 0.3, 1000, 0.99), (10000, 0.5, 512, 0.999),
                                                (50, 0.1, 100, 0.99), (5, 0.0, 100, 0.99), (6, 0.0, 50, 0.99)])
def test_pnp_ransac_vs_oracle(ctx, oracle, n, frac, iters, conf):
    K, X, x, rvec, tvec = _scene(n, seed=n + iters, outlier_frac=frac)
    ok, rv, tv, inl, hyps = ctx.solvePnPRansac(X, x, K, iters, 1.0, conf)
    oko, rvo, tvo, inlo, hypso = oracle.solve_pnp_ransac_epnp(X, x, K, iters, 1.0, conf)
    assert ok == oko
    assert hyps == hypso                     # same stopping iteration as the sequential loop
    assert np.array_equal(inl, inlo)         # inlier set bit-exact
    assert np.abs(rv - rvo).max() <= 1e-9 and np.abs(tv - tvo).max() <= 1e-9
    if n >= 50:
        # and the answer is right: rotation within 0.01 degree of the ground truth is not guaranteed by noise, but
        # the estimate must be close
        assert np.abs(rv - rvec).max() < 5e-3 and np.abs(tv - tvec).max() < 5e-2


def test_pnp_ransac_all_hypotheses(ctx, oracle):
    """config D style: confidence 1-2^-53 keeps every hypothesis alive"""
    K, X, x, _, _ = _scene(2000, seed=3, outlier_frac=0.75)
    conf = 1.0 - 2.0 ** -53
    ok, rv, tv, inl, hyps = ctx.solvePnPRansac(X, x, K, 512, 1.0, conf)
    oko, rvo, tvo, inlo, hypso = oracle.solve_pnp_ransac_epnp(X, x, K, 512, 1.0, conf)
    assert hyps == hypso == 512
    assert ok == oko and np.array_equal(inl, inlo)


def test_pnp_ransac_degenerate_inputs(ctx):
    import ergo_uvo_b200 as U
    K = np.array([[1300.0, 0, 640], [0, 1300.0, 512], [0, 0, 1]])
    ok, rv, tv, inl, hyps = ctx.solvePnPRansac(np.zeros((3, 3)), np.zeros((3, 2), np.float32), K)
    assert not ok and len(inl) == 0
    with pytest.raises(U.UvoError):          # CV_Assert(confidence > 0 && confidence < 1)
        ctx.solvePnPRansac(np.zeros((10, 3)), np.zeros((10, 2), np.float32), K, 10, 1.0, 1.0)
    # pure garbage: no hypothesis reaches 5 inliers -> failure, empty inlier list
    rs = np.random.RandomState(0)
    X = rs.uniform(-1, 1, (40, 3)) + [0, 0, 5]
    x = rs.uniform(0, 1000, (40, 2)).astype(np.float32)
    ok, rv, tv, inl, hyps = ctx.solvePnPRansac(X, x, K, 50, 0.01, 0.99)
    assert not ok and len(inl) == 0 and hyps == 50


@pytest.mark.parametrize("n", [1, 2, 3, 4, 7, 64, 255, 256, 257, 1000, 4001])
def test_select_estimation_method_stage(ctx, oracle, n):
    """K9 (VO_utility.cpp:725-748 + compute_median, math_utility.cpp:65-86) on its own: odd and even counts, medians
    far from, just below, exactly at and just above DISTANCE -- `median < DISTANCE` means homography"""
    rs = np.random.RandomState(n)
    p1 = rs.uniform(0, 1000, (n, 2)).astype(np.float32)
    ang = rs.uniform(0, 2 * np.pi, n)
    for scale in (2.0, 9.99, 10.0, 10.01, 40.0):
        disp = rs.uniform(0.5, 1.5, n) * scale
        p2 = (p1 + np.stack([np.cos(ang), np.sin(ang)], -1) * disp[:, None]).astype(np.float32)
        ctx.params.distance = 10
        assert ctx.select_estimation_method(p1, p2) == oracle.select_estimation_method(p1, p2, 10), (n, scale)
    # every displacement exactly DISTANCE (3-4-5 triangles, exact in f32): the median equals DISTANCE -> essential
    p1 = np.round(p1)
    p2 = p1 + np.array([6.0, 8.0], np.float32)
    assert ctx.select_estimation_method(p1, p2) is True and oracle.select_estimation_method(p1, p2, 10) is True
    # the median of an even count is the mean of the middle two: {8, 12} -> 10 -> essential, {8, 11.99} -> homography
    if n % 2 == 0:
        d = np.where(np.arange(n) < n // 2, 8.0, 12.0).astype(np.float32)
        p2 = p1 + np.stack([d, np.zeros(n, np.float32)], -1)
        assert ctx.select_estimation_method(p1, p2) is True and oracle.select_estimation_method(p1, p2, 10) is True
        d[n // 2:] = 11.75
        p2 = p1 + np.stack([d, np.zeros(n, np.float32)], -1)
        assert ctx.select_estimation_method(p1, p2) is False and oracle.select_estimation_method(p1, p2, 10) is False


def test_scale_factor_stage(ctx, oracle):
    """uvo_scale_factor (convert_3Dpoints_camera + compute_scale_factor, VO_utility.cpp:23-63) on its own: empty set,
    every point behind the camera, odd / even medians, f32 narrowing of the range (App. D-4), range == 0"""
    rs = np.random.RandomState(5)
    R = oracle.rodrigues_vec2mat(np.array([0.02, -0.01, 0.03]))
    t = np.array([0.05, -0.02, 0.1])
    assert ctx.compute_scale_factor(3.0, np.zeros((0, 3)), R, t, with_count=True) == (0.0, 0)
    for n in (1, 2, 5, 6, 255, 256, 257, 3000):
        pts = np.stack([rs.uniform(-2, 2, n), rs.uniform(-2, 2, n), rs.uniform(0.5, 9, n)], -1)
        pts[::3, 2] *= -1                      # a third of the points end up behind the current camera
        for rng in (2.9371, 0.0):
            sf, m = ctx.compute_scale_factor(rng, pts, R, t, with_count=True)
            sfo, mo = oracle.scale_factor(pts, R, t, np.float32(rng), with_count=True)
            assert m == mo and sf == sfo, (n, rng)       # the median is an order statistic: bit-exact
            if m > 0 and rng > 0:
                z = np.sort(pts[(pts @ R[2] + t[2]) > 0][:, 2])
                med = z[len(z) // 2] if len(z) % 2 else (z[len(z) // 2 - 1] + z[len(z) // 2]) / 2.0
                assert sf == float(np.float32(rng)) / med
    behind = np.stack([rs.uniform(-1, 1, 50), rs.uniform(-1, 1, 50), -rs.uniform(1, 5, 50)], -1)
    assert ctx.compute_scale_factor(3.0, behind, np.eye(3), np.zeros(3), with_count=True) == (0.0, 0)
    # range == 0 with points in front: SF = 0 but the set is non-empty (the node assigns SF = 0 and stays valid)
    front = np.abs(behind)
    assert ctx.compute_scale_factor(0.0, front, np.eye(3), np.zeros(3), with_count=True) == (0.0, 50)
