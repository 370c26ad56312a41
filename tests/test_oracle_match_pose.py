"""CPU: pins the matcher and pose restatements (oracle/match.cpp, oracle/pose.cpp) against the committed cv2 fixtures
and, when importable, live cv2."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")
try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def test_knn_golden(oracle):
    z = np.load(os.path.join(GOLD, "matcher_300x400.npz"))
    k = oracle.knn2(z["q"], z["t"])
    assert np.array_equal(k["trainIdx"], z["idx"])
    assert np.array_equal(k["distance"].view(np.uint32), z["dist"].view(np.uint32))  # bit-equal distances


def test_knn_golden_128(oracle):
    """extended-SURF rows (SURF_EXTENDED, VO_utility.h:86): fixture from cv2 4.13 (tools/make_golden.py matcher128)"""
    z = np.load(os.path.join(GOLD, "matcher128_120x160.npz"))
    k = oracle.knn2(z["q"], z["t"])
    assert np.array_equal(k["trainIdx"], z["idx"])
    assert np.array_equal(k["distance"].view(np.uint32), z["dist"].view(np.uint32))
    assert not (z["idx"][:, 0] == 9).any()  # train rows 5 and 9 are identical: the lower index always comes first


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
@pytest.mark.parametrize("dim", [64, 128])
def test_knn_live_cv2(oracle, dim):
    rs = np.random.RandomState(5)
    a = np.abs(rs.randn(700, dim)).astype(np.float32)
    b = np.abs(rs.randn(900, dim)).astype(np.float32)
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    kk = cv2.BFMatcher(cv2.NORM_L2).knnMatch(a, b, 2)
    k = oracle.knn2(a, b)
    assert np.array_equal(np.array([[m.trainIdx for m in r] for r in kk]), k["trainIdx"])
    assert np.array_equal(np.array([[m.distance for m in r] for r in kk], np.float32).view(np.uint32),
                          k["distance"].view(np.uint32))


def test_match_ratio_and_edges(oracle):
    rs = np.random.RandomState(2)
    t = np.abs(rs.randn(50, 64)).astype(np.float32)
    q = t[:10] + 0.01 * rs.randn(10, 64).astype(np.float32)
    m = oracle.match_features(q, t, 0.8)
    assert list(m["queryIdx"]) == list(range(10)) and list(m["trainIdx"]) == list(range(10))
    assert len(oracle.match_features(q, t[:1], 0.8)) == 0 and len(oracle.match_features(q[:0], t, 0.8)) == 0


def test_rng_stream_and_stopping_rule(oracle):
    s = oracle.rng_subsets(4096, 5, 4)
    assert s[0].tolist() == [3317, 924, 2956, 3909, 3271]  # observed through cv2.solvePnPRansac (tools/make_golden)
    assert oracle.rng_subsets(60, 5, 1)[0].tolist() == [45, 4, 20, 33, 51]
    for row in oracle.rng_subsets(7, 5, 200):
        assert len(set(row.tolist())) == 5
    assert [oracle.ransac_update_num_iters(0.99, e, 5, 1000) for e in (0.1, 0.3, 0.5, 0.7, 0.9)] == [5, 25, 145, 1000, 1000]
    assert oracle.ransac_update_num_iters(0.99, 0.45, 5, 2000) == 89   # LMedS budget, 5-point
    assert oracle.ransac_update_num_iters(0.99, 0.45, 4, 2000) == 48   # LMedS budget, 4-point


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_rng_stream_through_cv2(oracle):
    """only the predicted first subset is consistent with a pose: cv2 must return exactly those 5 inliers"""
    z = np.load(os.path.join(GOLD, "pose_600.npz"))
    for count in (60, 1000):
        rs = np.random.RandomState(count)
        X = np.stack([rs.uniform(-4, 4, count), rs.uniform(-3, 3, count), rs.uniform(4, 9, count)], -1)
        x = np.stack([rs.uniform(0, 1280, count), rs.uniform(0, 1024, count)], -1).astype(np.float32)
        S = oracle.rng_subsets(count, 5, 1)[0]
        pr, _ = cv2.projectPoints(X[S], z["rvec"], z["tvec"], z["K"], None)
        x[S] = pr.reshape(-1, 2)
        ok, rv, tv, inl = cv2.solvePnPRansac(X, x, z["K"], np.zeros(4), iterationsCount=1, reprojectionError=0.05,
                                             confidence=0.99, flags=cv2.SOLVEPNP_EPNP)
        assert ok and sorted(inl.ravel().tolist()) == sorted(S.tolist())


def test_pose_golden(oracle):
    z = np.load(os.path.join(GOLD, "pose_600.npz"))
    K = z["K"]
    R = oracle.rodrigues_vec2mat(z["rvec"])
    assert np.abs(oracle.project_points(z["X"], R, z["tvec"], K) - z["proj"]).max() < 1e-9
    assert np.abs(oracle.rodrigues_mat2vec(z["Rm"]) - z["rv_back"]).max() < 1e-12
    X4 = oracle.triangulate_points(z["P1"], z["P2"], z["x1"], z["x2"])
    a, b = X4[:3] / X4[3], z["X4"][:3] / z["X4"][3]
    assert np.abs(a - b).max() <= 2e-6 * np.abs(b).max()
    ok, rv, tv, inl, hyps = oracle.solve_pnp_ransac_epnp(z["X"], z["x"], K, 1000, 1.0, 0.99)
    assert ok and np.array_equal(inl, z["pnp_inliers"])           # same inlier set as cv2.solvePnPRansac
    assert np.abs(rv - z["pnp_rvec"]).max() < 1e-6 and np.abs(tv - z["pnp_tvec"]).max() < 1e-6


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_epnp_against_cv2_overdetermined(oracle):
    z = np.load(os.path.join(GOLD, "pose_600.npz"))
    X, x = z["X"][:80], z["x"][:80].astype(np.float64)
    R, t = oracle.epnp(X, x, z["K"])
    ok, rv, tv = cv2.solvePnP(X, x, z["K"], None, flags=cv2.SOLVEPNP_EPNP)
    Rc, _ = cv2.Rodrigues(rv)
    assert np.abs(R - Rc).max() < 1e-9 and np.abs(t - tv.ravel()).max() < 1e-8


def test_extract_3dpoints_semantics(oracle):
    z = np.load(os.path.join(GOLD, "pose_600.npz"))
    K = z["K"]
    R2 = oracle.rodrigues_vec2mat(z["rvec"])
    X4 = oracle.triangulate_points(z["P1"], z["P2"], z["x1"], z["x2"])
    pts, idx = oracle.extract_3dpoints(z["x1"], z["x2"], np.eye(3), np.zeros(3), R2, z["tvec"], K, K, X4, 3.0, 5)
    assert 400 < len(idx) <= 600 and np.all(np.diff(idx) > 0) and np.all(pts[:, 2] > 0)
    zs = pts[:, 2]
    assert np.abs(zs - zs.mean()).max() <= 3.2 * zs.std()
    none, _ = oracle.extract_3dpoints(z["x1"], z["x2"], np.eye(3), np.zeros(3), R2, z["tvec"], K, K, X4, 3.0, 601)
    assert len(none) == 0                                           # fewer than MIN_NUM_3DPOINTS rows: nothing


def test_median_scale_and_method(oracle):
    assert oracle.compute_median([3.0, 1.0, 2.0]) == 2.0 and oracle.compute_median([4.0, 1.0, 2.0, 3.0]) == 2.5
    assert oracle.compute_median([]) == 0.0
    pts = np.array([[0, 0, 2.0], [0, 0, 4.0], [0, 0, -1.0], [0, 0, 6.0]])
    assert oracle.scale_factor(pts, np.eye(3), np.zeros(3), 8.0) == 2.0       # median of (2,4,6) = 4
    assert oracle.scale_factor(pts[2:3], np.eye(3), np.zeros(3), 8.0) == 0.0  # nothing in front of the camera
    p1 = np.zeros((5, 2), np.float32)
    p2 = p1 + np.float32([3, 4])
    assert oracle.select_estimation_method(p1, p2, 10) is False and oracle.select_estimation_method(p1, p2, 5) is True
