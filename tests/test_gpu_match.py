"""K8 parity: CUDA brute-force kNN(k=2) + Lowe ratio vs the CPU oracle and the committed cv2 fixture -- indices,
order, distances and match lists bit-exact.  Reference path: match_features, VO_utility.cpp:515-573."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _descs(n, seed, dup_from=None, noise=0.03):
    rs = np.random.RandomState(seed)
    d = np.abs(rs.randn(n, 64)).astype(np.float32)
    if dup_from is not None:
        m = min(n, len(dup_from)) // 2
        idx = rs.permutation(len(dup_from))[:m]
        d[:m] = dup_from[idx] + noise * rs.randn(m, 64).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return d


def test_knn_golden_cv2(ctx):
    z = np.load(os.path.join(GOLD, "matcher_300x400.npz"))
    k = ctx.knn_match2(z["q"], z["t"])
    assert np.array_equal(k["trainIdx"], z["idx"])
    assert np.array_equal(k["distance"].view(np.uint32), z["dist"].view(np.uint32))
    assert np.array_equal(k["queryIdx"][:, 0], np.arange(300))


@pytest.mark.parametrize("nq,nt,ratio", [(4096, 4096, 0.8), (1000, 3333, 0.7), (8192, 8192, 0.7), (129, 65, 0.8), (5, 2, 0.9)])
def test_match_features_vs_oracle(ctx, oracle, nq, nt, ratio):
    t = _descs(nt, 1)
    q = _descs(nq, 2, dup_from=t)
    ctx.params.lowe_ratio = ratio
    m = ctx.match_features(None, None, q, t)
    mo = oracle.match_features(q, t, ratio)
    assert len(mo) > 0 or nq < 10
    assert m.tobytes() == mo.tobytes()
    k = ctx.knn_match2(q, t)
    ko = oracle.knn2(q, t)
    assert k.tobytes() == ko.tobytes()
    ctx.params.lowe_ratio = 0.8


def test_match_ties_pick_lower_train_index(ctx, oracle):
    t = _descs(300, 3)
    t[200] = t[10]
    t[250] = t[10]
    q = t[[10, 11, 12]].copy()
    k = ctx.knn_match2(q, t)
    assert list(k["trainIdx"][0]) == [10, 200]
    assert k.tobytes() == oracle.knn2(q, t).tobytes()


def test_match_edge_cases(ctx, oracle):
    q = _descs(10, 4)
    assert len(ctx.match_features(None, None, q[:0], q)) == 0          # empty query set
    assert len(ctx.match_features(None, None, q, q[:0])) == 0          # empty train set
    assert len(ctx.match_features(None, None, q, q[:1])) == 0          # 1 train row: reference UB, defined as no match
    k = ctx.knn_match2(q, q[:1])
    assert np.all(k["trainIdx"][:, 0] == 0) and np.all(k["trainIdx"][:, 1] == -1)
    # an empty cv::Mat as the query set (rows == 0 and cols == 0: the descriptors the node carries after a frame that
    # failed a gate, visual_odometry.h:592) is "no match", not UVO_ERR_UNSUPPORTED for the row length
    import ctypes as C
    n = C.c_int(7)
    rc = ctx.lib.uvo_match_features(ctx.h, None, 0, q.ctypes.data_as(C.c_void_p), 10, 0, C.c_float(0.8), None,
                                    C.byref(n))
    assert rc == 0 and n.value == 0


def test_match_7arg_overload_points(ctx, oracle):
    """7-arg overload (VO_utility.cpp:551-573) also returns the matched Point2f pairs"""
    import ergo_uvo_b200 as U
    t = _descs(500, 5)
    q = _descs(400, 6, dup_from=t)
    rs = np.random.RandomState(0)
    k1 = np.zeros(400, U.KEYPOINT_DTYPE)
    k2 = np.zeros(500, U.KEYPOINT_DTYPE)
    k1["x"], k1["y"] = rs.rand(400) * 640, rs.rand(400) * 480
    k2["x"], k2["y"] = rs.rand(500) * 640, rs.rand(500) * 480
    m, p1, p2 = ctx.match_features(k1, k2, q, t, with_points=True)
    assert np.array_equal(p1[:, 0], k1["x"][m["queryIdx"]]) and np.array_equal(p2[:, 1], k2["y"][m["trainIdx"]])


def test_match_forced_exact_scan(ctx, oracle):
    """train rows that the TF32 pass cannot separate (hundreds of near-duplicates of every query): the candidate set
    cannot be proven complete, the exact full scan must take over, and the answer is still bit-identical"""
    rs = np.random.RandomState(7)
    base = _descs(8, 8)
    t = np.repeat(base, 200, axis=0) + (1e-4 * rs.randn(1600, 64)).astype(np.float32)
    q = base + (1e-4 * rs.randn(8, 64)).astype(np.float32)
    k = ctx.knn_match2(q, t)
    assert ctx.match_last_fallbacks() > 0
    assert k.tobytes() == oracle.knn2(q, t).tobytes()


def test_match_unnormalised_descriptors(ctx, oracle):
    """the error bound scales with |q| max|t|: descriptors far from unit norm stay exact"""
    rs = np.random.RandomState(9)
    t = (rs.randn(700, 64) * 37.0).astype(np.float32)
    q = (t[rs.permutation(700)[:300]] + rs.randn(300, 64) * 5.0).astype(np.float32)
    k = ctx.knn_match2(q, t)
    assert k.tobytes() == oracle.knn2(q, t).tobytes()


def test_match_fallback_is_rare_on_descriptor_like_data(ctx):
    t = _descs(4096, 1)
    q = _descs(4096, 2, dup_from=t)
    ctx.knn_match2(q, t)
    assert ctx.match_last_fallbacks() <= 4096 // 20


@pytest.mark.parametrize("nt", [1, 2, 3, 4, 5, 127, 128, 129, 257])
def test_match_small_train_sets(ctx, oracle, nt):
    t = _descs(nt, 11)
    q = _descs(130, 12)
    k = ctx.knn_match2(q, t)
    ko = oracle.knn2(q, t)
    assert np.array_equal(k["trainIdx"], ko["trainIdx"])
    valid = ko["trainIdx"] >= 0
    assert np.array_equal(k["distance"][valid].view(np.uint32), ko["distance"][valid].view(np.uint32))


def test_match_features_gated(ctx, oracle):
    """uvo_match_features_gated == ratio-test matches filtered by the epipolar / disparity predicate (f32)."""
    import ergo_uvo_b200 as U
    from oracle.ref_stereo import stereo_gate
    rng = np.random.RandomState(5)
    t = _descs(700, 1)
    q = _descs(600, 2, dup_from=t)
    kq = np.zeros(600, U.KEYPOINT_DTYPE)
    kt = np.zeros(700, U.KEYPOINT_DTYPE)
    kq["x"], kq["y"] = rng.uniform(0, 640, 600), rng.uniform(0, 480, 600)
    kt["x"], kt["y"] = rng.uniform(0, 640, 700), rng.uniform(0, 480, 700)
    ctx.params.lowe_ratio = 0.9
    full = oracle.match_features(q, t, np.float32(0.9))
    # make about half of the matched pairs satisfy the gate
    for m in full[::2]:
        kt["y"][m["trainIdx"]] = kq["y"][m["queryIdx"]] + np.float32(0.75)
        kt["x"][m["trainIdx"]] = kq["x"][m["queryIdx"]] - np.float32(20.5)
    got = ctx.match_features(kq, kt, q, t, gate=(1.0, 2.0, 100.0))
    want = stereo_gate(full, kq, kt, 1.0, 2.0, 100.0)
    assert 0 < len(want) < len(full)
    assert got.tobytes() == want.tobytes()
    assert ctx.match_features(kq, kt, q, t).tobytes() == full.tobytes()
    ctx.params.lowe_ratio = 0.8
