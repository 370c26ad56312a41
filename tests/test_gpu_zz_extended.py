"""Extended (128-d) SURF descriptors -- SURF_EXTENDED (VO_utility.h:86), passed to SURF::create by detect_features
(VO_utility.cpp:117) and matched by the same BFMatcher(NORM_L2) (VO_utility.cpp:515-573).  CUDA path vs the CPU oracle
through the C ABI: keypoints and descriptors bit-exact, match lists and kNN distances bit-exact, with the tensor-core
route (k_knn_tc<128>) and the exact-scan route cross-checked against each other."""
import numpy as np
import pytest

from conftest import noise_image

pytestmark = pytest.mark.gpu


def _descs(n, seed, dim=128, dup_from=None, noise=0.03):
    rs = np.random.RandomState(seed)
    d = np.abs(rs.randn(n, dim)).astype(np.float32)
    if dup_from is not None:
        m = min(n, len(dup_from)) // 2
        idx = rs.permutation(len(dup_from))[:m]
        d[:m] = dup_from[idx] + noise * rs.randn(m, dim).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return d


@pytest.fixture
def ext_ctx(ctx):
    ctx.params.surf_extended = 1
    ctx.params.max_features = 1 << 16
    yield ctx
    ctx.params.surf_extended = 0
    ctx.params.surf_upright = 1
    ctx.match_exact_only(False)


@pytest.mark.parametrize("w,h,thr", [(640, 480, 50), (417, 303, 300)])
def test_surf_extended_matches_oracle(ext_ctx, oracle, w, h, thr):
    g = noise_image(h, w, seed=3 * w + h)
    ext_ctx.params.surf_min_hessian = thr
    k, d = ext_ctx.detect_features(g)
    ko, do = oracle.surf_detect_and_compute(g, thr, extended=True)
    assert len(k) > 50 and d.shape == (len(k), 128)
    assert k.tobytes() == ko.tobytes()
    assert np.abs(d - do).max() <= 1e-4 * np.abs(do).max()
    assert np.array_equal(d.view(np.uint32), do.view(np.uint32))  # stronger: bit-exact in practice
    # the flag does not change the keypoints, and the 64-d path still works on the same context afterwards
    ext_ctx.params.surf_extended = 0
    k64, d64 = ext_ctx.detect_features(g)
    ko64, do64 = oracle.surf_detect_and_compute(g, thr)
    assert k64.tobytes() == k.tobytes() and d64.shape == (len(k), 64)
    assert np.array_equal(d64.view(np.uint32), do64.view(np.uint32))


def test_surf_extended_oriented(ext_ctx, oracle):
    g = noise_image(480, 640, seed=77)
    ext_ctx.params.surf_min_hessian = 400
    ext_ctx.params.surf_upright = 0
    k, d = ext_ctx.detect_features(g)
    ko, do = oracle.surf_detect_and_compute(g, 400, extended=True, upright=False)
    assert len(k) == len(ko) and d.shape == do.shape
    for f in ("x", "y", "size", "response", "octave", "angle"):
        assert np.array_equal(k[f], ko[f]), f
    close = np.abs(d - do).max(axis=1) <= 1e-4 * np.abs(do).max()
    assert close.mean() > 0.999


def test_knn_128_golden_cv2(ctx):
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "matcher128_120x160.npz"))
    k = ctx.knn_match2(z["q"], z["t"])
    assert np.array_equal(k["trainIdx"], z["idx"])
    assert np.array_equal(k["distance"].view(np.uint32), z["dist"].view(np.uint32))


@pytest.mark.parametrize("exact_only", [False, True])
@pytest.mark.parametrize("nq,nt,ratio", [(4096, 4096, 0.8), (1000, 3333, 0.7), (129, 65, 0.8), (5, 2, 0.9)])
def test_match_128_vs_oracle(ext_ctx, oracle, nq, nt, ratio, exact_only):
    t = _descs(nt, 1)
    q = _descs(nq, 2, dup_from=t)
    ext_ctx.match_exact_only(exact_only)
    ext_ctx.params.lowe_ratio = ratio
    m = ext_ctx.match_features(None, None, q, t)
    ext_ctx.params.lowe_ratio = 0.8
    mo = oracle.match_features(q, t, ratio)
    assert len(mo) > 0 or nq < 10
    assert m.tobytes() == mo.tobytes()
    k = ext_ctx.knn_match2(q, t)
    assert k.tobytes() == oracle.knn2(q, t).tobytes()
    if exact_only:
        assert ext_ctx.match_last_fallbacks() == nq
    elif nq >= 1000:
        assert ext_ctx.match_last_fallbacks() <= nq // 20  # the tensor-core candidates are almost always provably complete


def test_match_64_exact_only_route(ext_ctx, oracle):
    """the same cross-check for the default 64-float rows"""
    t = _descs(1500, 1, dim=64)
    q = _descs(700, 2, dim=64, dup_from=t)
    ext_ctx.match_exact_only(True)
    k = ext_ctx.knn_match2(q, t)
    assert ext_ctx.match_last_fallbacks() == 700
    ext_ctx.match_exact_only(False)
    assert k.tobytes() == oracle.knn2(q, t).tobytes() == ext_ctx.knn_match2(q, t).tobytes()


def test_match_128_ties_duplicates_and_small_sets(ext_ctx, oracle):
    rs = np.random.RandomState(7)
    t = _descs(300, 3)
    t[200] = t[10]
    t[250] = t[10]
    k = ext_ctx.knn_match2(t[[10, 11, 12]].copy(), t)
    assert list(k["trainIdx"][0]) == [10, 200]
    # hundreds of near-duplicates per query: the exact scan must take over, result still bit-identical
    base = _descs(8, 8)
    t2 = np.repeat(base, 200, axis=0) + (1e-4 * rs.randn(1600, 128)).astype(np.float32)
    q2 = base + (1e-4 * rs.randn(8, 128)).astype(np.float32)
    k2 = ext_ctx.knn_match2(q2, t2)
    assert ext_ctx.match_last_fallbacks() > 0
    assert k2.tobytes() == oracle.knn2(q2, t2).tobytes()
    # unnormalised rows
    t3 = (rs.randn(700, 128) * 37.0).astype(np.float32)
    q3 = (t3[rs.permutation(700)[:300]] + rs.randn(300, 128) * 5.0).astype(np.float32)
    assert ext_ctx.knn_match2(q3, t3).tobytes() == oracle.knn2(q3, t3).tobytes()
    q = _descs(130, 12)
    for nt in (1, 2, 3, 127, 128, 129, 257):
        tt = _descs(nt, 11)
        kk, ko = ext_ctx.knn_match2(q, tt), oracle.knn2(q, tt)
        assert np.array_equal(kk["trainIdx"], ko["trainIdx"])
        valid = ko["trainIdx"] >= 0
        assert np.array_equal(kk["distance"][valid].view(np.uint32), ko["distance"][valid].view(np.uint32))
    assert len(ext_ctx.match_features(None, None, q[:0], q)) == 0 and len(ext_ctx.match_features(None, None, q, q[:0])) == 0


def test_unsupported_row_lengths(ctx):
    """rows other than 64 / 128 floats are refused with UVO_ERR_UNSUPPORTED -- never silently truncated"""
    import ergo_uvo_b200 as U
    d = _descs(20, 1, dim=32)
    with pytest.raises(U.UvoError) as e:
        ctx.knn_match2(d, d)
    assert e.value.code == -5


def test_stereo_handle_extended(ctx, oracle, small_stereo):
    """uvo_stereo created with surf_extended: lane buffers, the after-stereo gather and both matcher calls carry
    128-float rows; every product of the frame equals the CPU replay of stereo_VO (visual_odometry.h:526-740)"""
    import ergo_uvo_b200 as U
    from oracle.ref_stereo import RefStereoVO
    from test_gpu_stereo import _compare_frame
    seq = small_stereo
    p = U.default_params(True)
    p.surf_min_hessian = 3000
    p.max_features = 16384
    p.surf_extended = 1
    camL = U.make_camera(seq.KL, seq.DL, seq.newKL)
    camR = U.make_camera(seq.KR, seq.DR, seq.newKR)
    vo = U.StereoVO(ctx, seq.w, seq.h, camL, camR, seq.R_right, seq.t_right, p)
    ref = RefStereoVO(oracle, seq, p)
    for k, (L, R) in enumerate(seq.frames):
        res = vo.frame(L, R, 0.1)
        r = ref.frame(L, R, 0.1)
        assert r["dL"].shape[1] == 128
        _compare_frame(vo, res, r)
        assert res.valid == (1 if k else 0)
    assert res.n_temporal_matches > 100 and res.n_inliers > 50
    vo.close()


def test_mono_handle_extended(ctx, oracle):
    import ergo_uvo_b200 as U
    from oracle.ref_mono import RefMonoVO
    from tools import synth
    seq = synth.MonoSequence(640, 480, n_frames=3, tex_size=1024, velocity=(0.02, 0.004, 0.0))
    p = U.default_params(False)
    p.surf_extended = 1
    vo = U.MonoVO(ctx, 640, 480, U.make_camera(seq.K, seq.D, seq.newK), p)
    ref = RefMonoVO(oracle, seq, p)
    published = 0
    for k in range(3):
        r = vo.frame(seq.frames[k], 0.1, seq.ranges[k])
        o = ref.frame(seq.frames[k], 0.1, seq.ranges[k])
        for f in ("initialised", "skipped", "published", "valid", "n_keypoints", "n_matches", "n_inliers", "n_3d"):
            assert getattr(r, f) == o[f], (k, f, getattr(r, f), o[f])
        if o["published"]:
            published += 1
            assert np.abs(np.array(r.R).reshape(3, 3) - o["R"]).max() < 1e-6
            assert np.abs(np.array(r.velocity) - o["velocity"]).max() <= 1e-6 * np.abs(o["velocity"]).max()
    assert published == 2
    vo.close()
