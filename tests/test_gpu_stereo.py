"""Whole-frame parity: device-resident uvo_stereo pipeline vs the CPU replay of stereo_VO built from the oracle
(tests/ref_stereo.py).  Keypoints, match lists and inlier sets bit-exact; pose within 1e-9; velocity within 1e-4
relative of the oracle and close to the synthetic ground truth.  Reference: visual_odometry.h:406-741."""
import numpy as np
import pytest

from oracle.ref_stereo import RefStereoVO

pytestmark = pytest.mark.gpu


def _make(ctx, seq, thr):
    import ergo_uvo_b200 as U
    p = U.default_params(True)
    p.surf_min_hessian = thr
    p.max_features = 16384
    camL = U.make_camera(seq.KL, seq.DL, seq.newKL)
    camR = U.make_camera(seq.KR, seq.DR, seq.newKR)
    return U.StereoVO(ctx, seq.w, seq.h, camL, camR, seq.R_right, seq.t_right, p), p


def _compare_frame(vo, res, ref):
    assert res.initialised == ref["initialised"]
    assert res.n_left == ref["n_left"] and res.n_right == ref["n_right"]
    kL, dL = vo.last_keypoints(False)
    kR, dR = vo.last_keypoints(True)
    assert kL.tobytes() == ref["kL"].tobytes() and kR.tobytes() == ref["kR"].tobytes()
    assert dL.shape == ref["dL"].shape
    if len(dL):
        assert np.abs(dL - ref["dL"]).max() <= 1e-4 * np.abs(ref["dL"]).max()
    assert res.n_stereo_matches == ref["n_stereo"]
    if "m_stereo" in ref:
        assert vo.last_matches(False).tobytes() == ref["m_stereo"].tobytes()
    assert res.gate == ref["gate"] and res.valid == ref["valid"]
    assert res.n_temporal_matches == ref["n_temporal"]
    if "m_temporal" in ref:
        assert vo.last_matches(True).tobytes() == ref["m_temporal"].tobytes()
    assert res.n_3d == ref["n_3d"]
    assert res.n_inliers == ref["n_inliers"] and res.hyps_evaluated == ref["hyps"]
    if "inliers" in ref:
        assert np.array_equal(vo.last_inliers(), ref["inliers"])
        assert np.abs(np.array(res.rvec) - ref["rvec"]).max() <= 1e-9
        assert np.abs(np.array(res.tvec) - ref["tvec"]).max() <= 1e-9
    v, vr = np.array(res.velocity), ref["velocity"]
    assert np.abs(v - vr).max() <= 1e-4 * max(np.abs(vr).max(), 1e-12)


def test_stereo_sequence_with_epipolar_disparity_gate(ctx, oracle, small_stereo):
    """north_star's optional stereo gate (off by default, absent from the reference): the device pipeline with the gate
    on equals the CPU replay with the same filter applied to the left-right match list, and the gate does remove
    matches on this sequence."""
    seq = small_stereo
    vo, p = _make(ctx, seq, 3000)
    vo.close()
    import ergo_uvo_b200 as U
    p.stereo_gate = 1
    p.stereo_max_epipolar_dy = 1.5
    p.stereo_min_disparity = 1.0
    p.stereo_max_disparity = 400.0
    camL = U.make_camera(seq.KL, seq.DL, seq.newKL)
    camR = U.make_camera(seq.KR, seq.DR, seq.newKR)
    vo = U.StereoVO(ctx, seq.w, seq.h, camL, camR, seq.R_right, seq.t_right, p)
    ref = RefStereoVO(oracle, seq, p)
    p0 = U.default_params(True)
    assert p0.stereo_gate == 0
    removed = 0
    for k, (L, R) in enumerate(seq.frames[:3]):
        res = vo.frame(L, R, 0.1)
        r = ref.frame(L, R, 0.1)
        _compare_frame(vo, res, r)
        ungated = oracle.match_features(r["dL"], r["dR"], np.float32(p.lowe_ratio))
        removed += len(ungated) - r["n_stereo"]
    assert removed > 0


def test_stereo_sequence_small(ctx, oracle, small_stereo):
    seq = small_stereo
    vo, p = _make(ctx, seq, 3000)
    ref = RefStereoVO(oracle, seq, p)
    dt = 0.1
    for k, (L, R) in enumerate(seq.frames):
        res = vo.frame(L, R, dt)
        r = ref.frame(L, R, dt)
        _compare_frame(vo, res, r)
        if k == 0:
            assert res.initialised == 1 and res.valid == 0
        else:
            assert res.valid == 1 and res.gate == 0
            truth = seq.true_t_prev_curr(k)
            est = np.array(res.t_prev_curr)
            assert np.linalg.norm(est - truth) < 0.15 * np.linalg.norm(truth) + 2e-3
    ms = vo.stage_ms()
    assert set(ms) and all(v >= 0 for v in ms.values())
    vo.close()


def test_stereo_full_size_frame_pair(ctx, oracle, full_stereo):
    """BASELINE config B geometry (1280x1024); threshold chosen to give about 4k keypoints"""
    seq = full_stereo
    vo, p = _make(ctx, seq, 9000)
    ref = RefStereoVO(oracle, seq, p)
    for k in range(2):
        L, R = seq.frames[k]
        res = vo.frame(L, R, 0.1)
        r = ref.frame(L, R, 0.1)
        _compare_frame(vo, res, r)
    assert res.valid == 1 and 2000 < res.n_left < 8000
    vo.close()


def test_stereo_gates_and_constant_motion(ctx, oracle, small_stereo):
    """a featureless frame triggers gate 1: validity 0 and the stale t re-published with the new dt
    (visual_odometry.h:707-717); the frame after that has an empty 'previous' set (gate 3)"""
    seq = small_stereo
    vo, p = _make(ctx, seq, 3000)
    ref = RefStereoVO(oracle, seq, p)
    flat = np.full_like(seq.frames[0][0], 127)
    frames = [seq.frames[0], seq.frames[1], (flat, flat), seq.frames[0], seq.frames[1]]
    dts = [0.1, 0.1, 0.2, 0.1, 0.1]
    gates = []
    for (L, R), dt in zip(frames, dts):
        res = vo.frame(L, R, dt)
        r = ref.frame(L, R, dt)
        _compare_frame(vo, res, r)
        gates.append(res.gate)
    assert gates == [0, 0, 1, 3, 0]
    vo.close()


def test_stereo_async_matches_sync(ctx, small_stereo):
    """enqueue/collect with two frames in flight gives the same records as the synchronous call"""
    import ctypes as C
    import torch
    seq = small_stereo
    vo, p = _make(ctx, seq, 3000)
    sync = [vo.frame(L, R, 0.1) for (L, R) in seq.frames]
    vo.close()
    vo, p = _make(ctx, seq, 3000)
    dev = [(torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()) for (L, R) in seq.frames]
    torch.cuda.synchronize()
    for (L, R) in dev:
        vo.enqueue_device(L.data_ptr(), R.data_ptr(), 3 * seq.w, 0.1)
    for s in sync:
        r = vo.collect()
        assert bytes(r) == bytes(s)
    vo.close()


def test_stereo_capacity_overflow_is_an_error(ctx, small_stereo):
    import ergo_uvo_b200 as U
    seq = small_stereo
    p = U.default_params(True)
    p.surf_min_hessian = 100
    p.max_features = 256
    vo = U.StereoVO(ctx, seq.w, seq.h, U.make_camera(seq.KL, seq.DL, seq.newKL),
                    U.make_camera(seq.KR, seq.DR, seq.newKR), seq.R_right, seq.t_right, p)
    with pytest.raises(U.UvoError) as e:
        vo.frame(*seq.frames[0], 0.1)
    assert e.value.code == -4
    vo.close()


def test_stereo_host_async_pipeline_matches_sync(ctx, small_stereo):
    """uvo_stereo_enqueue_host with several frames in flight (frames overlap on separate lanes/streams inside the
    library) returns exactly the records of the one-frame-at-a-time synchronous call, over more frames than lanes"""
    import torch
    seq = small_stereo
    vo, p = _make(ctx, seq, 3000)
    order = [k % len(seq.frames) for k in range(vo.lanes() + 3)]
    sync = [vo.frame(*seq.frames[k], 0.1) for k in order]
    vo.close()
    vo, p = _make(ctx, seq, 3000)
    pinned = [(torch.from_numpy(L).pin_memory(), torch.from_numpy(R).pin_memory()) for (L, R) in seq.frames]
    got, q = [], 0
    for k in order:
        L, R = pinned[k]
        vo.enqueue_host(L.data_ptr(), R.data_ptr(), 3 * seq.w, 0.1)
        q += 1
        if q >= 4:
            got.append(vo.collect())
            q -= 1
    while q:
        got.append(vo.collect())
        q -= 1
    assert vo.max_in_flight() >= 4
    for a, b in zip(got, sync):
        assert bytes(a) == bytes(b)
    vo.close()


def test_bayer_input_equals_demosaiced_input(ctx, oracle, small_stereo):
    """uvo_stereo_enqueue_host_bayer(bayer) gives the same frames as uvo_stereo_frame on the images the CPU demosaic
    (cvtColor BayerBGGR2BGR, as from_ros_to_cv_image does) produces from the same bayer data."""
    seq = small_stereo

    def mosaic(img3):  # sample a BGGR mosaic out of a colour image: (even, even) blue ... (odd, odd) red
        m = img3[:, :, 1].copy()
        m[0::2, 0::2] = img3[0::2, 0::2, 0]
        m[1::2, 1::2] = img3[1::2, 1::2, 2]
        return np.ascontiguousarray(m)
    vo_a, p = _make(ctx, seq, 3000)
    vo_b, _ = _make(ctx, seq, 3000)
    for k, (L, R) in enumerate(seq.frames[:3]):
        bl, br = mosaic(L), mosaic(R)
        ra = vo_a.frame(oracle.bayer_bggr2bgr(bl), oracle.bayer_bggr2bgr(br), 0.1)
        vo_b.enqueue_host_bayer(bl.ctypes.data, br.ctypes.data, bl.strides[0], 0.1)
        rb = vo_b.collect()
        for f in ("initialised", "valid", "n_left", "n_right", "n_stereo_matches", "n_temporal_matches", "n_3d",
                  "n_inliers", "hyps_evaluated", "gate"):
            assert getattr(ra, f) == getattr(rb, f), (k, f)
        assert list(ra.rvec) == list(rb.rvec) and list(ra.tvec) == list(rb.tvec) and list(ra.velocity) == list(rb.velocity)
        ka, _ = vo_a.last_keypoints(False)
        kb, _ = vo_b.last_keypoints(False)
        assert ka.tobytes() == kb.tobytes() and len(ka) > 100
    vo_a.close()
    vo_b.close()


def test_stereo_graph_replay_matches_direct_launches(ctx, small_stereo):
    """the asynchronous path replays each lane's kernel runs as CUDA graphs from the lane's second frame on: five
    frames per lane (direct, capture + launch, three replays) give byte-identical result records and identical
    intermediate products to the synchronous one-frame-at-a-time path, which launches every kernel directly"""
    import torch
    seq = small_stereo
    vo, p = _make(ctx, seq, 3000)
    lanes = vo.lanes()
    order = [k % len(seq.frames) for k in range(5 * lanes)]
    sync = [vo.frame(*seq.frames[k], 0.1) for k in order]
    taps = (vo.last_keypoints(False), vo.last_keypoints(True), vo.last_matches(False), vo.last_matches(True),
            vo.last_inliers())
    assert vo.graph_launches == 0          # the synchronous call times its stages: no graphs
    vo.close()
    pinned = [(torch.from_numpy(L).pin_memory(), torch.from_numpy(R).pin_memory()) for (L, R) in seq.frames]

    def run(graphs):
        vo, _ = _make(ctx, seq, 3000)
        vo.set_graphs(graphs)
        got, q = [], 0
        for k in order:
            L, R = pinned[k]
            vo.enqueue_host(L.data_ptr(), R.data_ptr(), 3 * seq.w, 0.1)
            q += 1
            if q >= lanes:
                got.append(vo.collect())
                q -= 1
        while q:
            got.append(vo.collect())
            q -= 1
        t = (vo.last_keypoints(False), vo.last_keypoints(True), vo.last_matches(False), vo.last_matches(True),
             vo.last_inliers())
        n = vo.graph_launches
        vo.close()
        return got, t, n

    got, t, n = run(True)
    assert n == 3 * (len(order) - lanes)   # three graph launches per frame after each lane's first frame
    assert any(r.valid for r in got)
    for a, b in zip(got, sync):
        assert bytes(a) == bytes(b)
    for a, b in zip(t, taps):
        if isinstance(a, tuple):
            assert all(x.tobytes() == y.tobytes() for x, y in zip(a, b))
        else:
            assert a.tobytes() == b.tobytes()
    got2, _, n2 = run(False)
    assert n2 == 0
    for a, b in zip(got2, sync):
        assert bytes(a) == bytes(b)
