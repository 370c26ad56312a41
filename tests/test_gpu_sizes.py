"""Parity at the other BASELINE.json sizes: 1920x1080 (config C, ~8k features, ratio 0.7) and 2448x2048 (config E):
get_image, integral, SURF keypoints / descriptors and the matcher against the CPU oracle on one synthetic stereo pair
each.  Guards the size-dependent paths: CLAHE tile geometry, int32 integral head-room (1.28e9 at 5 MP), SURF tile
grid, windows larger than the descriptor row buffer, > 8k descriptors per set in the tensor-core matcher."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("w,h,target,ratio", [(1920, 1080, 8192, 0.7), (2448, 2048, 6000, 0.8)])
def test_front_end_and_matcher_at_size(ctx, oracle, w, h, target, ratio):
    from tools import synth
    seq = synth.StereoSequence(w, h, n_frames=1, tex_size=2048)
    L, R = seq.frames[0]
    gL = ctx.get_image(L, seq.KL, seq.DL, seq.newKL)
    gR = ctx.get_image(R, seq.KR, seq.DR, seq.newKR)
    assert np.array_equal(gL, oracle.get_image(L, seq.KL, seq.DL, seq.newKL, True, float(ctx.params.clip_limit)))
    assert np.array_equal(ctx.integral(gL), oracle.integral(gL))
    # threshold for roughly `target` keypoints (bisected on the oracle so both sides use the same value)
    lo, hi, thr = 100, 400000, None
    while lo < hi:
        mid = (lo + hi) // 2
        n = len(oracle.surf_detect_and_compute(gL, mid)[0])
        thr = mid
        if abs(n - target) <= 0.05 * target:
            break
        if n > target:
            lo = mid + 1
        else:
            hi = mid
    ctx.params.surf_min_hessian = thr
    ctx.params.max_features = 1 << 15
    kL, dL = ctx.detect_features(gL)
    kR, dR = ctx.detect_features(gR)
    koL, doL = oracle.surf_detect_and_compute(gL, thr)
    koR, doR = oracle.surf_detect_and_compute(gR, thr)
    assert kL.tobytes() == koL.tobytes() and kR.tobytes() == koR.tobytes()
    assert np.abs(dL - doL).max() <= 1e-4 * np.abs(doL).max()
    assert np.abs(dR - doR).max() <= 1e-4 * np.abs(doR).max()
    ctx.params.lowe_ratio = ratio
    m = ctx.match_features(None, None, dL, dR)
    assert m.tobytes() == oracle.match_features(doL, doR, np.float32(ratio)).tobytes()
    assert ctx.knn_match2(dL, dR).tobytes() == oracle.knn2(doL, doR).tobytes()
    assert ctx.match_last_fallbacks() <= len(dL) // 20
    ctx.params.lowe_ratio = 0.8
    ctx.params.max_features = 16384
