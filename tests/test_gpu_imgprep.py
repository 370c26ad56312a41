"""K1-K3 parity: CUDA get_image / integral vs the CPU oracle -- bit-exact (integer / exactly specified f32 work).
Reference path: get_image, VO_utility.cpp:337-379; integral inside SURF::detectAndCompute, VO_utility.cpp:118."""
import numpy as np
import pytest

from conftest import noise_image

pytestmark = pytest.mark.gpu


def _cams(w, h):
    from tools import synth
    K, D = synth.scaled_camera(synth.STEREO_YAML["left"], w, 1280)
    K[1, 2] = h * 0.5 + 5
    D = D.copy()
    D[2:] = (0.001, -0.002)  # exercise the tangential terms too
    return K, D, synth.optimal_new_camera_matrix(K, D, w, h)


@pytest.mark.parametrize("w,h,clip", [(640, 480, 3), (1280, 1024, 8), (333, 251, 8), (336, 251, 3), (64, 48, 2)])
def test_get_image_bit_exact(ctx, oracle, w, h, clip):
    img = noise_image(h, w, seed=w + h, channels=3)
    K, D, newK = _cams(w, h)
    ctx.params.clahe, ctx.params.clip_limit = 1, clip
    got = ctx.get_image(img, K, D, newK)
    ref = oracle.get_image(img, K, D, newK, True, float(clip))
    assert got.shape == ref.shape
    assert int((got != ref).sum()) == 0


def test_get_image_no_clahe(ctx, oracle):
    w, h = 640, 480
    img = noise_image(h, w, seed=5, channels=3)
    K, D, newK = _cams(w, h)
    ctx.params.clahe = 0
    got = ctx.get_image(img, K, D, newK)
    ctx.params.clahe = 1
    assert int((got != oracle.get_image(img, K, D, newK, False, 0.0)).sum()) == 0


def test_get_image_rejects_gray_input(ctx):
    with pytest.raises(ValueError):
        ctx.get_image(np.zeros((48, 64), np.uint8), np.eye(3), np.zeros(4), np.eye(3))


@pytest.mark.parametrize("w,h", [(1280, 1024), (641, 479), (31, 7), (2448, 2048)])
def test_integral_bit_exact(ctx, oracle, w, h):
    g = noise_image(h, w, seed=11)
    if w == 2448:
        g[:] = 255  # maximum-size case: largest sums stay below 2^31
    got = ctx.integral(g)
    assert int((got != oracle.integral(g)).sum()) == 0


def test_get_image_golden(ctx):
    """committed cv2 fixture (tools/make_golden.py): gray+undistort+CLAHE of a 320x240 image"""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "imgprep_320x240.npz"))
    ctx.params.clahe, ctx.params.clip_limit = 1, int(z["clip"])
    got = ctx.get_image(z["img"], z["K"], z["D"], z["newK"])
    assert int((got != z["out"]).sum()) == 0
    assert int((ctx.integral(z["out"]) != z["integral"]).sum()) == 0


@pytest.mark.parametrize("sw,sh,dw", [(2564, 2048, 640), (1282, 1024, 640), (1280, 1024, 640), (1000, 750, 333),
                                      (2448, 2048, 1280), (641, 480, 640), (1920, 1080, 640)])
def test_resize_area_matches_oracle(ctx, oracle, sw, sh, dw):
    """K0 (VO_utility.cpp:362-363): cv::resize(INTER_AREA), 3-channel; the oracle is bit-exact with cv2 on these sizes
    (tests/test_oracle_imgprep.py::test_resize_area_3ch_cv2)"""
    rs = np.random.RandomState(sw + sh)
    img = rs.randint(0, 256, (sh, sw, 3)).astype(np.uint8)
    dh = int(sh / (sw / dw))
    assert np.array_equal(ctx.resize_area(img, dw, dh), oracle.resize_area(img, dw, dh))
    g = img[:, :, 1].copy()
    assert np.array_equal(ctx.resize_area(g, dw, dh), oracle.resize_area(g, dw, dh))


def test_get_image_resized_branch(ctx, oracle):
    """whole pre-scaling branch of get_image: resize -> gray -> undistort -> CLAHE at DESIRED_WIDTH = 640"""
    from tools import synth
    seq = synth.MonoSequence(1282, 962, n_frames=1, tex_size=1024)
    small = synth.MonoSequence(640, 480, n_frames=1, tex_size=1024)  # camera model at the working resolution
    img = seq.frames[0]
    dh = int(962 / (1282 / 640))
    out = ctx.get_image_resized(img, 640, small.K, small.D, small.newK)
    ref = oracle.get_image(oracle.resize_area(img, 640, dh), small.K, small.D, small.newK, True,
                           float(ctx.params.clip_limit))
    assert out.shape == (dh, 640) and np.array_equal(out, ref)


@pytest.mark.parametrize("h,w", [(3, 3), (4, 5), (7, 9), (480, 640), (1024, 1280), (1081, 1921)])
def test_bayer_demosaic(ctx, oracle, h, w):
    """uvo_demosaic_bggr2bgr == cvtColor(COLOR_BayerBGGR2BGR) (oracle pinned to cv2): interior means and the copied
    border rows / columns, odd and even sizes"""
    b = np.random.RandomState(h * w).randint(0, 256, (h, w)).astype(np.uint8)
    assert np.array_equal(ctx.demosaic_bggr2bgr(b), oracle.bayer_bggr2bgr(b))
