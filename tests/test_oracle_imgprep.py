"""CPU: pins the oracle's image-prep restatement (oracle/imgprep.cpp) against OpenCV -- live cv2 when importable, and
the committed cv2 fixture otherwise.  Everything here is bit-exact.  """
import os

import numpy as np
import pytest

from conftest import noise_image

GOLD = os.path.join(os.path.dirname(__file__), "golden")
cv2 = pytest.importorskip("cv2") if os.environ.get("UVO_REQUIRE_CV2") else None
try:
    import cv2 as _cv2
    cv2 = _cv2
except Exception:  # pragma: no cover
    cv2 = None


def test_golden_fixture(oracle):
    z = np.load(os.path.join(GOLD, "imgprep_320x240.npz"))
    assert np.array_equal(oracle.gray(z["img"]), z["gray"])
    assert np.array_equal(oracle.undistort(z["gray"], z["K"], z["D"], z["newK"]), z["und"])
    assert np.array_equal(oracle.clahe(z["und"], float(z["clip"])), z["out"])
    assert np.array_equal(oracle.get_image(z["img"], z["K"], z["D"], z["newK"], True, float(z["clip"])), z["out"])
    assert np.array_equal(oracle.integral(z["out"]), z["integral"])
    for s in (25, 42, 57, 63, 100):
        assert np.array_equal(oracle.resize_area(np.ascontiguousarray(z["out"][:s, :s]), 21, 21), z[f"patch_{s}"])
    assert np.array_equal(oracle.gaussian_kernel_f32(13, 2.5).view(np.uint32), z["g13"].view(np.uint32))


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
@pytest.mark.parametrize("w,h", [(640, 480), (333, 251), (336, 251)])
def test_against_live_cv2(oracle, w, h):
    from tools import synth
    img = noise_image(h, w, seed=w, channels=3)
    K, D = synth.scaled_camera(synth.STEREO_YAML["left"], w, 1280)
    K[1, 2] = h / 2 + 3
    D[2:] = (0.001, -0.002)
    newK, _ = cv2.getOptimalNewCameraMatrix(K, D, (w, h), 0, (w, h), 0)
    gray = cv2.cvtColor(img, cv2.COLOR_RGB2GRAY)
    assert np.array_equal(oracle.gray(img), gray)
    m1, m2 = cv2.initUndistortRectifyMap(K, D, None, newK, (w, h), cv2.CV_16SC2)
    mxy, mfr = oracle.undistort_map(K, D, newK, w, h)
    assert np.array_equal(m1, mxy) and np.array_equal(m2, mfr)
    und = cv2.undistort(gray, K, D, None, newK)
    assert np.array_equal(oracle.undistort(gray, K, D, newK), und)
    for clip in (3.0, 8.0):
        cl = cv2.createCLAHE()
        cl.setClipLimit(clip)
        assert np.array_equal(oracle.clahe(und, clip), cl.apply(und))
    assert np.array_equal(oracle.integral(und), cv2.integral(und, sdepth=cv2.CV_32S))


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_resize_area_every_window_size(oracle):
    """INTER_AREA to 21x21 for every SURF window size 21..760 on white noise"""
    g = np.random.RandomState(3).randint(0, 256, (800, 800)).astype(np.uint8)
    for s in list(range(21, 200)) + list(range(200, 760, 7)):
        win = np.ascontiguousarray(g[:s, :s])
        assert np.array_equal(oracle.resize_area(win, 21, 21), cv2.resize(win, (21, 21), interpolation=cv2.INTER_AREA)), s


def test_optimal_new_camera_matrix_close_to_cv2():
    if cv2 is None:
        pytest.skip("cv2 not importable")
    from tools import synth
    K, D = synth.scaled_camera(synth.STEREO_YAML["left"], 1280, 1280)
    mine = synth.optimal_new_camera_matrix(K, D, 1280, 1024)
    ref, _ = cv2.getOptimalNewCameraMatrix(K, D, (1280, 1024), 0, (1280, 1024), 0)
    assert np.abs(mine - ref).max() < 1e-3 * ref[0, 0]  # cv2's undistortPoints stops after 5 iterations; ours converges


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
@pytest.mark.parametrize("sw,sh,dw", [(2564, 2048, 640), (1282, 1024, 640), (1280, 1024, 640), (1000, 750, 333),
                                      (641, 480, 640)])
def test_resize_area_3ch_cv2(oracle, sw, sh, dw):
    """K0: the pre-scaling cv::resize(INTER_AREA) of get_image (VO_utility.cpp:362-363), 3 interleaved channels,
    different x / y scales -- bit-exact with cv2"""
    rs = np.random.RandomState(sw)
    img = rs.randint(0, 256, (sh, sw, 3)).astype(np.uint8)
    dh = int(sh / (sw / dw))
    assert np.array_equal(oracle.resize_area(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_AREA))


def test_bayer_demosaic_matches_cv2(oracle):
    """cvtColor(COLOR_BayerBGGR2BGR) of from_ros_to_cv_image (math_utility.cpp:161-164): interior and border rule"""
    cv2 = pytest.importorskip("cv2")
    rs = np.random.RandomState(3)
    for (h, w) in [(3, 3), (4, 4), (5, 4), (7, 9), (12, 14), (13, 15), (480, 640), (1024, 1280)]:
        b = rs.randint(0, 256, (h, w)).astype(np.uint8)
        assert np.array_equal(oracle.bayer_bggr2bgr(b), cv2.cvtColor(b, cv2.COLOR_BayerBGGR2BGR)), (h, w)
