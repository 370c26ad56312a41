"""CPU: the SURF restatement (oracle/surf.cpp).  PARITY UNPINNED against real OpenCV SURF (contrib is not available);
what can be pinned is pinned: Gaussian tables, fastAtan2/phase, the INTER_AREA patch, the integral.  The rest are
self-consistency properties of the algorithm."""
import os

import numpy as np
import pytest

from conftest import noise_image

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def test_gaussian_tables(oracle):
    """DESC_SIGMA is the float constant 3.3f in OpenCV: the 20-tap table uses sigma = (double)3.3f"""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "imgprep_320x240.npz"))
    assert np.array_equal(oracle.gaussian_kernel_f32(13, 2.5).view(np.uint32), z["g13"].view(np.uint32))
    if cv2 is not None:
        s = float(np.float32(3.3))
        assert np.array_equal(oracle.gaussian_kernel_f32(20, s).view(np.uint32),
                              cv2.getGaussianKernel(20, s, cv2.CV_32F).ravel().view(np.uint32))
        for n, sg in ((7, 1.1), (8, 2.0), (13, 2.5), (21, 4.0)):
            assert np.array_equal(oracle.gaussian_kernel_f32(n, sg).view(np.uint32),
                                  cv2.getGaussianKernel(n, sg, cv2.CV_32F).ravel().view(np.uint32))


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_fast_atan2(oracle):
    rs = np.random.RandomState(1)
    for y, x in zip(rs.randn(2000).astype(np.float32), rs.randn(2000).astype(np.float32)):
        assert oracle.fast_atan2(float(y), float(x)) == np.float32(cv2.fastAtan2(float(y), float(x)))


def test_surf_basic_properties(oracle):
    g = noise_image(480, 640, seed=9)
    k, d = oracle.surf_detect_and_compute(g, 100)
    assert len(k) > 500 and d.shape == (len(k), 64)
    r = k["response"]
    assert np.all(r > 100) and np.all(np.diff(r) <= 0)           # thresholded, KeypointGreater order
    assert np.all(k["angle"] == 270) and np.all(k["class_id"] == -1)
    assert set(np.unique(k["octave"])) <= {0, 1, 2, 3}
    assert np.all(k["size"] == np.rint(k["size"])) and k["size"].min() >= 9
    assert np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-5)
    k2, _ = oracle.surf_detect_and_compute(g, 1000)                # raising the threshold selects a prefix
    assert len(k2) < len(k) and k2.tobytes() == k[:len(k2)].tobytes()


def test_surf_translation_covariance(oracle):
    """shifting the image by whole pixels shifts interior keypoints by the same amount"""
    g = noise_image(300, 400, seed=4)
    a, _ = oracle.surf_detect_and_compute(g[:, :360], 300)
    b, _ = oracle.surf_detect_and_compute(g[:, 16:376], 300)
    ia = {(round(float(x), 3), round(float(y), 3), float(s)) for x, y, s in zip(a["x"], a["y"], a["size"])
          if 120 < x < 240}
    ib = {(round(float(x) + 16, 3), round(float(y), 3), float(s)) for x, y, s in zip(b["x"], b["y"], b["size"])
          if 120 < x + 16 < 240}
    assert len(ia) > 20 and len(ia & ib) >= 0.95 * len(ia)


def test_det_trace_layer_against_brute_force(oracle):
    """box-filter responses of one layer vs a direct evaluation on the image (no integral)"""
    g = noise_image(64, 80, seed=2)
    sum_ = oracle.integral(g)
    det, tr = oracle.surf_det_trace_layer(sum_, 9, 1)
    gi = g.astype(np.int64)
    def box(y, x, x1, y1, x2, y2):
        return gi[y + y1:y + y2, x + x1:x + x2].sum()
    for (i, j) in [(0, 0), (10, 17), (55, 71)]:
        dx = (box(i, j, 0, 2, 3, 7) - 2 * box(i, j, 3, 2, 6, 7) + box(i, j, 6, 2, 9, 7)) / 15.0
        dy = (box(i, j, 2, 0, 7, 3) - 2 * box(i, j, 2, 3, 7, 6) + box(i, j, 2, 6, 7, 9)) / 15.0
        dxy = (box(i, j, 1, 1, 4, 4) - box(i, j, 5, 1, 8, 4) - box(i, j, 1, 5, 4, 8) + box(i, j, 5, 5, 8, 8)) / 9.0
        assert np.isclose(det[i + 4, j + 4], dx * dy - 0.81 * dxy * dxy, rtol=1e-5)
        assert np.isclose(tr[i + 4, j + 4], dx + dy, rtol=1e-5)
    assert det[0, 0] == 0 and det[63, 79] == 0  # never-written border


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_upright_patch_is_cv2_resize_of_the_window(oracle):
    g = noise_image(200, 240, seed=6)
    for (cx, cy, size) in [(100.3, 90.7, 15.0), (5.0, 6.0, 27.0), (120.0, 100.0, 66.0)]:
        patch, ws = oracle.surf_patch(g, cx, cy, size)
        s = np.float32(size) * np.float32(1.2) / np.float32(9.0)
        assert ws == int(np.float32(21) * s)
        off = -np.float32(ws - 1) / np.float32(2)
        sx, sy = int(np.rint(np.float32(cx) + off)), int(np.rint(np.float32(cy) - off))
        ii, jj = np.meshgrid(np.arange(ws), np.arange(ws), indexing="ij")
        win = g[np.clip(sy - jj, 0, 199), np.clip(sx + ii, 0, 239)]
        assert np.array_equal(patch, cv2.resize(np.ascontiguousarray(win), (21, 21), interpolation=cv2.INTER_AREA))


def test_extended_descriptor_refines_the_standard_one(oracle):
    """SURF_EXTENDED (VO_utility.h:86, VO_utility.cpp:117): the 128-d vector splits each of the four sums of a cell by
    the sign of the other gradient, so pairwise sums of its entries give back the 64-d vector up to the normalisation
    and f32 summation order; the keypoints do not depend on the flag."""
    g = noise_image(240, 320, seed=9)
    k64, d64 = oracle.surf_detect_and_compute(g, 200)
    k128, d128 = oracle.surf_detect_and_compute(g, 200, extended=True)
    assert len(k64) > 50 and k64.tobytes() == k128.tobytes()
    assert d128.shape == (len(k64), 128)
    assert np.allclose(np.linalg.norm(d128.astype(np.float64), axis=1), 1.0, atol=1e-5)
    e = d128.astype(np.float64).reshape(-1, 16, 8)
    s = np.stack([e[..., 0] + e[..., 2], e[..., 4] + e[..., 6], e[..., 1] + e[..., 3], e[..., 5] + e[..., 7]], -1)
    s = s.reshape(-1, 64)
    s /= np.linalg.norm(s, axis=1, keepdims=True)
    assert np.abs(s - d64).max() < 1e-5
    # |dx| sums dominate the signed ones, half by half
    assert np.all(e[..., 1] >= np.abs(e[..., 0]) - 1e-6) and np.all(e[..., 7] >= np.abs(e[..., 6]) - 1e-6)
