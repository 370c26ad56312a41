"""The screen of k_surf_detect (csrc/surf.cu, screen_tile0 / screen_at) may only discard samples the exact lazy rule
discards too.  This restates both sides in numpy -- the exact Dxx / Dyy of calcLayerDetAndTrace (f32 products, f64 sum,
f32 result) and the screen's single-product form -- and checks the error bound the kernel relies on
(|dx dy - ax ay| < 1) on noise, saturated and synthetic images, for every middle-layer size of four octaves."""
import numpy as np
import pytest

from tools import synth


def _round_half_even(v):
    return int(np.rint(np.float32(v)))


def _boxes(size):
    """resizeHaarPattern for the Dxx pattern {0,2,3,7,1},{3,2,6,7,-2},{6,2,9,7,1}: column edges, row edges, weights"""
    ratio = np.float32(size) / np.float32(9)
    e = [_round_half_even(ratio * np.float32(k)) for k in range(10)]
    cols = [e[0], e[3], e[6], e[9]]
    rows = [e[2], e[7]]
    w = []
    for k, wt in enumerate((1, -2, 1)):
        area = (cols[k + 1] - cols[k]) * (rows[1] - rows[0])
        w.append(np.float32(wt) / np.float32(area))
    return cols, rows, w


def _dxx_exact_and_screen(img, size, step):
    """exact dx and screened ax of every sample of a layer (Dyy is the same computation on the transposed image)"""
    S = np.zeros((img.shape[0] + 1, img.shape[1] + 1), np.int64)
    S[1:, 1:] = img.astype(np.int64).cumsum(0).cumsum(1)
    cols, rows, w = _boxes(size)
    ni, nj = 1 + (img.shape[0] - size) // step, 1 + (img.shape[1] - size) // step
    ii, jj = np.meshgrid(np.arange(ni) * step, np.arange(nj) * step, indexing="ij")
    v = []
    for k in range(3):
        v.append(S[ii + rows[0], jj + cols[k]] + S[ii + rows[1], jj + cols[k + 1]] - S[ii + rows[1], jj + cols[k]] -
                 S[ii + rows[0], jj + cols[k + 1]])
    d = np.zeros(ii.shape, np.float64)
    for k in range(3):
        d += (v[k].astype(np.float32) * w[k]).astype(np.float64)
    dx = d.astype(np.float32)
    assert w[2] == w[0] and w[1] == np.float32(-2) * w[0]
    comb = v[0] - 2 * v[1] + v[2]
    assert np.abs(comb).max() < 2 ** 24
    ax = comb.astype(np.float32) * w[0]
    return dx, ax


def _images():
    rs = np.random.RandomState(5)
    h, w = 240, 320
    yield "noise", rs.randint(0, 256, (h, w)).astype(np.uint8)
    yield "saturated checker", (((np.add.outer(np.arange(h) // 7, np.arange(w) // 5)) & 1) * 255).astype(np.uint8)
    yield "all 255", np.full((h, w), 255, np.uint8)
    seq = synth.StereoSequence(w, h, n_frames=1, tex_size=512)
    yield "synthetic frame", np.ascontiguousarray(seq.frames[0][0][..., 1])


@pytest.mark.parametrize("name,img", list(_images()), ids=lambda v: v if isinstance(v, str) else "")
def test_screen_never_discards_a_sample_the_lazy_rule_keeps(name, img):
    worst = 0.0
    for o in range(4):
        step = 1 << o
        for l in (1, 2, 3):
            size = (9 + 6 * l) << o
            if size > min(img.shape):
                continue
            dx, ax = _dxx_exact_and_screen(img, size, step)
            dy, ay = _dxx_exact_and_screen(np.ascontiguousarray(img.T), size, step)
            dy, ay = dy.T, ay.T
            assert np.abs(dx - ax).max() <= 1.3e-4 and np.abs(dy - ay).max() <= 1.3e-4
            exact = (dx * dy).astype(np.float32)
            screen = (ax * ay).astype(np.float32)
            worst = max(worst, float(np.abs(exact.astype(np.float64) - screen).max()))
            for thr in (0.5, 50.0, 1500.0, 11032.0):
                thr32 = np.float32(thr)
                thr_skip = np.nextafter(np.float32(thr32 - np.float32(1) - abs(thr32) * np.float32(1e-6)),
                                        np.float32(-np.inf))
                discarded = ~(screen > thr_skip)
                assert not np.any(discarded & (exact > thr32)), (name, size, thr)
    assert worst < 0.2  # the bound in the kernel's comment is 0.16; the margin used is 1
