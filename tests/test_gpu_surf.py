"""K4-K7 parity: CUDA SURF vs the CPU oracle (oracle/surf.cpp).  Keypoints bit-exact (set, order and
every field); descriptors within 1e-4 relative (observed: bit-exact).  Reference path: detect_features,
VO_utility.cpp:114-119."""
import numpy as np
import pytest

from conftest import noise_image

pytestmark = pytest.mark.gpu


def _check(ctx, oracle, g, thr, upright=True):
    ctx.params.surf_min_hessian = thr
    ctx.params.surf_upright = int(upright)
    ctx.params.max_features = 1 << 16
    k, d = ctx.detect_features(g)
    ctx.params.surf_upright = 1
    ko, do = oracle.surf_detect_and_compute(g, thr, upright=upright)
    assert len(k) == len(ko)
    return k, d, ko, do


@pytest.mark.parametrize("w,h,thr", [(640, 480, 50), (1280, 1024, 1500), (417, 303, 300), (160, 120, 10)])
def test_surf_upright_matches_oracle(ctx, oracle, w, h, thr):
    g = noise_image(h, w, seed=3 * w + h)
    k, d, ko, do = _check(ctx, oracle, g, thr)
    assert len(k) > 50
    for f in ("x", "y", "size", "angle", "response", "octave", "class_id"):
        assert np.array_equal(k[f], ko[f]), f
    rel = np.abs(d - do).max() / max(np.abs(do).max(), 1e-12)
    assert rel <= 1e-4
    assert np.array_equal(d.view(np.uint32), do.view(np.uint32))  # stronger: bit-exact in practice


def test_surf_synthetic_frame(ctx, oracle, full_stereo):
    L, _ = full_stereo.frames[0]
    g = oracle.get_image(L, full_stereo.KL, full_stereo.DL, full_stereo.newKL, True, 8.0)
    k, d, ko, do = _check(ctx, oracle, g, 1500)
    assert k.tobytes() == ko.tobytes()
    assert np.abs(d - do).max() <= 1e-4 * np.abs(do).max()
    # order is OpenCV's KeypointGreater: response descending
    assert np.all(np.diff(k["response"]) <= 0)
    # descriptors are unit vectors
    assert np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-5)


def test_surf_empty_and_tiny(ctx, oracle):
    g = np.full((120, 160), 128, np.uint8)  # flat image: no keypoints
    ctx.params.surf_min_hessian = 100
    k, d = ctx.detect_features(g)
    assert len(k) == 0 and d.shape == (0, 64)
    g = noise_image(40, 50, seed=2)  # smaller than most layers
    k, d, ko, do = _check(ctx, oracle, g, 1)
    assert k.tobytes() == ko.tobytes()


def test_surf_capacity_error(ctx):
    import ergo_uvo_b200 as U
    g = noise_image(480, 640, seed=4)
    ctx.params.surf_min_hessian = 1
    ctx.params.max_features = 64
    with pytest.raises(U.UvoError) as e:
        ctx.detect_features(g, capacity=64)
    assert e.value.code == -4
    ctx.params.max_features = 1 << 16


def test_surf_oriented_matches_oracle(ctx, oracle):
    """upright=0: orientation assignment + rotated window.  Keypoint sets identical; angles and descriptors within
    tolerance (sin/cos of the direction go through device libm)."""
    g = noise_image(480, 640, seed=77)
    k, d, ko, do = _check(ctx, oracle, g, 400, upright=False)
    for f in ("x", "y", "size", "response", "octave"):
        assert np.array_equal(k[f], ko[f]), f
    assert np.array_equal(k["angle"], ko["angle"])
    close = (np.abs(d - do).max(axis=1) <= 1e-4 * np.abs(do).max())
    assert close.mean() > 0.999


@pytest.mark.parametrize("n,capacity", [(1, 64), (2, 64), (37, 64), (1000, 4096), (4097, 8192), (9000, 16384),
                                        (3000, 20000)])
def test_keypoint_sort_with_ties(ctx, n, capacity):
    """K6 in isolation on adversarial input: responses quantised to a handful of values so that most of the order is
    decided by the tie-breakers (size desc, octave desc, y desc, x asc), plus fully identical keypoints, which must
    keep their input order.  capacity <= 16384 runs the single-block bitonic sort, above it the rank sort."""
    import ergo_uvo_b200 as U
    rs = np.random.RandomState(n)
    k = np.zeros(n, U.KEYPOINT_DTYPE)
    k["response"] = rs.randint(0, 6, n).astype(np.float32) * 1000 + 500
    k["size"] = rs.choice([15, 21, 27, 30], n).astype(np.float32)
    k["octave"] = rs.randint(0, 3, n)
    k["y"] = rs.randint(0, 8, n).astype(np.float32)
    k["x"] = rs.randint(0, 8, n).astype(np.float32)
    k["class_id"] = np.arange(n)  # not part of the key: identifies the input position
    k["angle"] = -1
    got = ctx.sort_keypoints(k, capacity)
    order = sorted(range(n), key=lambda i: (-k["response"][i], -k["size"][i], -k["octave"][i], -k["y"][i], k["x"][i], i))
    assert np.array_equal(got["class_id"], np.array(order))
    assert got.tobytes() == k[order].tobytes()
