"""CPU: pins the JPEG restatement (oracle/jpeg.cpp) -- the decode inside from_ros_to_cv_image (math_utility.cpp:154-173:
cv_bridge::toCvCopy -> cv::imdecode -> libjpeg-turbo, JDCT_ISLOW + fancy upsampling) -- bit-exact against the committed
cv2 fixture (tests/golden/jpeg_64x48.npz, tools/make_golden.py) and against live cv2 where importable."""
import io
import os

import numpy as np
import pytest

from conftest import noise_image

GOLD = os.path.join(os.path.dirname(__file__), "golden")
try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def test_jpeg_golden(oracle):
    z = np.load(os.path.join(GOLD, "jpeg_64x48.npz"))
    names = sorted(k[:-4] for k in z.files if k.endswith("_jpg"))
    assert names == ["c420_rst2", "c422", "c440", "c444_opt", "gray"]
    for n in names:
        got = oracle.jpeg_decode(z[n + "_jpg"].tobytes())
        assert got.shape == z[n + "_img"].shape, n
        assert np.array_equal(got, z[n + "_img"]), n


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
@pytest.mark.parametrize("sampling", ["444", "422", "420", "440", "411"])
def test_jpeg_live_cv2_sizes_and_qualities(oracle, sampling):
    """every MCU geometry, image sizes that leave partial MCUs, chroma planes of 1-2 samples (where libjpeg drops
    the triangle filter), qualities from heavy clipping to near-lossless"""
    sf = getattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR_" + sampling)
    rs = np.random.RandomState(int(sampling))
    for h, w in [(1, 1), (2, 2), (3, 5), (5, 3), (8, 8), (17, 33), (33, 17), (31, 47), (100, 6), (2, 37), (243, 317)]:
        for kind in ("noise", "smooth"):
            img = (rs.randint(0, 256, (h, w, 3)).astype(np.uint8) if kind == "noise"
                   else noise_image(max(h, 40), max(w, 40), seed=h * w, channels=3)[:h, :w])
            for q in (15, 75, 100):
                ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf])
                ref = cv2.imdecode(enc, cv2.IMREAD_UNCHANGED)
                got = oracle.jpeg_decode(enc.tobytes())
                assert got.shape == ref.shape and np.array_equal(got, ref), (h, w, kind, q)


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_jpeg_live_cv2_restarts_tables_gray_and_frame(oracle):
    img = noise_image(243, 317, seed=5, channels=3)
    cases = []
    for rst in (1, 2, 7, 100):
        for sf in (cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444):
            cases.append((img, [cv2.IMWRITE_JPEG_QUALITY, 80, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf,
                                cv2.IMWRITE_JPEG_RST_INTERVAL, rst]))
    cases.append((img, [cv2.IMWRITE_JPEG_QUALITY, 80, cv2.IMWRITE_JPEG_OPTIMIZE, 1]))
    cases.append((img[:, :, 1].copy(), [cv2.IMWRITE_JPEG_QUALITY, 50]))                     # the bayer-mosaic case: 1 component
    cases.append((noise_image(1024, 1280, seed=9, channels=3), [cv2.IMWRITE_JPEG_QUALITY, 90]))  # BASELINE config B size
    for src, flags in cases:
        ok, enc = cv2.imencode(".jpg", src, flags)
        ref = cv2.imdecode(enc, cv2.IMREAD_UNCHANGED)
        got = oracle.jpeg_decode(enc.tobytes())
        assert got.shape == ref.shape and np.array_equal(got, ref), flags


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_jpeg_other_encoder_and_refusals(oracle):
    """streams from a second encoder (Pillow) decode identically; progressive streams are refused, not mis-decoded"""
    from PIL import Image
    img = noise_image(120, 200, seed=6, channels=3)
    for sub in (0, 1, 2):
        for q in (30, 95):
            b = io.BytesIO()
            Image.fromarray(img[:, :, ::-1]).save(b, "JPEG", quality=q, subsampling=sub, optimize=bool(sub))
            ref = cv2.imdecode(np.frombuffer(b.getvalue(), np.uint8), cv2.IMREAD_UNCHANGED)
            assert np.array_equal(oracle.jpeg_decode(b.getvalue()), ref), (sub, q)
    ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    with pytest.raises(ValueError):
        oracle.jpeg_decode(enc.tobytes())
    with pytest.raises(ValueError):
        oracle.jpeg_decode(b"\xff\xd8\xff")
