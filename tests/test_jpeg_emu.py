"""CPU: the thread bodies of the two JPEG kernels (ergo_uvo_b200/csrc/jpeg_kernels.cuh), executed on the host over the
kernels' launch grid by the harness tests/emu/jpeg_emu.cpp, fed with the coefficients of the product's host-side Huffman
decoder, against the oracle -- bit-exact.  This covers the index arithmetic, the integer IDCT / upsampling / colour
pipeline (clear / scatter of the sparse coefficients, IDCT, upsampling, colour) and the __syncwarp() granularity of
k_jpeg_idct; it is not a GPU run (tests/test_gpu_zz_jpeg.py is).
Reference path: the cv::imdecode inside from_ros_to_cv_image, math_utility.cpp:154-173."""
import ctypes as C
import io
import os
import subprocess

import numpy as np
import pytest

from conftest import noise_image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "libjpeg_emu.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-x", "c++",
                           "-I" + os.path.join(ROOT, "ergo_uvo_b200", "csrc"), "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "emu", "jpeg_emu.cpp"), "-o", so])
    return C.CDLL(so)


def _decode(emu, data):
    import ergo_uvo_b200 as U
    lay, entries, first, count = U.jpeg_entropy_decode_sparse(data)
    ch = 1 if lay.components == 1 else 3
    out = np.full((lay.height, lay.width) if ch == 1 else (lay.height, lay.width, 3), 0xCD, np.uint8)
    if len(entries) == 0:
        entries = np.zeros(1, np.uint32)
    rc = emu.emu_jpeg_decode(entries.ctypes.data_as(C.c_void_p), first.ctypes.data_as(C.c_void_p),
                             count.ctypes.data_as(C.c_void_p), C.byref(lay), out.ctypes.data_as(C.c_void_p),
                             C.c_size_t(lay.width * ch))
    assert rc == 0
    return out


def test_kernel_bodies_golden_streams(emu, oracle):
    z = np.load(os.path.join(GOLD, "jpeg_64x48.npz"))
    for n in ("c420_rst2", "c422", "c440", "c444_opt", "gray"):
        data = z[n + "_jpg"].tobytes()
        got = _decode(emu, data)
        assert np.array_equal(got, oracle.jpeg_decode(data)), n
        assert np.array_equal(got, z[n + "_img"]), n  # cv2 / libjpeg-turbo itself


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
@pytest.mark.parametrize("sampling", ["444", "422", "420", "440", "411"])
def test_kernel_bodies_sizes_and_qualities(emu, oracle, sampling):
    sf = getattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR_" + sampling)
    rs = np.random.RandomState(int(sampling))
    for h, w in [(1, 1), (2, 2), (3, 5), (5, 3), (8, 8), (17, 33), (33, 17), (31, 47), (100, 6), (2, 37), (243, 317)]:
        img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8) if h * w < 2000 else noise_image(h, w, seed=h, channels=3)
        for q, rst in ((15, 0), (75, 3), (100, 1)):
            ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sf,
                                                 cv2.IMWRITE_JPEG_RST_INTERVAL, rst])
            assert np.array_equal(_decode(emu, enc.tobytes()), oracle.jpeg_decode(enc.tobytes())), (h, w, q, rst)


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_kernel_bodies_gray_other_encoder_and_full_frame(emu, oracle):
    from PIL import Image
    img = noise_image(243, 317, seed=5, channels=3)
    ok, enc = cv2.imencode(".jpg", img[:, :, 1].copy(), [cv2.IMWRITE_JPEG_QUALITY, 50])
    assert np.array_equal(_decode(emu, enc.tobytes()), oracle.jpeg_decode(enc.tobytes()))
    b = io.BytesIO()
    Image.fromarray(img[:, :, ::-1]).save(b, "JPEG", quality=85, subsampling=2, optimize=True)
    assert np.array_equal(_decode(emu, b.getvalue()), oracle.jpeg_decode(b.getvalue()))
    big = noise_image(1024, 1280, seed=9, channels=3)  # BASELINE config B size: 30 720 blocks, 960 thread blocks
    ok, enc = cv2.imencode(".jpg", big, [cv2.IMWRITE_JPEG_QUALITY, 90])
    got = _decode(emu, enc.tobytes())
    assert np.array_equal(got, oracle.jpeg_decode(enc.tobytes()))
    assert np.array_equal(got, cv2.imdecode(enc, cv2.IMREAD_UNCHANGED))
