"""CPU: the host half of the JPEG decode in the product library (csrc/jpeg.cu: uvo_jpeg_info and
uvo_jpeg_entropy_decode, host-only entry points like the camera set-up) against the oracle's coefficients -- exact.
This is the decode inside from_ros_to_cv_image (math_utility.cpp:154-173).  The GPU half (IDCT, upsampling, colour)
is compared with the oracle in tests/test_gpu_zz_jpeg.py."""
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


def _same(U, oracle, data):
    lay, coef = U.jpeg_entropy_decode(data)
    want = oracle.jpeg_coefficients(data)
    assert lay.coeff_total == len(want) == len(coef)
    assert np.array_equal(coef, want)
    # the sparse form (what travels to the GPU) expands to the same planes: non-zero coefficients only, per block
    _, entries, first, count = U.jpeg_entropy_decode_sparse(data)
    assert len(entries) == int(np.count_nonzero(want)) == int(count.sum())
    dense = np.zeros(len(want), np.int16)
    blk = np.repeat(np.arange(len(count)), count)                       # owning block of every entry, block by block
    idx = np.concatenate([np.arange(f, f + c) for f, c in zip(first, count)]) if len(entries) else np.zeros(0, int)
    e = entries[idx.astype(np.int64)]
    dense[blk * 64 + (e >> 16).astype(np.int64)] = (e & 0xFFFF).astype(np.uint16).view(np.int16)
    assert np.array_equal(dense, want)
    return lay


def test_entropy_decode_golden_streams(oracle):
    import ergo_uvo_b200 as U
    z = np.load(os.path.join(GOLD, "jpeg_64x48.npz"))
    for n in ("c420_rst2", "c422", "c440", "c444_opt", "gray"):
        data = z[n + "_jpg"].tobytes()
        lay = _same(U, oracle, data)
        img = z[n + "_img"]
        assert (lay.height, lay.width) == img.shape[:2] and lay.components == (1 if img.ndim == 2 else 3)
    lay = U.jpeg_info(z["c420_rst2_jpg"].tobytes())
    assert list(lay.h_samp) == [2, 1, 1] and list(lay.v_samp) == [2, 1, 1]
    assert list(lay.blocks_x) == [8, 4, 4] and list(lay.blocks_y) == [6, 3, 3]          # 61 x 45: 4 x 3 MCUs of 16 x 16
    assert list(lay.samples_x) == [61, 31, 31] and list(lay.samples_y) == [45, 23, 23]
    assert list(lay.coeff_offset) == [0, 48 * 64, 60 * 64] and lay.coeff_total == 72 * 64
    assert lay.quant[0][0] > 0 and lay.quant[1][63] > 0


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_entropy_decode_live_streams(oracle):
    import ergo_uvo_b200 as U
    from PIL import Image
    rs = np.random.RandomState(3)
    S = cv2.IMWRITE_JPEG_SAMPLING_FACTOR
    n = 0
    for h, w in [(1, 1), (3, 5), (17, 33), (31, 47), (100, 6), (243, 317)]:
        img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8) if h < 40 else noise_image(h, w, seed=h, channels=3)
        for sf in ("444", "422", "420", "440", "411"):
            for q, rst in ((15, 0), (75, 3), (100, 1)):
                ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q, S,
                                                     getattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR_" + sf),
                                                     cv2.IMWRITE_JPEG_RST_INTERVAL, rst])
                _same(U, oracle, enc.tobytes())
                n += 1
        ok, enc = cv2.imencode(".jpg", img[:, :, 0].copy(), [cv2.IMWRITE_JPEG_QUALITY, 60, cv2.IMWRITE_JPEG_OPTIMIZE, 1])
        _same(U, oracle, enc.tobytes())
    big = noise_image(1024, 1280, seed=9, channels=3)
    ok, enc = cv2.imencode(".jpg", big, [cv2.IMWRITE_JPEG_QUALITY, 90])
    lay = _same(U, oracle, enc.tobytes())
    assert (lay.width, lay.height) == (1280, 1024)
    b = io.BytesIO()
    Image.fromarray(big[:200, :300, ::-1]).save(b, "JPEG", quality=85, subsampling=2, optimize=True)
    _same(U, oracle, b.getvalue())
    assert n == 90


def test_entropy_decode_refusals():
    import ergo_uvo_b200 as U
    with pytest.raises(U.UvoError) as e:
        U.jpeg_info(b"not a jpeg at all")
    assert e.value.code == -3
    z = np.load(os.path.join(GOLD, "jpeg_64x48.npz"))
    data = z["c444_opt_jpg"].tobytes()
    with pytest.raises(U.UvoError) as e:
        U.jpeg_info(data.replace(b"\xff\xc0", b"\xff\xc2", 1))  # a progressive frame header
    assert e.value.code == -5
    with pytest.raises(U.UvoError):
        U.jpeg_info(data[:100])  # cut inside the tables: no frame header


def test_scan_header_naming_a_component_twice_is_refused():
    """libjpeg-turbo rejects a scan whose header names a component twice (JERR_BAD_COMPONENT_ID; cv::imdecode returns
    an empty Mat); here the GPU route would leave the other component's block table unwritten, so both routes refuse"""
    import ergo_uvo_b200 as U
    z = np.load(os.path.join(GOLD, "jpeg_64x48.npz"))
    data = bytearray(z["c444_opt_jpg"].tobytes())
    p = data.index(b"\xff\xda")
    assert data[p + 4] == 3                      # Ns: three components in the scan
    ids = [data[p + 5 + 2 * j] for j in range(3)]
    assert len(set(ids)) == 3
    data[p + 5 + 2] = ids[0]                     # the second entry now repeats the first component
    with pytest.raises(U.UvoError) as e:
        U.jpeg_entropy_decode(bytes(data))
    assert e.value.code == -3
    if cv2 is not None:
        assert cv2.imdecode(np.frombuffer(bytes(data), np.uint8), cv2.IMREAD_UNCHANGED) is None


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_parser_survives_mutated_streams():
    """the streams come off the network (ROS CompressedImage): byte edits, truncations and insertions must end in a
    status code, never in a crash or an out-of-bounds table (tools/jpeg_fuzz.py runs the long version, also under
    AddressSanitizer)"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "jpeg_fuzz.py"), "7", "2500"], capture_output=True,
                       text=True)
    assert r.returncode == 0, (r.returncode, r.stderr[-2000:])
    assert "('sparse', 0)" in r.stdout and "('info', -3)" in r.stdout  # both outcomes occurred


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_parallel_subsequence_scheme_reaches_the_sequential_decode():
    """design check for a GPU entropy stage: decoders started at arbitrary bit offsets lock onto the true
    decode, and the round-based sub-sequence scheme has the sequential decode as its fixed point"""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "jpeg_sync_probe.py"), "--size", "320", "256",
                        "--starts", "40", "--subsequence", "1024", "4096"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["not_locked_within_3000_symbols"] == 0 and d["lock_in_bits"]["max"] < 20000
    for S in ("1024", "4096"):
        assert d["parallel_scheme"][S]["equals_sequential_decode"] is True
        assert d["parallel_scheme"][S]["rounds_to_fixed_point"] <= 16
