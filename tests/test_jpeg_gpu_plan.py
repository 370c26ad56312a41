"""CPU: the HOST half of the GPU Huffman route (csrc/jpeg.cu build_gpu_plan, reached through the host-only entry point
uvo_jpeg_gpu_plan): the table plan and the unstuffed copy of the scan that k_jpeg_huff works from.  The scan is decoded
here sequentially in Python with NOTHING but the plan -- the two-level code tables, the canonical-code walk for the
prefixes that got no second-level table, the block-of-the-MCU tables and the scan-order -> plane-order block mapping,
i.e. the rules of jh_symbol and of the kernel's last phase (csrc/jpeg_huff.cuh) -- and must give the coefficients of the
host decoder (which tests/test_jpeg_host.py pins to the oracle and cv2).  The decode inside from_ros_to_cv_image,
math_utility.cpp:154-173."""
import os

import numpy as np
import pytest

from conftest import noise_image

cv2 = pytest.importorskip("cv2")

FAST_BITS, SUB_TABLES, MAX_BPM = 10, 16, 10
PLAN = np.dtype([("fast", "<u2", (4, 1 << FAST_BITS)), ("sub", "<u2", (4, SUB_TABLES * 64)), ("maxcode", "<i4", (4, 18)),
                 ("valoff", "<i4", (4, 17)), ("huffval", "u1", (4, 256)), ("bpm", "<i4"), ("mcus_x", "<i4"),
                 ("mcus_y", "<i4"), ("total_blocks", "<i4"), ("components", "<i4"), ("H", "<i4", 3), ("V", "<i4", 3),
                 ("blocks_x", "<i4", 3), ("block_off", "<i4", 3), ("blk_comp", "u1", MAX_BPM + 2),
                 ("blk_v", "u1", MAX_BPM + 2), ("blk_h", "u1", MAX_BPM + 2), ("dc_tab", "u1", 4), ("ac_tab", "u1", 4),
                 ("total_bits", "<u4"), ("entries_cap", "<u4")], align=True)
NATURAL = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21,
           28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61,
           54, 47, 55, 62, 63]


def _decode_with_plan(plan, scan, coeff_total):
    """sequential decode of the unstuffed scan with the plan's tables; dense int16 coefficients in plane order"""
    bits = int(plan["total_bits"])
    stream = int.from_bytes(bytes(scan[: (bits + 7) // 8 + 8]), "big")
    nbits_stream = 8 * ((bits + 7) // 8 + 8)

    def peek32(pos):
        return (stream >> (nbits_stream - pos - 32)) & 0xFFFFFFFF

    out = np.zeros(coeff_total, np.int16)
    bpm, mcus_x = int(plan["bpm"]), int(plan["mcus_x"])
    pred = [0, 0, 0]
    pos, sb = 0, 0
    while sb < int(plan["total_blocks"]):
        k, m = sb % bpm, sb // bpm
        c = int(plan["blk_comp"][k])
        mx, my = m % mcus_x, m // mcus_x
        X = mx * int(plan["H"][c]) + int(plan["blk_h"][k])
        Y = my * int(plan["V"][c]) + int(plan["blk_v"][k])
        blk = int(plan["block_off"][c]) + Y * int(plan["blocks_x"][c]) + X
        z = 0
        while z < 64:
            x = peek32(pos)
            t = int(plan["dc_tab"][c]) if z == 0 else int(plan["ac_tab"][c])
            e = int(plan["fast"][t][x >> (32 - FAST_BITS)])
            if e & 0x8000:
                e = int(plan["sub"][t][((e & 0x7FFF) << 6) | ((x >> (32 - FAST_BITS - 6)) & 63)])
            if e:
                ln, sym = e >> 8, e & 255
            else:
                ln, sym = 16, 0
                for l in range(FAST_BITS + 1, 17):
                    code = x >> (32 - l)
                    if code <= int(plan["maxcode"][t][l]):
                        ln, sym = l, int(plan["huffval"][t][(code + int(plan["valoff"][t][l])) & 255])
                        break
            mag = min(sym, 15) if z == 0 else sym & 15
            assert pos + ln + mag <= bits, "symbol runs past the end of the scan"
            v = 0
            if mag:
                v = ((x << ln) & 0xFFFFFFFF) >> (32 - mag)
                if v < (1 << (mag - 1)):
                    v += -(1 << mag) + 1
            pos += ln + mag
            if z == 0:
                pred[c] += v
                out[blk * 64] = np.int16(pred[c])
                z = 1
            else:
                run = sym >> 4
                if mag == 0:
                    z = z + 16 if run == 15 else 64
                else:
                    kk = z + run
                    assert kk <= 63
                    out[blk * 64 + NATURAL[kk]] = np.int16(v)
                    z = kk + 1
        sb += 1
    assert bits - pos < 8, "more than padding left after the last block"
    return out


def _streams():
    rgb = noise_image(45, 61, seed=3, channels=3)
    for sf in ("444", "420", "422", "411", "440"):
        for q, opt in ((35, 0), (90, 1)):
            ok, enc = cv2.imencode(".jpg", rgb, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_OPTIMIZE, opt,
                                                 cv2.IMWRITE_JPEG_SAMPLING_FACTOR,
                                                 getattr(cv2, "IMWRITE_JPEG_SAMPLING_FACTOR_" + sf)])
            assert ok
            yield f"{sf} q{q} opt{opt}", enc.tobytes()
    ok, enc = cv2.imencode(".jpg", np.ascontiguousarray(rgb[:40, :56, 1]), [cv2.IMWRITE_JPEG_QUALITY, 80])
    yield "gray 56x40", enc.tobytes()
    ok, enc = cv2.imencode(".jpg", rgb[:8, :8], [cv2.IMWRITE_JPEG_QUALITY, 95])
    yield "one MCU", enc.tobytes()


@pytest.mark.parametrize("name,data", list(_streams()), ids=lambda v: v if isinstance(v, str) else "")
def test_plan_decodes_to_the_host_decoders_coefficients(name, data):
    import ergo_uvo_b200 as U
    from ergo_uvo_b200 import vo_utility as V
    got = V.jpeg_gpu_plan(data)
    assert got is not None, "a single-scan stream without restart intervals takes the GPU route"
    lay, bits, upload, staging = got
    plan = np.frombuffer(staging[: PLAN.itemsize].tobytes(), PLAN)[0]
    off = (PLAN.itemsize + 255) & ~255
    assert int(plan["total_bits"]) == bits and upload >= off + (bits + 7) // 8 + 16 and upload <= len(staging)
    assert not staging[off + (bits + 7) // 8: upload].any()          # the padding the bit reader may run into is zero
    lay2, want = U.jpeg_entropy_decode(data)
    assert lay.coeff_total == lay2.coeff_total and int(plan["total_blocks"]) * 64 == lay.coeff_total
    dense = _decode_with_plan(plan, staging[off:], int(lay.coeff_total))
    assert np.array_equal(dense, want)


def test_streams_the_gpu_route_does_not_take():
    from ergo_uvo_b200 import vo_utility as V
    import ergo_uvo_b200 as U
    rgb = noise_image(45, 61, seed=3, channels=3)
    ok, enc = cv2.imencode(".jpg", rgb, [cv2.IMWRITE_JPEG_QUALITY, 70, cv2.IMWRITE_JPEG_RST_INTERVAL, 2])
    assert V.jpeg_gpu_plan(enc.tobytes()) is None          # restart intervals: the host decoder
    with pytest.raises(U.UvoError):
        V.jpeg_gpu_plan(b"\xff\xd8 not a jpeg")
    ok, enc = cv2.imencode(".jpg", rgb, [cv2.IMWRITE_JPEG_QUALITY, 70])
    with pytest.raises(U.UvoError):
        V.jpeg_gpu_plan(enc.tobytes()[:200])                # cut inside the headers
    data = bytearray(enc.tobytes())
    p = data.index(b"\xff\xda")
    data[p + 7] = data[p + 5]                               # the scan names its first component twice
    with pytest.raises(U.UvoError):
        V.jpeg_gpu_plan(bytes(data))
