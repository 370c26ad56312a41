"""Per-kernel times and bytes of the JPEG ingest (uvo_jpeg_decode) on one synthetic 1280x1024 frame: host entropy
decoding (wall clock, one thread), k_jpeg_idct / k_jpeg_color (CUDA events around each launch, kernel alone), the
bytes that cross PCIe, the kernels' algorithmic GB/s; output checked against cv2.imdecode.  Prints one JSON line.
Usage on the B200 box: python tools/jpeg_probe.py > gpurun_out/jpeg_probe.json"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import ergo_uvo_b200 as U  # noqa: E402
from tools import synth  # noqa: E402


def main():
    import cv2
    seq = synth.StereoSequence(1280, 1024, n_frames=1, tex_size=2048)
    img = seq.frames[0][0]
    ctx = U.Context(0)
    out = {"image": "1280x1024 synthetic frame", "unit": "us per launch (kernel alone, CUDA events)"}
    for q in (75, 90):
        ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q])
        data = enc.tobytes()
        got = ctx.jpeg_decode(data)  # warm-up: allocations
        exact = bool(np.array_equal(got, cv2.imdecode(enc, cv2.IMREAD_UNCHANGED)))
        t0 = time.perf_counter()
        for _ in range(5):
            lay, entries, first, count = U.jpeg_entropy_decode_sparse(data)
        host_ms = (time.perf_counter() - t0) / 5 * 1e3
        ctx.kernel_timing(True)
        t0 = time.perf_counter()
        for _ in range(5):
            ctx.jpeg_decode(data)
        call_ms = (time.perf_counter() - t0) / 5 * 1e3
        rep = ctx.kernel_report()
        ctx.kernel_timing(False)
        k = {name: round(1e3 * ms / cnt, 2) for name, (cnt, ms) in sorted(rep.items())}
        samples = sum(lay.blocks_x[c] * lay.blocks_y[c] * 64 for c in range(lay.components))
        h2d = 4 * len(entries) + 5 * len(first)
        idct_bytes = h2d + samples                       # sparse coefficients in, samples out
        color_bytes = samples + 3 * lay.width * lay.height  # planes in, BGR out
        out[f"q{q}"] = {"stream_bytes": len(data), "bit_exact_vs_cv2_imdecode": exact, "host_entropy_decode_ms": round(host_ms, 2),
                        "whole_call_ms_host_buffers": round(call_ms, 2), "h2d_bytes": h2d, "kernels_us": k,
                        "k_jpeg_idct_GBps": round(idct_bytes / (k.get("k_jpeg_idct", float("nan")) * 1e-6) / 1e9, 1),
                        "k_jpeg_color_GBps": round(color_bytes / (k.get("k_jpeg_color", float("nan")) * 1e-6) / 1e9, 1)}
    ctx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
