"""Per-kernel times (CUDA events around every launch, uvo_ctx_kernel_timing) of the stage-level SURF and matcher calls
with 64-float and extended 128-float descriptor rows, on one synthetic 1280x1024 frame (~4k keypoints).  Prints one
JSON line.  Usage on the B200 box: python tools/ext_probe.py > gpurun_out/ext_probe.json"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import ergo_uvo_b200 as U  # noqa: E402
from tools import synth  # noqa: E402


def timed(ctx, fn, reps=5):
    fn()  # warm-up (allocations, tensor maps, function attributes)
    ctx.kernel_timing(True)
    for _ in range(reps):
        fn()
    rep = ctx.kernel_report()
    ctx.kernel_timing(False)
    return {k: round(1e3 * ms / cnt, 2) for k, (cnt, ms) in sorted(rep.items())}


def main():
    seq = synth.StereoSequence(1280, 1024, n_frames=1, tex_size=2048)
    ctx = U.Context(0)
    ctx.params.max_features = 16384
    ctx.params.surf_min_hessian = 11032  # bench.py's frozen threshold for ~4k keypoints on this texture
    L, R = seq.frames[0]
    gL = ctx.get_image(L, seq.KL, seq.DL, seq.newKL)
    gR = ctx.get_image(R, seq.KR, seq.DR, seq.newKR)
    out = {"image": "1280x1024", "unit": "us per launch (kernel alone, CUDA events)"}
    for ext in (0, 1):
        ctx.params.surf_extended = ext
        kL, dL = ctx.detect_features(gL)
        kR, dR = ctx.detect_features(gR)
        tag = "dim128" if ext else "dim64"
        out[tag] = {"keypoints": [len(kL), len(kR)],
                    "detect_features": timed(ctx, lambda: ctx.detect_features(gL)),
                    "match_features": timed(ctx, lambda: ctx.match_features(None, None, dL, dR)),
                    "fallbacks": ctx.match_last_fallbacks()}
        ctx.match_exact_only(True)
        out[tag]["match_features_exact_only"] = timed(ctx, lambda: ctx.match_features(None, None, dL, dR), reps=2)
        ctx.match_exact_only(False)
    ctx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
