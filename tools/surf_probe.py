"""Event-timed image-preparation and SURF kernels of one 1280x1024 stereo pair's left image at the bench threshold (A/B of kernel variants:
rebuild with `make -C ergo_uvo_b200/csrc EXTRA=-D...`, run this).  Prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ergo_uvo_b200 as U  # noqa: E402
from tools import synth  # noqa: E402


def main():
    seq = synth.StereoSequence(1280, 1024, n_frames=1, seed=1300, tex_size=2048)
    ctx = U.Context(0)
    ctx.params.max_features = 16384
    ctx.params.surf_min_hessian = int(sys.argv[1]) if len(sys.argv) > 1 else 11032
    g = ctx.get_image(seq.frames[0][0], seq.KL, seq.DL, seq.newKL)
    k, d = ctx.detect_features(g)
    for _ in range(3):
        ctx.detect_features(g)
    ctx.kernel_timing(True)
    for _ in range(20):
        ctx.get_image(seq.frames[0][0], seq.KL, seq.DL, seq.newKL)
        ctx.detect_features(g)
    rep = ctx.kernel_report()
    print(json.dumps({"keypoints": len(k), "us_per_launch": {n: round(1e3 * ms / c, 2) for n, (c, ms) in sorted(rep.items())}}))


if __name__ == "__main__":
    main()
