"""Kernel list and wall time per frame of the mono path (uvo_mono, BASELINE config A): python tools/prof_mono.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import ergo_uvo_b200 as U
    from tools import synth
    n = 12
    seq = synth.MonoSequence(640, 480, n_frames=n, tex_size=2048, velocity=(0.05, 0.01, 0.005))
    ctx = U.Context(0)
    p = U.default_params(False)
    cam = U.make_camera(seq.K, seq.D, seq.newK)
    for rep in range(2):
        vo = U.MonoVO(ctx, 640, 480, cam, p)
        if rep == 1:
            ctx.kernel_timing(True)
        ts = []
        for k in range(n):
            t0 = time.perf_counter()
            r = vo.frame(seq.frames[k], 0.1, seq.ranges[k])
            ts.append(time.perf_counter() - t0)
        vo.close()
    rep = ctx.kernel_report()
    ctx.kernel_timing(False)
    print("wall ms per frame (with per-kernel events):", [round(1e3 * t, 2) for t in ts])
    tot = 0.0
    for k, (c, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:28s} launches/frame {c / n:6.1f}  ms/frame {ms / n:8.4f}  us/launch {1e3 * ms / c:8.1f}")
        tot += ms / n
    print("sum of kernel ms per frame", tot, " keypoints", r.n_keypoints, "matches", r.n_matches)


if __name__ == "__main__":
    main()
