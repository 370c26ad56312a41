"""Profiling driver for the K8 matcher (run under ncu on the GPU box): real SURF descriptors of one synthetic
1280x1024 stereo pair through uvo_knn_match2, a few repetitions; prints the exact-scan fallback counts."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import ergo_uvo_b200 as U  # noqa: E402
from tools import synth  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    seq = synth.StereoSequence(1280, 1024, n_frames=2)
    ctx = U.Context(0)
    ctx.params.surf_min_hessian = 11032
    descs = []
    for k in range(2):
        L, R = seq.frames[k]
        for img, K, D, nK in ((L, seq.KL, seq.DL, seq.newKL), (R, seq.KR, seq.DR, seq.newKR)):
            g = ctx.get_image(img, K, D, nK)
            _, d = ctx.detect_features(g)
            descs.append(d)
    print("descriptor sets:", [len(d) for d in descs])
    for name, q, t in (("stereo", descs[0], descs[1]), ("temporal", descs[0], descs[2])):
        for _ in range(reps):
            t0 = time.perf_counter()
            ctx.knn_match2(q, t)
            dt = time.perf_counter() - t0
        print(f"{name}: nq={len(q)} nt={len(t)} fallbacks={ctx.match_last_fallbacks()} host_call_ms={dt * 1e3:.3f}")
    rs = np.random.RandomState(0)
    t = np.abs(rs.randn(8192, 64)).astype(np.float32)
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    q = np.abs(rs.randn(8192, 64)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    ctx.knn_match2(q, t)
    print(f"random 8192x8192: fallbacks={ctx.match_last_fallbacks()}")
    ctx.close()


if __name__ == "__main__":
    main()
