"""Repeatability probe for the end-to-end (pinned host images) path: python tools/e2e_probe.py [reps] [frames]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import ergo_uvo_b200 as U
    from tools import synth
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    W, H = 1280, 1024
    seq = synth.StereoSequence(W, H, n_frames=8, seed=1300, tex_size=2048)
    ctx = U.Context(0)
    p = U.default_params(True)
    p.surf_min_hessian = 11032
    vo = U.StereoVO(ctx, W, H, U.make_camera(seq.KL, seq.DL, seq.newKL), U.make_camera(seq.KR, seq.DR, seq.newKR),
                    seq.R_right, seq.t_right, p)
    order = [0, 1, 2, 3, 4, 5, 6, 7, 6, 5, 4, 3, 2, 1]
    RING = 28
    host = [(torch.from_numpy(seq.frames[order[i % len(order)]][0]).pin_memory(),
             torch.from_numpy(seq.frames[order[i % len(order)]][1]).pin_memory()) for i in range(RING)]
    dev = [(a.cuda(), b.cuda()) for a, b in host]

    def mosaic(t):  # BGGR mosaic sampled from the colour image (1 byte per pixel over PCIe instead of 3)
        img = t.numpy()
        m = img[:, :, 1].copy()
        m[0::2, 0::2] = img[0::2, 0::2, 0]
        m[1::2, 1::2] = img[1::2, 1::2, 2]
        return torch.from_numpy(np.ascontiguousarray(m)).pin_memory()
    bayer = [(mosaic(a), mosaic(b)) for a, b in host]

    def run(n, ring, enq, inflight=8):
        q = 0
        t_enq = t_col = 0.0
        for k in range(n):
            L, R = ring[k % RING]
            t0 = time.perf_counter()
            enq(L.data_ptr(), R.data_ptr(), L.stride(0), 0.1)
            t_enq += time.perf_counter() - t0
            q += 1
            if q >= inflight:
                t0 = time.perf_counter()
                vo.collect()
                t_col += time.perf_counter() - t0
                q -= 1
        while q:
            vo.collect()
            q -= 1
        return t_enq, t_col
    run(32, host, vo.enqueue_host)
    run(32, dev, vo.enqueue_device)
    run(32, bayer, vo.enqueue_host_bayer)
    for r in range(reps):
        for name, ring, enq in (("device", dev, vo.enqueue_device), ("host", host, vo.enqueue_host),
                                ("bayer", bayer, vo.enqueue_host_bayer)):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            te, tc = run(n, ring, enq)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print(f"rep {r} {name:6s} {n / dt:8.1f} frames/s   enqueue {1e6 * te / n:6.1f} us/frame   "
                  f"collect wait {1e6 * tc / n:6.1f} us/frame", flush=True)
    vo.close()
    ctx.close()


if __name__ == "__main__":
    main()
