"""Per-kernel times of the stereo handle's frame (every kernel timed running alone: one frame at a time, CUDA events
around every launch) plus a short throughput figure, device-resident input: the quick A/B for kernel variants.
    python tools/pair_probe.py [frames_for_throughput] [frames_in_flight]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    import torch
    import ergo_uvo_b200 as U
    from tools import synth
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    inflight = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    W, H = 1280, 1024
    seq = synth.StereoSequence(W, H, n_frames=8, seed=1300, tex_size=2048)
    ctx = U.Context(0)
    p = U.default_params(True)
    p.surf_min_hessian = 11032
    p.max_features = 16384
    vo = U.StereoVO(ctx, W, H, U.make_camera(seq.KL, seq.DL, seq.newKL), U.make_camera(seq.KR, seq.DR, seq.newKR),
                    seq.R_right, seq.t_right, p)
    order = [0, 1, 2, 3, 4, 5, 6, 7, 6, 5, 4, 3, 2, 1]
    dev = [(torch.from_numpy(seq.frames[k][0]).cuda(), torch.from_numpy(seq.frames[k][1]).cuda()) for k in order]

    def run(n, inflight):
        q = 0
        for k in range(n):
            L, R = dev[k % len(dev)]
            vo.enqueue_device(L.data_ptr(), R.data_ptr(), L.stride(0), 0.1)
            q += 1
            if q >= inflight:
                vo.collect()
                q -= 1
        while q:
            vo.collect()
            q -= 1
    run(48, inflight)
    rates = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run(n, inflight)
        torch.cuda.synchronize()
        rates.append(n / (time.perf_counter() - t0))
    vo.set_graphs(False)
    ctx.kernel_timing(True)
    run(12, 1)
    rep = ctx.kernel_report()
    print(json.dumps({"frames_per_s": [round(r, 1) for r in rates],
                      "us_per_launch": {k: round(1e3 * ms / c, 1) for k, (c, ms) in rep.items()}}))


if __name__ == "__main__":
    main()
