"""Per-stage microbenchmarks for the other BASELINE.json configurations, GPU (through the C ABI, HOST buffers, copies
inside the timed region) beside the CPU (cv2 where the wheel has the function, the oracle port for SURF):

  A  mono UVO 640x480 synthetic seabed sequence + synthetic range, shipped mono YAML (SURF + BF match + essential /
     homography LMedS): uvo_mono frames/s beside the CPU replay (oracle port)
  C  SURF extract + BF kNN match: 1920x1080 frames, ~8k features, ratio 0.7
  D  batched RANSAC sweep: 4096 hypotheses x 10 000 correspondences for 5-point essential, homography and PnP
     (confidence 1 - 2^-53 and 75 % outliers keep every hypothesis alive)
  E  one stereo sequence at 2448x2048 on one GPU (the per-GPU unit of the 8-GPU configuration), frames/s

    python tools/microbench.py [--out profiles/microbench_rNN.json]

Wall clock (perf_counter) around synchronous calls, 2 warm-ups, median of `--reps`.  Not the headline metric
(bench.py is); this is the "per-stage ms vs CPU OpenCV" evidence for configs C and D.
"""
import argparse
import json
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def med(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return 1e3 * statistics.median(ts)


def config_a(ctx, n_frames=24):
    import ergo_uvo_b200 as U
    from oracle import oracle as O
    from oracle.ref_mono import RefMonoVO
    from tools import synth
    seq = synth.MonoSequence(640, 480, n_frames=n_frames, tex_size=2048, velocity=(0.05, 0.01, 0.005))
    p = U.default_params(False)
    cam = U.make_camera(seq.K, seq.D, seq.newK)
    out = {"width": 640, "height": 480, "frames": n_frames, "params": "mono_VO_parameters.yaml (SURF 50, LMedS)"}
    for rep in range(2):  # second pass is the measured one (first warms allocations)
        vo = U.MonoVO(ctx, 640, 480, cam, p)
        t0 = time.perf_counter()
        res = [vo.frame(seq.frames[k], 0.1, seq.ranges[k]) for k in range(n_frames)]
        dt = time.perf_counter() - t0
        vo.close()
    out["gpu_frames_per_s"] = n_frames / dt
    out["keypoints_per_frame"] = float(np.mean([r.n_keypoints for r in res]))
    out["published"] = int(sum(r.published for r in res))
    out["used_essential"] = int(sum(r.used_essential for r in res if r.published))
    n_cpu = min(n_frames, 6)
    ref = RefMonoVO(O, seq, p)
    t0 = time.perf_counter()
    for k in range(n_cpu):
        ref.frame(seq.frames[k], 0.1, seq.ranges[k])
    out["cpu_frames_per_s(oracle port)"] = n_cpu / (time.perf_counter() - t0)
    out["cpu_sample"] = f"first {n_cpu} frames"
    return out


def config_e(ctx, n_frames=60):
    import ergo_uvo_b200 as U
    from tools import synth
    W, H = 2448, 2048
    seq = synth.StereoSequence(W, H, n_frames=6, seed=1300, tex_size=4096)
    p = U.default_params(True)
    # threshold for ~4k keypoints per image, as in config B
    g = ctx.get_image(seq.frames[0][0], seq.KL, seq.DL, seq.newKL)
    lo, hi, thr = 100, 400000, None
    ctx.params.max_features = 1 << 15
    while lo < hi:
        thr = (lo + hi) // 2
        ctx.params.surf_min_hessian = thr
        n = len(ctx.detect_features(g)[0])
        if abs(n - 4096) <= 0.03 * 4096:
            break
        lo, hi = (thr + 1, hi) if n > 4096 else (lo, thr)
    ctx.params.max_features = 16384
    p.surf_min_hessian = thr
    p.max_features = 16384
    vo = U.StereoVO(ctx, W, H, U.make_camera(seq.KL, seq.DL, seq.newKL), U.make_camera(seq.KR, seq.DR, seq.newKR),
                    seq.R_right, seq.t_right, p)
    import torch
    order = [0, 1, 2, 3, 4, 5, 4, 3, 2, 1]
    host = [(torch.from_numpy(seq.frames[i][0]).pin_memory(), torch.from_numpy(seq.frames[i][1]).pin_memory())
            for i in range(6)]

    def run(n):
        q, valid = 0, 0
        for k in range(n):
            L, R = host[order[k % len(order)]]
            vo.enqueue_host(L.data_ptr(), R.data_ptr(), 3 * W, 0.1)
            q += 1
            if q >= 8:
                valid += vo.collect().valid
                q -= 1
        while q:
            valid += vo.collect().valid
            q -= 1
        return valid
    run(16)
    t0 = time.perf_counter()
    valid = run(n_frames)
    dt = time.perf_counter() - t0
    vo.close()
    return {"width": W, "height": H, "surf_min_hessian": thr, "frames": n_frames, "valid": int(valid),
            "e2e_frames_per_s": n_frames / dt, "h2d_bytes_per_frame": 2 * 3 * W * H,
            "api": "uvo_stereo_enqueue_host + uvo_stereo_collect, 8 frames in flight, pinned host images"}


def config_c(ctx, reps):
    from oracle import oracle as O
    from tools import synth
    try:
        import cv2
    except ImportError:
        cv2 = None
    seq = synth.StereoSequence(1920, 1080, n_frames=2, tex_size=2048)
    (L0, _), (L1, _) = seq.frames[0], seq.frames[1]
    g0 = ctx.get_image(L0, seq.KL, seq.DL, seq.newKL)
    g1 = ctx.get_image(L1, seq.KL, seq.DL, seq.newKL)
    lo, hi, thr = 100, 400000, None
    ctx.params.max_features = 1 << 15
    while lo < hi:  # threshold for ~8192 keypoints, bisected on the GPU path
        thr = (lo + hi) // 2
        ctx.params.surf_min_hessian = thr
        n = len(ctx.detect_features(g0)[0])
        if abs(n - 8192) <= 0.02 * 8192:
            break
        lo, hi = (thr + 1, hi) if n > 8192 else (lo, thr)
    ctx.params.lowe_ratio = 0.7
    k0, d0 = ctx.detect_features(g0)
    k1, d1 = ctx.detect_features(g1)
    out = {"width": 1920, "height": 1080, "surf_min_hessian": thr, "keypoints": [len(k0), len(k1)], "ratio": 0.7}
    out["gpu_ms"] = {
        "get_image": med(lambda: ctx.get_image(L0, seq.KL, seq.DL, seq.newKL), reps),
        "surf_detect_and_compute": med(lambda: ctx.detect_features(g0), reps),
        "knn_match_ratio": med(lambda: ctx.match_features(None, None, d0, d1), reps),
    }
    out["matches"] = int(len(ctx.match_features(None, None, d0, d1)))
    cpu = {"surf_detect_and_compute(oracle port, all cores)": med(lambda: O.surf_detect_and_compute(g0, thr), 3, 1),
           "get_image(oracle port)": med(lambda: O.get_image(L0, seq.KL, seq.DL, seq.newKL, True, 8.0), 3, 1)}
    if cv2 is not None:
        bf = cv2.BFMatcher(cv2.NORM_L2, False)
        cpu[f"knn_match_ratio(cv2 {cv2.__version__}, {cv2.getNumThreads()} threads)"] = med(
            lambda: bf.knnMatch(d0, d1, 2), 3, 1)
    else:
        cpu["knn_match_ratio(oracle port)"] = med(lambda: O.match_features(d0, d1, np.float32(0.7)), 3, 1)
    out["cpu_ms"] = cpu
    ctx.params.max_features = 16384
    ctx.params.lowe_ratio = 0.8
    return out


def config_d(ctx, reps):
    from tools.make_golden_twoview import scene
    try:
        import cv2
    except ImportError:
        cv2 = None
    conf = 1.0 - 2.0 ** -53
    K4 = np.array([1300.0, 1300.0, 640.0, 512.0])
    KM = np.array([[1300.0, 0, 640], [0, 1300.0, 512], [0, 0, 1]])
    out = {"hypotheses": 4096, "correspondences": 10000, "outlier_fraction": 0.75}
    gpu, cpu, hyp = {}, {}, {}
    p1, p2, _ = scene(10000, 11, True, 0.75, 0.3)
    hyp["homography"] = ctx.findHomography(p1, p2, 8, 3.0, 4096, conf)[2]
    gpu["findHomography"] = med(lambda: ctx.findHomography(p1, p2, 8, 3.0, 4096, conf), reps)
    if cv2 is not None:
        cpu["findHomography"] = med(lambda: cv2.findHomography(p1, p2, cv2.RANSAC, 3.0, None, 4096, conf), 3, 1)
    p1, p2, _ = scene(10000, 11, False, 0.75, 0.3)
    hyp["essential"] = ctx.findEssentialMat(p1, p2, K4, 8, conf, 1.0, 4096)[2]
    gpu["findEssentialMat"] = med(lambda: ctx.findEssentialMat(p1, p2, K4, 8, conf, 1.0, 4096), reps)
    if cv2 is not None:
        cpu["findEssentialMat"] = med(lambda: cv2.findEssentialMat(p1, p2, KM, cv2.RANSAC, conf, 1.0, 4096), 2, 1)
    # PnP: 3-D points in front of the camera, 75 % of the image points replaced by uniform outliers
    rs = np.random.RandomState(11)
    X = np.stack([rs.uniform(-4, 4, 10000), rs.uniform(-3, 3, 10000), rs.uniform(4, 9, 10000)], -1)
    rvec, tvec = np.array([0.01, -0.02, 0.015]), np.array([0.3, 0.05, 0.1])
    th = np.linalg.norm(rvec)
    k = rvec / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
    Xc = X @ R.T + tvec
    x = (Xc[:, :2] / Xc[:, 2:]) * [1300.0, 1300.0] + [640.0, 512.0] + rs.normal(0, 0.3, (10000, 2))
    bad = rs.rand(10000) < 0.75
    x[bad] = np.stack([rs.uniform(0, 1280, bad.sum()), rs.uniform(0, 1024, bad.sum())], -1)
    x = x.astype(np.float32)
    hyp["pnp"] = ctx.solvePnPRansac(X, x, KM, 4096, 1.0, conf)[4]
    gpu["solvePnPRansac(EPNP)"] = med(lambda: ctx.solvePnPRansac(X, x, KM, 4096, 1.0, conf), reps)
    if cv2 is not None:
        cpu["solvePnPRansac(EPNP)"] = med(lambda: cv2.solvePnPRansac(X, x, KM, None, None, None, False, 4096, 1.0, conf,
                                                                     None, cv2.SOLVEPNP_EPNP), 2, 1)
        out["cpu"] = f"cv2 {cv2.__version__}, {cv2.getNumThreads()} threads"
    out.update(gpu_ms=gpu, cpu_ms=cpu, hypotheses_evaluated=hyp)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    import ergo_uvo_b200 as U
    ctx = U.Context(0)
    res = {"host_cpus": os.cpu_count(),
           "timing": "wall clock around synchronous C-ABI calls with host buffers (H2D/D2H inside); median"}
    if args.only in ("", "A"):
        res["config_A"] = config_a(ctx)
    if args.only in ("", "E"):
        res["config_E"] = config_e(ctx)
    if args.only in ("", "C"):
        res["config_C"] = config_c(ctx, args.reps)
    if args.only in ("", "D"):
        res["config_D"] = config_d(ctx, args.reps)
    ctx.close()
    txt = json.dumps(res, indent=1)
    print(txt)
    if args.out:
        open(args.out, "w").write(txt + "\n")


if __name__ == "__main__":
    main()
