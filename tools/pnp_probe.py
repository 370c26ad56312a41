"""GPU probe of the solvePnPRansac stage (visual_odometry.h:647-648): per-kernel CUDA-event times and the SM clock
stamps of the phases of one hypothesis and of the refit, on a scene shaped like the stereo benchmark's (about 3500
correspondences, a few percent outliers, iterationsCount 1000).  Prints one JSON object.
    python tools/pnp_probe.py [n] [outlier_fraction]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ergo_uvo_b200 as U  # noqa: E402


def scene(n, frac, seed=7):
    rs = np.random.RandomState(seed)
    K = np.array([[1300.0, 0, 640], [0, 1300.0, 512], [0, 0, 1]])
    X = np.stack([rs.uniform(-4, 4, n), rs.uniform(-3, 3, n), rs.uniform(4, 9, n)], -1)
    rvec, tvec = np.array([0.01, -0.02, 0.015]), np.array([0.3, 0.05, 0.1])
    th = np.linalg.norm(rvec)
    k = rvec / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
    Xc = X @ R.T + tvec
    x = (Xc[:, :2] / Xc[:, 2:]) * 1300.0 + [640, 512] + rs.randn(n, 2) * 0.3
    out = rs.rand(n) < frac
    x[out] = np.stack([rs.uniform(0, 1280, out.sum()), rs.uniform(0, 1024, out.sum())], -1)
    return K, X, x.astype(np.float32)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3500
    frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.03
    K, X, x = scene(n, frac)
    ctx = U.Context(0)
    for _ in range(3):
        ctx.solvePnPRansac(X, x, K, 1000, 1.0, 0.99)
    ctx.pnp_profile(True)
    ctx.kernel_timing(True)
    reps = 20
    for _ in range(reps):
        ok, rv, tv, inl, hyps = ctx.solvePnPRansac(X, x, K, 1000, 1.0, 0.99)
    rep = ctx.kernel_report()
    st = ctx.pnp_profile(False)
    ctx.kernel_timing(False)
    names_h = ["read", "control_points", "MtM", "eigh", "betas(lane2)", "pose(lane2)", "model", "score", "bookkeeping"]
    names_f = ["inlier_list", "control_points", "MtM", "eigh", "betas", "poses", "reproj+rodrigues"]
    clk = 1.965e3  # cycles per us at the nominal SM clock
    out = {"n": n, "outlier_fraction": frac, "hyps_evaluated": int(hyps), "inliers": int(len(inl)), "ok": bool(ok),
           "kernels_us": {k: round(1e3 * v[1] / v[0] * (v[0] / reps) , 2) for k, v in rep.items()},
           "kernel_launches_per_call": {k: v[0] / reps for k, v in rep.items()},
           "hypothesis0_phase_us": {nm: round((st[i + 1] - st[i]) / clk, 2) for i, nm in enumerate(names_h)
                                     if st[i + 1] and st[i]},
           "refit_phase_us": {nm: round((st[17 + i] - st[16 + i]) / clk, 2) for i, nm in enumerate(names_f)
                              if st[17 + i] and st[16 + i]}}
    if st[13]:
        out["eigh_breakdown_cycles_per_round"] = {"rotation": round(st[10] / st[13]), "apply": round(st[11] / st[13]),
                                                  "write": round(st[12] / st[13]), "rounds": int(st[13])}
    if st[24]:
        out["betas_breakdown_us"] = {"L_rho_init": round((st[24] - st[4]) / clk, 2), "sweeps": round((st[25] - st[24]) / clk, 2),
                                     "finish+solve": round((st[26] - st[25]) / clk, 2),
                                     "gauss_newton": round((st[5] - st[26]) / clk, 2)}
    out["stage_us_sum"] = round(sum(out["kernels_us"].values()), 2)
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
