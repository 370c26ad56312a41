"""Generates tests/golden/twoview.npz: synthetic two-view correspondences + the outputs of cv2 (4.13 in the build
container) for findHomography / findEssentialMat / recoverPose / decomposeHomographyMat, RANSAC and LMedS, with the
parameters of the shipped mono YAML (mono_VO_parameters.yaml:18-26) and of BASELINE config D.  The oracle
(oracle/twoview.py) and the CUDA path are both checked against these vectors.  Run: python tools/make_golden_twoview.py"""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def scene(n, seed, planar, outl, noise, K4=(1300., 1300., 640., 512.)):
    """points U([-4,4]x[-3,3]x[4,9]) (planar: z = 6), pose rvec (0.01,-0.02,0.015), t (0.3,0.05,0.1), pixel noise
    N(0, noise), a fraction `outl` of the second view replaced by uniform outliers"""
    rs = np.random.RandomState(seed)
    X = np.c_[rs.uniform(-4, 4, n), rs.uniform(-3, 3, n), np.full(n, 6.0) if planar else rs.uniform(4, 9, n)]
    R, _ = cv2.Rodrigues(np.array([0.01, -0.02, 0.015]))
    t = np.array([0.3, 0.05, 0.1])
    fx, fy, cx, cy = K4

    def proj(P):
        return np.c_[P[:, 0] / P[:, 2] * fx + cx, P[:, 1] / P[:, 2] * fy + cy]

    p1 = proj(X) + noise * rs.randn(n, 2)
    p2 = proj(X @ R.T + t) + noise * rs.randn(n, 2)
    no = int(outl * n)
    idx = rs.permutation(n)[:no]
    p2[idx] = np.c_[rs.uniform(0, 1280, no), rs.uniform(0, 1024, no)]
    return p1.astype(np.float32), p2.astype(np.float32), np.array(K4)


CASES = [  # name, n, seed, outlier fraction, noise
    ("a", 60, 3, 0.2, 0.3), ("b", 500, 1, 0.3, 1.0), ("c", 1200, 2, 0.5, 1.0), ("d", 800, 4, 0.7, 0.5)]


def main():
    out = {}
    for name, n, seed, outl, noise in CASES:
        for planar in (True, False):
            tag = f"{name}_{'h' if planar else 'e'}"
            p1, p2, K4 = scene(n, seed, planar, outl, noise)
            Km = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.]])
            out[tag + "_p1"], out[tag + "_p2"] = p1, p2
            if planar:
                for meth, mn in ((cv2.RANSAC, "ransac"), (cv2.LMEDS, "lmeds")):
                    # shipped mono config: threshold 1.0 (px), maxIters 2000, confidence 0.99 (mono yaml :23-26)
                    H, m = cv2.findHomography(p1, p2, meth, 1.0, maxIters=2000, confidence=0.99)
                    out[f"{tag}_{mn}_H"] = H if H is not None else np.zeros((3, 3))
                    out[f"{tag}_{mn}_mask"] = m.ravel().astype(np.uint8)
                    if H is not None and m.sum() >= 4:
                        ns, Rs, ts, ns_ = cv2.decomposeHomographyMat(H, Km)
                        out[f"{tag}_{mn}_dec_R"] = np.array(Rs)
                        out[f"{tag}_{mn}_dec_t"] = np.array(ts).reshape(-1, 3)
                        out[f"{tag}_{mn}_dec_n"] = np.array(ns_).reshape(-1, 3)
            else:
                for meth, mn, thr, conf, mi in ((cv2.RANSAC, "ransac", 1.0, 0.999, 1000),
                                                (cv2.LMEDS, "lmeds", 0.1, 0.99, 2000)):  # mono yaml :18-21
                    E, m = cv2.findEssentialMat(p1, p2, Km, meth, conf, thr, mi)
                    E = E[:3]
                    out[f"{tag}_{mn}_E"] = E
                    out[f"{tag}_{mn}_mask"] = m.ravel().astype(np.uint8)
                    good, R, t, m2 = cv2.recoverPose(E, p1, p2, Km, mask=m.copy())
                    out[f"{tag}_{mn}_rp_good"] = np.array(good)
                    out[f"{tag}_{mn}_rp_R"], out[f"{tag}_{mn}_rp_t"] = R, t.ravel()
                    out[f"{tag}_{mn}_rp_mask"] = (m2.ravel() > 0).astype(np.uint8)
    out["K4"] = np.array([1300., 1300., 640., 512.])
    out["cv2_version"] = np.array(cv2.__version__)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "twoview.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
