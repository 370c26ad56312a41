"""Generates tests/golden/*.npz from the importable OpenCV (cv2 4.13) -- the only executable piece of the reference's
arithmetic available offline.  Run from the repo root:  python -m tools.make_golden
The fixtures pin the CPU oracle (tests/test_oracle_*.py) and, through it, the CUDA path."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def noise_image(h, w, seed, channels=1):
    rs = np.random.RandomState(seed)
    a = cv2.GaussianBlur(rs.rand(h, w).astype(np.float32), (0, 0), 1.5)
    a = (a - a.min()) / (a.max() - a.min()) * 255
    g = a.astype(np.uint8)
    if channels == 1:
        return g
    return np.stack([np.clip(g * 0.9, 0, 255).astype(np.uint8), g, np.clip(g * 0.8 + 10, 0, 255).astype(np.uint8)], -1)


def imgprep():
    w, h, clip = 320, 240, 3
    img = noise_image(h, w, 7, 3)
    K, D = synth.scaled_camera(synth.STEREO_YAML["left"], w, 1280)
    K[1, 2] = h / 2 + 2
    D[2:] = (0.001, -0.002)
    newK, _ = cv2.getOptimalNewCameraMatrix(K, D, (w, h), 0, (w, h), 0)
    gray = cv2.cvtColor(img, cv2.COLOR_RGB2GRAY)
    und = cv2.undistort(gray, K, D, None, newK)
    cl = cv2.createCLAHE()
    cl.setClipLimit(clip)
    out = cl.apply(und)
    integral = cv2.integral(out, sdepth=cv2.CV_32S)
    patches = {}
    for s in (25, 42, 57, 63, 100):
        patches[f"patch_{s}"] = cv2.resize(np.ascontiguousarray(out[:s, :s]), (21, 21), interpolation=cv2.INTER_AREA)
    np.savez_compressed(os.path.join(OUT, "imgprep_320x240.npz"), img=img, K=K, D=D, newK=newK, clip=clip, gray=gray,
                        und=und, out=out, integral=integral, g13=cv2.getGaussianKernel(13, 2.5, cv2.CV_32F).ravel(),
                        g20=cv2.getGaussianKernel(20, 3.3, cv2.CV_32F).ravel(), **patches)


def matcher():
    rs = np.random.RandomState(21)
    q = np.abs(rs.randn(300, 64)).astype(np.float32)
    t = np.abs(rs.randn(400, 64)).astype(np.float32)
    # make a third of the queries near-duplicates of train rows so the ratio test passes for them
    idx = rs.permutation(400)[:100]
    q[:100] = t[idx] + 0.02 * rs.randn(100, 64).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    t[7] = t[3]  # exact duplicate train rows -> tie must resolve to the lower trainIdx
    knn = cv2.BFMatcher(cv2.NORM_L2, False).knnMatch(q, t, 2)
    idx2 = np.array([[m.trainIdx for m in row] for row in knn], np.int32)
    dist2 = np.array([[m.distance for m in row] for row in knn], np.float32)
    np.savez_compressed(os.path.join(OUT, "matcher_300x400.npz"), q=q, t=t, idx=idx2, dist=dist2)


def matcher128():
    """extended-SURF row length (SURF_EXTENDED, VO_utility.h:86): the same 4 x 4-lane accumulators run over 8 groups"""
    rs = np.random.RandomState(22)
    q = np.abs(rs.randn(120, 128)).astype(np.float32)
    t = np.abs(rs.randn(160, 128)).astype(np.float32)
    idx = rs.permutation(160)[:40]
    q[:40] = t[idx] + 0.02 * rs.randn(40, 128).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    t[9] = t[5]
    knn = cv2.BFMatcher(cv2.NORM_L2, False).knnMatch(q, t, 2)
    idx2 = np.array([[m.trainIdx for m in row] for row in knn], np.int32)
    dist2 = np.array([[m.distance for m in row] for row in knn], np.float32)
    np.savez_compressed(os.path.join(OUT, "matcher128_120x160.npz"), q=q, t=t, idx=idx2, dist=dist2)


def jpeg():
    """JPEG streams (encoded by cv2 / libjpeg-turbo) with cv2.imdecode(IMREAD_UNCHANGED) as the expected output:
    the decode inside from_ros_to_cv_image (math_utility.cpp:154-173)"""
    from scipy import ndimage
    rs = np.random.RandomState(31)
    a = ndimage.gaussian_filter(rs.rand(48, 64, 3).astype(np.float32), (1.2, 1.2, 0))
    img = ((a - a.min()) / (a.max() - a.min()) * 255).astype(np.uint8)
    img[10:20, 30:40] = (255, 0, 0)  # saturated patches exercise the range limit of the colour conversion
    img[30:40, 5:15] = (0, 255, 255)
    S = cv2.IMWRITE_JPEG_SAMPLING_FACTOR
    out = {}
    for name, src, flags in [
        ("c420_rst2", img[:45, :61], [cv2.IMWRITE_JPEG_QUALITY, 80, S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420,
                                      cv2.IMWRITE_JPEG_RST_INTERVAL, 2]),
        ("c422", img[:45, :61], [cv2.IMWRITE_JPEG_QUALITY, 60, S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422]),
        ("c444_opt", img, [cv2.IMWRITE_JPEG_QUALITY, 95, S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444,
                           cv2.IMWRITE_JPEG_OPTIMIZE, 1]),
        ("c440", img[:47, :63], [cv2.IMWRITE_JPEG_QUALITY, 30, S, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440]),
        ("gray", img[:45, :61, 1].copy(), [cv2.IMWRITE_JPEG_QUALITY, 85]),
    ]:
        ok, enc = cv2.imencode(".jpg", src, flags)
        assert ok
        out[name + "_jpg"] = enc.ravel()
        out[name + "_img"] = cv2.imdecode(enc, cv2.IMREAD_UNCHANGED)
    np.savez_compressed(os.path.join(OUT, "jpeg_64x48.npz"), **out)


def pose():
    rs = np.random.RandomState(11)
    n = 600
    K = np.array([[1300.0, 0, 640], [0, 1300.0, 512], [0, 0, 1]])
    X = np.stack([rs.uniform(-4, 4, n), rs.uniform(-3, 3, n), rs.uniform(4, 9, n)], -1)
    rvec = np.array([0.01, -0.02, 0.015])
    tvec = np.array([0.3, 0.05, 0.1])
    x, _ = cv2.projectPoints(X, rvec, tvec, K, None)
    x = x.reshape(-1, 2) + rs.randn(n, 2) * 0.3
    out = rs.rand(n) < 0.3
    x[out] = np.stack([rs.uniform(0, 1280, out.sum()), rs.uniform(0, 1024, out.sum())], -1)
    x = x.astype(np.float32)
    ok, rv, tv, inl = cv2.solvePnPRansac(X, x, K, np.zeros(4), useExtrinsicGuess=False, iterationsCount=1000,
                                         reprojectionError=1.0, confidence=0.99, flags=cv2.SOLVEPNP_EPNP)
    # triangulation fixture: two views of the same points
    P1 = K @ np.hstack([np.eye(3), np.zeros((3, 1))])
    R2, _ = cv2.Rodrigues(rvec)
    P2 = K @ np.hstack([R2, tvec.reshape(3, 1)])
    x1, _ = cv2.projectPoints(X, np.zeros(3), np.zeros(3), K, None)
    x1 = (x1.reshape(-1, 2) + rs.randn(n, 2) * 0.2).astype(np.float32)
    x2, _ = cv2.projectPoints(X, rvec, tvec, K, None)
    x2 = (x2.reshape(-1, 2) + rs.randn(n, 2) * 0.2).astype(np.float32)
    X4 = cv2.triangulatePoints(P1, P2, x1.T.copy(), x2.T.copy())
    proj, _ = cv2.projectPoints(X, rvec, tvec, K, None)
    Rm, _ = cv2.Rodrigues(np.array([0.3, -0.2, 0.5]))
    rv_back, _ = cv2.Rodrigues(Rm)
    np.savez_compressed(os.path.join(OUT, "pose_600.npz"), K=K, X=X, x=x, pnp_ok=ok, pnp_rvec=rv.ravel(),
                        pnp_tvec=tv.ravel(), pnp_inliers=inl.ravel().astype(np.int32), P1=P1, P2=P2, x1=x1, x2=x2,
                        X4=X4.astype(np.float32), proj=proj.reshape(-1, 2), rvec=rvec, tvec=tvec, Rm=Rm,
                        rv_back=rv_back.ravel())


def main():
    os.makedirs(OUT, exist_ok=True)
    imgprep()
    matcher()
    matcher128()
    jpeg()
    pose()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
