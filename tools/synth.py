"""Synthetic textured-seabed sequences for tests and bench.py.  numpy/scipy only.

The generator is test infrastructure: it produces the *inputs* (distorted 3-channel camera images, camera matrices,
ground-truth motion) that both the CUDA path and the CPU oracle consume.  Nothing here is on the measured path.
"""
import numpy as np
from scipy import ndimage

# uvo/config/stereo_VO_intrinsics.yaml:5-53 (calibrated for a ~1288 px wide image)
STEREO_YAML = dict(
    left=dict(fx=1.335036735254999e+03, fy=1.332419247540885e+03, cx=0.644564474737301e+03, cy=0.357685235527149e+03,
              k1=0.475667186716851, k2=0.126480045385593, p1=0.0, p2=0.0),
    right=dict(fx=1.330461901943011e+03, fy=1.328225165048530e+03, cx=0.684598875987595e+03, cy=0.382841174819059e+03,
               k1=0.493006394402676, k2=0.037112494470407, p1=0.0, p2=0.0),
    R_right=np.eye(3), t_right=np.array([-0.33, 0.0, 0.0]))
# uvo/config/mono_VO_intrinsics.yaml:5-20 (downward camera, ~2564 px wide)
MONO_YAML = dict(fx=2305.660253962050, fy=2303.950911497790, cx=1281.944364189583, cy=1028.352241411627, k1=0.08,
                 k2=0.45, p1=0.0, p2=0.0)


def make_texture(size=2048, seed=1234):
    """normalise(sum_k 2^(k/2) * blur(U[0,1], sigma=2^k)) to [16, 240]."""
    rs = np.random.RandomState(seed)
    base = rs.rand(size, size).astype(np.float32)
    acc = np.zeros_like(base)
    for k in range(6):
        acc += (2.0 ** (k / 2)) * ndimage.gaussian_filter(base, 2.0 ** k, mode="wrap")
    acc -= acc.min()
    acc /= acc.max()
    return (acc * 224.0 + 16.0).astype(np.float32)


def make_relief(size=512, seed=1235, sigma=16.0):
    rs = np.random.RandomState(seed)
    r = ndimage.gaussian_filter(rs.randn(size, size).astype(np.float32), sigma, mode="wrap")
    r /= np.abs(r).max()
    return r


def distort(x, y, k1, k2, p1, p2):
    r2 = x * x + y * y
    kr = 1 + (k2 * r2 + k1) * r2
    xd = x * kr + p1 * 2 * x * y + p2 * (r2 + 2 * x * x)
    yd = y * kr + p1 * (r2 + 2 * y * y) + p2 * 2 * x * y
    return xd, yd


def undistort_normalized(xd, yd, k1, k2, p1, p2, iters=30):
    x, y = xd.copy(), yd.copy()
    for _ in range(iters):
        r2 = x * x + y * y
        icdist = 1.0 / (1 + (k2 * r2 + k1) * r2)
        dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
        dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
        x = (xd - dx) * icdist
        y = (yd - dy) * icdist
    return x, y


def optimal_new_camera_matrix(K, D, w, h):
    """getOptimalNewCameraMatrix(K, D, (w,h), alpha=0, (w,h), centerPrincipalPoint=0) (VO_utility.cpp:674): inner
    rectangle of the undistorted 9x9 grid.  Own restatement (validated against cv2 to 1e-6 in tests); the result is
    an *input* of the parity tests, so OpenCV bit-equality is not needed."""
    N = 9
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    xs = np.array([[x * (w - 1) / (N - 1) for x in range(N)] for _ in range(N)], dtype=np.float64)
    ys = np.array([[y * (h - 1) / (N - 1)] * N for y in range(N)], dtype=np.float64)
    xn, yn = undistort_normalized((xs - cx) / fx, (ys - cy) / fy, D[0], D[1], D[2], D[3], iters=50)
    ix0 = xn[:, 0].max()
    ix1 = xn[:, N - 1].min()
    iy0 = yn[0, :].max()
    iy1 = yn[N - 1, :].min()
    fx0 = (w - 1) / (ix1 - ix0)
    fy0 = (h - 1) / (iy1 - iy0)
    cx0 = -fx0 * ix0
    cy0 = -fy0 * iy0
    return np.array([[fx0, 0, cx0], [0, fy0, cy0], [0, 0, 1]], dtype=np.float64)


def scaled_camera(y, w, native_w):
    s = w / float(native_w)
    K = np.array([[y["fx"] * s, 0, y["cx"] * s], [0, y["fy"] * s, y["cy"] * s], [0, 0, 1]], dtype=np.float64)
    D = np.array([y["k1"], y["k2"], y["p1"], y["p2"]], dtype=np.float64)
    return K, D


class Scene:
    """Plane z = z0 (metres, world frame) with relief, textured; world x,y span `extent` metres."""

    def __init__(self, tex_size=2048, extent=8.0, z0=3.0, relief_amp=0.3, seed=1234):
        self.tex = make_texture(tex_size, seed)
        self.relief = make_relief(512, seed + 1) * relief_amp
        self.extent = extent
        self.z0 = z0

    def _sample(self, img, X, Y):
        n = img.shape[0]
        u = (X / self.extent + 0.5) * n
        v = (Y / self.extent + 0.5) * n
        return ndimage.map_coordinates(img, [v, u], order=1, mode="wrap")

    def render(self, K, D, w, h, R_wc, p_w, gains=(0.9, 1.0, 0.8)):
        """Image seen by a camera with pose X_world = R_wc @ X_cam + p_w, through the *distorted* model (K, D)."""
        jj, ii = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
        xd = (jj - K[0, 2]) / K[0, 0]
        yd = (ii - K[1, 2]) / K[1, 1]
        x, y = undistort_normalized(xd, yd, D[0], D[1], D[2], D[3])
        d = np.stack([x, y, np.ones_like(x)], -1) @ R_wc.T  # ray directions in world
        zs = np.full_like(x, self.z0)
        for _ in range(4):
            t = (zs - p_w[2]) / d[..., 2]
            X = p_w[0] + t * d[..., 0]
            Y = p_w[1] + t * d[..., 1]
            zs = self.z0 + self._sample(self.relief, X, Y)
        g = self._sample(self.tex, X, Y)
        out = np.stack([np.clip(g * gains[0], 0, 255), np.clip(g * gains[1], 0, 255), np.clip(g * gains[2], 0, 255)],
                       -1)
        return np.rint(out).astype(np.uint8)


def rot_z(deg):
    a = np.deg2rad(deg)
    return np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])


class StereoSequence:
    """Constant-velocity stereo sequence: per-frame left/right 3-channel images."""

    def __init__(self, w=1280, h=1024, n_frames=4, seed=1234, tex_size=2048, velocity=(0.010, 0.002, 0.001),
                 yaw_deg=0.2, relief_amp=0.3):
        self.w, self.h = w, h
        self.scene = Scene(tex_size=tex_size, seed=seed, relief_amp=relief_amp)
        self.KL, self.DL = scaled_camera(STEREO_YAML["left"], w, 1280)
        self.KR, self.DR = scaled_camera(STEREO_YAML["right"], w, 1280)
        # the YAML principal points are for a 1288x720-ish sensor; recentre vertically for the synthetic geometry
        self.KL[1, 2] = h * 0.5 + 6.0
        self.KR[1, 2] = h * 0.5 - 4.0
        self.newKL = optimal_new_camera_matrix(self.KL, self.DL, w, h)
        self.newKR = optimal_new_camera_matrix(self.KR, self.DR, w, h)
        self.R_right = STEREO_YAML["R_right"].copy()
        self.t_right = STEREO_YAML["t_right"].copy()
        self.velocity = np.asarray(velocity, dtype=np.float64)
        self.yaw_deg = yaw_deg
        self.frames = []
        for k in range(n_frames):
            self.frames.append(self.render_pair(k))

    def pose(self, k):
        R = rot_z(self.yaw_deg * k)
        p = self.velocity * k
        return R, p

    def render_pair(self, k):
        R, p = self.pose(k)
        left = self.scene.render(self.KL, self.DL, self.w, self.h, R, p)
        # X_right = R_right X_left + t_right  =>  right camera centre in left coords = -R_right^T t_right
        Rr = R @ self.R_right.T
        pr = p + R @ (-self.R_right.T @ self.t_right)
        right = self.scene.render(self.KR, self.DR, self.w, self.h, Rr, pr)
        return left, right

    def true_t_prev_curr(self, k):
        """translation of camera k expressed in camera k-1 (what stereo_VO publishes times dt)."""
        R0, p0 = self.pose(k - 1)
        _, p1 = self.pose(k)
        return R0.T @ (p1 - p0)


class MonoSequence:
    def __init__(self, w=640, h=480, n_frames=4, seed=1234, tex_size=2048, velocity=(0.02, 0.004, 0.0),
                 yaw_deg=0.2, relief_amp=0.3):
        self.w, self.h = w, h
        self.scene = Scene(tex_size=tex_size, seed=seed, relief_amp=relief_amp)
        self.K, self.D = scaled_camera(MONO_YAML, w, 2564)
        self.K[1, 2] = h * 0.5 + 3.0
        self.newK = optimal_new_camera_matrix(self.K, self.D, w, h)
        self.velocity = np.asarray(velocity, dtype=np.float64)
        self.yaw_deg = yaw_deg
        self.frames = [self.scene.render(self.K, self.D, w, h, rot_z(yaw_deg * k), self.velocity * k)
                       for k in range(n_frames)]
        rs = np.random.RandomState(seed + 2)
        self.ranges = [self.scene.z0 + 0.01 * rs.randn() for _ in range(n_frames)]
