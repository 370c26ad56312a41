"""ctypes binding of the CPU oracle (oracle/libuvo_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module
(see oracle/uvo_oracle.h).  The product package ergo_uvo_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libuvo_oracle.so")

KEYPOINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"),
     ("class_id", "<i4")])
DMATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        for name, res in (("orc_compute_median", C.c_double), ("orc_scale_factor", C.c_double),
                          ("orc_fast_atan2", C.c_float)):
            if hasattr(_lib, name):
                getattr(_lib, name).restype = res
    return _lib


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def _d4(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(-1))


def _k4(K):
    K = np.asarray(K, dtype=np.float64)
    if K.shape == (3, 3):
        return np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]], dtype=np.float64)
    return _d4(K)


# ------------------------------------------------------------------ image prep
def gray(img3):
    img3 = np.ascontiguousarray(img3, dtype=np.uint8)
    h, w, _ = img3.shape
    out = np.empty((h, w), np.uint8)
    lib().orc_gray(_p(img3), w, h, _p(out))
    return out


def undistort_map(K, D, newK, w, h):
    mxy = np.empty((h, w, 2), np.int16)
    mfr = np.empty((h, w), np.uint16)
    lib().orc_undistort_map(_p(_k4(K)), _p(_d4(D)), _p(_k4(newK)), w, h, _p(mxy), _p(mfr))
    return mxy, mfr


def undistort(g, K, D, newK):
    g = np.ascontiguousarray(g, dtype=np.uint8)
    h, w = g.shape
    out = np.empty_like(g)
    lib().orc_undistort(_p(g), w, h, _p(_k4(K)), _p(_d4(D)), _p(_k4(newK)), _p(out))
    return out


def clahe(g, clip_limit, tiles=(8, 8)):
    g = np.ascontiguousarray(g, dtype=np.uint8)
    h, w = g.shape
    out = np.empty_like(g)
    lib().orc_clahe(_p(g), w, h, C.c_double(clip_limit), tiles[0], tiles[1], _p(out))
    return out


def get_image(img3, K, D, newK, clahe_on=True, clip_limit=8.0):
    img3 = np.ascontiguousarray(img3, dtype=np.uint8)
    h, w, _ = img3.shape
    out = np.empty((h, w), np.uint8)
    lib().orc_get_image(_p(img3), w, h, _p(_k4(K)), _p(_d4(D)), _p(_k4(newK)), int(clahe_on),
                        C.c_double(clip_limit), _p(out))
    return out


def integral(g):
    g = np.ascontiguousarray(g, dtype=np.uint8)
    h, w = g.shape
    out = np.empty((h + 1, w + 1), np.int32)
    lib().orc_integral(_p(g), w, h, _p(out))
    return out


def bayer_bggr2bgr(bayer):
    """cv2.cvtColor(bayer, COLOR_BayerBGGR2BGR) (from_ros_to_cv_image, math_utility.cpp:161-164)"""
    b = np.ascontiguousarray(bayer, dtype=np.uint8)
    h, w = b.shape
    out = np.empty((h, w, 3), np.uint8)
    lib().orc_bayer_bggr2bgr(_p(b), w, h, _p(out))
    return out


def jpeg_decode(data):
    """cv2.imdecode(data, IMREAD_UNCHANGED) for baseline JPEG (from_ros_to_cv_image, math_utility.cpp:154-173)"""
    buf = np.frombuffer(bytes(data), np.uint8)
    w, h, c = C.c_int(0), C.c_int(0), C.c_int(0)
    rc = lib().orc_jpeg_info(_p(buf), C.c_size_t(len(buf)), C.byref(w), C.byref(h), C.byref(c))
    if rc < 0:
        raise ValueError(f"orc_jpeg_info: {rc}")
    out = np.empty((h.value, w.value) if c.value == 1 else (h.value, w.value, 3), np.uint8)
    rc = lib().orc_jpeg_decode(_p(buf), C.c_size_t(len(buf)), _p(out))
    if rc < 0:
        raise ValueError(f"orc_jpeg_decode: {rc}")
    return out


def jpeg_coefficients(data):
    """quantised coefficients after entropy decoding (parity tap for the product's host-side Huffman decoder)"""
    buf = np.frombuffer(bytes(data), np.uint8)
    cap = 1 << 20
    while True:
        out = np.empty(cap, np.int16)
        lib().orc_jpeg_coefficients.restype = C.c_long
        n = lib().orc_jpeg_coefficients(_p(buf), C.c_size_t(len(buf)), _p(out), C.c_size_t(cap))
        if n == -3:
            cap *= 4
            continue
        if n < 0:
            raise ValueError(f"orc_jpeg_coefficients: {n}")
        return out[:n].copy()


def resize_area(src, dw, dh):
    src = np.ascontiguousarray(src, dtype=np.uint8)
    sh, sw = src.shape[:2]
    cn = 1 if src.ndim == 2 else src.shape[2]
    out = np.empty((dh, dw) if src.ndim == 2 else (dh, dw, cn), np.uint8)
    lib().orc_resize_area(_p(src), sw, sh, cn, _p(out), dw, dh)
    return out


# ------------------------------------------------------------------ SURF
def surf_detect_and_compute(img, hessian_threshold, n_octaves=4, n_octave_layers=3, extended=False, upright=True,
                            capacity=1 << 16):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    dim = 128 if extended else 64
    kps = np.zeros(capacity, KEYPOINT_DTYPE)
    desc = np.zeros((capacity, dim), np.float32)
    n = lib().orc_surf_detect_and_compute(_p(img), w, h, C.c_double(hessian_threshold), n_octaves, n_octave_layers,
                                          int(extended), int(upright), _p(kps), _p(desc), capacity)
    if n < 0:
        return surf_detect_and_compute(img, hessian_threshold, n_octaves, n_octave_layers, extended, upright,
                                       capacity=-n)
    return kps[:n].copy(), desc[:n].copy()


def surf_det_trace_layer(sum_, size, step):
    sum_ = np.ascontiguousarray(sum_, dtype=np.int32)
    h, w = sum_.shape[0] - 1, sum_.shape[1] - 1
    det = np.zeros((h // step, w // step), np.float32)
    tr = np.zeros_like(det)
    lib().orc_surf_det_trace_layer(_p(sum_), w, h, size, step, _p(det), _p(tr))
    return det, tr


def surf_patch(img, cx, cy, size, angle=270.0, upright=True):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    patch = np.zeros((21, 21), np.uint8)
    ws = C.c_int(0)
    lib().orc_surf_patch(_p(img), w, h, C.c_float(cx), C.c_float(cy), C.c_float(size), C.c_float(angle),
                         int(upright), _p(patch), C.byref(ws))
    return patch, ws.value


def gaussian_kernel_f32(n, sigma):
    out = np.empty(n, np.float32)
    lib().orc_gaussian_kernel_f32(n, C.c_double(sigma), _p(out))
    return out


def fast_atan2(y, x):
    return float(lib().orc_fast_atan2(C.c_float(y), C.c_float(x)))


# ------------------------------------------------------------------ matcher
def knn2(q, t):
    q = np.ascontiguousarray(q, dtype=np.float32)
    t = np.ascontiguousarray(t, dtype=np.float32)
    out = np.zeros((q.shape[0], 2), DMATCH_DTYPE)
    lib().orc_knn2(_p(q), q.shape[0], _p(t), t.shape[0], q.shape[1] if q.ndim == 2 else 64, _p(out))
    return out


def match_features(q, t, ratio):
    q = np.ascontiguousarray(q, dtype=np.float32)
    t = np.ascontiguousarray(t, dtype=np.float32)
    out = np.zeros(q.shape[0], DMATCH_DTYPE)
    n = lib().orc_match_features(_p(q), q.shape[0], _p(t), t.shape[0], q.shape[1], C.c_float(ratio), _p(out))
    return out[:n].copy()


# ------------------------------------------------------------------ RNG / RANSAC bookkeeping
def rng_subsets(count, model_points, n_subsets):
    out = np.empty((n_subsets, model_points), np.int32)
    lib().orc_rng_subsets(count, model_points, n_subsets, _p(out))
    return out


def ransac_update_num_iters(p, ep, model_points, max_iters):
    return lib().orc_ransac_update_num_iters(C.c_double(p), C.c_double(ep), model_points, max_iters)


# ------------------------------------------------------------------ pose
def triangulate_points(P1, P2, pts1, pts2):
    P1 = np.ascontiguousarray(P1, np.float64)
    P2 = np.ascontiguousarray(P2, np.float64)
    pts1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
    pts2 = np.ascontiguousarray(pts2, np.float32).reshape(-1, 2)
    n = pts1.shape[0]
    out = np.empty((4, n), np.float32)
    lib().orc_triangulate_points(_p(P1), _p(P2), _p(pts1), _p(pts2), n, _p(out))
    return out


def extract_3dpoints(kp1, kp2, R1, t1, R2, t2, K1, K2, points4d, reproj_tol, min_num_3dpoints):
    kp1 = np.ascontiguousarray(kp1, np.float32).reshape(-1, 2)
    kp2 = np.ascontiguousarray(kp2, np.float32).reshape(-1, 2)
    n = kp1.shape[0]
    p4 = np.ascontiguousarray(points4d, np.float32)
    pts = np.empty((max(n, 1), 3), np.float64)
    idx = np.empty(max(n, 1), np.int32)
    m = lib().orc_extract_3dpoints(_p(kp1), _p(kp2), n, _p(_d4(R1)), _p(_d4(t1)), _p(_d4(R2)), _p(_d4(t2)),
                                   _p(_k4(K1)), _p(_k4(K2)), _p(p4), C.c_double(reproj_tol), min_num_3dpoints,
                                   _p(pts), _p(idx))
    return pts[:m].copy(), idx[:m].copy()


def project_points(X, R, t, K):
    X = np.ascontiguousarray(X, np.float64).reshape(-1, 3)
    out = np.empty((X.shape[0], 2), np.float64)
    lib().orc_project_points(_p(X), X.shape[0], _p(_d4(R)), _p(_d4(t)), _p(_k4(K)), _p(out))
    return out


def rodrigues_vec2mat(r):
    R = np.empty((3, 3), np.float64)
    lib().orc_rodrigues_vec2mat(_p(_d4(r)), _p(R))
    return R


def rodrigues_mat2vec(R):
    r = np.empty(3, np.float64)
    lib().orc_rodrigues_mat2vec(_p(_d4(R)), _p(r))
    return r


def compute_median(v):
    v = np.ascontiguousarray(v, np.float64)
    return float(lib().orc_compute_median(_p(v), v.shape[0]))


def scale_factor(pts, R, t, rng, with_count=False):
    pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
    m = C.c_int(0)
    lib().orc_scale_factor_front.restype = C.c_double
    sf = float(lib().orc_scale_factor_front(_p(pts), pts.shape[0], _p(_d4(R)), _p(_d4(t)), C.c_float(rng), C.byref(m)))
    return (sf, m.value) if with_count else sf


def select_estimation_method(p1, p2, distance):
    p1 = np.ascontiguousarray(p1, np.float32).reshape(-1, 2)
    p2 = np.ascontiguousarray(p2, np.float32).reshape(-1, 2)
    return bool(lib().orc_select_estimation_method(_p(p1), _p(p2), p1.shape[0], int(distance)))


def solve_pnp_ransac_epnp(X, x, K, iterations=1000, reproj_err=1.0, confidence=0.99):
    X = np.ascontiguousarray(X, np.float64).reshape(-1, 3)
    x = np.ascontiguousarray(x, np.float32).reshape(-1, 2)
    n = X.shape[0]
    rvec = np.zeros(3)
    tvec = np.zeros(3)
    inl = np.empty(max(n, 1), np.int32)
    hyp = C.c_int(0)
    m = lib().orc_solve_pnp_ransac_epnp(_p(X), _p(x), n, _p(_k4(K)), iterations, C.c_float(reproj_err),
                                        C.c_double(confidence), _p(rvec), _p(tvec), _p(inl), C.byref(hyp))
    return m > 0, rvec, tvec, inl[:max(m, 0)].copy(), hyp.value


def epnp(X, x, K):
    X = np.ascontiguousarray(X, np.float64).reshape(-1, 3)
    x = np.ascontiguousarray(x, np.float64).reshape(-1, 2)
    R = np.empty((3, 3))
    t = np.empty(3)
    lib().orc_epnp(_p(X), _p(x), X.shape[0], _p(_k4(K)), _p(R), _p(t))
    return R, t
