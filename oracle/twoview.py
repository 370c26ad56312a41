"""CPU ORACLE (test infrastructure only) for the mono two-view step: findHomography, findEssentialMat, recoverPose,
decomposeHomographyMat as the reference calls them from estimate_relative_pose / recover_pose_homography
(uvo_libraries/src/VO_utility.cpp:134-180, :581-624).

The arithmetic lives in OpenCV (un-vendored; the reference pins "OpenCV 4.5"): calib3d/src/ptsetreg.cpp (RANSAC and
LMedS drivers, subset stream), fundam.cpp (homography kernel, f32 reprojection error, refinement), five-point.cpp
(Nister 5-point, Sampson error, recoverPose), homography_decomp.cpp, levmarq.cpp.  This module restates those
algorithms in numpy and is PINNED against cv2 4.13 run in this container
(tests/test_oracle_twoview.py live when cv2 is importable + tests/golden/twoview_*.npz made by
tools/make_golden_twoview.py):
  * subset stream, checkSubset, stopping rule, LMedS sigma rule: exact (same inlier masks, same hypothesis count);
  * findHomography's returned mask is that of the refined H at the reprojection threshold (cv2 4.13 behaviour for
    RANSAC and LMedS; SURVEY C.8 observed only cases where it coincides with the best-hypothesis mask);
  * model values (H, E, R, t): to 1e-8 relative -- minimal-set solvers go through LAPACK in the wheel and through
    numpy's LAPACK here; they are not bit-reproducible across builds.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np

F32 = np.float32
FLT_EPS = float(np.finfo(np.float32).eps)
DBL_EPS = float(np.finfo(np.float64).eps)
DBL_MIN = float(np.finfo(np.float64).tiny)
RANSAC, LMEDS = 8, 4


# ------------------------------------------------------------------------------------------------ cv::RNG (B.9)
class CvRng:
    def __init__(self, state=0xFFFFFFFFFFFFFFFF):
        self.state = state if state else 0xFFFFFFFF

    def next(self):
        self.state = ((self.state & 0xFFFFFFFF) * 4164903690 + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a, b):
        return a if a == b else a + self.next() % (b - a)


def ransac_update_num_iters(p, ep, model_points, max_iters):
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    num = max(1.0 - p, DBL_MIN)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < DBL_MIN:
        return 0
    num = np.log(num)
    denom = np.log(denom)
    return max_iters if (denom >= 0 or -num >= max_iters * (-denom)) else int(np.rint(num / denom))


def get_subset(rng, count, model_points, check_subset, m1, m2, max_attempts=1000):
    """RANSACPointSetRegistrator::getSubset: draw distinct indices, accept iff checkSubset; None after 1000 tries"""
    for _ in range(max_attempts):
        idx = []
        for _i in range(model_points):
            while True:
                v = rng.uniform(0, count)
                if v not in idx:
                    break
            idx.append(v)
        if check_subset is None or check_subset(m1[idx], m2[idx]):
            return idx
    return None


# ------------------------------------------------------------------------------------------------ homography (B.5)
def _have_collinear(p):
    """haveCollinearPoints for the last point of the subset (f32 points, f64 arithmetic)"""
    p = p.astype(np.float64)
    i = len(p) - 1
    for j in range(i):
        dx1, dy1 = p[j, 0] - p[i, 0], p[j, 1] - p[i, 1]
        for k in range(j):
            dx2, dy2 = p[k, 0] - p[i, 0], p[k, 1] - p[i, 1]
            if abs(dx2 * dy1 - dy2 * dx1) <= FLT_EPS * (abs(dx1) + abs(dy1) + abs(dx2) + abs(dy2)):
                return True
    return False


def _det3(a):
    return (a[0, 0] * (a[1, 1] * a[2, 2] - a[1, 2] * a[2, 1]) - a[0, 1] * (a[1, 0] * a[2, 2] - a[1, 2] * a[2, 0]) +
            a[0, 2] * (a[1, 0] * a[2, 1] - a[1, 1] * a[2, 0]))


def homography_check_subset(ms1, ms2):
    if _have_collinear(ms1) or _have_collinear(ms2):
        return False
    if len(ms1) == 4:
        neg = 0
        for t in ((0, 1, 2), (1, 2, 3), (0, 2, 3), (0, 1, 3)):
            A = np.c_[ms1[list(t)].astype(np.float64), np.ones(3)]
            B = np.c_[ms2[list(t)].astype(np.float64), np.ones(3)]
            neg += _det3(A) * _det3(B) < 0
        if neg != 0 and neg != 4:
            return False
    return True


def homography_kernel(M, m):
    """HomographyEstimatorCallback::runKernel: normalised DLT through the 9x9 LtL eigen-decomposition.  Returns H
    (H[2,2] == 1) or None."""
    M = M.astype(np.float64)
    m = m.astype(np.float64)
    n = len(M)
    cM, cm = M.sum(0) / n, m.sum(0) / n
    sM, sm = np.abs(M - cM).sum(0), np.abs(m - cm).sum(0)
    if min(abs(sM[0]), abs(sM[1]), abs(sm[0]), abs(sm[1])) < DBL_EPS:
        return None
    sM, sm = n / sM, n / sm
    invHnorm = np.array([[1. / sm[0], 0, cm[0]], [0, 1. / sm[1], cm[1]], [0, 0, 1]])
    Hnorm2 = np.array([[sM[0], 0, -cM[0] * sM[0]], [0, sM[1], -cM[1] * sM[1]], [0, 0, 1]])
    x, y = (m[:, 0] - cm[0]) * sm[0], (m[:, 1] - cm[1]) * sm[1]
    X, Y = (M[:, 0] - cM[0]) * sM[0], (M[:, 1] - cM[1]) * sM[1]
    one, zero = np.ones(n), np.zeros(n)
    Lx = np.stack([X, Y, one, zero, zero, zero, -x * X, -x * Y, -x], 1)
    Ly = np.stack([zero, zero, zero, X, Y, one, -y * X, -y * Y, -y], 1)
    LtL = Lx.T @ Lx + Ly.T @ Ly
    w, V = np.linalg.eigh(LtL)
    h = V[:, 0].reshape(3, 3)  # eigenvector of the smallest eigenvalue
    H = invHnorm @ h @ Hnorm2
    return H / H[2, 2]


def homography_error_f32(H, M, m):
    """HomographyEstimatorCallback::computeError: pure f32, separate multiply/add (no contraction)"""
    Hf = H.astype(F32).reshape(-1)
    Mx, My = M[:, 0].astype(F32), M[:, 1].astype(F32)
    one = F32(1)
    ww = one / ((Hf[6] * Mx + Hf[7] * My) + one)
    dx = ((Hf[0] * Mx + Hf[1] * My) + Hf[2]) * ww - m[:, 0].astype(F32)
    dy = ((Hf[3] * Mx + Hf[4] * My) + Hf[5]) * ww - m[:, 1].astype(F32)
    return dx * dx + dy * dy


def _refine_compute(h, M, m, want_j):
    Mx, My = M[:, 0], M[:, 1]
    ww = h[6] * Mx + h[7] * My + 1.
    ww = np.where(np.abs(ww) > DBL_EPS, 1. / np.where(ww == 0, 1, ww), 0.)
    xi = (h[0] * Mx + h[1] * My + h[2]) * ww
    yi = (h[3] * Mx + h[4] * My + h[5]) * ww
    err = np.empty(2 * len(M))
    err[0::2] = xi - m[:, 0]
    err[1::2] = yi - m[:, 1]
    if not want_j:
        return err, None
    J = np.zeros((2 * len(M), 8))
    J[0::2, 0], J[0::2, 1], J[0::2, 2] = Mx * ww, My * ww, ww
    J[0::2, 6], J[0::2, 7] = -Mx * ww * xi, -My * ww * xi
    J[1::2, 3], J[1::2, 4], J[1::2, 5] = Mx * ww, My * ww, ww
    J[1::2, 6], J[1::2, 7] = -Mx * ww * yi, -My * ww * yi
    return err, J


def _solve_eig(A, b):
    """cv::solve(A, b, x, DECOMP_EIG) for symmetric A: x = V diag(1/w) V^T b with tiny eigenvalues dropped"""
    w, V = np.linalg.eigh(A)
    t = V.T @ b
    thr = np.abs(w).sum() * DBL_EPS * 2  # cvSVBkSb threshold on the spectrum
    inv = np.where(np.abs(w) > thr, 1. / np.where(w == 0, 1, w), 0.)
    return V @ (t * inv)


def lm_refine_homography(H, M, m, max_iters=10):
    """LMSolver(HomographyRefineCallback, 10 iterations) on the first 8 entries of H (H[2,2] stays 1); levmarq.cpp"""
    M = M.astype(np.float64)
    m = m.astype(np.float64)
    x = H.reshape(-1)[:8].copy()
    r, J = _refine_compute(x, M, m, True)
    S = float(r @ r)
    A = J.T @ J
    v = J.T @ r
    D = np.diag(A).copy()
    Rlo, Rhi = 0.25, 0.75
    lam, lc = 1.0, 0.75
    it = 0
    while True:
        Ap = A + np.diag(lam * D)
        d = _solve_eig(Ap, v)
        xd = x - d
        rd, _ = _refine_compute(xd, M, m, False)
        Sd = float(rd @ rd)
        temp_d = -(A @ d) + 2 * v
        dS = float(d @ temp_d)
        R = (S - Sd) / (dS if abs(dS) > DBL_EPS else 1)
        if R > Rhi:
            lam *= 0.5
            if lam < lc:
                lam = 0
        elif R < Rlo:
            t = float(d @ v)
            nu = (Sd - S) / (t if abs(t) > DBL_EPS else 1) + 2
            nu = min(max(nu, 2.), 10.)
            if lam == 0:
                w, V = np.linalg.eigh(A)
                thr = np.abs(w).sum() * DBL_EPS * 2
                inv = np.where(np.abs(w) > thr, 1. / np.where(w == 0, 1, w), 0.)
                Ainv = (V * inv) @ V.T
                maxval = max(DBL_EPS, float(np.abs(np.diag(Ainv)).max()))
                lam = lc = 1. / maxval
                nu *= 0.5
            lam *= nu
        if Sd < S:
            S = Sd
            x = xd
            r, J = _refine_compute(x, M, m, True)
            A = J.T @ J
            v = J.T @ r
        it += 1
        if not (it < max_iters and np.abs(d).max() >= FLT_EPS and np.abs(r).max() >= FLT_EPS):
            break
    out = np.append(x, 1.0).reshape(3, 3)
    return out


# ------------------------------------------------------------------------------------------------ drivers (B.1, B.2)
def _run_ransac(m1, m2, model_points, kernel, error_fn, check_subset, threshold, confidence, max_iters):
    """RANSACPointSetRegistrator::run.  Returns (model, mask, hypotheses evaluated) or (None, zeros, n)"""
    count = len(m1)
    rng = CvRng()
    niters = max(max_iters, 1)
    best_model, best_mask, max_good = None, np.zeros(count, np.uint8), 0
    thr = F32(threshold * threshold)
    if count == model_points:
        models = kernel(m1, m2)
        if not models:
            return None, best_mask, 0
        return models[0], np.ones(count, np.uint8), 1
    it = 0
    while it < niters:
        idx = get_subset(rng, count, model_points, check_subset, m1, m2)
        if idx is None:
            if it == 0:
                return None, best_mask, 0
            break
        for mdl in kernel(m1[idx], m2[idx]):
            err = error_fn(mdl, m1, m2)
            mask = err <= thr
            good = int(mask.sum())
            if good > max(max_good, model_points - 1):
                best_model, best_mask, max_good = mdl, mask.astype(np.uint8), good
                niters = ransac_update_num_iters(confidence, (count - good) / count, model_points, niters)
        it += 1
    return best_model, best_mask, it


def _run_lmeds(m1, m2, model_points, kernel, error_fn, check_subset, confidence, max_iters):
    """LMeDSPointSetRegistrator::run: fixed iteration count, median of the f32 errors at count/2, sigma rule"""
    count = len(m1)
    rng = CvRng()
    best_model, min_median = None, np.inf
    if count == model_points:
        models = kernel(m1, m2)
        if not models:
            return None, np.zeros(count, np.uint8), 0
        return models[0], np.ones(count, np.uint8), 1
    niters = ransac_update_num_iters(confidence, 0.45, model_points, max(max_iters, 1))
    it = 0
    while it < niters:
        idx = get_subset(rng, count, model_points, check_subset, m1, m2)
        if idx is None:
            if it == 0:
                return None, np.zeros(count, np.uint8), 0
            break
        for mdl in kernel(m1[idx], m2[idx]):
            err = error_fn(mdl, m1, m2)
            med = float(np.partition(err.astype(F32), count // 2)[count // 2])
            if med < min_median:
                min_median, best_model = med, mdl
        it += 1
    if min_median < np.inf:
        sigma = 2.5 * 1.4826 * (1 + 5. / (count - model_points)) * np.sqrt(min_median)
        sigma = max(sigma, 0.001)
        err = error_fn(best_model, m1, m2)
        mask = (err <= F32(sigma * sigma)).astype(np.uint8)
        return best_model, mask, it
    return None, np.zeros(count, np.uint8), it


def find_homography(p1, p2, method=RANSAC, threshold=3.0, max_iters=2000, confidence=0.995):
    """cv::findHomography(p1, p2, method, threshold, mask, maxIters, confidence) -> (H or None, mask, hypotheses)"""
    p1 = np.ascontiguousarray(p1, F32).reshape(-1, 2)
    p2 = np.ascontiguousarray(p2, F32).reshape(-1, 2)
    n = len(p1)
    if threshold <= 0:
        threshold = 3.0

    def kernel(a, b):
        H = homography_kernel(a, b)
        return [] if H is None else [H]

    if method == 0 or n == 4:
        H = homography_kernel(p1, p2)
        mask, hyp = np.ones(n, np.uint8), 1
    elif method == RANSAC:
        H, mask, hyp = _run_ransac(p1, p2, 4, kernel, homography_error_f32, homography_check_subset, threshold,
                                   confidence, max_iters)
    elif method == LMEDS:
        H, mask, hyp = _run_lmeds(p1, p2, 4, kernel, homography_error_f32, homography_check_subset, confidence,
                                  max_iters)
    else:
        raise ValueError("method")
    if H is None:
        return None, np.zeros(n, np.uint8), hyp
    if n > 4:
        sel = mask.astype(bool)
        if sel.any():
            a, b = p1[sel], p2[sel]
            if method in (RANSAC, LMEDS):
                H2 = homography_kernel(a, b)
                if H2 is not None:
                    H = H2
            H = lm_refine_homography(H, a, b, 10)
            # cv2 4.13 (the build this oracle is pinned to) returns the mask of the REFINED model at the
            # reprojection threshold, for RANSAC and LMedS alike (probed: 12/12 cases, tools/make_golden_twoview.py);
            # the hypothesis mask above only selects the points of the refit.
            mask = (homography_error_f32(H, p1, p2) <= F32(threshold * threshold)).astype(np.uint8)
    return H, mask, hyp


# ------------------------------------------------------------------------------------------------ 5-point (B.3)
# monomial order of the 10 x 20 constraint matrix (Nister / Stewenius): degree-3 monomials of (x, y, z) with
# E = x*E0 + y*E1 + z*E2 + E3
_MONO = [(3, 0, 0), (0, 3, 0), (2, 1, 0), (1, 2, 0), (2, 0, 1), (2, 0, 0), (0, 2, 1), (0, 2, 0), (1, 1, 1), (1, 1, 0),
         (1, 0, 2), (1, 0, 1), (1, 0, 0), (0, 1, 2), (0, 1, 1), (0, 1, 0), (0, 0, 3), (0, 0, 2), (0, 0, 1), (0, 0, 0)]
_MONO_IDX = {m: i for i, m in enumerate(_MONO)}


class _Poly:
    """polynomial in (x, y, z) of total degree <= 3 as a dict monomial -> coefficient"""

    def __init__(self, c=None):
        self.c = c or {}

    def __add__(self, o):
        r = dict(self.c)
        for k, v in o.c.items():
            r[k] = r.get(k, 0.0) + v
        return _Poly(r)

    def __sub__(self, o):
        r = dict(self.c)
        for k, v in o.c.items():
            r[k] = r.get(k, 0.0) - v
        return _Poly(r)

    def __mul__(self, o):
        if not isinstance(o, _Poly):
            return _Poly({k: v * o for k, v in self.c.items()})
        r = {}
        for k1, v1 in self.c.items():
            for k2, v2 in o.c.items():
                k = (k1[0] + k2[0], k1[1] + k2[1], k1[2] + k2[2])
                r[k] = r.get(k, 0.0) + v1 * v2
        return _Poly(r)

    def row(self):
        out = np.zeros(20)
        for k, v in self.c.items():
            out[_MONO_IDX[k]] += v
        return out


def five_point_constraints(EE):
    """EE: 4 x 9 null-space basis (rows E0..E3, row-major 3x3).  Returns the 10 x 20 matrix of det(E) = 0 and
    2 E E^T E - trace(E E^T) E = 0 in the monomial order above (getCoeffMat, restated by polynomial arithmetic)."""
    x, y, z, one = (1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, 0)
    E = [[_Poly({x: EE[0][3 * i + j], y: EE[1][3 * i + j], z: EE[2][3 * i + j], one: EE[3][3 * i + j]})
          for j in range(3)] for i in range(3)]
    det = (E[0][0] * (E[1][1] * E[2][2] - E[1][2] * E[2][1]) - E[0][1] * (E[1][0] * E[2][2] - E[1][2] * E[2][0]) +
           E[0][2] * (E[1][0] * E[2][1] - E[1][1] * E[2][0]))
    EEt = [[E[i][0] * E[j][0] + E[i][1] * E[j][1] + E[i][2] * E[j][2] for j in range(3)] for i in range(3)]
    tr = EEt[0][0] + EEt[1][1] + EEt[2][2]
    rows = [det.row()]
    for i in range(3):
        for j in range(3):
            p = (EEt[i][0] * E[0][j] + EEt[i][1] * E[1][j] + EEt[i][2] * E[2][j]) * 2.0 - tr * E[i][j]
            rows.append(p.row())
    return np.array(rows)


def _polymul(a, b):
    return np.convolve(a, b)


def five_point_kernel(q1, q2):
    """EMEstimatorCallback::runKernel on normalised f64 points: up to 10 essential matrices (unit Frobenius norm)"""
    q1 = np.asarray(q1, np.float64)
    q2 = np.asarray(q2, np.float64)
    n = len(q1)
    # x2^T E x1 = 0 with E row-major: column 3 i + j holds x2_i * x1_j (homogeneous third coordinate 1)
    x1h = np.c_[q1, np.ones(n)]
    x2h = np.c_[q2, np.ones(n)]
    Q = (x2h[:, :, None] * x1h[:, None, :]).reshape(n, 9)
    _, _, Vt = np.linalg.svd(Q, full_matrices=True)
    EE = Vt[5:9]  # null-space basis: rows 5..8 of Vt
    A = five_point_constraints(EE)
    try:
        A = np.linalg.solve(A[:, :10], A[:, 10:])
    except np.linalg.LinAlgError:
        return []
    # rows 4..9: x^2 z, x^2, y^2 z, y^2, x y z, x y expressed in (x z^2, x z, x, y z^2, y z, y, z^3, z^2, z, 1)
    B = np.zeros((3, 13))
    for i in range(3):
        r1, r2 = A[2 * i + 4], A[2 * i + 5]
        row1, row2 = np.zeros(13), np.zeros(13)
        row1[1:4], row1[5:8], row1[9:13] = r1[0:3], r1[3:6], r1[6:10]
        row2[0:3], row2[4:7], row2[8:12] = r2[0:3], r2[3:6], r2[6:10]
        B[i] = row1 - row2
    # B(z) [x, y, 1]^T = 0: entries are polynomials in z (highest power first): degree 3, 3, 4
    P = [[B[i, 0:4], B[i, 4:8], B[i, 8:13]] for i in range(3)]

    def det2(a, b, c, d):
        return np.polysub(_polymul(a, d), _polymul(b, c))

    det = np.polyadd(np.polysub(_polymul(P[0][0], det2(P[1][1], P[1][2], P[2][1], P[2][2])),
                                _polymul(P[0][1], det2(P[1][0], P[1][2], P[2][0], P[2][2]))),
                     _polymul(P[0][2], det2(P[1][0], P[1][1], P[2][0], P[2][1])))
    det = np.concatenate([np.zeros(11 - len(det)), det]) if len(det) < 11 else det[-11:]
    roots = np.roots(det)
    out = []
    for rt in roots:
        if abs(rt.imag) > 1e-10:
            continue
        z1 = rt.real
        zp = np.array([z1 ** 3, z1 ** 2, z1, 1.0])
        zp4 = np.array([z1 ** 4, z1 ** 3, z1 ** 2, z1, 1.0])
        Bz = np.array([[B[j, 0:4] @ zp, B[j, 4:8] @ zp, B[j, 8:13] @ zp4] for j in range(3)])
        _, _, vt = np.linalg.svd(Bz)
        xy1 = vt[2]
        if abs(xy1[2]) < 1e-10:
            continue
        xs, ys = xy1[0] / xy1[2], xy1[1] / xy1[2]
        Evec = EE[0] * xs + EE[1] * ys + EE[2] * z1 + EE[3]
        Evec = Evec / np.linalg.norm(Evec)
        out.append(Evec.reshape(3, 3))
    return out


def sampson_error_f32(E, x1, x2):
    """EMEstimatorCallback::computeError on normalised f64 points: f64 Sampson distance cast to f32"""
    x1h = np.c_[x1, np.ones(len(x1))]
    x2h = np.c_[x2, np.ones(len(x2))]
    Ex1 = x1h @ E.T
    Etx2 = x2h @ E
    x2tEx1 = (x2h * Ex1).sum(1)
    a = Ex1[:, 0] ** 2
    b = Ex1[:, 1] ** 2
    c = Etx2[:, 0] ** 2
    d = Etx2[:, 1] ** 2
    return (x2tEx1 * x2tEx1 / (a + b + c + d)).astype(F32)


def find_essential_mat(p1, p2, K4, method=RANSAC, prob=0.999, threshold=1.0, max_iters=1000):
    """cv::findEssentialMat(p1, p2, K, method, prob, threshold, maxIters, mask) -> (E or None, mask, hypotheses).
    K4 = (fx, fy, cx, cy)."""
    fx, fy, cx, cy = [float(v) for v in K4]
    p1 = np.ascontiguousarray(p1, F32).reshape(-1, 2).astype(np.float64)
    p2 = np.ascontiguousarray(p2, F32).reshape(-1, 2).astype(np.float64)
    q1 = np.stack([(p1[:, 0] - cx) / fx, (p1[:, 1] - cy) / fy], 1)
    q2 = np.stack([(p2[:, 0] - cx) / fx, (p2[:, 1] - cy) / fy], 1)
    threshold = threshold / ((fx + fy) / 2)
    if len(q1) < 5:
        return None, np.zeros(len(q1), np.uint8), 0
    if method == RANSAC:
        return _run_ransac(q1, q2, 5, five_point_kernel, sampson_error_f32, None, threshold, prob, max_iters)
    return _run_lmeds(q1, q2, 5, five_point_kernel, sampson_error_f32, None, prob, max_iters)


# ------------------------------------------------------------------------------------------------ recoverPose (B.4)
def decompose_essential(E):
    U, _, Vt = np.linalg.svd(E)
    if np.linalg.det(U) < 0:
        U = -U
    if np.linalg.det(Vt) < 0:
        Vt = -Vt
    W = np.array([[0., 1, 0], [-1, 0, 0], [0, 0, 1]])
    return U @ W @ Vt, U @ W.T @ Vt, U[:, 2].copy()


def _triangulate_h(P0, P1, x0, x1):
    """cv::triangulatePoints (DLT, smallest right singular vector), f64; returns 4 x N"""
    out = np.empty((4, len(x0)))
    for i in range(len(x0)):
        A = np.stack([x0[i, 0] * P0[2] - P0[0], x0[i, 1] * P0[2] - P0[1], x1[i, 0] * P1[2] - P1[0],
                      x1[i, 1] * P1[2] - P1[1]])
        out[:, i] = np.linalg.svd(A)[2][3]
    return out


def recover_pose(E, p1, p2, K4, mask_in=None, distance_thresh=50.0):
    """cv::recoverPose(E, p1, p2, K, R, t, mask): cheirality vote over the 4 decompositions.  Returns
    (good, R, t, mask_out)."""
    fx, fy, cx, cy = [float(v) for v in K4]
    p1 = np.ascontiguousarray(p1, F32).reshape(-1, 2).astype(np.float64)
    p2 = np.ascontiguousarray(p2, F32).reshape(-1, 2).astype(np.float64)
    q1 = np.stack([(p1[:, 0] - cx) / fx, (p1[:, 1] - cy) / fy], 1)
    q2 = np.stack([(p2[:, 0] - cx) / fx, (p2[:, 1] - cy) / fy], 1)
    R1, R2, t = decompose_essential(np.asarray(E, np.float64))
    P0 = np.eye(3, 4)
    cands = [(R1, t), (R2, t), (R1, -t), (R2, -t)]
    masks = []
    for R, tt in cands:
        P = np.c_[R, tt]
        Q = _triangulate_h(P0, P, q1, q2)
        m = (Q[2] * Q[3]) > 0
        Qn = Q / Q[3]
        m &= Qn[2] < distance_thresh
        Q2 = P @ Qn
        m &= (Q2[2] > 0) & (Q2[2] < distance_thresh)
        if mask_in is not None:
            m &= np.asarray(mask_in).astype(bool).reshape(-1)
        masks.append(m)
    goods = [int(m.sum()) for m in masks]
    # OpenCV's selection order: good1 >= others -> (R1, t); good2 -> (R2, t); good3 -> (R1, -t); else (R2, -t)
    g1, g2, g3, g4 = goods
    if g1 >= g2 and g1 >= g3 and g1 >= g4:
        k = 0
    elif g2 >= g1 and g2 >= g3 and g2 >= g4:
        k = 1
    elif g3 >= g1 and g3 >= g2 and g3 >= g4:
        k = 2
    else:
        k = 3
    return goods[k], cands[k][0], cands[k][1], masks[k].astype(np.uint8)


# ------------------------------------------------------------------------------------------------ decomposeHomographyMat (B.8)
def _opp_minor(M, row, col):
    x1 = 1 if col == 0 else 0
    x2 = 1 if col == 2 else 2
    y1 = 1 if row == 0 else 0
    y2 = 1 if row == 2 else 2
    return M[y1, x2] * M[y2, x1] - M[y1, x1] * M[y2, x2]


def _signd(x):
    return 1.0 if x >= 0 else -1.0


def decompose_homography_mat(H, K):
    """cv::decomposeHomographyMat (Malis-Vargas analytical, HomographyDecompInria): list of (R, t, n), 4 entries
    (1 when H is a pure rotation).  K is the 3x3 camera matrix."""
    H = np.asarray(H, np.float64).reshape(3, 3)
    K = np.asarray(K, np.float64).reshape(3, 3)
    Hn = np.linalg.inv(K) @ H @ K
    w = np.linalg.svd(Hn, compute_uv=False)
    Hn = Hn / w[1]
    S = Hn.T @ Hn - np.eye(3)
    if np.abs(S).sum(1).max() < 0.001:  # NORM_INF of a matrix in OpenCV: max |entry|; either is far below for real H
        if np.abs(S).max() < 0.001:
            return [(Hn, np.zeros(3), np.zeros(3))]
    M00, M11, M22 = _opp_minor(S, 0, 0), _opp_minor(S, 1, 1), _opp_minor(S, 2, 2)
    rtM00, rtM11, rtM22 = np.sqrt(M00), np.sqrt(M11), np.sqrt(M22)
    M01, M12, M02 = _opp_minor(S, 0, 1), _opp_minor(S, 1, 2), _opp_minor(S, 0, 2)
    e12, e02, e01 = _signd(M12), _signd(M02), _signd(M01)
    nS = [abs(S[0, 0]), abs(S[1, 1]), abs(S[2, 2])]
    indx = 0
    if nS[0] < nS[1]:
        indx = 1
        if nS[1] < nS[2]:
            indx = 2
    elif nS[0] < nS[2]:
        indx = 2
    if indx == 0:
        npa = np.array([S[0, 0], S[0, 1] + rtM22, S[0, 2] + e12 * rtM11])
        npb = np.array([S[0, 0], S[0, 1] - rtM22, S[0, 2] - e12 * rtM11])
    elif indx == 1:
        npa = np.array([S[0, 1] + rtM22, S[1, 1], S[1, 2] - e02 * rtM00])
        npb = np.array([S[0, 1] - rtM22, S[1, 1], S[1, 2] + e02 * rtM00])
    else:
        npa = np.array([S[0, 2] + e01 * rtM11, S[1, 2] + rtM00, S[2, 2]])
        npb = np.array([S[0, 2] - e01 * rtM11, S[1, 2] - rtM00, S[2, 2]])
    traceS = S[0, 0] + S[1, 1] + S[2, 2]
    v = 2.0 * np.sqrt(1 + traceS - M00 - M11 - M22)
    ESii = _signd(S[indx, indx])
    r = np.sqrt(2 + traceS + v)
    n_t = np.sqrt(2 + traceS - v)
    na = npa / np.linalg.norm(npa)
    nb = npb / np.linalg.norm(npb)
    half_nt = 0.5 * n_t
    esii_t_r = ESii * r
    ta_star = half_nt * (esii_t_r * nb - n_t * na)
    tb_star = half_nt * (esii_t_r * na - n_t * nb)

    def rmat(tstar, n):
        R = Hn @ (np.eye(3) - (2.0 / v) * np.outer(tstar, n))
        if np.linalg.det(R) < 0:
            R = -R
        return R

    Ra = rmat(ta_star, na)
    ta = Ra @ ta_star
    Rb = rmat(tb_star, nb)
    tb = Rb @ tb_star
    return [(Ra, ta, na), (Ra, -ta, -na), (Rb, tb, nb), (Rb, -tb, -nb)]


def recover_pose_homography(H, p1, p2, K, homography_distance):
    """recover_pose_homography (VO_utility.cpp:581-624): decompose, triangulate the matches for every candidate and
    keep the one with most points at 0 < z < HOMOGRAPHY_DISTANCE; t normalised.  INTENT parity: the reference reads
    the f32 depth through at<double> (undefined behaviour); this follows the evident intent.
    Returns (max_good_points, R, t) with R, t None when no candidate has a good point."""
    K = np.asarray(K, np.float64).reshape(3, 3)
    p1 = np.ascontiguousarray(p1, F32).reshape(-1, 2).astype(np.float64)
    p2 = np.ascontiguousarray(p2, F32).reshape(-1, 2).astype(np.float64)
    P0 = K @ np.eye(3, 4)
    best, best_good = -1, 0
    cands = decompose_homography_mat(H, K)
    for i, (R, t, _n) in enumerate(cands):
        X = _triangulate_h(P0, K @ np.c_[R, t], p1, p2).astype(F32)  # triangulatePoints returns f32 for f32 input
        with np.errstate(divide="ignore", invalid="ignore"):
            z = (X[2] / X[3]).astype(np.float64)  # convert_from_homogeneous_coords: plain f32 division
        good = int(((z > 0) & (z < homography_distance)).sum())
        if good > best_good:
            best, best_good = i, good
    if best < 0:
        return 0, None, None
    R, t, _ = cands[best]
    return best_good, R, t / np.linalg.norm(t)
