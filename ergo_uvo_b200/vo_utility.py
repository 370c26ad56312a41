"""Python mirror of the reference's `uvo_libraries` function API over the C ABI (include/uvo_c.h).

Function names and argument meaning follow uvo_libraries/include/uvo_libraries/VO_utility.h:96-117 so the parity
tests read like calls into the reference; numpy arrays stand in for cv::Mat / std::vector.  The configuration
globals of VO_utility.h:25-89 are the fields of `Params`.  Every function runs on the GPU through libuvo_b200.so;
nothing here computes on the CPU.
"""
import ctypes as C

import numpy as np

from . import _lib as L

KEYPOINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"),
     ("class_id", "<i4")])
DMATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])

__all__ = ["Context", "default_params", "make_camera", "resize_camera_matrix", "jpeg_info", "jpeg_entropy_decode", "jpeg_entropy_decode_sparse", "SparseImage", "KEYPOINT_DTYPE", "DMATCH_DTYPE", "StereoVO", "MonoVO"]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f64(v, n=None):
    a = np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(-1))
    if n is not None and a.size != n:
        raise ValueError(f"expected {n} values, got {a.size}")
    return a


def _k4(K):
    K = np.asarray(K, dtype=np.float64)
    if K.shape == (3, 3):
        return np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]], dtype=np.float64)
    return _f64(K, 4)


def default_params(stereo=True):
    p = L.Params()
    L.load().uvo_default_params(1 if stereo else 0, C.byref(p))
    return p


def make_camera(K, D, newK):
    """cameraMatrix (3x3), distortionCoeff (k1,k2,p1,p2), newCamMatrix (3x3) -> uvo_camera."""
    K = np.asarray(K, dtype=np.float64)
    newK = np.asarray(newK, dtype=np.float64)
    D = _f64(D, 4)
    return L.Camera(K[0, 0], K[1, 1], K[0, 2], K[1, 2], D[0], D[1], D[2], D[3], newK[0, 0], newK[1, 1], newK[0, 2],
                    newK[1, 2])


def resize_camera_matrix(original_width, original_height, desired_width, cameraMatrix, distortionCoeff):
    """resize_camera_matrix (VO_utility.h:112, VO_utility.cpp:658-675), host-only: returns (scaled cameraMatrix,
    newCamMatrix, (width, height)); getOptimalNewCameraMatrix(alpha=0) is restated in the library (camera.cu)."""
    lib = L.load()
    K = np.array(cameraMatrix, dtype=np.float64).reshape(3, 3).copy()
    D = _f64(distortionCoeff, 4)
    newK = np.zeros(9)
    ow, oh = C.c_int(0), C.c_int(0)
    rc = lib.uvo_resize_camera_matrix(int(original_width), int(original_height), int(desired_width), _p(K), _p(D),
                                      _p(newK), C.byref(ow), C.byref(oh))
    if rc != L.UVO_OK:
        raise L.UvoError(rc, "uvo_resize_camera_matrix: bad camera / size")
    return K, newK.reshape(3, 3), (ow.value, oh.value)


def jpeg_info(data):
    """header of a JPEG stream (uvo_jpeg_info, host-only): the uvo_jpeg_layout of the decode inside
    from_ros_to_cv_image (math_utility.cpp:154-173)"""
    buf = data if isinstance(data, np.ndarray) else np.frombuffer(bytes(data), np.uint8)
    lay = L.JpegLayout()
    rc = L.load().uvo_jpeg_info(_p(buf), C.c_size_t(len(buf)), C.byref(lay))
    if rc != L.UVO_OK:
        raise L.UvoError(rc, "uvo_jpeg_info: not a supported JPEG stream")
    return lay


def jpeg_entropy_decode(data):
    """host half of the JPEG decode (uvo_jpeg_entropy_decode): (layout, int16 quantised coefficients)"""
    buf = np.frombuffer(bytes(data), np.uint8)
    lay = jpeg_info(data)
    coef = np.empty(int(lay.coeff_total), np.int16)
    rc = L.load().uvo_jpeg_entropy_decode(_p(buf), C.c_size_t(len(buf)), _p(coef), C.c_size_t(len(coef)), C.byref(lay))
    if rc != L.UVO_OK:
        raise L.UvoError(rc, "uvo_jpeg_entropy_decode: corrupt or unsupported JPEG stream")
    return lay, coef


def jpeg_entropy_decode_sparse(data):
    """uvo_jpeg_entropy_decode_sparse: (layout, entries u32, block_first u32, block_count u8) -- the form that travels
    to the GPU: one entry per non-zero coefficient"""
    buf = np.frombuffer(bytes(data), np.uint8)
    lay = jpeg_info(data)
    nb = int(lay.coeff_total) // 64
    entries = np.empty(int(lay.coeff_total), np.uint32)
    first = np.empty(nb, np.uint32)
    count = np.empty(nb, np.uint8)
    n = C.c_size_t(0)
    rc = L.load().uvo_jpeg_entropy_decode_sparse(_p(buf), C.c_size_t(len(buf)), _p(entries), C.c_size_t(len(entries)),
                                                 _p(first), _p(count), C.byref(n), C.byref(lay))
    if rc != L.UVO_OK:
        raise L.UvoError(rc, "uvo_jpeg_entropy_decode_sparse: corrupt or unsupported JPEG stream")
    return lay, entries[:n.value].copy(), first, count


def jpeg_gpu_plan(data):
    """uvo_jpeg_gpu_plan (host-only): the host half of the GPU Huffman route.  Returns None when the stream takes the
    host decoder, else (layout, scan bits, upload bytes, staging bytes as a numpy array: [table plan][unstuffed scan])."""
    lib = L.load()
    lib.uvo_jpeg_gpu_staging_bytes.restype = C.c_size_t
    buf = np.frombuffer(bytes(data), np.uint8)
    staging = np.zeros(int(lib.uvo_jpeg_gpu_staging_bytes(C.c_size_t(len(buf)))), np.uint8)
    ok, up, bits, lay = C.c_int(0), C.c_size_t(0), C.c_uint32(0), L.JpegLayout()
    rc = lib.uvo_jpeg_gpu_plan(_p(buf), C.c_size_t(len(buf)), _p(staging), C.c_size_t(len(staging)), C.byref(ok),
                               C.byref(up), C.byref(bits), C.byref(lay))
    if rc != L.UVO_OK:
        raise L.UvoError(rc, "uvo_jpeg_gpu_plan: corrupt or unsupported JPEG stream")
    if not ok.value:
        return None
    return lay, int(bits.value), int(up.value), staging


class SparseImage:
    """A compressed image after the host half of the decode (uvo_jpeg_entropy_decode_sparse), in buffers that stay
    alive with the object -- pinned host memory when `pinned` (uvo_host_alloc), so that the upload inside
    uvo_stereo_enqueue_host_sparse is asynchronous.  Building one touches no GPU state and releases the GIL for the
    duration of the Huffman decode: several can be built in parallel from Python threads."""

    def __init__(self, data=None, pinned=True):
        self._pinned = None
        self._base = None
        self._use_pinned = pinned
        self._lib = L.load()
        if data is not None:
            self.decode(data)

    def decode(self, data):
        """(re)fill this object from a JPEG stream; the buffers are reused when they are large enough, so a ring of
        SparseImage objects decodes frame after frame without allocating"""
        lib = self._lib
        buf = data if isinstance(data, np.ndarray) else np.frombuffer(bytes(data), np.uint8)
        lay = jpeg_info(buf)
        nb, total = int(lay.coeff_total) // 64, int(lay.coeff_total)
        words = nb + total + (nb + 3) // 4
        if self._base is None or len(self._base) < words:
            self.close()
            if self._use_pinned:
                ptr = lib.uvo_host_alloc(C.c_size_t(4 * words))
                if not ptr:
                    raise MemoryError("uvo_host_alloc failed")
                self._pinned = ptr
                self._base = np.ctypeslib.as_array((C.c_uint32 * words).from_address(ptr))
            else:
                self._base = np.empty(words, np.uint32)
        base = self._base
        first, entries = base[:nb], base[nb:nb + total]
        count = base[nb + total:].view(np.uint8)[:nb]
        n = C.c_size_t(0)
        rc = lib.uvo_jpeg_entropy_decode_sparse(_p(buf), C.c_size_t(len(buf)), _p(entries), C.c_size_t(total),
                                                _p(first), _p(count), C.byref(n), C.byref(lay))
        if rc != L.UVO_OK:
            self.close()
            raise L.UvoError(rc, "uvo_jpeg_entropy_decode_sparse: corrupt or unsupported JPEG stream")
        self.layout, self.n_entries = lay, int(n.value)
        self.c = L.JpegSparse(entries.ctypes.data, n.value, first.ctypes.data, count.ctypes.data, lay)
        self.nbytes = 4 * nb + 4 * self.n_entries + nb  # what travels to the GPU
        return self

    def close(self):
        self._base = None
        if self._pinned:
            self._lib.uvo_host_free(C.c_void_p(self._pinned))
            self._pinned = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One uvo_ctx per (GPU, stream)."""

    def __init__(self, device=0, stream=None):
        self.lib = L.load()
        h = C.c_void_p()
        rc = self.lib.uvo_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != L.UVO_OK:
            raise L.UvoError(rc, "uvo_ctx_create failed (no CUDA device? there is no CPU fallback)")
        self.h = h
        self.params = default_params(True)

    def close(self):
        if getattr(self, "h", None):
            self.lib.uvo_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != L.UVO_OK:
            raise L.UvoError(rc, self.lib.uvo_last_error(self.h).decode())

    @property
    def stream(self):
        return self.lib.uvo_ctx_stream(self.h)

    @property
    def launches(self):
        return int(self.lib.uvo_ctx_launch_count(self.h))

    def kernel_timing(self, enable):
        self._ck(self.lib.uvo_ctx_kernel_timing(self.h, int(enable)))

    def kernel_report(self):
        """{kernel name: (launches, total_ms)} since timing was enabled / last report"""
        buf = C.create_string_buffer(1 << 16)
        self._ck(self.lib.uvo_ctx_kernel_report(self.h, buf, C.c_size_t(len(buf))))
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.split()
            out[name] = (int(cnt), float(ms))
        return out

    def synchronize(self):
        self._ck(self.lib.uvo_ctx_synchronize(self.h))

    # ------------------------------------------------------------------ VO_utility.h:105
    def get_image(self, current_img, cameraMatrix, distortionCoeff, newCamMatrix):
        """Mat get_image(current_img, cameraMatrix, distortionCoeff, newCamMatrix); CLAHE_CORRECTION / CLIP_LIMIT are
        read from self.params like the reference reads its globals (VO_utility.cpp:352-357)."""
        img = np.ascontiguousarray(current_img, dtype=np.uint8)
        if img.ndim != 3 or img.shape[2] != 3:
            raise ValueError("get_image expects a 3-channel image (cvtColor RGB2GRAY throws otherwise)")
        h, w, _ = img.shape
        cam = make_camera(cameraMatrix, distortionCoeff, newCamMatrix)
        out = np.empty((h, w), np.uint8)
        self._ck(self.lib.uvo_get_image(self.h, _p(img), w, h, C.c_size_t(3 * w), C.byref(cam),
                                        int(self.params.clahe), int(self.params.clip_limit), _p(out),
                                        C.c_size_t(w)))
        return out

    def demosaic_bggr2bgr(self, bayer):
        """cv::cvtColor(bayer, COLOR_BayerBGGR2BGR) as from_ros_to_cv_image applies it (math_utility.cpp:161-164)"""
        b = np.ascontiguousarray(bayer, np.uint8)
        h, w = b.shape
        out = np.empty((h, w, 3), np.uint8)
        self._ck(self.lib.uvo_demosaic_bggr2bgr(self.h, _p(b), w, h, C.c_size_t(b.strides[0]), _p(out),
                                                C.c_size_t(out.strides[0])))
        return out

    def resize_area(self, img, dw, dh):
        """cv::resize(img, (dw, dh), interpolation=INTER_AREA), u8, 1 or 3 channels"""
        img = np.ascontiguousarray(img, np.uint8)
        cn = 1 if img.ndim == 2 else img.shape[2]
        out = np.empty((dh, dw) if cn == 1 else (dh, dw, cn), np.uint8)
        self._ck(self.lib.uvo_resize_area(self.h, _p(img), img.shape[1], img.shape[0], C.c_size_t(img.strides[0]), cn,
                                          _p(out), dw, dh, C.c_size_t(out.strides[0])))
        return out

    def get_image_resized(self, current_img, desired_width, cameraMatrix, distortionCoeff, newCamMatrix):
        """get_image with DESIRED_WIDTH != image width (VO_utility.cpp:339-376); the camera is the resized one"""
        img = np.ascontiguousarray(current_img, np.uint8)
        h, w = img.shape[:2]
        dh = int(h / (w / desired_width))
        out = np.empty((dh, desired_width), np.uint8)
        cam = make_camera(cameraMatrix, distortionCoeff, newCamMatrix)
        ow, oh = C.c_int(0), C.c_int(0)
        self._ck(self.lib.uvo_get_image_resized(self.h, _p(img), w, h, C.c_size_t(img.strides[0]), int(desired_width),
                                                C.byref(cam), int(self.params.clahe), int(self.params.clip_limit),
                                                _p(out), C.c_size_t(out.strides[0]), C.byref(ow), C.byref(oh)))
        assert (ow.value, oh.value) == (desired_width, dh)
        return out

    def integral(self, gray):
        g = np.ascontiguousarray(gray, dtype=np.uint8)
        h, w = g.shape
        out = np.empty((h + 1, w + 1), np.int32)
        self._ck(self.lib.uvo_integral(self.h, _p(g), w, h, C.c_size_t(w), _p(out)))
        return out

    # ------------------------------------------------------------------ math_utility.h:27
    def jpeg_decode(self, data, bayer=False):
        """the cv::imdecode(IMREAD_UNCHANGED) inside from_ros_to_cv_image: h x w (1 component) or h x w x 3 BGR;
        bayer=True: a 1-component stream is the BGGR mosaic and comes back demosaiced (math_utility.cpp:161-164)"""
        buf = np.frombuffer(bytes(data), np.uint8)
        lay = jpeg_info(data)
        ch = 1 if (lay.components == 1 and not bayer) else 3
        out = np.empty((lay.height, lay.width) if ch == 1 else (lay.height, lay.width, 3), np.uint8)
        w, h, c = C.c_int(0), C.c_int(0), C.c_int(0)
        self._ck(self.lib.uvo_jpeg_decode(self.h, _p(buf), C.c_size_t(len(buf)), int(bool(bayer)), _p(out),
                                          C.c_size_t(lay.width * ch),
                                          C.c_size_t(out.nbytes), C.byref(w), C.byref(h), C.byref(c)))
        return out

    def jpeg_gpu_entropy(self, enable=None):
        """switch the GPU Huffman decoder on / off (None: leave); returns (route of the last jpeg_decode: 1 = GPU
        entropy decoder, synchronisation rounds it took)"""
        route, rounds = C.c_int(0), C.c_int(0)
        self._ck(self.lib.uvo_jpeg_gpu_entropy(self.h, -1 if enable is None else int(bool(enable)), C.byref(route),
                                               C.byref(rounds)))
        return route.value, rounds.value

    def jpeg_decode_device(self, data, out_ptr, out_pitch, out_capacity, bayer=False):
        """as jpeg_decode with the image left in device memory at out_ptr (ordered on the context stream);
        returns (width, height, channels)"""
        buf = np.frombuffer(bytes(data), np.uint8)
        w, h, c = C.c_int(0), C.c_int(0), C.c_int(0)
        self._ck(self.lib.uvo_jpeg_decode_device(self.h, _p(buf), C.c_size_t(len(buf)), int(bool(bayer)),
                                                 C.c_void_p(out_ptr),
                                                 C.c_size_t(out_pitch), C.c_size_t(out_capacity), C.byref(w),
                                                 C.byref(h), C.byref(c)))
        return w.value, h.value, c.value

    # ------------------------------------------------------------------ VO_utility.h:100
    def detect_features(self, img, capacity=None):
        """void detect_features(Mat img, vector<KeyPoint>&, Mat&) -> (keypoints, descriptors)."""
        g = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = g.shape
        cap = int(capacity or self.params.max_features)
        dim = 128 if self.params.surf_extended else 64
        kps = np.zeros(cap, KEYPOINT_DTYPE)
        desc = np.zeros((cap, dim), np.float32)
        n = C.c_int(0)
        self._ck(self.lib.uvo_detect_features(self.h, _p(g), w, h, C.c_size_t(w), C.byref(self.params), _p(kps),
                                              _p(desc), cap, C.byref(n)))
        return kps[:n.value].copy(), desc[:n.value].copy()

    def sort_keypoints(self, keypoints, capacity=None):
        """K6 parity tap: keypoints in SURF's final order (uvo_sort_keypoints)"""
        k = np.ascontiguousarray(keypoints, dtype=KEYPOINT_DTYPE)
        out = np.zeros(max(len(k), 1), KEYPOINT_DTYPE)
        self._ck(self.lib.uvo_sort_keypoints(self.h, _p(k), len(k), int(capacity or max(len(k), 64)), _p(out)))
        return out[:len(k)]

    # ------------------------------------------------------------------ VO_utility.h:109-110
    def match_features(self, keypoints1, keypoints2, descriptors1, descriptors2, with_points=False, gate=None):
        """5-arg overload returns matches; with_points=True is the 7-arg overload (also keypoints*_conv).
        gate=(max_dy, min_disparity, max_disparity) adds the stereo epipolar / disparity gate (not in the reference,
        off by default; uvo_match_features_gated)."""
        d1 = np.ascontiguousarray(descriptors1, dtype=np.float32)
        d2 = np.ascontiguousarray(descriptors2, dtype=np.float32)
        n1, n2 = d1.shape[0], d2.shape[0]
        dim = d1.shape[1] if d1.ndim == 2 else 64
        out = np.zeros(max(n1, 1), DMATCH_DTYPE)
        n = C.c_int(0)
        if gate is not None:
            k1 = np.ascontiguousarray(keypoints1, dtype=KEYPOINT_DTYPE)
            k2 = np.ascontiguousarray(keypoints2, dtype=KEYPOINT_DTYPE)
            self._ck(self.lib.uvo_match_features_gated(
                self.h, _p(k1), _p(d1), n1, _p(k2), _p(d2), n2, dim, C.c_float(self.params.lowe_ratio),
                C.c_float(gate[0]), C.c_float(gate[1]), C.c_float(gate[2]), _p(out), C.byref(n)))
        else:
            self._ck(self.lib.uvo_match_features(self.h, _p(d1), n1, _p(d2), n2, dim,
                                                 C.c_float(self.params.lowe_ratio), _p(out), C.byref(n)))
        m = out[:n.value].copy()
        if not with_points:
            return m
        p1 = np.stack([keypoints1["x"][m["queryIdx"]], keypoints1["y"][m["queryIdx"]]], -1).astype(np.float32)
        p2 = np.stack([keypoints2["x"][m["trainIdx"]], keypoints2["y"][m["trainIdx"]]], -1).astype(np.float32)
        return m, p1, p2

    def knn_match2(self, descriptors1, descriptors2):
        d1 = np.ascontiguousarray(descriptors1, dtype=np.float32)
        d2 = np.ascontiguousarray(descriptors2, dtype=np.float32)
        out = np.zeros((d1.shape[0], 2), DMATCH_DTYPE)
        self._ck(self.lib.uvo_knn_match2(self.h, _p(d1), d1.shape[0], _p(d2), d2.shape[0], d1.shape[1], _p(out)))
        return out

    def match_exact_only(self, enable):
        """diagnostics: route every query of the stage-level matcher calls through the exact full scan"""
        self._ck(self.lib.uvo_match_exact_only(self.h, int(bool(enable))))

    def match_last_fallbacks(self):
        n = C.c_int(0)
        self._ck(self.lib.uvo_match_last_fallbacks(self.h, C.byref(n)))
        return n.value

    # ------------------------------------------------------------------ K10a / K10b (VO_utility.cpp:134-180, :581-624)
    @staticmethod
    def _pts(p):
        return np.ascontiguousarray(p, np.float32).reshape(-1, 2)

    def findHomography(self, srcPoints, dstPoints, method=8, ransacReprojThreshold=3.0, maxIters=2000,
                       confidence=0.995):
        """cv::findHomography -> (H or None, mask uint8 n, hypotheses evaluated)"""
        p1, p2 = self._pts(srcPoints), self._pts(dstPoints)
        n = len(p1)
        H = np.zeros(9)
        mask = np.zeros(max(n, 1), np.uint8)
        ni, hy, ok = C.c_int(0), C.c_int(0), C.c_int(0)
        self._ck(self.lib.uvo_find_homography(self.h, _p(p1), _p(p2), n, int(method), C.c_double(ransacReprojThreshold),
                                              int(maxIters), C.c_double(confidence), _p(H), _p(mask), C.byref(ni),
                                              C.byref(hy), C.byref(ok)))
        return (H.reshape(3, 3) if ok.value else None), mask[:n], hy.value

    def findEssentialMat(self, points1, points2, cameraMatrix, method=8, prob=0.999, threshold=1.0, maxIters=1000):
        """cv::findEssentialMat -> (E or None, mask uint8 n, hypotheses evaluated)"""
        p1, p2 = self._pts(points1), self._pts(points2)
        n = len(p1)
        E = np.zeros(9)
        mask = np.zeros(max(n, 1), np.uint8)
        ni, hy, ok = C.c_int(0), C.c_int(0), C.c_int(0)
        self._ck(self.lib.uvo_find_essential_mat(self.h, _p(p1), _p(p2), n, _p(_k4(cameraMatrix)), int(method),
                                                 C.c_double(prob), C.c_double(threshold), int(maxIters), _p(E),
                                                 _p(mask), C.byref(ni), C.byref(hy), C.byref(ok)))
        return (E.reshape(3, 3) if ok.value else None), mask[:n], hy.value

    def recoverPose(self, E, points1, points2, cameraMatrix, mask=None):
        """cv::recoverPose -> (good, R, t, mask)"""
        p1, p2 = self._pts(points1), self._pts(points2)
        n = len(p1)
        E = np.ascontiguousarray(E, np.float64).reshape(9)
        R, t = np.zeros(9), np.zeros(3)
        m = np.ascontiguousarray(mask, np.uint8).reshape(-1).copy() if mask is not None else np.ones(max(n, 1), np.uint8)
        good = C.c_int(0)
        self._ck(self.lib.uvo_recover_pose(self.h, _p(E), _p(p1), _p(p2), n, _p(_k4(cameraMatrix)), _p(m), _p(R),
                                           _p(t), C.byref(good)))
        return good.value, R.reshape(3, 3), t, m[:n]

    def recover_pose_homography(self, H, inliers1, inliers2, cameraMatrix):
        """VO_utility.cpp:581-624 -> (max_good_points, R, t) with R, t None when no candidate has a good point"""
        p1, p2 = self._pts(inliers1), self._pts(inliers2)
        H = np.ascontiguousarray(H, np.float64).reshape(9)
        R, t = np.zeros(9), np.zeros(3)
        good, found = C.c_int(0), C.c_int(0)
        self._ck(self.lib.uvo_recover_pose_homography(self.h, _p(H), _p(p1), _p(p2), len(p1), _p(_k4(cameraMatrix)),
                                                      C.c_double(self.params.homography_distance), _p(R), _p(t),
                                                      C.byref(good), C.byref(found)))
        if not found.value:
            return good.value, None, None
        return good.value, R.reshape(3, 3), t

    def estimate_relative_pose(self, keypoints1_conv, keypoints2_conv, cameraMatrix, use_essential):
        """VO_utility.cpp:134-180 -> (success, R, t, inlier mask of extract_inliers, use_essential after the call)"""
        p1, p2 = self._pts(keypoints1_conv), self._pts(keypoints2_conv)
        n = len(p1)
        R, t = np.zeros(9), np.zeros(3)
        mask = np.zeros(max(n, 1), np.uint8)
        ue, ni, ok = C.c_int(1 if use_essential else 0), C.c_int(0), C.c_int(0)
        self._ck(self.lib.uvo_estimate_relative_pose(self.h, _p(p1), _p(p2), n, _p(_k4(cameraMatrix)),
                                                     C.byref(self.params), C.byref(ue), _p(R), _p(t), _p(mask),
                                                     C.byref(ni), C.byref(ok)))
        return bool(ok.value), R.reshape(3, 3), t, mask[:n], bool(ue.value)

    # ------------------------------------------------------------------ VO_utility.h:116
    def select_estimation_method(self, keypoints1_conv, keypoints2_conv):
        p1 = np.ascontiguousarray(keypoints1_conv, np.float32).reshape(-1, 2)
        p2 = np.ascontiguousarray(keypoints2_conv, np.float32).reshape(-1, 2)
        r = C.c_int(0)
        self._ck(self.lib.uvo_select_estimation_method(self.h, _p(p1), _p(p2), p1.shape[0],
                                                       int(self.params.distance), C.byref(r)))
        return bool(r.value)

    # ------------------------------------------------------------------ node-direct cv:: calls
    def triangulatePoints(self, P1, P2, pts1, pts2):
        P1 = _f64(P1, 12)
        P2 = _f64(P2, 12)
        a = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
        b = np.ascontiguousarray(pts2, np.float32).reshape(-1, 2)
        out = np.empty((4, a.shape[0]), np.float32)
        self._ck(self.lib.uvo_triangulate_points(self.h, _p(P1), _p(P2), _p(a), _p(b), a.shape[0], _p(out)))
        return out

    def solvePnPRansac(self, objectPoints, imagePoints, cameraMatrix, iterationsCount=None, reprojectionError=None,
                       confidence=None):
        X = np.ascontiguousarray(objectPoints, np.float64).reshape(-1, 3)
        x = np.ascontiguousarray(imagePoints, np.float32).reshape(-1, 2)
        n = X.shape[0]
        rvec = np.zeros(3)
        tvec = np.zeros(3)
        inl = np.empty(max(n, 1), np.int32)
        ni = C.c_int(0)
        hyp = C.c_int(0)
        p = self.params
        self._ck(self.lib.uvo_solve_pnp_ransac(
            self.h, _p(X), _p(x), n, _p(_k4(cameraMatrix)), int(iterationsCount or p.iterations_count),
            C.c_float(reprojectionError if reprojectionError is not None else p.reprojection_error),
            C.c_double(confidence if confidence is not None else p.confidence), _p(rvec), _p(tvec), _p(inl),
            C.byref(ni), C.byref(hyp)))
        return ni.value > 0, rvec, tvec, inl[:ni.value].copy(), hyp.value

    def pnp_profile(self, enable=True):
        """diagnostics: SM clock stamps of the previous solvePnPRansac call (uvo_pnp_profile)"""
        st = np.zeros(32, np.int64)
        self._ck(self.lib.uvo_pnp_profile(self.h, int(bool(enable)), _p(st)))
        return st

    # ------------------------------------------------------------------ VO_utility.h:102
    def extract_3Dpoints(self, keypoints1_conv, keypoints2_conv, R1, t1, R2, t2, cameraMatrix1, cameraMatrix2,
                         points4D):
        a = np.ascontiguousarray(keypoints1_conv, np.float32).reshape(-1, 2)
        b = np.ascontiguousarray(keypoints2_conv, np.float32).reshape(-1, 2)
        n = a.shape[0]
        p4 = np.ascontiguousarray(points4D, np.float32)
        pts = np.empty((max(n, 1), 3), np.float64)
        idx = np.empty(max(n, 1), np.int32)
        m = C.c_int(0)
        p = self.params
        self._ck(self.lib.uvo_extract_3dpoints(self.h, _p(a), _p(b), n, _p(_f64(R1, 9)), _p(_f64(t1, 3)),
                                               _p(_f64(R2, 9)), _p(_f64(t2, 3)), _p(_k4(cameraMatrix1)),
                                               _p(_k4(cameraMatrix2)), _p(p4), C.c_double(p.reprojection_tolerance),
                                               int(p.min_num_3dpoints), _p(pts), _p(idx), C.byref(m)))
        return pts[:m.value].copy(), idx[:m.value].copy()

    # ------------------------------------------------------------------ VO_utility.h:97-98
    def compute_scale_factor(self, distance, good_prevCam_points, R, t, with_count=False):
        """convert_3Dpoints_camera + compute_scale_factor (visual_odometry.h:365-368); with_count also returns the
        number of points convert_3Dpoints_camera kept (uvo_scale_factor_front)."""
        pts = np.ascontiguousarray(good_prevCam_points, np.float64).reshape(-1, 3)
        sf = C.c_double(0)
        m = C.c_int(0)
        self._ck(self.lib.uvo_scale_factor_front(self.h, _p(pts), pts.shape[0], _p(_f64(R, 9)), _p(_f64(t, 3)),
                                                 C.c_float(distance), C.byref(sf), C.byref(m)))
        return (sf.value, m.value) if with_count else sf.value


class StereoVO:
    """Device-resident replay of visual_odometry_node::stereo_VO (visual_odometry.h:406-741)."""

    def __init__(self, ctx, width, height, cam_left, cam_right, R_right, t_right, params=None):
        self.ctx = ctx
        self.lib = ctx.lib
        self.params = params or ctx.params
        h = C.c_void_p()
        ctx._ck(self.lib.uvo_stereo_create(ctx.h, width, height, C.byref(cam_left), C.byref(cam_right),
                                           _p(_f64(R_right, 9)), _p(_f64(t_right, 3)), C.byref(self.params),
                                           C.byref(h)))
        self.h = h
        self.width, self.height = width, height

    def close(self):
        if getattr(self, "h", None):
            self.lib.uvo_stereo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def frame(self, left3, right3, dt):
        """Host images (h x w x 3 u8) in, StereoResult out."""
        l = np.ascontiguousarray(left3, np.uint8)
        r = np.ascontiguousarray(right3, np.uint8)
        res = L.StereoResult()
        self.ctx._ck(self.lib.uvo_stereo_frame(self.h, _p(l), _p(r), C.c_size_t(3 * self.width), C.c_double(dt),
                                               C.byref(res)))
        return res

    def frame_device(self, left_ptr, right_ptr, pitch, dt):
        res = L.StereoResult()
        self.ctx._ck(self.lib.uvo_stereo_frame_device(self.h, C.c_void_p(left_ptr), C.c_void_p(right_ptr),
                                                      C.c_size_t(pitch), C.c_double(dt), C.byref(res)))
        return res

    def enqueue_device(self, left_ptr, right_ptr, pitch, dt):
        self.ctx._ck(self.lib.uvo_stereo_enqueue_device(self.h, C.c_void_p(left_ptr), C.c_void_p(right_ptr),
                                                        C.c_size_t(pitch), C.c_double(dt)))

    def enqueue_host(self, left_ptr, right_ptr, pitch, dt):
        """host (pinned) image pointers; the H2D copies are enqueued with the frame"""
        self.ctx._ck(self.lib.uvo_stereo_enqueue_host(self.h, C.c_void_p(left_ptr), C.c_void_p(right_ptr),
                                                      C.c_size_t(pitch), C.c_double(dt)))

    def enqueue_host_bayer(self, left_ptr, right_ptr, pitch, dt):
        """host pointers to 1-channel BGGR bayer images (demosaiced on the device, uvo_stereo_enqueue_host_bayer)"""
        self.ctx._ck(self.lib.uvo_stereo_enqueue_host_bayer(self.h, C.c_void_p(left_ptr), C.c_void_p(right_ptr),
                                                            C.c_size_t(pitch), C.c_double(dt)))

    def enqueue_host_jpeg(self, left_jpeg, right_jpeg, dt, bayer=False):
        """both images as JPEG byte strings (uvo_stereo_enqueue_host_jpeg: Huffman decoding inside, on two threads)"""
        l = left_jpeg if isinstance(left_jpeg, np.ndarray) else np.frombuffer(bytes(left_jpeg), np.uint8)
        r = right_jpeg if isinstance(right_jpeg, np.ndarray) else np.frombuffer(bytes(right_jpeg), np.uint8)
        self.ctx._ck(self.lib.uvo_stereo_enqueue_host_jpeg(self.h, _p(l), C.c_size_t(len(l)), _p(r), C.c_size_t(len(r)),
                                                           int(bool(bayer)), C.c_double(dt)))

    def set_gpu_entropy(self, enable):
        """Huffman decoding of enqueue_host_jpeg's input on the GPU (default) or on the host"""
        self.ctx._ck(self.lib.uvo_stereo_set_gpu_entropy(self.h, int(bool(enable))))

    @property
    def gpu_entropy_frames(self):
        self.lib.uvo_stereo_gpu_entropy_frames.restype = C.c_int64
        self.lib.uvo_stereo_gpu_entropy_frames.argtypes = [C.c_void_p]
        return int(self.lib.uvo_stereo_gpu_entropy_frames(self.h))

    def enqueue_host_sparse(self, left, right, dt, bayer=False):
        """both images as SparseImage objects (uvo_stereo_enqueue_host_sparse); they must stay alive until the frame
        has been collected"""
        self.ctx._ck(self.lib.uvo_stereo_enqueue_host_sparse(self.h, C.byref(left.c), C.byref(right.c),
                                                             int(bool(bayer)), C.c_double(dt)))

    def max_in_flight(self):
        return int(self.lib.uvo_stereo_max_in_flight())

    def lanes(self):
        return int(self.lib.uvo_stereo_lanes())

    def set_graphs(self, enable):
        """replay each lane's fixed kernel runs as CUDA graphs (default) or launch every kernel directly"""
        self.ctx._ck(self.lib.uvo_stereo_set_graphs(self.h, int(bool(enable))))

    @property
    def graph_launches(self):
        self.lib.uvo_stereo_graph_launches.restype = C.c_int64
        self.lib.uvo_stereo_graph_launches.argtypes = [C.c_void_p]
        return int(self.lib.uvo_stereo_graph_launches(self.h))

    def collect(self):
        res = L.StereoResult()
        self.ctx._ck(self.lib.uvo_stereo_collect(self.h, C.byref(res)))
        return res

    def last_keypoints(self, right=False):
        cap = int(self.params.max_features)
        kps = np.zeros(cap, KEYPOINT_DTYPE)
        desc = np.zeros((cap, 128 if self.params.surf_extended else 64), np.float32)
        n = C.c_int(0)
        self.ctx._ck(self.lib.uvo_stereo_last_keypoints(self.h, int(right), _p(kps), _p(desc), cap, C.byref(n)))
        return kps[:n.value].copy(), desc[:n.value].copy()

    def last_matches(self, temporal=False):
        cap = int(self.params.max_features)
        m = np.zeros(cap, DMATCH_DTYPE)
        n = C.c_int(0)
        self.ctx._ck(self.lib.uvo_stereo_last_matches(self.h, int(temporal), _p(m), cap, C.byref(n)))
        return m[:n.value].copy()

    def last_inliers(self):
        cap = int(self.params.max_features)
        a = np.zeros(cap, np.int32)
        n = C.c_int(0)
        self.ctx._ck(self.lib.uvo_stereo_last_inliers(self.h, _p(a), cap, C.byref(n)))
        return a[:n.value].copy()

    def stage_ms(self):
        ms = (C.c_float * L.UVO_N_STAGES)()
        self.ctx._ck(self.lib.uvo_stereo_stage_ms(self.h, ms))
        names = [self.lib.uvo_stage_name(i).decode() for i in range(L.UVO_N_STAGES)]
        return dict(zip(names, list(ms)))


class MonoVO:
    """visual_odometry_node::mono_VO's per-frame body (visual_odometry.h:247-397) behind uvo_mono."""

    def __init__(self, ctx, width, height, cam, params=None):
        self.ctx = ctx
        self.lib = ctx.lib
        self.params = params or default_params(False)
        h = C.c_void_p()
        ctx._ck(self.lib.uvo_mono_create(ctx.h, width, height, C.byref(cam), C.byref(self.params), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.uvo_mono_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def frame(self, img3, dt, distance):
        """Host image (h x w x 3 u8) + altimeter range in, MonoResult out."""
        img = np.ascontiguousarray(img3, np.uint8)
        res = L.MonoResult()
        self.ctx._ck(self.lib.uvo_mono_frame(self.h, _p(img), C.c_size_t(img.strides[0]), C.c_double(dt),
                                             C.c_double(distance), C.byref(res)))
        return res
