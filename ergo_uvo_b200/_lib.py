"""ctypes loader for the in-tree CUDA library (ergo_uvo_b200/libuvo_b200.so).

There is no CPU path: if the shared library is missing, importing anything that needs it raises, and creating a
context without a CUDA device fails with UVO_ERR_NO_DEVICE.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libuvo_b200.so")

UVO_OK = 0
UVO_ERR_NO_DEVICE = -1
UVO_ERR_CUDA = -2
UVO_ERR_INVALID = -3
UVO_ERR_CAPACITY = -4
UVO_ERR_UNSUPPORTED = -5
UVO_N_STAGES = 8


class UvoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"uvo error {code}: {msg}")
        self.code = code


class Camera(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "nfx", "nfy", "ncx", "ncy")]


class Params(C.Structure):
    _fields_ = [
        ("clahe", C.c_int32), ("clip_limit", C.c_int32), ("distance", C.c_int32), ("lowe_ratio", C.c_double),
        ("essential_method", C.c_int32), ("essential_max_iters", C.c_double), ("essential_confidence", C.c_double),
        ("essential_threshold", C.c_double), ("homography_method", C.c_int32), ("homography_max_iters", C.c_double),
        ("homography_confidence", C.c_double), ("homography_threshold", C.c_double),
        ("homography_distance", C.c_double), ("vpf_threshold", C.c_double), ("reprojection_tolerance", C.c_double),
        ("min_num_features", C.c_int32), ("min_num_3dpoints", C.c_int32), ("min_num_inliers", C.c_int32),
        ("iterations_count", C.c_int32), ("reprojection_error", C.c_double), ("confidence", C.c_double),
        ("pnp_method_flag", C.c_int32), ("surf_min_hessian", C.c_int32), ("surf_octaves", C.c_int32),
        ("surf_octave_layers", C.c_int32), ("surf_extended", C.c_int32), ("surf_upright", C.c_int32),
        ("max_features", C.c_int32),
        ("stereo_gate", C.c_int32), ("stereo_max_epipolar_dy", C.c_double), ("stereo_min_disparity", C.c_double),
        ("stereo_max_disparity", C.c_double),
    ]


class StereoResult(C.Structure):
    _fields_ = [
        ("initialised", C.c_int32), ("valid", C.c_int32), ("n_left", C.c_int32), ("n_right", C.c_int32),
        ("n_stereo_matches", C.c_int32), ("n_temporal_matches", C.c_int32), ("n_3d", C.c_int32),
        ("n_inliers", C.c_int32), ("hyps_evaluated", C.c_int32), ("gate", C.c_int32),
        ("rvec", C.c_double * 3), ("tvec", C.c_double * 3), ("t_prev_curr", C.c_double * 3),
        ("velocity", C.c_double * 3),
    ]


class MonoResult(C.Structure):
    _fields_ = [
        ("initialised", C.c_int32), ("skipped", C.c_int32), ("published", C.c_int32), ("valid", C.c_int32),
        ("used_essential", C.c_int32), ("n_keypoints", C.c_int32), ("n_matches", C.c_int32), ("n_inliers", C.c_int32),
        ("n_3d", C.c_int32), ("reserved", C.c_int32),
        ("R", C.c_double * 9), ("t", C.c_double * 3), ("scale_factor", C.c_double), ("velocity", C.c_double * 3),
    ]


class JpegLayout(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("components", C.c_int32),
        ("h_samp", C.c_int32 * 3), ("v_samp", C.c_int32 * 3), ("blocks_x", C.c_int32 * 3), ("blocks_y", C.c_int32 * 3),
        ("samples_x", C.c_int32 * 3), ("samples_y", C.c_int32 * 3),
        ("coeff_offset", C.c_int64 * 3), ("coeff_total", C.c_int64), ("quant", (C.c_uint16 * 64) * 3),
    ]


class JpegSparse(C.Structure):
    """uvo_jpeg_sparse: one compressed image as entropy-decoded sparse coefficients"""
    _fields_ = [("entries", C.c_void_p), ("n_entries", C.c_size_t), ("block_first", C.c_void_p),
                ("block_count", C.c_void_p), ("layout", JpegLayout)]


_lib = None


def load():
    """Load libuvo_b200.so (built by `make -C ergo_uvo_b200/csrc` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: build it with `make -C ergo_uvo_b200/csrc` (nvcc, sm_100a). "
            "ergo_uvo_b200 has no CPU fallback.")
    lib = C.CDLL(SO_PATH)
    lib.uvo_version.restype = C.c_char_p
    lib.uvo_last_error.restype = C.c_char_p
    lib.uvo_last_error.argtypes = [C.c_void_p]
    lib.uvo_ctx_create.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.uvo_ctx_destroy.argtypes = [C.c_void_p]
    lib.uvo_ctx_destroy.restype = None
    lib.uvo_ctx_stream.restype = C.c_void_p
    lib.uvo_ctx_stream.argtypes = [C.c_void_p]
    lib.uvo_ctx_synchronize.argtypes = [C.c_void_p]
    lib.uvo_ctx_launch_count.restype = C.c_int64
    lib.uvo_ctx_launch_count.argtypes = [C.c_void_p]
    lib.uvo_ctx_kernel_timing.argtypes = [C.c_void_p, C.c_int]
    lib.uvo_ctx_kernel_report.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    lib.uvo_default_params.argtypes = [C.c_int, C.POINTER(Params)]
    lib.uvo_default_params.restype = None
    lib.uvo_host_alloc.restype = C.c_void_p
    lib.uvo_host_alloc.argtypes = [C.c_size_t]
    lib.uvo_host_free.argtypes = [C.c_void_p]
    lib.uvo_host_free.restype = None
    if hasattr(lib, "uvo_stage_name"):
        lib.uvo_stage_name.restype = C.c_char_p
    _lib = lib
    return lib
