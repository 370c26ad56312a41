"""CPU: the C-ABI shared library builds for sm_100a, loads without a GPU, exports every symbol include/uvo_c.h
declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "uvo_c.h")).read()
    return sorted(set(re.findall(r"UVO_API\s+[\w\s\*]+?\b(uvo_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import ergo_uvo_b200 as U
    lib = U.load()
    names = _declared()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_only_declared_symbols_are_exported():
    import ergo_uvo_b200 as U
    out = subprocess.check_output(["nm", "-D", "--defined-only", U.SO_PATH], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert {e for e in exported if e.startswith("uvo_")} == set(_declared())


def test_sass_is_sm100a_only():
    import ergo_uvo_b200 as U
    out = subprocess.run(["cuobjdump", "-lelf", U.SO_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_struct_layouts_match_header():
    import ergo_uvo_b200 as U
    import numpy as np
    assert U.KEYPOINT_DTYPE.itemsize == 28 and U.DMATCH_DTYPE.itemsize == 16  # cv::KeyPoint / cv::DMatch
    assert C.sizeof(U.Camera) == 96
    p = U.default_params(True)
    assert (p.clip_limit, p.lowe_ratio, p.reprojection_tolerance, p.surf_min_hessian) == (8, 0.8, 3.0, 1500)
    assert p.stereo_gate == 0 and U.default_params(False).stereo_gate == 0  # the reference has no stereo gate
    m = U.default_params(False)
    assert (m.clip_limit, m.lowe_ratio, m.reprojection_tolerance, m.surf_min_hessian, m.min_num_features) == \
        (3, 0.7, 0.1, 50, 20)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import ergo_uvo_b200 as U
    with pytest.raises(U.UvoError) as e:
        U.Context(0)
    assert e.value.code == -1  # UVO_ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ergo_uvo_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f == "surf.cu" or "the oracle" in txt.lower(), f
                assert "import oracle" not in txt and "from oracle" not in txt and "uvo_oracle" not in txt, f


def test_optimal_new_camera_matrix_matches_cv2():
    """uvo_optimal_new_camera_matrix / uvo_resize_camera_matrix (host-only, VO_utility.cpp:658-675) against
    cv2.getOptimalNewCameraMatrix(alpha=0): bit-exact on random 4-coefficient cameras at every BASELINE resolution."""
    import numpy as np
    cv2 = pytest.importorskip("cv2")
    import ergo_uvo_b200 as U
    lib = U.load()
    rs = np.random.RandomState(0)
    for (w, h) in [(640, 480), (1280, 1024), (1920, 1080), (2448, 2048), (641, 513)]:
        for _ in range(10):
            f = w * rs.uniform(0.7, 1.3)
            K = np.array([[f, 0, w / 2 + rs.uniform(-20, 20)], [0, f * rs.uniform(0.95, 1.05), h / 2 + rs.uniform(-20, 20)],
                          [0, 0, 1.0]])
            D = np.array([rs.uniform(-0.3, 0.1), rs.uniform(-0.05, 0.1), rs.uniform(-2e-3, 2e-3), rs.uniform(-2e-3, 2e-3)])
            ref, _ = cv2.getOptimalNewCameraMatrix(K, D, (w, h), 0, (w, h), False)
            out = np.zeros(9)
            assert lib.uvo_optimal_new_camera_matrix(K.ctypes.data_as(C.c_void_p), D.ctypes.data_as(C.c_void_p), w, h,
                                                     out.ctypes.data_as(C.c_void_p)) == 0
            assert np.array_equal(out.reshape(3, 3), ref)
    # resize_camera_matrix: 2560x2048 camera scaled to DESIRED_WIDTH 640 (the shipped mono set-up)
    K = np.array([[2400.0, 0.5, 1275.0], [0, 2410.0, 1030.0], [0, 0, 1.0]])
    D = np.array([-0.2, 0.05, 1e-3, -5e-4])
    Kio, newK = K.copy(), np.zeros(9)
    ow, oh = C.c_int(0), C.c_int(0)
    assert lib.uvo_resize_camera_matrix(2560, 2048, 640, Kio.ctypes.data_as(C.c_void_p), D.ctypes.data_as(C.c_void_p),
                                        newK.ctypes.data_as(C.c_void_p), C.byref(ow), C.byref(oh)) == 0
    assert (ow.value, oh.value) == (640, 512)
    Ks = K * (1.0 / 4.0)
    Ks[0, 1], Ks[2, 2] = 0.5, 1.0
    assert np.array_equal(Kio, Ks)
    ref, _ = cv2.getOptimalNewCameraMatrix(Ks, D, (640, 512), 0, (640, 512), False)
    assert np.array_equal(newK.reshape(3, 3), ref)


def test_ctypes_mirrors_have_the_header_sizes(tmp_path):
    """the ctypes mirrors in ergo_uvo_b200/_lib.py against sizeof / offsetof from include/uvo_c.h compiled with gcc"""
    import ergo_uvo_b200 as U
    from ergo_uvo_b200 import _lib as L
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "uvo_c.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(uvo_params), sizeof(uvo_camera), sizeof(uvo_stereo_result), sizeof(uvo_mono_result),'
                   'sizeof(uvo_keypoint), sizeof(uvo_dmatch), offsetof(uvo_params, stereo_gate),'
                   'offsetof(uvo_params, max_features), sizeof(uvo_jpeg_layout), offsetof(uvo_jpeg_layout, coeff_offset),'
                   'offsetof(uvo_jpeg_layout, quant));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    want = [C.sizeof(L.Params), C.sizeof(L.Camera), C.sizeof(L.StereoResult), C.sizeof(L.MonoResult),
            U.KEYPOINT_DTYPE.itemsize, U.DMATCH_DTYPE.itemsize, L.Params.stereo_gate.offset,
            L.Params.max_features.offset, C.sizeof(L.JpegLayout), L.JpegLayout.coeff_offset.offset,
            L.JpegLayout.quant.offset]
    assert got == want


def test_lazy_dxy_bound_holds_in_f32():
    """k_surf_detect skips the Dxy half of a sample when fl(dx dy) <= threshold.  That is exact because
    det = fl(fl(dx dy) - fl(fl(0.81f dxy) dxy)) <= fl(dx dy): the subtrahend is >= +0 and IEEE subtraction is
    monotonic.  Checked here on random and adversarial f32 values."""
    import numpy as np
    rs = np.random.RandomState(0)
    f = np.float32
    dx = np.concatenate([rs.randn(200000) * 300, rs.randn(1000) * 1e-3, [0, -0.0, 1e30, -1e30]]).astype(f)
    dy = np.concatenate([rs.randn(200000) * 300, rs.randn(1000) * 1e3, [5, 7, 1e-30, 1e-30]]).astype(f)
    dxy = np.concatenate([rs.randn(200000) * 200, rs.randn(1000) * 1e-20, [0, -0.0, 3, -3]]).astype(f)
    p = (dx * dy).astype(f)
    sub = ((f(0.81) * dxy).astype(f) * dxy).astype(f)
    det = (p - sub).astype(f)
    assert (sub >= 0).all() and (det <= p).all()


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md is the map from the C ABI to the reference interface: every exported function is named there"""
    import re
    h = open(os.path.join(ROOT, "include", "uvo_c.h")).read()
    syms = sorted(set(re.findall(r"UVO_API\s+[\w\s\*]+?\b(uvo_\w+)\s*\(", h)))
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    assert len(syms) > 40
    assert [s for s in syms if s not in doc] == []


def test_cpp_example_compiles_links_and_fails_loudly_without_a_gpu(tmp_path):
    """examples/stereo_sequence.cpp: a C++ consumer of the C ABI (no ROS, no OpenCV) builds against include/uvo_c.h and
    the in-tree library; without arguments it prints its usage, and on a machine without a GPU it stops at
    uvo_ctx_create instead of computing anything on the CPU"""
    import torch
    exe = str(tmp_path / "stereo_sequence")
    lib_dir = os.path.join(ROOT, "ergo_uvo_b200")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "stereo_sequence.cpp"), "-L" + lib_dir, "-luvo_b200",
                           "-Wl,-rpath," + lib_dir, "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "usage:" in r.stderr and "sm_100a" in r.stderr
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "640", "512", "0.1", "/nonexistent/l%04d", "/nonexistent/r%04d", "0", "2"],
                           capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stderr


def test_rodrigues_host_matches_cv2():
    """uvo_rodrigues (host-only; cv::Rodrigues(rvec, R) of visual_odometry.h:673) against cv2, no GPU needed"""
    import numpy as np
    cv2 = pytest.importorskip("cv2")
    import ergo_uvo_b200 as U
    lib = U.load()
    rs = np.random.RandomState(3)
    vecs = [np.zeros(3), np.array([1e-20, 0, 0]), np.array([np.pi, 0, 0]), np.array([0.01, -0.02, 0.015])]
    vecs += [rs.uniform(-3, 3, 3) for _ in range(50)]
    for r in vecs:
        R = np.zeros(9)
        assert lib.uvo_rodrigues(r.ctypes.data_as(C.c_void_p), R.ctypes.data_as(C.c_void_p)) == 0
        assert np.abs(R.reshape(3, 3) - cv2.Rodrigues(r)[0]).max() <= 4e-16, r
    assert lib.uvo_rodrigues(None, None) != 0
