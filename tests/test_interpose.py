"""The cv:: interposer (shim/cv_interpose.cpp) linked and RUN without OpenCV or ROS: the three calls the
unchanged node makes by itself -- cv::triangulatePoints (visual_odometry.h:631), cv::solvePnPRansac (:647-648),
cv::Rodrigues (:673) -- must bind to the interposer when `uvo_libraries` precedes OpenCV on the link line
(uvo/CMakeLists.txt:43-47), and argument shapes it does not take must reach OpenCV's definition through
dlsym(RTLD_NEXT, ...).

"OpenCV" here is a toy: tests/interpose/fake_cv_core.cpp implements the stand-in classes of tests/stubs, and
tests/interpose/fake_cv_calib3d.cpp exports the three functions with OpenCV's exact signatures (= the mangled names
the real libopencv_calib3d exports), counting its calls.  What is demonstrated is the LINKING behaviour (symbol
precedence, RTLD_NEXT fall-through, the argument-shape tests); the arithmetic behind the GPU route is covered by
tests/test_gpu_pose.py.  The CPU test runs without a GPU (uvo_ctx_create fails, so the two GPU routes fall through
and only Rodrigues -- host-only -- is answered by the interposer); the `gpu` test runs the same executable on the GPU box."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "interpose")
CXX = shutil.which("g++")
SO = os.path.join(ROOT, "ergo_uvo_b200", "libuvo_b200.so")
MANGLED = {"triangulatePoints": "_ZN2cv17triangulatePointsERKNS_11_InputArrayES2_S2_S2_RKNS_12_OutputArrayE",
           "solvePnPRansac": "_ZN2cv14solvePnPRansacERKNS_11_InputArrayES2_S2_S2_RKNS_12_OutputArrayES5_bifdS5_i",
           "Rodrigues": "_ZN2cv9RodriguesERKNS_11_InputArrayERKNS_12_OutputArrayES5_"}

pytestmark = pytest.mark.skipif(CXX is None or not os.path.exists(SO), reason="needs g++ and the built libuvo_b200.so")


def _build(out):
    inc = ["-I" + os.path.join(ROOT, "tests", "stubs"), "-I" + os.path.join(ROOT, "include")]
    cxx = [CXX, "-std=c++14", "-O1", "-fPIC", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter"]
    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, " ".join(cmd) + "\n" + r.stderr[-3000:]
    # "OpenCV": core (the classes) and calib3d (the three functions)
    run(cxx + inc + ["-shared", "-o", os.path.join(out, "libfake_opencv_core.so"), os.path.join(SRC, "fake_cv_core.cpp")])
    run(cxx + inc + ["-shared", "-o", os.path.join(out, "libfake_opencv_calib3d.so"),
                     os.path.join(SRC, "fake_cv_calib3d.cpp"), "-L" + out, "-lfake_opencv_core"])
    # the replacement library: the interposer + the C ABI it calls (what `uvo_libraries` becomes)
    run(cxx + inc + ["-shared", "-o", os.path.join(out, "libuvo_interpose.so"), os.path.join(ROOT, "shim", "cv_interpose.cpp"),
                     SO, "-ldl", "-L" + out, "-lfake_opencv_core"])
    # the node: uvo_libraries BEFORE OpenCV, as ${catkin_LIBRARIES} precedes ${OpenCV_LIBRARIES} in uvo/CMakeLists.txt
    run(cxx + inc + ["-o", os.path.join(out, "node_demo"), os.path.join(SRC, "node_demo.cpp"), "-L" + out,
                     "-luvo_interpose", "-lfake_opencv_calib3d", "-lfake_opencv_core",
                     "-Wl,-rpath," + out + ":" + os.path.dirname(SO)])
    return os.path.join(out, "node_demo")


def _run(exe, n=200):
    env = dict(os.environ, LD_DEBUG="bindings", LD_DEBUG_OUTPUT="")
    env.pop("LD_DEBUG_OUTPUT")
    r = subprocess.run([exe, str(n)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = dict(l.split(": ", 1) for l in r.stdout.strip().splitlines())
    return lines, r.stderr


def _bound_to(stderr, sym, frm):
    """library a `binding file <frm> ... to <lib>: normal symbol `<sym>'` line of LD_DEBUG=bindings names"""
    for l in stderr.splitlines():
        if "symbol `" + sym + "'" in l and "binding file" in l and frm in l.split(" to ")[0]:
            return os.path.basename(l.split(" to ")[1].split(" ")[0].rstrip(":"))
    return None


def test_node_calls_bind_to_the_interposer_and_fall_through(tmp_path):
    exe = _build(str(tmp_path))
    out, dbg = _run(exe)
    # symbol precedence: the executable's references resolve to the interposer, not to "OpenCV"
    for name, sym in MANGLED.items():
        assert _bound_to(dbg, sym, "node_demo") == "libuvo_interpose.so", (name, _bound_to(dbg, sym, "node_demo"))
    # Rodrigues in the node's shape is answered by the interposer on any machine (host-only)
    assert out["Rodrigues node-shape"].startswith("interposer")
    import numpy as np
    cv2 = pytest.importorskip("cv2")
    R = cv2.Rodrigues(np.array([0.01, -0.02, 0.015]))[0]
    f = dict(kv.split("=") for kv in out["Rodrigues node-shape"].split()[1:])
    assert abs(float(f["R00"]) - R[0, 0]) < 1e-15 and abs(float(f["R01"]) - R[0, 1]) < 1e-15
    # argument shapes the interposer does not take reach the next definition (RTLD_NEXT): the toy writes 42
    assert out["triangulatePoints f32-projections"] == "opencv X4[0]=42"
    assert out["solvePnPRansac iterative"] == "opencv r[0]=42"
    assert out["Rodrigues matrix-input"] == "opencv"
    import torch
    if not torch.cuda.is_available():
        # no GPU: uvo_ctx_create fails, the interposer hands the node-shaped calls to OpenCV as well (no CPU
        # fallback inside the library -- the fall-through IS OpenCV)
        assert out["triangulatePoints node-shape"].startswith("opencv")
        assert out["solvePnPRansac node-shape"].startswith("opencv")


@pytest.mark.gpu
def test_node_calls_reach_the_gpu_through_the_interposer(tmp_path):
    """on the GPU box the node-shaped calls are answered by the interposer with the library's results"""
    exe = _build(str(tmp_path))
    out, dbg = _run(exe, n=500)
    t = out["triangulatePoints node-shape"]
    assert t.startswith("interposer") and float(t.split("maxerr=")[1]) < 1e-3
    s = out["solvePnPRansac node-shape"]
    assert s.startswith("interposer") and "ok=1" in s and "inliers=500" in s
    # the pose of the synthetic scene is the identity: rvec = tvec = 0 up to the f32 pixel coordinates
    assert float(s.split("=")[-1]) < 1e-3
    assert out["Rodrigues node-shape"].startswith("interposer")
    assert out["triangulatePoints f32-projections"].startswith("opencv")
    assert out["solvePnPRansac iterative"].startswith("opencv")
