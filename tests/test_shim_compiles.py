"""CPU: the reference-side binding (shim/VO_utility_shim.cpp, shim/cv_interpose.cpp) type-checks against the
reference's OWN header uvo_libraries/VO_utility.h (VO_utility.h:96-117 function declarations, :25-89 globals), with the
declaration-only ROS / OpenCV stand-ins of tests/stubs.  This is a compile check, not a link or a run: the real build
needs ROS + OpenCV.  /root/reference is read where it lies and only in this container."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_INC = "/root/reference/uvo_libraries/include"
CXX = shutil.which("g++")
FLAGS = ["-std=c++14", "-fsyntax-only", "-Wall", "-Wextra", "-Wformat=2", "-Werror", "-Wno-unused-parameter",
         "-I" + os.path.join(ROOT, "tests", "stubs"), "-I" + REF_INC, "-I" + os.path.join(ROOT, "include")]

needs_ref = pytest.mark.skipif(CXX is None or not os.path.exists(os.path.join(REF_INC, "uvo_libraries", "VO_utility.h")),
                               reason="needs g++ and the reference headers (this container only)")


def _check(src, extra=()):
    return subprocess.run([CXX, *FLAGS, *extra, src], capture_output=True, text=True)


@needs_ref
@pytest.mark.parametrize("src,extra", [("VO_utility_shim.cpp", ()), ("VO_utility_shim.cpp", ("-DUVO_SHIM_GPU_IMAGE_DECODE",)),
                                       ("cv_interpose.cpp", ())])
def test_shim_type_checks_against_reference_header(src, extra):
    r = _check(os.path.join(ROOT, "shim", src), extra=extra)
    assert r.returncode == 0, r.stderr[-4000:]


@needs_ref
def test_shim_defines_the_reference_signatures(tmp_path):
    """a definition whose signature drifted from VO_utility.h would be a new overload, not an error: call every
    replaced function through the reference header's declaration and require the shim to be the only candidate"""
    probe = tmp_path / "probe.cpp"
    probe.write_text('#include "' + os.path.join(ROOT, "shim", "VO_utility_shim.cpp") + '''"
// taking the address with the reference's exact type fails to compile if the shim's definition is a different overload
static Mat (*p_get_image)(const Mat&, const Mat&, const Mat&, const Mat&) = &get_image;
static void (*p_detect)(Mat, vector<KeyPoint>&, Mat&) = &detect_features;
static void (*p_match5)(vector<KeyPoint>, vector<KeyPoint>, Mat, Mat, vector<DMatch>&) = &match_features;
static void (*p_match7)(vector<KeyPoint>, vector<KeyPoint>, Mat, Mat, vector<DMatch>&, vector<Point2f>&,
                        vector<Point2f>&) = &match_features;
static void (*p_pose)(vector<Point2f>, vector<Point2f>, Mat, Mat&, Mat&, vector<Point2f>&, vector<Point2f>&,
                      vector<DMatch>&, bool&) = &estimate_relative_pose;
static void (*p_x3d)(vector<Point2f>, vector<Point2f>, Mat, Mat, Mat, Mat, Mat, Mat, Mat, Mat&, Mat&) = &extract_3Dpoints;
static bool (*p_sel)(const vector<Point2f>&, const vector<Point2f>&) = &select_estimation_method;
static int (*p_rph)(Mat, vector<Point2f>, vector<Point2f>, Mat, Mat&, Mat&) = &recover_pose_homography;
static void (*p_rcm)(Mat, Mat&, Mat, Mat&) = &resize_camera_matrix;
int main() { return p_get_image && p_detect && p_match5 && p_match7 && p_pose && p_x3d && p_sel && p_rph && p_rcm ? 0 : 1; }
''')
    r = _check(str(probe), extra=["-Wno-unused-variable"])
    assert r.returncode == 0, r.stderr[-4000:]
    # and the number of function definitions in the shim equals the number it claims to replace: no stray overloads
    import re
    src = open(os.path.join(ROOT, "shim", "VO_utility_shim.cpp")).read()
    body = src.split("namespace uvo_shim")[0].split("}  // namespace\n", 1)[1]  # after the file-local helpers
    names = re.findall(r"^(?:Mat|void|bool|int) (\w+)\(", body, re.M)
    assert sorted(names) == sorted(["get_image", "resize_camera_matrix", "detect_features", "match_features",
                                    "match_features", "extract_3Dpoints", "select_estimation_method",
                                    "estimate_relative_pose", "recover_pose_homography"])


def test_stub_pods_have_opencv_layouts(tmp_path):
    """the stand-in cv::KeyPoint / DMatch / Point2f must have the sizes the C ABI structs mirror"""
    if CXX is None:
        pytest.skip("no g++")
    probe = tmp_path / "sizes.cpp"
    probe.write_text('''#include <opencv2/opencv.hpp>
#include "uvo_c.h"
static_assert(sizeof(cv::KeyPoint) == 28 && sizeof(uvo_keypoint) == 28, "KeyPoint");
static_assert(sizeof(cv::DMatch) == 16 && sizeof(uvo_dmatch) == 16, "DMatch");
static_assert(sizeof(cv::Point2f) == 8, "Point2f");
int main() { return 0; }
''')
    r = subprocess.run([CXX, "-std=c++14", "-fsyntax-only", "-I" + os.path.join(ROOT, "tests", "stubs"),
                        "-I" + os.path.join(ROOT, "include"), str(probe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
