"""Generates tests/golden/surf_{640x480,1280x1024}.npz -- the pin of SURF against the reference's
own detector, `cv::xfeatures2d::SURF::create(hessian, 4, 3, extended, upright)->detectAndCompute`
(uvo_libraries/src/VO_utility.cpp:114-119).

This is the ONE piece of the hot path whose arithmetic cannot be executed in the build container: SURF lives in
opencv_contrib's non-free module and the only OpenCV here is the opencv-python-headless 4.13 wheel ("Non-free
algorithms: NO", no cv2.xfeatures2d; searched: no contrib wheel, source tree, conda package or pip cache anywhere in
the image).  Until this script has been run on a machine that has it, tests/test_oracle_surf_pin.py and
tests/test_gpu_surf_pin.py report "UNPINNED" (xfail) and DESIGN.md says "parity unpinned" for K4-K7.

Run on any machine with `cv2.xfeatures2d.SURF_create` (opencv-contrib-python built with OPENCV_ENABLE_NONFREE=ON),
numpy and scipy, from the repo root:

    python -m tools.make_golden_surf                 # writes tests/golden/surf_640x480.npz, surf_1280x1024.npz
    python -m pytest tests/test_oracle_surf_pin.py   # the CPU restatement against the fixture
    python -m pytest tests/test_gpu_surf_pin.py -m gpu   # the CUDA kernels against the fixture (on a B200)

then commit the two .npz files.  `--detector oracle` writes the same file layout from the repository's own CPU
restatement instead (used by the tests to exercise the fixture path end to end; such a file is marked
`source = "oracle"` and is NOT a pin -- the tests refuse to count it as one).

Fixture layout (one .npz per image size; `gray` is the exact uint8 image the detector saw, so the pin does not depend
on the image generator or on the cv2 version that prepared it):
    gray                         (h, w) u8: left frame 0 of tools/synth.StereoSequence after the reference's get_image
    thresholds                   the hessianThreshold values
    source, cv2_version, build   provenance
    k_{thr}_{u|o}                (N, 7) f64: pt.x, pt.y, size, angle, response, octave, class_id in output order
                                 (u = upright, o = oriented)
    rows_{thr}                   (R,) i32: the keypoint rows whose descriptors are stored (all, or an even subsample)
    d_{thr}_{u|o}_{64|128}       (R, 64|128) f32: descriptor rows `rows_{thr}` (extended = 128)
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SIZES = ((640, 480), (1280, 1024))
THRESHOLDS = (50, 1500)      # mono_VO_parameters.yaml:45, stereo_VO_parameters.yaml:43
MAX_DESC_ROWS = {(640, 480): 512, (1280, 1024): 256}
FIELDS = ("x", "y", "size", "angle", "response", "octave", "class_id")


def fixture_path(w, h, out_dir=OUT):
    return os.path.join(out_dir, f"surf_{w}x{h}.npz")


def prepared_gray_cv2(w, h):
    """left image of frame 0 through the reference's get_image (VO_utility.cpp:346-357): cvtColor(RGB2GRAY) ->
    undistort(K, D, newK) -> CLAHE(clip 8, stereo_VO_parameters.yaml:14-15), with cv2."""
    import cv2
    seq = synth.StereoSequence(w, h, n_frames=1, tex_size=2048 if w > 640 else 1024)
    L = seq.frames[0][0]
    g = cv2.cvtColor(L, cv2.COLOR_RGB2GRAY)
    g = cv2.undistort(g, seq.KL, seq.DL, None, seq.newKL)
    cl = cv2.createCLAHE()
    cl.setClipLimit(8)
    return cl.apply(g)


def prepared_gray_oracle(w, h):
    from oracle import oracle as O
    seq = synth.StereoSequence(w, h, n_frames=1, tex_size=2048 if w > 640 else 1024)
    return O.get_image(seq.frames[0][0], seq.KL, seq.DL, seq.newKL, True, 8.0)


def detect_cv2(gray, thr, extended, upright):
    import cv2
    surf = cv2.xfeatures2d.SURF_create(float(thr), 4, 3, bool(extended), bool(upright))
    kps, desc = surf.detectAndCompute(gray, None)
    k = np.array([[p.pt[0], p.pt[1], p.size, p.angle, p.response, p.octave, p.class_id] for p in kps], np.float64)
    k = k.reshape(-1, 7)
    d = np.zeros((0, 128 if extended else 64), np.float32) if desc is None else np.asarray(desc, np.float32)
    return k, d


def detect_oracle(gray, thr, extended, upright):
    from oracle import oracle as O
    k, d = O.surf_detect_and_compute(gray, thr, extended=bool(extended), upright=bool(upright))
    return np.stack([k[f].astype(np.float64) for f in FIELDS], -1).reshape(-1, 7), d


def build_fixture(w, h, detector="cv2", thresholds=THRESHOLDS, gray=None, max_rows=None):
    detect = detect_cv2 if detector == "cv2" else detect_oracle
    if gray is None:
        gray = prepared_gray_cv2(w, h) if detector == "cv2" else prepared_gray_oracle(w, h)
    out = dict(gray=np.ascontiguousarray(gray, np.uint8), thresholds=np.array(thresholds, np.int32),
               source=np.array(detector))
    if detector == "cv2":
        import cv2
        out["cv2_version"] = np.array(cv2.__version__)
        out["build"] = np.array("\n".join(l for l in cv2.getBuildInformation().splitlines()
                                          if "Non-free" in l or "Extra modules" in l or "Version control" in l))
    max_rows = max_rows or MAX_DESC_ROWS.get((w, h), 512)
    for thr in thresholds:
        rows = None
        for upright in (1, 0):
            tag = "u" if upright else "o"
            for extended in (0, 1):
                k, d = detect(gray, thr, extended, upright)
                assert len(k) == len(d)
                if extended == 0:
                    out[f"k_{thr}_{tag}"] = k
                else:  # the keypoints do not depend on `extended`
                    assert np.array_equal(k, out[f"k_{thr}_{tag}"]), "keypoints changed with the extended flag"
                if rows is None:  # the keypoint SET does not depend on `upright` either (only the angle field does)
                    n = len(k)
                    rows = np.arange(n) if n <= max_rows else np.linspace(0, n - 1, max_rows).astype(np.int64)
                    out[f"rows_{thr}"] = rows.astype(np.int32)
                if len(k) == len(out[f"k_{thr}_u"]):
                    out[f"d_{thr}_{tag}_{128 if extended else 64}"] = d[rows]
                else:  # an oriented run dropped keypoints (no orientation sample inside the image): store every row
                    out[f"d_{thr}_{tag}_{128 if extended else 64}"] = d
                    out[f"rows_{thr}_{tag}"] = np.arange(len(k), dtype=np.int32)
    return out


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--detector", choices=("cv2", "oracle"), default="cv2")
    ap.add_argument("--out", default=OUT)
    args = ap.parse_args()
    if args.detector == "cv2":
        import cv2
        if not hasattr(cv2, "xfeatures2d") or not hasattr(cv2.xfeatures2d, "SURF_create"):
            sys.exit("this cv2 has no xfeatures2d.SURF_create (needs opencv-contrib built with OPENCV_ENABLE_NONFREE): "
                     "SURF stays UNPINNED")
    os.makedirs(args.out, exist_ok=True)
    for (w, h) in SIZES:
        fx = build_fixture(w, h, args.detector)
        path = fixture_path(w, h, args.out)
        np.savez_compressed(path, **fx)
        print(path, {t: len(fx[f"k_{t}_u"]) for t in THRESHOLDS}, f"{os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
