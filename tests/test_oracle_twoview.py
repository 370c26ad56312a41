"""Pins oracle/twoview.py (numpy restatement of findHomography / findEssentialMat / recoverPose /
decomposeHomographyMat as estimate_relative_pose calls them, VO_utility.cpp:134-180, :581-624) against the committed
cv2 4.13 vectors (tests/golden/twoview.npz, tools/make_golden_twoview.py): inlier masks identical, models to 1e-7."""
import os

import numpy as np
import pytest

from oracle import twoview as T
from tools.make_golden_twoview import CASES

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "twoview.npz"))
K4 = GOLD["K4"]
KM = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.]])


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("case", [c[0] for c in CASES])
@pytest.mark.parametrize("method", ["ransac", "lmeds"])
def test_find_homography_matches_cv2(case, method):
    tag = f"{case}_h"
    p1, p2 = GOLD[tag + "_p1"], GOLD[tag + "_p2"]
    H, mask, hyp = T.find_homography(p1, p2, T.RANSAC if method == "ransac" else T.LMEDS, 1.0, 2000, 0.99)
    gm = GOLD[f"{tag}_{method}_mask"]
    assert np.array_equal(mask, gm)
    if gm.sum() >= 8:  # LMedS on > 50 % outliers returns a meaningless (ill-conditioned) model in cv2 as well
        assert _rel(H, GOLD[f"{tag}_{method}_H"]) < 1e-7
    if method == "lmeds":
        assert hyp == 48


@pytest.mark.parametrize("case", [c[0] for c in CASES])
@pytest.mark.parametrize("method", ["ransac", "lmeds"])
def test_find_essential_and_recover_pose_match_cv2(case, method):
    tag = f"{case}_e"
    p1, p2 = GOLD[tag + "_p1"], GOLD[tag + "_p2"]
    if method == "ransac":
        E, mask, hyp = T.find_essential_mat(p1, p2, K4, T.RANSAC, 0.999, 1.0, 1000)
    else:
        E, mask, hyp = T.find_essential_mat(p1, p2, K4, T.LMEDS, 0.99, 0.1, 2000)
        assert hyp == 89
    gE = GOLD[f"{tag}_{method}_E"]
    assert np.array_equal(mask, GOLD[f"{tag}_{method}_mask"])
    s = np.sign((E * gE).sum())
    assert _rel(s * E / np.linalg.norm(E), gE / np.linalg.norm(gE)) < 1e-5  # roots of an ill-conditioned degree-10 polynomial
    good, R, t, m2 = T.recover_pose(gE, p1, p2, K4, GOLD[f"{tag}_{method}_mask"])
    assert good == int(GOLD[f"{tag}_{method}_rp_good"])
    assert np.array_equal(m2, GOLD[f"{tag}_{method}_rp_mask"])
    assert _rel(R, GOLD[f"{tag}_{method}_rp_R"]) < 1e-9 and _rel(t, GOLD[f"{tag}_{method}_rp_t"]) < 1e-9


@pytest.mark.parametrize("case", [c[0] for c in CASES])
def test_decompose_homography_matches_cv2(case):
    key = f"{case}_h_ransac_dec_R"
    if key not in GOLD:
        pytest.skip("no decomposition stored for this case")
    dec = T.decompose_homography_mat(GOLD[f"{case}_h_ransac_H"], KM)
    assert len(dec) == len(GOLD[key])
    for i, (R, t, n) in enumerate(dec):
        assert _rel(R, GOLD[key][i]) < 1e-9
        assert _rel(t, GOLD[f"{case}_h_ransac_dec_t"][i]) < 1e-9
        assert _rel(n, GOLD[f"{case}_h_ransac_dec_n"][i]) < 1e-8


def test_live_cv2_if_available():
    """same comparison against the cv2 importable in this environment (skipped where cv2 is absent)"""
    cv2 = pytest.importorskip("cv2")
    from tools.make_golden_twoview import scene
    p1, p2, k4 = scene(400, 11, True, 0.35, 0.8)
    H, m = cv2.findHomography(p1, p2, cv2.RANSAC, 2.0, maxIters=500, confidence=0.98)
    Ho, mo, _ = T.find_homography(p1, p2, T.RANSAC, 2.0, 500, 0.98)
    assert np.array_equal(m.ravel(), mo) and _rel(Ho, H) < 1e-7
    p1, p2, k4 = scene(400, 12, False, 0.35, 0.8)
    E, m = cv2.findEssentialMat(p1, p2, KM, cv2.LMEDS, 0.99, 0.1, 2000)
    Eo, mo, _ = T.find_essential_mat(p1, p2, K4, T.LMEDS, 0.99, 0.1, 2000)
    assert np.array_equal(m.ravel(), mo)
