"""K10a / K10b parity: CUDA findHomography / findEssentialMat / recoverPose / recover_pose_homography /
estimate_relative_pose (through the C ABI) against the cv2 golden vectors and the numpy oracle (oracle/twoview.py).
Inlier masks and hypothesis counts exact; models to 1e-6 (minimal solvers are not bit-reproducible across LAPACK
builds even on the CPU).  Reference: VO_utility.cpp:134-180, :581-624."""
import os

import numpy as np
import pytest

from oracle import twoview as T
from tools.make_golden_twoview import CASES, scene

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "twoview.npz"))
K4 = GOLD["K4"]
KM = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1.]])


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("case", [c[0] for c in CASES])
@pytest.mark.parametrize("method", ["ransac", "lmeds"])
def test_find_homography_golden(ctx, case, method):
    tag = f"{case}_h"
    p1, p2 = GOLD[tag + "_p1"], GOLD[tag + "_p2"]
    H, mask, hyp = ctx.findHomography(p1, p2, 8 if method == "ransac" else 4, 1.0, 2000, 0.99)
    gm = GOLD[f"{tag}_{method}_mask"]
    assert np.array_equal(mask, gm)
    if gm.sum() >= 8:
        assert _rel(H, GOLD[f"{tag}_{method}_H"]) < 1e-6
    _, _, hyp_o = T.find_homography(p1, p2, 8 if method == "ransac" else 4, 1.0, 2000, 0.99)
    assert hyp == hyp_o


@pytest.mark.parametrize("case", [c[0] for c in CASES])
@pytest.mark.parametrize("method", ["ransac", "lmeds"])
def test_find_essential_and_recover_pose_golden(ctx, case, method):
    tag = f"{case}_e"
    p1, p2 = GOLD[tag + "_p1"], GOLD[tag + "_p2"]
    if method == "ransac":
        E, mask, hyp = ctx.findEssentialMat(p1, p2, K4, 8, 0.999, 1.0, 1000)
        _, _, hyp_o = T.find_essential_mat(p1, p2, K4, 8, 0.999, 1.0, 1000)
    else:
        E, mask, hyp = ctx.findEssentialMat(p1, p2, K4, 4, 0.99, 0.1, 2000)
        hyp_o = 89
    gE = GOLD[f"{tag}_{method}_E"]
    assert np.array_equal(mask, GOLD[f"{tag}_{method}_mask"])
    assert hyp == hyp_o
    s = np.sign((E * gE).sum())
    assert _rel(s * E, gE / np.linalg.norm(gE)) < 1e-5
    good, R, t, m2 = ctx.recoverPose(gE, p1, p2, K4, GOLD[f"{tag}_{method}_mask"])
    assert good == int(GOLD[f"{tag}_{method}_rp_good"])
    assert np.array_equal(m2, GOLD[f"{tag}_{method}_rp_mask"])
    assert _rel(R, GOLD[f"{tag}_{method}_rp_R"]) < 1e-9 and _rel(t, GOLD[f"{tag}_{method}_rp_t"]) < 1e-9


@pytest.mark.parametrize("seed,n,outl", [(21, 300, 0.3), (22, 3000, 0.45), (23, 5, 0.0), (24, 4, 0.0), (25, 40, 0.6)])
def test_two_view_vs_oracle_other_sizes(ctx, seed, n, outl):
    """sizes / seeds outside the fixture, including count == modelPoints"""
    p1, p2, k4 = scene(n, seed, True, outl, 0.7)
    for method in (8, 4):
        H, mask, hyp = ctx.findHomography(p1, p2, method, 2.0, 500, 0.99)
        Ho, mo, hyp_o = T.find_homography(p1, p2, method, 2.0, 500, 0.99)
        assert np.array_equal(mask, mo) and hyp == hyp_o
        if Ho is not None and mo.sum() >= 8:
            assert _rel(H, Ho) < 1e-6
    if n >= 5:
        p1, p2, k4 = scene(n, seed, False, outl, 0.7)
        for method, thr, conf, mi in ((8, 1.0, 0.999, 300), (4, 0.1, 0.99, 2000)):
            E, mask, hyp = ctx.findEssentialMat(p1, p2, K4, method, conf, thr, mi)
            Eo, mo, hyp_o = T.find_essential_mat(p1, p2, K4, method, conf, thr, mi)
            assert np.array_equal(mask, mo) and hyp == hyp_o


def test_recover_pose_homography_vs_oracle(ctx):
    for case in ("a", "b", "c"):
        p1, p2 = GOLD[f"{case}_h_p1"], GOLD[f"{case}_h_p2"]
        H = GOLD[f"{case}_h_ransac_H"]
        ctx.params.homography_distance = 50.0
        good, R, t = ctx.recover_pose_homography(H, p1, p2, K4)
        go, Ro, to = T.recover_pose_homography(H, p1, p2, KM, 50.0)
        assert good == go
        assert _rel(R, Ro) < 1e-9 and _rel(t, to) < 1e-9


def test_estimate_relative_pose_flow(ctx):
    """both branches + the switch: a planar scene fails the essential gate less often than it fails homography on a
    3-D scene; compare the whole decision with the oracle replay"""
    import ergo_uvo_b200 as U
    prm = U.default_params(stereo=False)
    ctx.params = prm
    for planar, start in ((False, True), (True, False), (False, False)):
        p1, p2, k4 = scene(600, 31, planar, 0.25, 0.4)
        ok, R, t, mask, ue = ctx.estimate_relative_pose(p1, p2, K4, start)
        # oracle replay of VO_utility.cpp:134-180
        use_e, switched, ok_o = start, False, False
        while True:
            if use_e:
                E, m, _ = T.find_essential_mat(p1, p2, K4, prm.essential_method, prm.essential_confidence,
                                               prm.essential_threshold, int(prm.essential_max_iters))
                good, Ro, to, m2 = T.recover_pose(E, p1, p2, K4, m)
                valid = int(m2.sum())
            else:
                H, m, _ = T.find_homography(p1, p2, prm.homography_method, prm.homography_threshold,
                                            int(prm.homography_max_iters), prm.homography_confidence)
                good, Ro, to = T.recover_pose_homography(H, p1, p2, KM, prm.homography_distance)
                valid = int(m.sum())
            if valid / len(p1) >= prm.vpf_threshold and valid >= prm.min_num_inliers:
                ok_o = True
                break
            if switched:
                break
            switched, use_e = True, not use_e
        assert ok == ok_o and ue == use_e
        assert np.array_equal(mask, m)
        if ok_o:
            assert _rel(R, Ro) < 1e-6 and _rel(t, to) < 1e-6
    ctx.params = U.default_params(stereo=True)


def test_batched_ransac_sweep_config_d(ctx):
    """BASELINE config D: 4096 hypotheses x 10 000 correspondences (outlier fraction 0.75 keeps every hypothesis
    alive); masks must be reproducible call to call and the homography mask must equal the oracle's"""
    p1, p2, k4 = scene(10000, 11, True, 0.75, 0.3)
    H, mask, hyp = ctx.findHomography(p1, p2, 8, 3.0, 4096, 1 - 2.0 ** -53)
    H2, mask2, hyp2 = ctx.findHomography(p1, p2, 8, 3.0, 4096, 1 - 2.0 ** -53)
    assert hyp == 4096 and hyp2 == 4096 and np.array_equal(mask, mask2) and np.array_equal(H, H2)
    Ho, mo, hyp_o = T.find_homography(p1, p2, 8, 3.0, 4096, 1 - 2.0 ** -53)
    assert hyp_o == 4096 and np.array_equal(mask, mo)
    p1, p2, k4 = scene(10000, 11, False, 0.75, 0.3)
    E, m, hyp = ctx.findEssentialMat(p1, p2, K4, 8, 1 - 2.0 ** -53, 1.0, 4096)
    E2, m2, _ = ctx.findEssentialMat(p1, p2, K4, 8, 1 - 2.0 ** -53, 1.0, 4096)
    assert hyp == 4096 and np.array_equal(m, m2)
    assert m.sum() > 2000  # the true model (2500 inliers) is found


@pytest.mark.parametrize("n,noise", [(4, 0.0), (5, 0.3), (60, 0.5), (2000, 0.3)])
def test_find_homography_method_0_all_points(ctx, n, noise):
    """cv::findHomography(p1, p2, 0): "a regular method using all the points" (mono_VO_parameters.yaml:23 lists it) --
    the DLT kernel on every point, the 10 Levenberg-Marquardt iterations for n > 4, and the mask cv2 4.13 returns (the
    refined model's at the threshold).  Against the numpy oracle and cv2 itself."""
    cv2 = pytest.importorskip("cv2")
    from oracle import twoview as T
    rs = np.random.RandomState(n)
    p1 = rs.uniform(0, 1000, (n, 2)).astype(np.float32)
    Ht = np.array([[1.01, 0.02, 5], [-0.01, 0.99, -3], [1e-5, 2e-5, 1]])
    q = np.c_[p1, np.ones(n)] @ Ht.T
    p2 = (q[:, :2] / q[:, 2:] + rs.randn(n, 2) * noise).astype(np.float32)
    H, mask, hyps = ctx.findHomography(p1, p2, 0, 3.0, 2000, 0.995)
    Ho, mo, _ = T.find_homography(p1, p2, 0, 3.0)
    Hc, mc = cv2.findHomography(p1, p2, 0, 3.0)
    assert H is not None and hyps == 1
    assert np.array_equal(mask, mo) and np.array_equal(mask, mc.ravel())
    assert np.abs(H - Ho).max() <= 1e-6 * max(1.0, np.abs(Ho).max()) and np.abs(H - Hc).max() <= 1e-5 * np.abs(Hc).max()


def test_find_homography_method_0_degenerate_and_unsupported(ctx):
    import ergo_uvo_b200 as U
    p = np.zeros((10, 2), np.float32)                       # all points identical: runKernel fails -> no model
    H, mask, hyps = ctx.findHomography(p, p, 0, 3.0, 2000, 0.995)
    assert H is None and not mask.any()
    rs = np.random.RandomState(0)
    a, b = rs.rand(20, 2).astype(np.float32) * 100, rs.rand(20, 2).astype(np.float32) * 100
    for method in (16, 32, 38):                             # RHO, USAC_DEFAULT, USAC_MAGSAC: not implemented
        with pytest.raises(U.UvoError) as e:
            ctx.findHomography(a, b, method, 3.0, 2000, 0.995)
        assert e.value.code == -5
    with pytest.raises(U.UvoError):                         # findEssentialMat has no method 0
        ctx.findEssentialMat(a, b, np.array([[1300.0, 0, 640], [0, 1300.0, 512], [0, 0, 1]]), 0, 0.99, 1.0, 1000)
