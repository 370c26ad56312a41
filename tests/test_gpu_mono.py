"""Whole-frame parity of the mono path: uvo_mono (C ABI) against the CPU replay of visual_odometry_node::mono_VO
(oracle/ref_mono.py; reference visual_odometry.h:247-397) on synthetic sequences -- counts, the branch taken and
the inlier sets exact, R / t / SF / velocity to 1e-6 (homography) / 1e-4 (essential) -- BASELINE's own tolerance."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("velocity,expect_essential", [((0.02, 0.004, 0.0), False), ((0.12, 0.03, 0.0), True)])
def test_mono_sequence_matches_cpu_replay(ctx, oracle, velocity, expect_essential):
    import ergo_uvo_b200 as U
    from oracle.ref_mono import RefMonoVO
    from tools import synth
    seq = synth.MonoSequence(640, 480, n_frames=4, tex_size=1024, velocity=velocity)
    p = U.default_params(False)
    cam = U.make_camera(seq.K, seq.D, seq.newK)
    vo = U.MonoVO(ctx, 640, 480, cam, p)
    ref = RefMonoVO(oracle, seq, p)
    published = 0
    for k in range(4):
        r = vo.frame(seq.frames[k], 0.1, seq.ranges[k])
        o = ref.frame(seq.frames[k], 0.1, seq.ranges[k])
        for f in ("initialised", "skipped", "published", "valid", "n_keypoints", "n_matches", "n_inliers", "n_3d"):
            assert getattr(r, f) == o[f], (k, f, getattr(r, f), o[f])
        if o["published"]:
            published += 1
            assert r.used_essential == o["used_essential"]
            assert bool(r.used_essential) == expect_essential
            # north_star tolerances: rotation within 0.01 deg (1.7e-4 rad), velocity within 1e-4 relative.  The
            # homography branch agrees to 1e-6; the essential matrix comes out of a degree-10 root finder whose last
            # digits differ between Durand-Kerner (GPU, OpenCV) and the companion-matrix solver of the numpy oracle.
            tol = 1e-4 if expect_essential else 1e-6
            assert np.abs(np.array(r.R).reshape(3, 3) - o["R"]).max() < tol
            assert np.abs(np.array(r.t) - o["t"]).max() < tol
            assert abs(r.scale_factor - o["scale_factor"]) <= tol * abs(o["scale_factor"])
            assert np.abs(np.array(r.velocity) - o["velocity"]).max() <= tol * np.abs(o["velocity"]).max()
    assert published == 3
    vo.close()


def test_mono_gates(ctx):
    """too few features: the first frame does not initialise; later frames are skipped, nothing is published"""
    import ergo_uvo_b200 as U
    from tools import synth
    seq = synth.MonoSequence(640, 480, n_frames=2, tex_size=1024)
    p = U.default_params(False)
    p.surf_min_hessian = 10 ** 9
    vo = U.MonoVO(ctx, 640, 480, U.make_camera(seq.K, seq.D, seq.newK), p)
    r = vo.frame(seq.frames[0], 0.1, 3.0)
    assert r.initialised == 0 and r.n_keypoints == 0 and r.published == 0
    vo.close()


def test_mono_zero_range_assigns_zero_scale_and_stays_valid(ctx, oracle):
    """before the first altimeter message `range` is 0: compute_scale_factor returns 0 with a NON-empty converted set,
    the node assigns SF = 0, keeps successful_estimate = 1 and publishes zero velocity (visual_odometry.h:365-374,
    :382) -- the gate is the emptiness of the set, not the value of SF"""
    import ergo_uvo_b200 as U
    from oracle.ref_mono import RefMonoVO
    from tools import synth
    seq = synth.MonoSequence(640, 480, n_frames=3, tex_size=1024, velocity=(0.02, 0.004, 0.0))
    p = U.default_params(False)
    vo = U.MonoVO(ctx, 640, 480, U.make_camera(seq.K, seq.D, seq.newK), p)
    ref = RefMonoVO(oracle, seq, p)
    ranges = [seq.ranges[0], seq.ranges[1], 0.0]
    for k in range(3):
        r = vo.frame(seq.frames[k], 0.1, ranges[k])
        o = ref.frame(seq.frames[k], 0.1, ranges[k])
        assert (r.valid, r.published, r.n_3d) == (o["valid"], o["published"], o["n_3d"])
    assert o["valid"] == 1 and o["scale_factor"] == 0.0
    assert r.valid == 1 and r.scale_factor == 0.0 and np.all(np.array(r.velocity) == 0.0)
    vo.close()
