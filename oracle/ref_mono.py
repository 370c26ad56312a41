"""CPU replay of visual_odometry_node::mono_VO (reference visual_odometry.h:164-397) built from the oracle's functions
(oracle.py for the image / SURF / matcher / 3-D point steps, twoview.py for the two-view step).  Test infrastructure:
the checker for uvo_mono."""
import numpy as np

from . import twoview as T


class RefMonoVO:
    def __init__(self, O, seq, params):
        self.O, self.seq, self.p = O, seq, params
        self.init = False
        self.prev = None
        self.use_essential = True
        self.R, self.t, self.SF = np.eye(3), np.zeros(3), 1.0
        K = seq.newK
        self.K4 = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]])

    def estimate_relative_pose(self, p1, p2):
        """VO_utility.cpp:134-180"""
        p = self.p
        Km = np.array([[self.K4[0], 0, self.K4[2]], [0, self.K4[1], self.K4[3]], [0, 0, 1.]])
        switched = False
        while True:
            if self.use_essential:
                E, mask, _ = T.find_essential_mat(p1, p2, self.K4, p.essential_method, p.essential_confidence,
                                                  p.essential_threshold, int(p.essential_max_iters))
                if E is not None:
                    _, self.R, self.t, m2 = T.recover_pose(E, p1, p2, self.K4, mask)
                    valid = int(m2.sum())
                else:
                    valid = 0
            else:
                H, mask, _ = T.find_homography(p1, p2, p.homography_method, p.homography_threshold,
                                               int(p.homography_max_iters), p.homography_confidence)
                if H is not None:
                    _, R, t = T.recover_pose_homography(H, p1, p2, Km, p.homography_distance)
                    if R is not None:
                        self.R, self.t = R, t
                valid = int(mask.sum())
            if len(p1) and valid / len(p1) >= p.vpf_threshold and valid >= p.min_num_inliers:
                return True, mask
            if switched:
                return False, mask
            switched = True
            self.use_essential = not self.use_essential

    def frame(self, img, dt, rng):
        O, p, s = self.O, self.p, self.seq
        out = dict(initialised=0, skipped=0, published=0, valid=0, n_matches=0, n_inliers=0, n_3d=0)
        g = O.get_image(img, s.K, s.D, s.newK, bool(p.clahe), float(p.clip_limit))
        k, d = O.surf_detect_and_compute(g, p.surf_min_hessian, p.surf_octaves, p.surf_octave_layers,
                                         bool(p.surf_extended), bool(p.surf_upright))
        out["n_keypoints"] = len(k)
        prev, self.prev = self.prev, (k, d)
        if not self.init:
            if len(k) >= p.min_num_features:
                self.init = True
            out["initialised"] = int(self.init)
            return out
        out["initialised"] = 1
        if len(k) < p.min_num_features:
            out["skipped"] = 1
            return out
        pk, pd = prev
        m = O.match_features(pd, d, np.float32(p.lowe_ratio)) if len(pd) else np.zeros(0, O.DMATCH_DTYPE)
        out["n_matches"] = len(m)
        if len(m) < p.min_num_features:
            out["skipped"] = 1
            return out
        p1 = np.stack([pk["x"][m["queryIdx"]], pk["y"][m["queryIdx"]]], 1).astype(np.float32)
        p2 = np.stack([k["x"][m["trainIdx"]], k["y"][m["trainIdx"]]], 1).astype(np.float32)
        self.use_essential = bool(O.select_estimation_method(p1, p2, p.distance))
        success, mask = self.estimate_relative_pose(p1, p2)
        out["used_essential"] = int(self.use_essential)
        out["n_inliers"] = int(mask.sum())
        out["mask"] = mask
        valid = success
        if success:
            i1, i2 = p1[mask.astype(bool)], p2[mask.astype(bool)]
            Km = s.newK
            P0 = Km @ np.eye(3, 4)
            P1 = Km @ np.c_[self.R, self.t]
            X4 = O.triangulate_points(P0, P1, i1, i2)
            good, gidx = O.extract_3dpoints(i1, i2, np.eye(3), np.zeros(3), self.R, self.t, Km, Km, X4,
                                            p.reprojection_tolerance, p.min_num_3dpoints)
            out["n_3d"] = len(good)
            if len(good) < p.min_num_3dpoints:
                valid = False
            else:
                sf, n_front = O.scale_factor(good, self.R, self.t, np.float32(rng), with_count=True)
                if n_front > 0:  # `if(!good_currCam_points.empty())`, visual_odometry.h:366 -- also when range == 0
                    self.SF = sf
                else:
                    valid = False
        out["valid"] = int(valid)
        out["published"] = 1
        out["R"], out["t"], out["scale_factor"] = self.R.copy(), self.t.copy(), self.SF
        out["velocity"] = -self.SF * (self.R.T @ self.t) / dt
        return out
