"""CPU replay of visual_odometry_node::stereo_VO (reference visual_odometry.h:406-741) built from the oracle's
functions.  Test infrastructure: the checker for the device-resident uvo_stereo pipeline."""
import time

import numpy as np


def stereo_gate(matches, kq, kt, max_dy, min_disp, max_disp):
    """keep iff |y_q - y_t| <= max_dy and min_disp <= x_q - x_t <= max_disp, all in f32 (uvo_params.stereo_gate)"""
    q, t = matches["queryIdx"], matches["trainIdx"]
    dy = np.abs(kq["y"][q].astype(np.float32) - kt["y"][t].astype(np.float32))
    disp = kq["x"][q].astype(np.float32) - kt["x"][t].astype(np.float32)
    keep = (dy <= np.float32(max_dy)) & (disp >= np.float32(min_disp)) & (disp <= np.float32(max_disp))
    return matches[keep]


class RefStereoVO:
    def __init__(self, O, seq, params):
        self.O = O
        self.seq = seq
        self.p = params
        self.init = False
        self.prev = None  # dict(kL, kR, dL) after stereo match
        self.t_prev_curr = np.zeros(3)
        # wall-clock seconds per stage, accumulated over frames (bench.py's per-stage CPU column); same stage names as
        # uvo_stage_name()
        self.stage_s = {k: 0.0 for k in ("get_image", "surf", "match_stereo", "match_temporal",
                                         "triangulate+extract3d", "pnp_ransac")}
        KL, KR = seq.newKL, seq.newKR
        self.P_left = KL @ np.hstack([np.eye(3), np.zeros((3, 1))])
        self.P_right = KR @ np.hstack([seq.R_right, seq.t_right.reshape(3, 1)])

    def frame(self, left, right, dt):
        O, p, s = self.O, self.p, self.seq
        out = dict(valid=0, gate=0, n_stereo=0, n_temporal=0, n_3d=0, n_inliers=0, hyps=0)
        t0 = time.perf_counter()
        gL = O.get_image(left, s.KL, s.DL, s.newKL, bool(p.clahe), float(p.clip_limit))
        gR = O.get_image(right, s.KR, s.DR, s.newKR, bool(p.clahe), float(p.clip_limit))
        t1 = time.perf_counter()
        kL, dL = O.surf_detect_and_compute(gL, p.surf_min_hessian, p.surf_octaves, p.surf_octave_layers,
                                           bool(p.surf_extended), bool(p.surf_upright))
        kR, dR = O.surf_detect_and_compute(gR, p.surf_min_hessian, p.surf_octaves, p.surf_octave_layers,
                                           bool(p.surf_extended), bool(p.surf_upright))
        t2 = time.perf_counter()
        self.stage_s["get_image"] += t1 - t0
        self.stage_s["surf"] += t2 - t1
        out.update(kL=kL, kR=kR, dL=dL, dR=dR, n_left=len(kL), n_right=len(kR))
        was_init = self.init
        cur = None
        gate = 0
        if len(kL) >= p.min_num_features and len(kR) >= p.min_num_features:
            t3 = time.perf_counter()
            ms = O.match_features(dL, dR, np.float32(p.lowe_ratio))
            self.stage_s["match_stereo"] += time.perf_counter() - t3
            if getattr(p, "stereo_gate", 0):  # optional epipolar / disparity gate (not in the reference; off by default)
                ms = stereo_gate(ms, kL, kR, p.stereo_max_epipolar_dy, p.stereo_min_disparity, p.stereo_max_disparity)
            out["m_stereo"] = ms
            out["n_stereo"] = len(ms)
            if len(ms) > p.min_num_features:
                cur = dict(kL=kL[ms["queryIdx"]], kR=kR[ms["trainIdx"]], dL=dL[ms["queryIdx"]])
                if was_init:
                    gate = self._pose(out, cur, kL, dL)
            else:
                gate = 2
        else:
            gate = 1
        if was_init:
            out["gate"] = gate
            out["valid"] = int(gate == 0)
        if cur is not None:
            self.init = True
        self.prev = cur if cur is not None else dict(kL=kL[:0], kR=kR[:0], dL=dL[:0])
        out["initialised"] = int(self.init)
        out["t_prev_curr"] = self.t_prev_curr.copy()
        out["velocity"] = self.t_prev_curr / dt
        return out

    def _pose(self, out, cur, kL, dL):
        O, p, s = self.O, self.p, self.seq
        prev = self.prev
        t0 = time.perf_counter()
        mt = O.match_features(prev["dL"], dL, np.float32(p.lowe_ratio)) if len(prev["dL"]) else prev["dL"][:0]
        t1 = time.perf_counter()
        self.stage_s["match_temporal"] += t1 - t0
        out["m_temporal"] = mt
        out["n_temporal"] = len(mt)
        if len(mt) <= p.min_num_features:
            return 3
        pl = np.stack([prev["kL"]["x"][mt["queryIdx"]], prev["kL"]["y"][mt["queryIdx"]]], -1)
        pr = np.stack([prev["kR"]["x"][mt["queryIdx"]], prev["kR"]["y"][mt["queryIdx"]]], -1)
        X4 = O.triangulate_points(self.P_left, self.P_right, pl, pr)
        pts, idx = O.extract_3dpoints(pl, pr, np.eye(3), np.zeros(3), s.R_right, s.t_right, s.newKL, s.newKR, X4,
                                      p.reprojection_tolerance, p.min_num_3dpoints)
        t2 = time.perf_counter()
        self.stage_s["triangulate+extract3d"] += t2 - t1
        out["n_3d"] = len(idx)
        out["good_idx"] = idx
        if len(idx) <= p.min_num_3dpoints:
            return 4
        tr = mt["trainIdx"][idx]
        x = np.stack([kL["x"][tr], kL["y"][tr]], -1).astype(np.float32)
        ok, rvec, tvec, inl, hyps = O.solve_pnp_ransac_epnp(pts, x, s.newKL, p.iterations_count,
                                                            np.float32(p.reprojection_error), p.confidence)
        self.stage_s["pnp_ransac"] += time.perf_counter() - t2
        out.update(n_inliers=len(inl), hyps=hyps, rvec=rvec, tvec=tvec, inliers=inl)
        if len(inl) < p.min_num_inliers:
            return 5
        R = O.rodrigues_vec2mat(rvec)
        self.t_prev_curr = -R.T @ tvec
        return 0
