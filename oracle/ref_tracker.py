"""CPU restatement of textureless_3d_reconstruction_b200/tracking.py on top of the C oracle.

TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py CPU legs).  PARITY UNPINNED: the
reference has no tracker (its poses come from SIFT + essential matrix,
depth_enhanced_reconstruction.py:1084-1099); the loop is defined by this repo as
frame-to-model point-to-plane ICP against the R6 surface of the blocks
visible from the predicted pose, followed by R4/R5 integration.
"""
from __future__ import annotations

import numpy as np

from . import capi


def predict_pose(T_prev, T_prev2, model="constant_velocity"):
    if T_prev2 is None or model == "previous":
        return T_prev.copy()
    return (T_prev @ np.linalg.inv(T_prev2)) @ T_prev


class FrameToModelTracker:
    def __init__(self, K, H, W, voxel_size=0.01, sdf_trunc=0.04, depth_max=5.0, min_depth=0.1, icp_subsample=4,
                 icp_max_corr=0.05, icp_max_iter=30, weight_threshold=1.0, motion_model="constant_velocity",
                 min_points=100):
        self.K, self.H, self.W = tuple(float(k) for k in K), H, W
        self.depth_max, self.min_depth = depth_max, min_depth
        self.icp_subsample, self.icp_max_corr, self.icp_max_iter = icp_subsample, icp_max_corr, icp_max_iter
        self.weight_threshold, self.motion_model, self.min_points = weight_threshold, motion_model, min_points
        self.volume = capi.TSDFVolume(voxel_size, sdf_trunc)
        self.poses, self.icp_log = [], []

    def add_frame(self, depth, bgr, init_pose=None, known_pose=None):
        res = None
        if known_pose is not None:
            T = np.eye(4); T[:3, :4] = np.asarray(known_pose, np.float64)[:3, :4]
        elif not self.poses:
            T = np.eye(4)
            if init_pose is not None:
                T[:3, :4] = np.asarray(init_pose, np.float64)[:3, :4]
        else:
            if init_pose is not None:
                guess = np.eye(4); guess[:3, :4] = np.asarray(init_pose, np.float64)[:3, :4]
            else:
                guess = predict_pose(self.poses[-1], self.poses[-2] if len(self.poses) > 1 else None,
                                     self.motion_model)
            tgt, tgt_n, _ = self.volume.extract_points(self.weight_threshold,
                                                       view=(self.K, guess, self.H, self.W, self.depth_max))
            src, _ = capi.backproject(depth, None, *self.K, min_depth=self.min_depth, max_depth=self.depth_max,
                                      pose=None, subsample=self.icp_subsample)
            T = guess
            if len(tgt) >= self.min_points and len(src) >= self.min_points:
                res = capi.icp_point_to_plane(src, tgt, tgt_n, self.icp_max_corr, T0=np.linalg.inv(guess),
                                              max_iter=self.icp_max_iter)
                T = np.linalg.inv(res["T"])
        self.volume.integrate(depth, bgr, self.K, T, 1.0, self.depth_max)
        self.poses.append(T)
        self.icp_log.append(res)
        return T
