"""Frame-to-model tracking + fusion (BASELINE config 3: point-to-plane ICP + TSDF).

north_star replaces the reference's pose step — SIFT/ORB/LSD matching + essential
matrix (depth_enhanced_reconstruction.py:1084-1099, 1175-1197) — by frame-to-model
point-to-plane ICP, and its vstack merge (der:1149-1156, 1236-1237) by TSDF fusion.
The reference has no tracker, so the loop is DEFINED here (and restated on the CPU in
oracle/ref_tracker.py for parity):

  frame 0: pose = given initial pose (identity by default); integrate.
  frame i: guess  = constant-velocity prediction  (T_{i-1} T_{i-2}^-1) T_{i-1}   (i >= 2)
                    or the previous pose                                          (i == 1)
           target = R6 surface points + normals (weight >= weight_threshold) of the model
                    blocks visible from `guess`            (t3d_tsdf_extract_points_view)
           source = K1 back-projection of the frame in the camera frame, stride
                    `icp_subsample`, depth in (min_depth, depth_max)
           R8 ICP (max_corr, 30 iterations, 1e-6/1e-6) from inv(guess) -> T_wc; pose = inv(T_wc)
           integrate the frame with that pose (K4 + K5).

All heavy steps run in libt3d.so; this module only sequences them.
"""
from __future__ import annotations

import numpy as np

from .runtime import TSDFVolume, get_context


def predict_pose(T_prev, T_prev2, model="constant_velocity"):
    """World->camera 4x4 guess for the next frame."""
    if T_prev2 is None or model == "previous":
        return T_prev.copy()
    return (T_prev @ np.linalg.inv(T_prev2)) @ T_prev


def as4x4(T):
    T = np.asarray(T, np.float64)
    if T.shape == (4, 4):
        return T.copy()
    out = np.eye(4)
    out[:3, :4] = T[:3, :4]
    return out


class FrameToModelTracker:
    def __init__(self, K, H, W, voxel_size=0.01, sdf_trunc=0.04, depth_max=5.0, min_depth=0.1,
                 block_capacity=400_000, icp_subsample=4, icp_max_corr=0.05, icp_max_iter=30,
                 weight_threshold=1.0, motion_model="constant_velocity", min_points=100, ctx=None):
        self.ctx = ctx or get_context()
        self.K = tuple(float(k) for k in K)
        self.H, self.W = int(H), int(W)
        self.depth_max, self.min_depth = float(depth_max), float(min_depth)
        self.icp_subsample, self.icp_max_corr, self.icp_max_iter = int(icp_subsample), float(icp_max_corr), int(icp_max_iter)
        self.weight_threshold = float(weight_threshold)
        self.motion_model = motion_model
        self.min_points = int(min_points)
        self.volume = TSDFVolume(voxel_size, sdf_trunc, block_capacity, ctx=self.ctx)
        self.poses = []          # world->camera 4x4 per fused frame
        self.icp_log = []        # IcpOutput | None per frame
        self._tgt = {}           # reusable target buffers
        self._src = None
        # a private (capturable) stream: the registration's launch-bound inner loop runs as a cached
        # CUDA graph, which the legacy default stream cannot capture
        self.stream = __import__("torch").cuda.Stream(device=self.ctx.device)
        self.stage_ms = None     # set to {} to collect synchronised per-stage wall times (profiling aid)

    def reset(self):
        """Forget the model and the trajectory (buffers are kept)."""
        torch = __import__("torch")
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            self.volume.reset()
        torch.cuda.current_stream().wait_stream(self.stream)
        self.poses, self.icp_log = [], []

    def _stage(self, name, t0):
        if self.stage_ms is not None:
            import time
            __import__("torch").cuda.synchronize()
            self.stage_ms[name] = self.stage_ms.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
            return time.perf_counter()
        return t0

    def _source_cloud(self, depth):
        """K1 in the camera frame into a reusable buffer; the point count stays on the device."""
        torch = __import__("torch")
        s = self.icp_subsample
        cap = (-(-self.H // s)) * (-(-self.W // s))
        if self._src is None:
            self._src = (torch.empty((cap, 3), dtype=torch.float32, device=self.ctx.device),
                         torch.zeros(1, dtype=torch.int64, device=self.ctx.device))
        xyz, n = self._src
        self.ctx.backproject(depth, None, fx=self.K[0], fy=self.K[1], cx=self.K[2], cy=self.K[3], subsample=s,
                             min_depth=self.min_depth, max_depth=self.depth_max, pose=None, out_xyz=xyz, out_n=n)
        return xyz, n

    def track(self, depth, guess):
        """ICP of one frame against the model seen from `guess` (world->camera 4x4).
        Returns (T_cw, IcpOutput | None).  The model surface (K6), the frame's cloud (K1) and the whole
        registration (K8) are enqueued back to back; cloud sizes stay in device memory and the host
        synchronises once, on the registration result."""
        import time
        vol = self.volume
        while True:
            t0 = time.perf_counter()
            vol.ensure_view_buffers(self._tgt, int(self._tgt.get("cap", 1 << 18)))
            vol.extract_points_view_async(self.K, guess, self.H, self.W, self.depth_max, self.weight_threshold, self._tgt)
            t0 = self._stage("extract_view", t0)
            src, n_src = self._source_cloud(depth)
            t0 = self._stage("backproject", t0)
            res, skipped = self.ctx.icp_point_to_plane_dev(src, n_src, self._tgt["xyz"], self._tgt["nrm"], self._tgt["n"],
                                                           self.icp_max_corr, init=np.linalg.inv(guess),
                                                           max_iter=self.icp_max_iter, min_points=self.min_points)
            self._stage("icp", t0)
            n_tgt = int(self._tgt["n"].item())          # stream is idle here: the ICP result was just read
            if n_tgt <= self._tgt["cap"]:
                break
            self._tgt["cap"] = int(n_tgt * 1.25) + 1024   # surface grew past the buffer: enlarge, redo this frame
            self._tgt["xyz"] = None
        self.last_target = (n_tgt, None)
        if skipped:
            return guess, None
        return np.linalg.inv(res.transformation), res

    def add_frame(self, depth, bgr, init_pose=None, known_pose=None):
        """Track (unless known_pose is given) and fuse one frame.  depth (H,W) f32 and
        bgr (H,W,3) u8 are CUDA tensors.  Returns the frame's world->camera 4x4."""
        import time
        torch = __import__("torch")
        from .runtime import _check_frame
        if _check_frame(depth, bgr, None, "add_frame") != (self.H, self.W):
            raise ValueError(f"add_frame: frame is {tuple(depth.shape)}, the tracker was built for {(self.H, self.W)}")
        caller = torch.cuda.current_stream()
        self.stream.wait_stream(caller)                 # the frame was produced on the caller's stream
        depth.record_stream(self.stream)
        if bgr is not None:
            bgr.record_stream(self.stream)
        res = None
        with torch.cuda.stream(self.stream):
            if known_pose is not None:
                T = as4x4(known_pose)
            elif not self.poses:
                T = as4x4(init_pose) if init_pose is not None else np.eye(4)
            else:
                guess = as4x4(init_pose) if init_pose is not None else predict_pose(
                    self.poses[-1], self.poses[-2] if len(self.poses) > 1 else None, self.motion_model)
                T, res = self.track(depth, guess)
            t0 = time.perf_counter()
            self.volume.integrate(depth, bgr, self.K, T, depth_scale=1.0, depth_max=self.depth_max)
            self._stage("integrate", t0)
        caller.wait_stream(self.stream)                 # later work of the caller sees the updated model
        self.poses.append(T)
        self.icp_log.append(res)
        return T
