"""NumPy restatement of the reference's own back-projection arithmetic (K1).

TEST INFRASTRUCTURE ONLY — only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.

Pinned: tests/test_oracle_golden.py checks every function here against
tests/golden/k1_*.npz, which oracle/gen_golden.py produced by importing the
UNMODIFIED reference modules from /root/reference in the build container.

Each function states the NumPy expression sequence of the reference (file:line
relative to the upstream repo) so that dtype promotion — the only subtle part —
is reproduced by NumPy itself rather than re-derived:

* d2r = depth_to_reconstruction.py, der = depth_enhanced_reconstruction.py,
  dp = depth_processor.py
"""
from __future__ import annotations

import numpy as np


def projection_factors(h: int, w: int, fx: float, fy: float, cx: float, cy: float):
    """x_factor=(u-cx)/fx, y_factor=(v-cy)/fy as float64 HxW maps (d2r:287-295, der:545-552)."""
    cols, rows = np.meshgrid(np.arange(w), np.arange(h))
    return (cols - cx) / fx, (rows - cy) / fy


def _finish(depth_s, color, xf, yf, lo, hi, pose):
    keep = (depth_s > lo) & (depth_s < hi) & np.isfinite(depth_s)          # d2r:359-361
    zc = depth_s[keep]
    cam = np.stack([xf[keep] * zc, yf[keep] * zc, zc], axis=-1)            # d2r:364-368
    if pose is not None:
        rot, trans = pose
        world = (rot.T @ cam.T).T - (rot.T @ trans).ravel()                 # d2r:376 / der:605
    else:
        world = cam
    rgb = color[keep][:, ::-1]                                              # d2r:381-382
    return world.astype(np.float32), rgb.astype(np.uint8)                   # d2r:384


def d2r_depth_to_pointcloud(depth, color, fx, fy, cx, cy, pose=None, scale=1.0, subsample=1,
                            min_depth=0.1, max_depth=50.0):
    """DenseReconstructor.depth_to_pointcloud (d2r:328-384).

    `depth * scale` keeps float32 for a Python-float scale and becomes float64 for
    an np.float64 scale (d2r:356, NEP 50); the thresholds follow that dtype.
    """
    h, w = depth.shape
    xf, yf = projection_factors(h, w, fx, fy, cx, cy)
    if subsample > 1:                                                       # d2r:348-353
        sl = (slice(None, None, subsample), slice(None, None, subsample))
        depth, color, xf, yf = depth[sl], color[sl], xf[sl], yf[sl]
    return _finish(depth * scale, color, xf, yf, min_depth, max_depth, pose)


def der_depth_to_pointcloud(depth, color, fx, fy, cx, cy, pose=None, min_depth=0.1,
                            max_depth=100.0, subsample=1):
    """DensePointCloudGenerator.depth_to_pointcloud (der:554-613): no scale argument,
    the caller pre-multiplies (der:1135)."""
    h, w = depth.shape
    xf, yf = projection_factors(h, w, fx, fy, cx, cy)
    if subsample > 1:
        sl = (slice(None, None, subsample), slice(None, None, subsample))
        depth, color, xf, yf = depth[sl], color[sl], xf[sl], yf[sl]
    return _finish(depth, color, xf, yf, min_depth, max_depth, pose)


def dp_generate(depth, rgb, fx, fy, cx, cy, downsample=1, max_depth=100.0, min_depth=0.1):
    """PointCloudGenerator.generate (dp:371-422): strided grids, no pose, colours as
    float32 in [0,1] (u8.astype(f32)/255 then channel reversal, dp:417-420)."""
    h, w = depth.shape
    cols, rows = np.meshgrid(np.arange(0, w, downsample), np.arange(0, h, downsample))  # dp:361-363
    xn = (cols - cx) / fx
    yn = (rows - cy) / fy
    if downsample > 1:
        depth = depth[::downsample, ::downsample]
    keep = (depth > min_depth) & (depth < max_depth) & np.isfinite(depth)   # dp:401
    pts = np.stack([xn * depth, yn * depth, depth], axis=-1)[keep]          # dp:404-410
    cols_out = None
    if rgb is not None:
        if downsample > 1:
            rgb = rgb[::downsample, ::downsample]
        cols_out = rgb[keep].astype(np.float32) / 255.0
        if cols_out.shape[1] == 3:
            cols_out = cols_out[:, ::-1]
    return pts.astype(np.float32), cols_out


def merge_without_open3d(clouds):
    """merge_pointclouds with O3D_AVAILABLE=False (d2r:386-420): skip empties, vstack."""
    pts = [p for p, _ in clouds if len(p) > 0]
    cols = [c for p, c in clouds if len(p) > 0]
    if not pts:
        return np.array([]), np.array([])
    return np.vstack(pts), np.vstack(cols)


def ascii_ply_lines(points, colors):
    """The fallback writer's body (d2r:700-701): str() of NumPy scalars."""
    return [f"{p[0]} {p[1]} {p[2]} {c[0]} {c[1]} {c[2]}\n" for p, c in zip(points, colors)]


ASCII_PLY_HEADER = (  # d2r:690-699
    "ply\nformat ascii 1.0\nelement vertex {n}\nproperty float x\nproperty float y\n"
    "property float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n"
)


def intrinsics_from_json_dict(data: dict) -> dict:
    """CameraIntrinsics.from_json key handling (dp:89-102).  `width`/`height` are
    required even when cx/cy are present, because the defaults are evaluated eagerly."""
    return dict(
        fx=data.get("fx", data.get("focal_length_x", 470.4)),
        fy=data.get("fy", data.get("focal_length_y", 470.4)),
        cx=data.get("cx", data.get("principal_point_x", data["width"] / 2)),
        cy=data.get("cy", data.get("principal_point_y", data["height"] / 2)),
        width=data["width"],
        height=data["height"],
        depth_scale=data.get("depth_scale", 1.0),
    )
