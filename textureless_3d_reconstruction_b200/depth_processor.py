"""Drop-in for the geometric part of the reference's depth_processor.py:
CameraIntrinsics (dp:78-135) and PointCloudGenerator (dp:339-450).  The depth
network, image sources and ROS 2 publisher (dp:138-336, 453-792) are out of scope.
"""
from __future__ import annotations

import json
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from . import _lib
from .runtime import get_context, to_host, write_ply


@dataclass
class CameraIntrinsics:
    """dp:78-135 — same fields, JSON key aliases and defaults."""
    fx: float
    fy: float
    cx: float
    cy: float
    width: int
    height: int
    depth_scale: float = 1.0

    @classmethod
    def from_json(cls, json_path: str) -> "CameraIntrinsics":
        with open(json_path, "r") as f:
            data = json.load(f)
        # width/height are required even if cx/cy are given: the reference evaluates
        # data['width'] / 2 eagerly as the .get() default (dp:98-99) -> KeyError.
        width, height = data["width"], data["height"]
        return cls(
            fx=data.get("fx", data.get("focal_length_x", 470.4)),
            fy=data.get("fy", data.get("focal_length_y", 470.4)),
            cx=data.get("cx", data.get("principal_point_x", width / 2)),
            cy=data.get("cy", data.get("principal_point_y", height / 2)),
            width=width, height=height, depth_scale=data.get("depth_scale", 1.0))

    @classmethod
    def default(cls, width: int = 640, height: int = 480) -> "CameraIntrinsics":
        return cls(fx=width * 0.8, fy=width * 0.8, cx=width / 2, cy=height / 2, width=width, height=height)

    @classmethod
    def realsense_d455(cls) -> "CameraIntrinsics":
        return cls(fx=382.193, fy=382.193, cx=320.819, cy=237.683, width=640, height=480, depth_scale=0.001)

    def to_matrix(self) -> np.ndarray:
        return np.array([[self.fx, 0, self.cx], [0, self.fy, self.cy], [0, 0, 1]], dtype=np.float64)


class PointCloudGenerator:
    """dp:339-450 — generate(depth, rgb, max_depth, min_depth) -> (N×3 f32, N×3 f32 in [0,1] | None)."""

    def __init__(self, intrinsics: CameraIntrinsics, downsample_factor: int = 1):
        self.intrinsics = intrinsics
        self.downsample = downsample_factor
        self._ctx = None

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = get_context()
        return self._ctx

    def generate_device(self, depth, rgb=None, max_depth=100.0, min_depth=0.1):
        K = self.intrinsics
        xyz, cols, n = self.ctx.backproject(depth, rgb, fx=K.fx, fy=K.fy, cx=K.cx, cy=K.cy,
                                            subsample=self.downsample, min_depth=min_depth, max_depth=max_depth,
                                            rgb_out_f32=True)
        k = int(n.item())
        return xyz[:k], (cols[:k] if cols is not None else None)

    def generate(self, depth: np.ndarray, rgb: Optional[np.ndarray] = None, max_depth: float = 100.0,
                 min_depth: float = 0.1) -> Tuple[np.ndarray, Optional[np.ndarray]]:
        import torch
        dev = self.ctx.device
        d = np.ascontiguousarray(depth)
        if d.dtype not in (np.float32, np.float64):
            d = d.astype(np.float32)
        c = None if rgb is None else torch.from_numpy(np.ascontiguousarray(rgb, np.uint8)).to(dev)
        xyz, cols = self.generate_device(torch.from_numpy(d).to(dev), c, max_depth, min_depth)
        return to_host(xyz), to_host(cols)

    def save_ply(self, filepath: str, points: np.ndarray, colors: Optional[np.ndarray] = None):
        """dp:424-440 — Open3D layout; colours arrive as floats in [0,1] here."""
        cols = None
        if colors is not None:
            cols = np.clip(np.asarray(colors, np.float64), 0.0, 1.0)
            cols = np.floor(cols * 255.0 + 0.5).astype(np.uint8)     # Open3D ColorToUint8 (R9)
        write_ply(filepath, points, cols, layout=_lib.PLY_O3D_BINARY)

    def pointcloud2_records(self, points, colors):
        """The PointCloud2 payload of ROS2DepthPublisher.publish_pointcloud (dp:726-764): one
        16-byte record x, y, z, rgb per point with rgb = bytes (b, g, r, 0) viewed as float32 —
        packed by one kernel instead of the reference's per-point Python loop (dp:750-756).
        points (N,3) f32 / colors (N,3) f32 in [0,1]; host arrays or CUDA tensors.  Returns
        an (N,4) f32 array of the same kind."""
        import torch
        host = not isinstance(points, torch.Tensor)
        dev = self.ctx.device
        p = torch.from_numpy(np.ascontiguousarray(points, np.float32)).to(dev) if host else points.contiguous()
        c = torch.from_numpy(np.ascontiguousarray(colors, np.float32)).to(dev) if host else colors.contiguous()
        rec = self.ctx.pack_pointcloud2(p, c)
        return to_host(rec) if host else rec

    def save_pcd(self, filepath: str, points: np.ndarray, colors: Optional[np.ndarray] = None):
        self.save_ply(filepath.replace(".ply", ".pcd"), points, colors)   # dp:442-450 (same quirk)


def save_depth_files(depth, identifier: str, depth_dir, save_raw_depth: bool = True, ctx=None):
    """The depth outputs of DepthProcessor._save_depth (dp:905-921) that depth_to_reconstruction reads
    back: `{id}_depth.npy` (float32 metres) and `{id}_depth.png` (16-bit millimetres,
    `(depth*1000).astype(uint16)` computed on the GPU).  depth: (H,W) f32 host array or CUDA tensor."""
    import cv2
    import torch
    from pathlib import Path
    ctx = ctx or get_context()
    d = depth if isinstance(depth, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(depth, np.float32)).to(ctx.device)
    depth_dir = Path(depth_dir)
    depth_dir.mkdir(parents=True, exist_ok=True)
    if save_raw_depth:
        np.save(depth_dir / f"{identifier}_depth.npy", d.cpu().numpy())
    mm = ctx.depth_f32_to_u16(d.contiguous()).cpu().numpy()
    cv2.imwrite(str(depth_dir / f"{identifier}_depth.png"), mm)
