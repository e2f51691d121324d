"""Drop-in for the dense path of the reference's depth_enhanced_reconstruction.py.

CameraIntrinsics (der:57-80) and DensePointCloudGenerator (der:535-645) keep their
signatures.  DepthEnhancedReconstruction (der:896-1311) keeps load_images /
reconstruct(output_dir) / reconstruction.ply, but — as north_star asks — poses come
from frame-to-model point-to-plane ICP and the merge is TSDF fusion instead of the
reference's SIFT/essential-matrix odometry + vstack.  Depth maps are an input
(`depths=` / a depth folder); the Depth-Anything network (der:87-171) is out of scope.
"""
from __future__ import annotations

import argparse
from dataclasses import dataclass
from pathlib import Path
from typing import List, Tuple

import numpy as np

from . import _lib
from .depth_to_reconstruction import DepthImageLoader
from .runtime import TSDFVolume, get_context, to_host, write_ply


@dataclass
class CameraIntrinsics:
    """der:57-80."""
    fx: float
    fy: float
    cx: float
    cy: float
    width: int
    height: int

    def to_matrix(self) -> np.ndarray:
        return np.array([[self.fx, 0, self.cx], [0, self.fy, self.cy], [0, 0, 1]], dtype=np.float64)

    @classmethod
    def from_matrix(cls, K: np.ndarray, width: int, height: int) -> "CameraIntrinsics":
        return cls(fx=K[0, 0], fy=K[1, 1], cx=K[0, 2], cy=K[1, 2], width=width, height=height)


class DensePointCloudGenerator:
    """der:535-645."""

    def __init__(self, intrinsics: CameraIntrinsics):
        self.K = intrinsics
        self._ctx = None

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = get_context()
        return self._ctx

    def depth_to_pointcloud_device(self, depth, color, pose=None, min_depth=0.1, max_depth=100.0, subsample=1):
        xyz, rgb, n = self.ctx.backproject(depth, color, fx=self.K.fx, fy=self.K.fy, cx=self.K.cx, cy=self.K.cy,
                                           subsample=subsample, min_depth=min_depth, max_depth=max_depth, pose=pose)
        k = int(n.item())
        return xyz[:k], rgb[:k]

    def depth_to_pointcloud(self, depth: np.ndarray, color: np.ndarray,
                            pose: Tuple[np.ndarray, np.ndarray] = None, min_depth: float = 0.1,
                            max_depth: float = 100.0, subsample: int = 1) -> Tuple[np.ndarray, np.ndarray]:
        """der:554-613 — the caller pre-scales depth (der:1135), so an f64 depth array keeps
        f64 mask arithmetic and an f32 one keeps f32 thresholds."""
        import torch
        dev = self.ctx.device
        d = np.ascontiguousarray(depth)
        if d.dtype not in (np.float32, np.float64):
            d = d.astype(np.float32)
        xyz, rgb = self.depth_to_pointcloud_device(
            torch.from_numpy(d).to(dev), torch.from_numpy(np.ascontiguousarray(color, np.uint8)).to(dev),
            pose=pose, min_depth=min_depth, max_depth=max_depth, subsample=subsample)
        return to_host(xyz), to_host(rgb)

    def merge_pointclouds(self, pointclouds: List[Tuple[np.ndarray, np.ndarray]],
                          voxel_size: float = 0.01) -> Tuple[np.ndarray, np.ndarray]:
        """der:615-645 — vstack + voxel_down_sample only (no outlier removal in this copy)."""
        import torch
        pts = [p for p, _ in pointclouds if len(p) > 0]
        cols = [c for p, c in pointclouds if len(p) > 0]
        if not pts:
            return np.array([]), np.array([])
        points, colors = np.vstack(pts), np.vstack(cols)
        if voxel_size > 0:
            dev = self.ctx.device
            ds = self.ctx.voxel_downsample(torch.from_numpy(np.ascontiguousarray(points)).to(dev),
                                           torch.from_numpy(np.ascontiguousarray(colors, np.uint8)).to(dev),
                                           voxel_size, sorted_output=True, want_idx=False)
            points, colors = to_host(ds["points"]), to_host(ds["colors"])
        return points, colors


class DepthScaleEstimator:
    """der:652-697 — median of Z/d over sparse points (no sanity gate in this copy; at least 5
    sparse points and 3 accepted samples, else 1.0)."""

    @staticmethod
    def estimate_scale(sparse_points: np.ndarray, sparse_pts2d: np.ndarray, depth_map, K: np.ndarray = None) -> float:
        import torch
        ctx = get_context()
        d = depth_map if isinstance(depth_map, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(depth_map, np.float32)).to(ctx.device)
        scale, n = ctx.estimate_scale(d.contiguous(), sparse_points, sparse_pts2d, gate=False, min_input_points=5)
        if n < 3:
            return 1.0
        scale = np.float64(scale)
        print(f"  Depth scale: {scale:.4f} (from {n} points)")
        return scale


class DepthEnhancedReconstruction:
    """der:896-1311 with the pose step = frame-to-model ICP and the merge = TSDF fusion."""

    def __init__(self, K: np.ndarray, image_size: Tuple[int, int] = None, use_depth: bool = True,
                 use_hybrid_features: bool = True, voxel_size: float = 0.01, sdf_trunc: float = 0.04,
                 depth_max: float = 5.0, block_capacity: int = 200_000, icp_subsample: int = 4,
                 icp_max_corr: float = 0.05, ply_layout=_lib.PLY_O3D_BINARY):
        self.K = np.asarray(K, np.float64)
        self.image_size = image_size
        if image_size:
            self.intrinsics = CameraIntrinsics.from_matrix(self.K, image_size[0], image_size[1])
        self.use_depth = use_depth
        self.use_hybrid = use_hybrid_features          # accepted for signature parity; no feature front-end
        self.images, self.image_names, self.depths = [], [], []
        self.camera_poses = []
        self.voxel_size, self.sdf_trunc, self.depth_max = voxel_size, sdf_trunc, depth_max
        self.block_capacity = block_capacity
        self.icp_subsample, self.icp_max_corr = icp_subsample, icp_max_corr
        self.ply_layout = ply_layout
        self.icp_log = []
        print("Reconstruction system initialized")

    def load_images(self, folder_path: str, extensions: List[str] = [".png", ".jpg", ".jpeg"]) -> int:
        """der:944-967."""
        import cv2
        folder = Path(folder_path)
        files = sorted(f for f in folder.iterdir() if f.suffix.lower() in extensions and "_depth" not in f.stem)
        print(f"Found {len(files)} images")
        for p in files:
            img = cv2.imread(str(p))
            if img is not None:
                self.images.append(img)
                self.image_names.append(p.name)
                print(f"  Loaded: {p.name} - Shape: {img.shape}")
        if self.images and self.image_size is None:
            h, w = self.images[0].shape[:2]
            self.image_size = (w, h)
            self.intrinsics = CameraIntrinsics.from_matrix(self.K, w, h)
        return len(self.images)

    def load_depths(self, depth_folder: str) -> int:
        """Depth maps replace the network (north_star): same naming rules as d2r:100-119."""
        self.depths = []
        for name in self.image_names:
            f = DepthImageLoader.find_matching_depth(name, Path(depth_folder))
            self.depths.append(None if f is None else DepthImageLoader.load_depth(f))
        return sum(d is not None for d in self.depths)

    def set_frames(self, images, depths, names=None):
        self.images, self.depths = list(images), list(depths)
        self.image_names = names or [f"frame_{i:04d}" for i in range(len(images))]
        if self.images and self.image_size is None:
            h, w = self.images[0].shape[:2]
            self.image_size = (w, h)
            self.intrinsics = CameraIntrinsics.from_matrix(self.K, w, h)

    def reconstruct(self, output_dir: str = "./output", init_poses=None):
        """Frame-to-model tracking + fusion.  Returns (points, colors, poses) or None.
        init_poses: optional list of world->camera 4x4 used as ICP initial guesses
        (default: previous frame's pose)."""
        import torch
        print("\n" + "=" * 70)
        print("STARTING DEPTH-ENHANCED 3D RECONSTRUCTION")
        print("=" * 70)
        out = Path(output_dir)
        out.mkdir(parents=True, exist_ok=True)
        if len(self.images) < 2 or len(self.depths) != len(self.images):
            print("Failed to initialize - need >= 2 images with depth")
            return None
        ctx = get_context()
        dev = ctx.device
        K4 = (self.K[0, 0], self.K[1, 1], self.K[0, 2], self.K[1, 2])
        h, w = self.images[0].shape[:2]
        from .tracking import FrameToModelTracker
        tracker = FrameToModelTracker(K4, h, w, voxel_size=self.voxel_size, sdf_trunc=self.sdf_trunc,
                                      depth_max=self.depth_max, block_capacity=self.block_capacity,
                                      icp_subsample=self.icp_subsample, icp_max_corr=self.icp_max_corr, ctx=ctx)
        vol = tracker.volume
        self.camera_poses, self.icp_log = [], []
        for i, (img, depth) in enumerate(zip(self.images, self.depths)):
            if depth is None:
                continue
            d = torch.from_numpy(np.ascontiguousarray(depth, np.float32)).to(dev)
            if img.shape[:2] != (h, w):
                raise ValueError(f"image {i} is {img.shape[:2]}, the first image is {(h, w)}")
            if tuple(d.shape) != (h, w):
                # the reference's estimate() resizes the network output to the image (der:157-166); depth files of
                # another size get the same treatment as d2r:465-467 (cv2.INTER_LINEAR semantics, on the GPU)
                d = ctx.resize_bilinear(d, h, w)
            c = torch.from_numpy(np.ascontiguousarray(img, np.uint8)).to(dev)
            T_cw = tracker.add_frame(d, c, init_pose=None if init_poses is None else init_poses[i])
            if tracker.icp_log[-1] is not None:
                self.icp_log.append(tracker.icp_log[-1])
            self.camera_poses.append((T_cw[:3, :3].copy(), T_cw[:3, 3:4].copy()))
        self.tracker = tracker
        pts, _, cols = vol.extract_points(weight_threshold=1.0, with_normals=False)
        all_points, all_colors = to_host(pts), to_host(cols)
        self.volume = vol
        print("\n" + "=" * 70)
        print("RECONSTRUCTION COMPLETE")
        print("=" * 70)
        print(f"Total points: {len(all_points)}")
        print(f"Total cameras: {len(self.camera_poses)}")
        self._save_pointcloud(all_points, all_colors, out / "reconstruction.ply")
        return all_points, all_colors, self.camera_poses

    def save_mesh(self, filepath, weight_threshold: float = 1.0):
        """(extension) Triangle mesh of the fused volume -> .ply in Open3D's write_triangle_mesh
        layout (K10 marching cubes; north_star "surface/point extraction to .ply").  Call after
        reconstruct().  Returns (n_vertices, n_triangles)."""
        from .runtime import write_ply_mesh
        v, n, c, t = self.volume.extract_mesh(weight_threshold)
        if t.shape[0] == 0:
            print("No surface to save")
            return 0, 0
        write_ply_mesh(filepath, v, t, colors=c, normals=n)
        print(f"Saved mesh with {v.shape[0]} vertices and {t.shape[0]} triangles to {filepath}")
        return int(v.shape[0]), int(t.shape[0])

    def _save_pointcloud(self, points: np.ndarray, colors: np.ndarray, filepath: Path):
        """der:1283-1311."""
        if len(points) == 0:
            print("No points to save")
            return
        write_ply(filepath, points, colors, layout=self.ply_layout)
        print(f"Saved {len(points)} points to {filepath}")


def main(argv=None):
    """der:1418-1468 — same flags (+ --depth-folder, since the depth network is out of scope)."""
    parser = argparse.ArgumentParser(description="Depth-Enhanced 3D Reconstruction")
    parser.add_argument("--input", type=str, default="./input_folder/buddha_images", help="Input folder with images")
    parser.add_argument("--output", type=str, default="./output", help="Output directory")
    parser.add_argument("--fx", type=float, default=1719.0, help="Focal length X")
    parser.add_argument("--fy", type=float, default=1719.0, help="Focal length Y")
    parser.add_argument("--cx", type=float, default=540.0, help="Principal point X")
    parser.add_argument("--cy", type=float, default=960.0, help="Principal point Y")
    parser.add_argument("--no-depth", action="store_true", help="Disable depth estimation")
    parser.add_argument("--no-hybrid", action="store_true", help="Disable hybrid features")
    parser.add_argument("--depth-folder", type=str, default=None, help="(extension) folder with depth maps")
    parser.add_argument("--mesh", action="store_true",
                        help="(extension) also write <output>/reconstruction_mesh.ply (marching-cubes triangle mesh)")
    args = parser.parse_args(argv)
    K = np.array([[args.fx, 0, args.cx], [0, args.fy, args.cy], [0, 0, 1]], dtype=np.float64)
    rec = DepthEnhancedReconstruction(K=K, use_depth=not args.no_depth, use_hybrid_features=not args.no_hybrid)
    if rec.load_images(args.input) < 2:
        print("Need at least 2 images for reconstruction")
        raise SystemExit(1)
    rec.load_depths(args.depth_folder or args.input)
    if rec.reconstruct(output_dir=args.output) is None:
        print("Reconstruction failed")
    elif args.mesh:
        rec.save_mesh(Path(args.output) / "reconstruction_mesh.ply")


if __name__ == "__main__":
    main()
