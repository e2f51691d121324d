"""Drop-in for the dense path of the reference's depth_to_reconstruction.py.

Same class names, signatures, defaults, printed lines, CLI flags and PLY layout
(reference file:line cited per item); the bodies run on the B200 through
libt3d.so.  What is NOT carried over: the SIFT/essential-matrix pose front-end
(SparseReconstructor, d2r:122-271) — poses come from `poses=` / a poses file or
from point-to-plane ICP registration of consecutive frames (north_star).
"""
from __future__ import annotations

import argparse
import json
from dataclasses import dataclass
from pathlib import Path
from typing import List, Optional, Tuple

import numpy as np

from . import _lib
from .runtime import get_context, to_host, write_ply


@dataclass
class ReconstructionConfig:
    """d2r:45-73 — identical fields and defaults."""
    fx: float = 1719.0
    fy: float = 1719.0
    cx: float = 540.0
    cy: float = 960.0
    depth_scale: float = 1.0
    min_depth: float = 0.1
    max_depth: float = 50.0
    match_ratio: float = 0.75
    ransac_threshold: float = 3.0
    voxel_size: float = 0.005
    subsample_factor: int = 2

    @property
    def K(self) -> np.ndarray:
        return np.array([[self.fx, 0, self.cx], [0, self.fy, self.cy], [0, 0, 1]], dtype=np.float64)


class DepthImageLoader:
    """d2r:76-119 — .npy -> f32; 16-bit .png -> mm/1000; .exr; same name search order."""

    @staticmethod
    def load_depth(filepath: Path) -> Optional[np.ndarray]:
        filepath = Path(filepath)
        if filepath.suffix == ".npy":
            return np.load(str(filepath)).astype(np.float32)
        if filepath.suffix == ".png":
            import cv2
            raw = cv2.imread(str(filepath), cv2.IMREAD_ANYDEPTH)
            if raw is not None:
                return raw.astype(np.float32) / 1000.0
        elif filepath.suffix in (".exr", ".EXR"):
            import cv2
            raw = cv2.imread(str(filepath), cv2.IMREAD_ANYDEPTH)
            if raw is not None:
                return raw.astype(np.float32)
        return None

    @staticmethod
    def load_depth_device(filepath: Path, ctx=None):
        """Same files, decoded on the GPU: a 16-bit millimetre PNG is uploaded raw (2 B/pixel)
        and divided by 1000 there (t3d_depth_u16_to_f32) — bit-identical to load_depth().
        Returns an (H,W) f32 CUDA tensor or None."""
        import torch
        filepath = Path(filepath)
        ctx = ctx or get_context()
        if filepath.suffix == ".png":
            import cv2
            raw = cv2.imread(str(filepath), cv2.IMREAD_ANYDEPTH)
            if raw is None:
                return None
            if raw.dtype == np.uint16:
                return ctx.depth_u16_to_f32(torch.from_numpy(raw.view(np.int16)).to(ctx.device), 1000.0)
            return torch.from_numpy(raw.astype(np.float32) / 1000.0).to(ctx.device)
        d = DepthImageLoader.load_depth(filepath)
        return None if d is None else torch.from_numpy(np.ascontiguousarray(d)).to(ctx.device)

    @staticmethod
    def find_matching_depth(rgb_name: str, depth_folder: Path) -> Optional[Path]:
        stem = Path(rgb_name).stem
        for cand in (f"{stem}_depth.npy", f"{stem}_depth.png", f"{stem}.npy", f"{stem}.png",
                     f"depth_{stem}.npy", f"depth_{stem}.png"):
            p = Path(depth_folder) / cand
            if p.exists():
                return p
        return None


def _scale_is_f64(scale) -> bool:
    """NumPy promotion of `depth_f32 * scale` (d2r:356): a Python scalar is weak (stays
    f32); an np.float64 / 0-d f64 array promotes the whole expression to f64."""
    return isinstance(scale, (np.floating, np.ndarray)) and np.asarray(scale).dtype == np.float64


class DenseReconstructor:
    """d2r:274-420.  depth_to_pointcloud / merge_pointclouds keep the reference's
    host-NumPy in/out contract; *_device variants keep everything on the GPU."""

    def __init__(self, config: ReconstructionConfig):
        self.config = config
        self.K = config.K
        self._ctx = None

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = get_context()
        return self._ctx

    def estimate_scale(self, sparse_points, sparse_pts2d, depth_map) -> float:
        """d2r:297-326 — median of Z/d over the sparse points with the 0.001 < s < 1000 gate
        (t3d_estimate_scale; depth_map may be a host array or an (H,W) f32 CUDA tensor)."""
        import torch
        if min(len(sparse_points), len(sparse_pts2d)) < 3:     # cannot reach 3 samples: no device work
            print("Warning: Too few scale samples, using default scale=1.0")
            return 1.0
        d = depth_map if isinstance(depth_map, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(depth_map, np.float32)).to(self.ctx.device)
        scale, n = self.ctx.estimate_scale(d.contiguous(), sparse_points, sparse_pts2d, gate=True)
        if n < 3:
            print("Warning: Too few scale samples, using default scale=1.0")
            return 1.0
        scale = np.float64(scale)          # np.median returns np.float64: keeps `depth * scale` in f64 (d2r:356)
        print(f"Estimated depth scale: {scale:.6f} (from {n} samples)")
        return scale

    def depth_to_pointcloud_device(self, depth, color, pose=None, scale=1.0, subsample=1):
        """Device in / device out: returns (xyz f32 [N,3], rgb u8 [N,3]) CUDA tensors."""
        xyz, rgb, n = self.ctx.backproject(
            depth, color, fx=self.config.fx, fy=self.config.fy, cx=self.config.cx, cy=self.config.cy,
            subsample=subsample, scale=float(scale), scale_is_f64=_scale_is_f64(scale),
            min_depth=self.config.min_depth, max_depth=self.config.max_depth, pose=pose)
        k = int(n.item())
        return xyz[:k], rgb[:k]

    def depth_to_pointcloud(self, depth: np.ndarray, color: np.ndarray,
                            pose: Tuple[np.ndarray, np.ndarray] = None, scale: float = 1.0,
                            subsample: int = 1) -> Tuple[np.ndarray, np.ndarray]:
        """d2r:328-384 — same arguments, returns host (N×3 f32, N×3 u8 RGB)."""
        import torch
        dev = self.ctx.device
        d = np.ascontiguousarray(depth)
        if d.dtype not in (np.float32, np.float64):
            d = d.astype(np.float32)
        xyz, rgb = self.depth_to_pointcloud_device(
            torch.from_numpy(d).to(dev), torch.from_numpy(np.ascontiguousarray(color, np.uint8)).to(dev),
            pose=pose, scale=scale, subsample=subsample)
        return to_host(xyz), to_host(rgb)

    def merge_pointclouds_device(self, points, colors, voxel_size=0.005, remove_outliers=True):
        """points f32|f64 [N,3], colors u8 [N,3] CUDA tensors -> (points f64, colors u8)."""
        ds = self.ctx.voxel_downsample(points, colors, voxel_size, sorted_output=True, want_idx=False)
        pts, cols = ds["points"].contiguous(), ds["colors"].contiguous()
        if remove_outliers and pts.shape[0] > 0:
            keep, _, _, _ = self.ctx.statistical_outlier(pts, 20, 2.0)          # d2r:413-415
            pts = self.ctx.compact_rows(pts, keep)
            cols = self.ctx.compact_rows(cols, keep)
        return pts, cols

    def merge_pointclouds(self, clouds: List[Tuple[np.ndarray, np.ndarray]],
                          voxel_size: float = 0.005) -> Tuple[np.ndarray, np.ndarray]:
        """d2r:386-420 with Open3D present: vstack -> voxel_down_sample -> statistical
        outlier removal (20, 2.0) -> (M×3 f64, M×3 u8)."""
        import torch
        pts = [p for p, _ in clouds if len(p) > 0]
        cols = [c for p, c in clouds if len(p) > 0]
        if not pts:
            return np.array([]), np.array([])
        points, colors = np.vstack(pts), np.vstack(cols)
        if voxel_size > 0:
            dev = self.ctx.device
            p, c = self.merge_pointclouds_device(
                torch.from_numpy(np.ascontiguousarray(points)).to(dev),
                torch.from_numpy(np.ascontiguousarray(colors, np.uint8)).to(dev), voxel_size)
            points, colors = to_host(p), to_host(c)
        return points, colors


def load_poses(path) -> List[Tuple[np.ndarray, np.ndarray]]:
    """poses file: .npy of shape (N,4,4)/(N,3,4) world->camera, or JSON list of 4x4."""
    path = Path(path)
    arr = np.load(path) if path.suffix == ".npy" else np.array(json.loads(path.read_text()), np.float64)
    return [(np.ascontiguousarray(T[:3, :3], np.float64), np.ascontiguousarray(T[:3, 3:4], np.float64))
            for T in arr]


class DepthToReconstructionPipeline:
    """d2r:423-703 — load_data / reconstruct / save_reconstruction."""

    def __init__(self, config: ReconstructionConfig = None, poses=None, ply_layout=_lib.PLY_O3D_BINARY):
        self.config = config or ReconstructionConfig()
        self.dense = DenseReconstructor(self.config)
        self.images, self.image_names, self.depths = [], [], []
        self.camera_poses = []
        self.given_poses = poses
        self.ply_layout = ply_layout

    def load_data(self, rgb_folder: str, depth_folder: str) -> int:
        """d2r:439-477 (same messages, same depth->RGB bilinear resize)."""
        import cv2
        rgb_path, depth_path = Path(rgb_folder), Path(depth_folder)
        rgb_files = sorted(f for f in rgb_path.iterdir() if f.suffix.lower() in (".png", ".jpg", ".jpeg"))
        print(f"Found {len(rgb_files)} RGB images")
        for rf in rgb_files:
            img = cv2.imread(str(rf))
            if img is None:
                continue
            df = DepthImageLoader.find_matching_depth(rf.name, depth_path)
            if df is None:
                print(f"  Warning: No depth found for {rf.name}")
                continue
            depth = DepthImageLoader.load_depth(df)
            if depth is None:
                continue
            if depth.shape[:2] != img.shape[:2]:          # d2r:465-467, cv2.INTER_LINEAR semantics on the GPU
                import torch
                ctx = self.dense.ctx
                depth = ctx.resize_bilinear(torch.from_numpy(np.ascontiguousarray(depth, np.float32)).to(ctx.device),
                                            img.shape[0], img.shape[1]).cpu().numpy()
            self.images.append(img)
            self.depths.append(depth)
            self.image_names.append(rf.name)
            print(f"  Loaded: {rf.name} with depth")
        print(f"Loaded {len(self.images)} image-depth pairs")
        return len(self.images)

    def _register(self, prev_cloud, cloud, T_init):
        """Pose of the new frame by point-to-plane ICP against the previous frame's cloud."""
        ctx = self.dense.ctx
        nrm = ctx.estimate_normals(prev_cloud, knn=30)
        res = ctx.icp_point_to_plane(cloud, prev_cloud, nrm, max_corr_dist=0.05 * max(1.0, self.config.voxel_size / 0.005),
                                     init=T_init, max_iter=30)
        return res

    def reconstruct(self):
        """d2r:479-671 — returns (points, colors, camera_poses) or (None, None, None)."""
        import torch
        if len(self.images) < 2:
            print("Need at least 2 images")
            return None, None, None
        print("\n" + "=" * 70)
        print("DEPTH-ENHANCED RECONSTRUCTION PIPELINE")
        print("=" * 70)
        ctx = self.dense.ctx
        dev = ctx.device
        s = self.config.subsample_factor
        all_xyz, all_rgb = [], []
        self.camera_poses = []
        prev_cam_cloud = None
        T_wc = np.eye(4)  # camera -> world of the current frame
        ds, cs = [], []
        for i, (img, depth) in enumerate(zip(self.images, self.depths)):
            d = torch.from_numpy(np.ascontiguousarray(depth, np.float32)).to(dev)
            c = torch.from_numpy(np.ascontiguousarray(img, np.uint8)).to(dev)
            ds.append(d)
            cs.append(c)
            if self.given_poses is not None:
                pose = self.given_poses[i]
            else:
                # ICP odometry on 1/4-resolution camera-frame clouds (north_star: ICP registration)
                cam, _ = self.dense.depth_to_pointcloud_device(d, c, pose=None, subsample=max(s, 4))
                cam = cam.contiguous()
                if prev_cam_cloud is not None and cam.shape[0] > 100 and prev_cam_cloud.shape[0] > 100:
                    res = self._register(prev_cam_cloud, cam, np.eye(4))
                    T_wc = T_wc @ res.transformation  # cam_i -> cam_{i-1} -> world
                prev_cam_cloud = cam
                T_cw = np.linalg.inv(T_wc)
                pose = (np.ascontiguousarray(T_cw[:3, :3]), np.ascontiguousarray(T_cw[:3, 3:4]))
            self.camera_poses.append(pose)
        same_size = all(d.shape == ds[0].shape for d in ds)
        if same_size:
            # the per-frame loop of d2r:566-658 as ONE batched call (32 frames per launch, frame-ordered output ==
            # the reference's list + np.vstack): no launch and no host sync per frame; the per-camera counts come
            # back with a single read of the offsets
            H, W = ds[0].shape
            frames = ctx.make_backproject_frames(ds, cs, self.camera_poses)
            xyz, rgb, offs = ctx.backproject_batch(
                frames, len(ds), H, W, fx=self.config.fx, fy=self.config.fy, cx=self.config.cx, cy=self.config.cy,
                subsample=s, scale=1.0, min_depth=self.config.min_depth, max_depth=self.config.max_depth)
            offs = offs.cpu().tolist()
            all_xyz, all_rgb = [xyz[:offs[-1]]], [rgb[:offs[-1]]]
            counts = [offs[i + 1] - offs[i] for i in range(len(ds))]
        else:
            counts = []
            for d, c, pose in zip(ds, cs, self.camera_poses):
                xyz, rgb = self.dense.depth_to_pointcloud_device(d, c, pose=pose, scale=1.0, subsample=s)
                all_xyz.append(xyz)
                all_rgb.append(rgb)
                counts.append(xyz.shape[0])
        for i, n in enumerate(counts):
            print(f"Camera {i}: {n} points" if i < 2 else f"  Camera {i}: {n} points")
        print("\n--- Step 5: Merge and clean point cloud ---")
        pts = (all_xyz[0] if len(all_xyz) == 1 else torch.cat(all_xyz)).contiguous()
        cols = (all_rgb[0] if len(all_rgb) == 1 else torch.cat(all_rgb)).contiguous()
        if pts.shape[0] == 0:
            return np.array([]), np.array([]), self.camera_poses
        if self.config.voxel_size > 0:
            pts, cols = self.dense.merge_pointclouds_device(pts, cols, self.config.voxel_size)
        final_points, final_colors = to_host(pts), to_host(cols)
        print(f"\nFinal reconstruction: {len(final_points)} points, {len(self.camera_poses)} cameras")
        return final_points, final_colors, self.camera_poses

    def save_reconstruction(self, points: np.ndarray, colors: np.ndarray, output_path: str):
        """d2r:673-703 — Open3D-binary layout by default, reference ASCII fallback on request."""
        if len(points) == 0:
            print("No points to save")
            return
        filepath = Path(output_path)
        filepath.parent.mkdir(parents=True, exist_ok=True)
        write_ply(filepath, points, colors, layout=self.ply_layout)
        print(f"Saved to {filepath}")


def main(argv=None):
    """d2r:770-818 — same flags; --poses and --ply-layout are additions."""
    parser = argparse.ArgumentParser(description="Depth to 3D Reconstruction")
    parser.add_argument("--rgb-folder", type=str, required=True, help="Folder with RGB images")
    parser.add_argument("--depth-folder", type=str, required=True, help="Folder with depth images")
    parser.add_argument("--output", type=str, default="./output/reconstruction.ply", help="Output PLY file path")
    parser.add_argument("--fx", type=float, default=1719.0)
    parser.add_argument("--fy", type=float, default=1719.0)
    parser.add_argument("--cx", type=float, default=540.0)
    parser.add_argument("--cy", type=float, default=960.0)
    parser.add_argument("--voxel-size", type=float, default=0.005)
    parser.add_argument("--subsample", type=int, default=2)
    parser.add_argument("--no-vis", action="store_true")
    parser.add_argument("--poses", type=str, default=None, help="(extension) world->camera poses .npy/.json")
    parser.add_argument("--ply-layout", choices=["open3d", "ascii"], default="open3d",
                        help="(extension) open3d = binary LE doubles; ascii = the reference's fallback writer")
    args = parser.parse_args(argv)
    config = ReconstructionConfig(fx=args.fx, fy=args.fy, cx=args.cx, cy=args.cy, voxel_size=args.voxel_size,
                                  subsample_factor=args.subsample)
    pipeline = DepthToReconstructionPipeline(
        config, poses=load_poses(args.poses) if args.poses else None,
        ply_layout=_lib.PLY_O3D_BINARY if args.ply_layout == "open3d" else _lib.PLY_REF_ASCII)
    if pipeline.load_data(args.rgb_folder, args.depth_folder) < 2:
        print("Failed to load sufficient data")
        return
    points, colors, poses = pipeline.reconstruct()
    if points is not None and len(points) > 0:
        pipeline.save_reconstruction(points, colors, args.output)
    else:
        print("Reconstruction failed")


if __name__ == "__main__":
    main()
