"""Device-tensor API over the C ABI: torch owns the memory and the streams,
libt3d.so does the work.  Every method takes/returns CUDA torch tensors and is
the layer the drop-in classes (depth_to_reconstruction.py etc.) and bench.py
call.  No method here computes anything on the CPU or with torch ops.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from pathlib import Path

import numpy as np

from . import _lib
from ._lib import BackprojectFrame, BackprojectParams, FrameView, IcpResult, T3DError, TsdfParams, check

_contexts: dict[int, "Context"] = {}


def _torch():
    import torch
    return torch


def get_context(device: int | None = None) -> "Context":
    torch = _torch()
    if not torch.cuda.is_available():
        raise RuntimeError(
            "textureless_3d_reconstruction_b200 needs a CUDA device (B200, sm_100a); "
            "there is no CPU fallback.")
    if device is None:
        device = torch.cuda.current_device()
    ctx = _contexts.get(device)
    if ctx is None:
        ctx = Context(device)
        _contexts[device] = ctx
    return ctx


def _stream():
    return C.c_void_p(_torch().cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _check_frame(depth, bgr=None, conf_mask=None, what="frame"):
    """The kernels index colour / confidence with the depth map's H and W: a mismatch would read out of
    bounds on the device (the reference raises IndexError there), so it is rejected here — with real
    exceptions, not asserts, which vanish under python -O."""
    torch = _torch()
    if depth.dim() != 2 or not depth.is_cuda or not depth.is_contiguous():
        raise ValueError(f"{what}: depth must be a contiguous (H, W) CUDA tensor, got {tuple(depth.shape)}")
    H, W = depth.shape
    if bgr is not None:
        if tuple(bgr.shape) != (H, W, 3) or bgr.dtype != torch.uint8 or not bgr.is_cuda or not bgr.is_contiguous():
            raise ValueError(f"{what}: colour must be a contiguous uint8 CUDA tensor of shape {(H, W, 3)} matching the "
                             f"depth map, got {tuple(bgr.shape)} {bgr.dtype}")
    if conf_mask is not None:
        if tuple(conf_mask.shape) != (H, W) or conf_mask.dtype != torch.uint8 or not conf_mask.is_cuda \
                or not conf_mask.is_contiguous():
            raise ValueError(f"{what}: conf_mask must be a contiguous uint8 CUDA tensor of shape {(H, W)}, "
                             f"got {tuple(conf_mask.shape)} {conf_mask.dtype}")
    return H, W


def _np_ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class _NpDtypes(dict):
    def __missing__(self, key):
        torch = _torch()
        table = {torch.float32: np.float32, torch.float64: np.float64, torch.uint8: np.uint8, torch.int32: np.int32,
                 torch.int64: np.int64, torch.uint16: np.uint16, torch.int16: np.int16}
        self.update(table)
        return table[key]


_NP_DTYPES = _NpDtypes()


_pinned_stage: dict = {}
_pinned_out = [0]                      # bytes of pinned memory currently backing arrays handed to callers
PINNED_RESULT_LIMIT = 1 << 30          # beyond this, results go to ordinary pageable memory


def _pinned_release(nbytes):
    _pinned_out[0] -= nbytes


def to_host(t, chunk_bytes: int = 32 << 20):
    """CUDA tensor -> NumPy array owned by the caller (what the reference's host-NumPy API returns).
    `tensor.cpu()` lands in freshly faulted pageable memory at ~2 GB/s.  Results of >= 1 MiB are
    instead copied straight into page-locked memory from torch's caching host allocator (one PCIe
    copy at link speed, no page faults; the array keeps its block alive and returns it to the cache
    when it is garbage-collected), as long as less than PINNED_RESULT_LIMIT bytes are outstanding;
    past that limit they go to pageable memory through a pinned staging buffer in two alternating
    halves (the PCIe copy of chunk i+1 overlaps the host copy of chunk i)."""
    import weakref
    torch = _torch()
    if t is None:
        return None
    if not t.is_cuda:
        return t.numpy()
    t = t.contiguous()
    nbytes = t.numel() * t.element_size()
    if nbytes < (1 << 20):
        return t.cpu().numpy()
    if _pinned_out[0] + nbytes <= PINNED_RESULT_LIMIT:
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        _pinned_out[0] += nbytes
        weakref.finalize(h, _pinned_release, nbytes)
        return h.numpy()
    flat = t.reshape(-1).view(torch.uint8)
    out = np.empty(nbytes, np.uint8)
    key = t.device.index
    st = _pinned_stage.get(key)
    if st is None or st[0].numel() < 2 * chunk_bytes:
        st = (torch.empty(2 * chunk_bytes, dtype=torch.uint8, pin_memory=True),
              [torch.cuda.Event(), torch.cuda.Event()])
        _pinned_stage[key] = st
    buf, ev = st
    host = buf.numpy()
    chunks = [(a, min(a + chunk_bytes, nbytes)) for a in range(0, nbytes, chunk_bytes)]
    stream = torch.cuda.current_stream()

    def issue(i):
        a, b = chunks[i]
        h = (i & 1) * chunk_bytes
        buf[h:h + (b - a)].copy_(flat[a:b], non_blocking=True)
        ev[i & 1].record(stream)
    issue(0)
    for i, (a, b) in enumerate(chunks):
        if i + 1 < len(chunks):
            issue(i + 1)
        ev[i & 1].synchronize()
        h = (i & 1) * chunk_bytes
        out[a:b] = host[h:h + (b - a)]
    return out.view(_NP_DTYPES[t.dtype]).reshape(tuple(t.shape))


@dataclass
class IcpOutput:
    transformation: np.ndarray
    fitness: float
    inlier_rmse: float
    iterations: int
    converged: bool
    correspondences: int


class Context:
    """One libt3d context on one GPU."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device; libt3d has no CPU fallback")
        torch.cuda.init()
        self.device = torch.device("cuda", device)
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.device)  # make sure the primary context exists
            h = self.lib.t3d_create(device)
        if not h:
            raise _lib.T3DError(_lib.T3D_E_CUDA, _lib.last_error())
        self.handle = C.c_void_p(h)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.t3d_destroy(self.handle)
            self.handle = None

    def launch_count(self) -> int:
        return int(self.lib.t3d_launch_count(self.handle))

    # ------------------------------------------------------------------ K1
    def backproject(self, depth, bgr, *, fx, fy, cx, cy, subsample=1, scale=1.0, scale_is_f64=False,
                    min_depth=0.1, max_depth=50.0, pose=None, rgb_out_f32=False, conf_mask=None,
                    out_xyz=None, out_rgb=None, out_n=None):
        """depth: (H,W) f32|f64 CUDA tensor, bgr: (H,W,3) u8 CUDA tensor or None.
        Returns (xyz[cap,3] f32, rgb[cap,3] u8|f32 or None, n int64[1]) — all on device;
        only the first n rows are valid."""
        torch = _torch()
        H, W = _check_frame(depth, bgr, conf_mask, "backproject")
        if depth.dtype not in (torch.float32, torch.float64):
            raise ValueError(f"backproject: depth must be float32 or float64, got {depth.dtype}")
        p = BackprojectParams()
        p.H, p.W, p.subsample = H, W, int(subsample)
        p.depth_is_f64 = int(depth.dtype == torch.float64)
        p.scale_is_f64 = int(bool(scale_is_f64))
        p.has_pose = int(pose is not None)
        p.rgb_out_f32 = int(bool(rgb_out_f32))
        p.has_color = int(bgr is not None)
        p.fx, p.fy, p.cx, p.cy = float(fx), float(fy), float(cx), float(cy)
        p.scale, p.min_depth, p.max_depth = float(scale), float(min_depth), float(max_depth)
        if pose is not None:
            R = np.ascontiguousarray(pose[0], np.float64).reshape(9)
            t = np.ascontiguousarray(pose[1], np.float64).reshape(3)
            p.R[:] = R.tolist()
            p.t[:] = t.tolist()
        s = int(subsample)
        cap = (-(-H // s)) * (-(-W // s)) if s >= 1 else 0
        if out_xyz is None:
            out_xyz = torch.empty((max(cap, 1), 3), dtype=torch.float32, device=depth.device)
        if bgr is not None and out_rgb is None:
            out_rgb = torch.empty((max(cap, 1), 3), dtype=torch.float32 if rgb_out_f32 else torch.uint8,
                                  device=depth.device)
        if out_n is None:
            out_n = torch.zeros(1, dtype=torch.int64, device=depth.device)
        check(self.lib.t3d_backproject(self.handle, _ptr(depth), _ptr(bgr), _ptr(conf_mask), C.byref(p),
                                       _ptr(out_xyz), _ptr(out_rgb), out_xyz.shape[0], _ptr(out_n), _stream()))
        return out_xyz, out_rgb, out_n

    def make_backproject_frames(self, depths, bgrs, poses=None, conf_masks=None):
        """Host-side descriptor array for backproject_batch (build once, reuse)."""
        n = len(depths)
        arr = (BackprojectFrame * n)()
        for i in range(n):
            _check_frame(depths[i], bgrs[i] if bgrs is not None else None,
                         conf_masks[i] if conf_masks is not None else None, f"backproject frame {i}")
            if depths[i].shape != depths[0].shape or depths[i].dtype != depths[0].dtype:
                raise ValueError("backproject_batch: every frame of a batch must have the same size and dtype")
            arr[i].depth = depths[i].data_ptr()
            arr[i].bgr = bgrs[i].data_ptr() if bgrs is not None and bgrs[i] is not None else None
            arr[i].conf_mask = conf_masks[i].data_ptr() if conf_masks is not None and conf_masks[i] is not None else None
            if poses is not None:
                arr[i].R[:] = np.ascontiguousarray(poses[i][0], np.float64).reshape(9).tolist()
                arr[i].t[:] = np.ascontiguousarray(poses[i][1], np.float64).reshape(3).tolist()
        return arr

    def backproject_batch(self, frames, n_frames, H, W, *, fx, fy, cx, cy, subsample=1, scale=1.0,
                          scale_is_f64=False, depth_is_f64=False, min_depth=0.1, max_depth=50.0, has_pose=True,
                          has_color=True, rgb_out_f32=False, out_xyz=None, out_rgb=None, out_offsets=None):
        """K1 over `n_frames` frames (descriptors from make_backproject_frames) with one launch
        per 32 frames; the output is the frame-ordered concatenation (the reference's per-frame
        loop + np.vstack).  Returns (xyz[cap,3], rgb[cap,3]|None, offsets int64[n+1]) on device."""
        torch = _torch()
        p = BackprojectParams()
        p.H, p.W, p.subsample = int(H), int(W), int(subsample)
        p.depth_is_f64, p.scale_is_f64 = int(bool(depth_is_f64)), int(bool(scale_is_f64))
        p.has_pose, p.rgb_out_f32, p.has_color = int(bool(has_pose)), int(bool(rgb_out_f32)), int(bool(has_color))
        p.fx, p.fy, p.cx, p.cy = float(fx), float(fy), float(cx), float(cy)
        p.scale, p.min_depth, p.max_depth = float(scale), float(min_depth), float(max_depth)
        s = int(subsample)
        cap = (-(-H // s)) * (-(-W // s)) * n_frames if s >= 1 else 0
        if out_xyz is None:
            out_xyz = torch.empty((max(cap, 1), 3), dtype=torch.float32, device=self.device)
        if has_color and out_rgb is None:
            out_rgb = torch.empty((max(cap, 1), 3), dtype=torch.float32 if rgb_out_f32 else torch.uint8,
                                  device=self.device)
        if out_offsets is None:
            out_offsets = torch.zeros(n_frames + 1, dtype=torch.int64, device=self.device)
        check(self.lib.t3d_backproject_batch(self.handle, frames, int(n_frames), C.byref(p), _ptr(out_xyz),
                                             _ptr(out_rgb), out_xyz.shape[0], _ptr(out_offsets), _stream()))
        return out_xyz, out_rgb, out_offsets

    # ------------------------------------------------------------------ K2
    def voxel_downsample(self, xyz, rgb, voxel_size, *, min_bound=None, sorted_output=True,
                         want_idx=True, capacity=None):
        """xyz: (N,3) f32|f64, rgb: (N,3) u8 or None.  Returns a dict of device tensors
        (points f64, colors u8, rgb_sum u32, count u32, idx i32) trimmed to M rows."""
        torch = _torch()
        n = xyz.shape[0]
        dev = xyz.device
        assert xyz.is_contiguous()
        # The outputs hold one row per VOXEL, which is known only after the first pass.  Allocating them for the worst
        # case (one voxel per point: 59 bytes x N, half a gigabyte for a merged cloud) costs more than the kernels, so the
        # first attempt sizes them from the last call's ratio (1/4 of N the first time); the library reports
        # T3D_E_CAPACITY after its first pass (before anything is written) if that was too small, and the call is redone
        # with the worst case.
        ratio = getattr(self, "_k2_ratio", 0.25)
        attempts = [int(capacity)] if capacity is not None else [min(n, max(int(n * ratio * 1.25) + 1024, 4096)), n]
        for cap in attempts:
            o_xyz = torch.empty((max(cap, 1), 3), dtype=torch.float64, device=dev)
            o_rgb = torch.empty((max(cap, 1), 3), dtype=torch.uint8, device=dev) if rgb is not None else None
            o_sum = torch.empty((max(cap, 1), 3), dtype=torch.int32, device=dev) if rgb is not None else None
            o_cnt = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
            o_idx = torch.empty((max(cap, 1), 3), dtype=torch.int32, device=dev) if want_idx else None
            o_m = torch.zeros(1, dtype=torch.int64, device=dev)
            mb_in = None if min_bound is None else np.ascontiguousarray(min_bound, np.float64)
            mb_out = np.zeros(3, np.float64)
            rc = self.lib.t3d_voxel_downsample(
                self.handle, _ptr(xyz), int(xyz.dtype == torch.float64), _ptr(rgb), n, float(voxel_size),
                _np_ptr(mb_in), int(bool(sorted_output)), _ptr(o_xyz), _ptr(o_rgb), _ptr(o_sum), _ptr(o_cnt),
                _ptr(o_idx), cap, _ptr(o_m), _np_ptr(mb_out), _stream())
            if rc == _lib.T3D_E_CAPACITY and cap < n and capacity is None:
                continue
            check(rc)
            break
        m = int(o_m.item())
        if n > 0:
            self._k2_ratio = max(m / n, 1e-3)
        return dict(points=o_xyz[:m], colors=None if o_rgb is None else o_rgb[:m],
                    rgb_sum=None if o_sum is None else o_sum[:m], count=o_cnt[:m],
                    idx=None if o_idx is None else o_idx[:m], min_bound=mb_out, m=m)

    VOXEL_RECORD_BYTES = 56

    def voxel_partials(self, xyz, rgb, voxel_size, min_bound, max_bound, world):
        """Sharded K2, sender side: (records uint8[M,56] grouped by owner, counts int32[world]) on the
        device.  min_bound / max_bound: the GLOBAL grid origin (min - voxel/2) and data maximum."""
        torch = _torch()
        n = xyz.shape[0]
        dev = xyz.device
        rec = torch.empty((max(n, 1), self.VOXEL_RECORD_BYTES), dtype=torch.uint8, device=dev)
        cnt = torch.zeros(world, dtype=torch.int32, device=dev)
        o_m = torch.zeros(1, dtype=torch.int64, device=dev)
        mb = np.ascontiguousarray(min_bound, np.float64)
        xb = np.ascontiguousarray(max_bound, np.float64)
        check(self.lib.t3d_voxel_partials(self.handle, _ptr(xyz), int(xyz.dtype == torch.float64), _ptr(rgb), n,
                                          float(voxel_size), _np_ptr(mb), _np_ptr(xb), int(world), _ptr(rec),
                                          rec.shape[0], _ptr(cnt), _ptr(o_m), _stream()))
        return rec[: int(o_m.item())], cnt

    def voxel_merge_partials(self, records, has_rgb, voxel_size, min_bound, max_bound, sorted_output=True):
        """Sharded K2, owner side: merge the records received from every rank; same outputs as
        voxel_downsample."""
        torch = _torch()
        n = records.shape[0]
        dev = records.device
        assert records.dtype == torch.uint8 and records.is_contiguous()
        cap = max(n, 1)
        o_xyz = torch.empty((cap, 3), dtype=torch.float64, device=dev)
        o_rgb = torch.empty((cap, 3), dtype=torch.uint8, device=dev) if has_rgb else None
        o_sum = torch.empty((cap, 3), dtype=torch.int32, device=dev) if has_rgb else None
        o_cnt = torch.empty(cap, dtype=torch.int32, device=dev)
        o_idx = torch.empty((cap, 3), dtype=torch.int32, device=dev)
        o_m = torch.zeros(1, dtype=torch.int64, device=dev)
        mb = np.ascontiguousarray(min_bound, np.float64)
        xb = np.ascontiguousarray(max_bound, np.float64)
        check(self.lib.t3d_voxel_merge_partials(self.handle, _ptr(records), n, int(bool(has_rgb)), float(voxel_size),
                                                _np_ptr(mb), _np_ptr(xb), int(bool(sorted_output)), _ptr(o_xyz),
                                                _ptr(o_rgb), _ptr(o_sum), _ptr(o_cnt), _ptr(o_idx), cap, _ptr(o_m),
                                                _stream()))
        m = int(o_m.item())
        if n > 0:
            self._k2_ratio = max(m / n, 1e-3)
        return dict(points=o_xyz[:m], colors=None if o_rgb is None else o_rgb[:m],
                    rgb_sum=None if o_sum is None else o_sum[:m], count=o_cnt[:m], idx=o_idx[:m],
                    min_bound=mb, m=m)

    def bounds(self, xyz):
        torch = _torch()
        mn, mx = np.zeros(3), np.zeros(3)
        check(self.lib.t3d_bounds(self.handle, _ptr(xyz), int(xyz.dtype == torch.float64), xyz.shape[0],
                                  _np_ptr(mn), _np_ptr(mx), _stream()))
        return mn, mx

    # ------------------------------------------------------------------ K3
    def statistical_outlier(self, xyz, nb_neighbors=20, std_ratio=2.0):
        """xyz: (N,3) f64.  Returns (keep u8[N], mean_dist f64[N], (mu, sigma, thr), kept)."""
        torch = _torch()
        n = xyz.shape[0]
        assert xyz.dtype == torch.float64 and xyz.is_contiguous()
        keep = torch.zeros(max(n, 1), dtype=torch.uint8, device=xyz.device)
        mean = torch.empty(max(n, 1), dtype=torch.float64, device=xyz.device)
        kept = torch.zeros(1, dtype=torch.int64, device=xyz.device)
        stats = np.zeros(3, np.float64)
        check(self.lib.t3d_statistical_outlier(self.handle, _ptr(xyz), n, int(nb_neighbors), float(std_ratio),
                                               _ptr(mean), _ptr(keep), _ptr(kept), _np_ptr(stats), _stream()))
        return keep[:n], mean[:n], tuple(stats), int(kept.item())

    def sor_mean_distances_part(self, xyz, nb_neighbors, part, parts, out_mean):
        """Sharded K3, query side: fill out_mean (f64[N]) for this part's share of the cloud."""
        torch = _torch()
        assert xyz.dtype == torch.float64 and xyz.is_contiguous() and out_mean.dtype == torch.float64
        check(self.lib.t3d_sor_mean_distances_part(self.handle, _ptr(xyz), xyz.shape[0], int(nb_neighbors), int(part),
                                                   int(parts), _ptr(out_mean), _stream()))
        return out_mean

    def sor_from_mean_distances(self, mean, std_ratio=2.0):
        """Sharded K3, after the all_reduce: (keep u8[N], (mu, sigma, thr), kept)."""
        torch = _torch()
        n = mean.shape[0]
        keep = torch.zeros(max(n, 1), dtype=torch.uint8, device=mean.device)
        kept = torch.zeros(1, dtype=torch.int64, device=mean.device)
        stats = np.zeros(3, np.float64)
        check(self.lib.t3d_sor_from_mean_distances(self.handle, _ptr(mean), n, float(std_ratio), _ptr(keep), _ptr(kept),
                                                   _np_ptr(stats), _stream()))
        return keep[:n], tuple(stats), int(kept.item())

    def compact_rows(self, rows, keep):
        torch = _torch()
        n = rows.shape[0]
        assert rows.is_contiguous() and keep.dtype == torch.uint8
        row_bytes = rows.element_size() * (rows.numel() // max(n, 1)) if n else rows.element_size()
        out = torch.empty_like(rows)
        out_n = torch.zeros(1, dtype=torch.int64, device=rows.device)
        check(self.lib.t3d_compact_rows(self.handle, _ptr(rows), n, row_bytes, _ptr(keep), _ptr(out), _ptr(out_n),
                                        _stream()))
        return out[: int(out_n.item())]

    # ------------------------------------------------------------------ K7
    def estimate_normals(self, xyz, knn=30, orient_to=None):
        torch = _torch()
        assert xyz.dtype == torch.float32 and xyz.is_contiguous()
        nrm = torch.empty_like(xyz)
        o = None if orient_to is None else np.ascontiguousarray(orient_to, np.float64)
        check(self.lib.t3d_estimate_normals(self.handle, _ptr(xyz), xyz.shape[0], int(knn), _np_ptr(o), _ptr(nrm),
                                            _stream()))
        return nrm

    # ------------------------------------------------------------------ K8
    def icp_point_to_plane(self, src, tgt, tgt_nrm, max_corr_dist, init=None, max_iter=30,
                           relative_fitness=1e-6, relative_rmse=1e-6) -> IcpOutput:
        torch = _torch()
        for t in (src, tgt, tgt_nrm):
            assert t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda
        T0 = None if init is None else np.ascontiguousarray(init, np.float64).reshape(16)
        res = IcpResult()
        check(self.lib.t3d_icp_point_to_plane(self.handle, _ptr(src), src.shape[0], _ptr(tgt), _ptr(tgt_nrm),
                                              tgt.shape[0], float(max_corr_dist), _np_ptr(T0), int(max_iter),
                                              float(relative_fitness), float(relative_rmse), C.byref(res),
                                              _stream()))
        return IcpOutput(np.array(res.T[:], np.float64).reshape(4, 4), res.fitness, res.inlier_rmse,
                         res.iterations, bool(res.converged), int(res.correspondences))

    def icp_point_to_plane_dev(self, src, n_src_dev, tgt, tgt_nrm, n_tgt_dev, max_corr_dist, init=None, max_iter=30,
                               relative_fitness=1e-6, relative_rmse=1e-6, min_points=0):
        """Same registration with the cloud sizes in device memory (int64[1] tensors): src / tgt are
        capacity-sized buffers.  Returns (IcpOutput, skipped) — skipped=True when a cloud had fewer
        than min_points points (no registration, transformation = init)."""
        torch = _torch()
        for t in (src, tgt, tgt_nrm):
            assert t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda
        assert n_src_dev.dtype == torch.int64 and n_tgt_dev.dtype == torch.int64
        T0 = None if init is None else np.ascontiguousarray(init, np.float64).reshape(16)
        res = IcpResult()
        skipped = C.c_int(0)
        check(self.lib.t3d_icp_point_to_plane_dev(
            self.handle, _ptr(src), src.shape[0], _ptr(n_src_dev), _ptr(tgt), _ptr(tgt_nrm), tgt.shape[0],
            _ptr(n_tgt_dev), int(min_points), float(max_corr_dist), _np_ptr(T0), int(max_iter),
            float(relative_fitness), float(relative_rmse), C.byref(res), C.byref(skipped), _stream()))
        out = IcpOutput(np.array(res.T[:], np.float64).reshape(4, 4), res.fitness, res.inlier_rmse,
                        res.iterations, bool(res.converged), int(res.correspondences))
        return out, bool(skipped.value)

    def icp_correspondences(self, src, tgt, tgt_nrm, max_corr_dist, T0, T1=None):
        """The registration's correspondence search on its own: (original target index or -1, squared distance
        f64) per source point at pose T0; with T1, a second search at T1 seeded with the first one's answers."""
        torch = _torch()
        for t in (src, tgt, tgt_nrm):
            assert t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda
        n = src.shape[0]
        idx = torch.empty(max(n, 1), dtype=torch.int32, device=src.device)
        d2 = torch.empty(max(n, 1), dtype=torch.float64, device=src.device)
        T0h = np.ascontiguousarray(T0, np.float64).reshape(16)
        T1h = None if T1 is None else np.ascontiguousarray(T1, np.float64).reshape(16)
        check(self.lib.t3d_icp_correspondences(self.handle, _ptr(src), n, _ptr(tgt), _ptr(tgt_nrm), tgt.shape[0],
                                               float(max_corr_dist), _np_ptr(T0h), _np_ptr(T1h), _ptr(idx), _ptr(d2),
                                               _stream()))
        return idx[:n], d2[:n]

    def icp_linearize(self, src, tgt, tgt_nrm, max_corr_dist, T):
        """One linearisation: returns (acc27, sum_d2, count) as host values."""
        Th = np.ascontiguousarray(T, np.float64).reshape(16)
        a27 = np.zeros(27, np.float64)
        st = np.zeros(2, np.float64)
        check(self.lib.t3d_icp_linearize(self.handle, _ptr(src), src.shape[0], _ptr(tgt), _ptr(tgt_nrm),
                                         tgt.shape[0], float(max_corr_dist), _np_ptr(Th), _np_ptr(a27),
                                         _np_ptr(st), _stream()))
        return a27, float(st[0]), float(st[1])

    def nearest_neighbor(self, query, ref, radius):
        torch = _torch()
        idx = torch.empty(max(query.shape[0], 1), dtype=torch.int32, device=query.device)
        d2 = torch.empty(max(query.shape[0], 1), dtype=torch.float32, device=query.device)
        check(self.lib.t3d_nearest_neighbor(self.handle, _ptr(query), query.shape[0], _ptr(ref), ref.shape[0],
                                            float(radius), _ptr(idx), _ptr(d2), _stream()))
        return idx[: query.shape[0]], d2[: query.shape[0]]

    # ------------------------------------------------------------------ formats
    def depth_u16_to_f32(self, raw, divisor=1000.0):
        """raw: (H,W) uint16|int16-typed CUDA tensor of millimetres -> f32 metres (d2r:85-90)."""
        torch = _torch()
        assert raw.is_cuda and raw.is_contiguous() and raw.element_size() == 2
        out = torch.empty(raw.shape, dtype=torch.float32, device=raw.device)
        check(self.lib.t3d_depth_u16_to_f32(self.handle, _ptr(raw), raw.numel(), float(divisor), _ptr(out), _stream()))
        return out

    def depth_f32_to_u16(self, depth, factor=1000.0):
        """(depth * 1000).astype(np.uint16) on the GPU (dp:919-921); returns a uint16 tensor."""
        torch = _torch()
        assert depth.is_cuda and depth.is_contiguous() and depth.dtype == torch.float32
        out = torch.empty(depth.shape, dtype=torch.uint16, device=depth.device)
        check(self.lib.t3d_depth_f32_to_u16(self.handle, _ptr(depth), depth.numel(), float(factor), _ptr(out), _stream()))
        return out

    def resize_bilinear(self, src, dst_h, dst_w):
        """cv2.resize(src, (dst_w, dst_h), interpolation=cv2.INTER_LINEAR) for (H,W) f32 (d2r:465-467)."""
        torch = _torch()
        assert src.is_cuda and src.is_contiguous() and src.dtype == torch.float32 and src.dim() == 2
        out = torch.empty((int(dst_h), int(dst_w)), dtype=torch.float32, device=src.device)
        check(self.lib.t3d_resize_bilinear_f32(self.handle, _ptr(src), src.shape[0], src.shape[1], _ptr(out),
                                               int(dst_h), int(dst_w), _stream()))
        return out

    def estimate_scale(self, depth, sparse_points, sparse_pts2d, gate=True, min_input_points=0):
        """Median of Z/d over sparse points (d2r:297-326 gate=True; der:652-697 gate=False,
        min_input_points=5).  depth: (H,W) f32 CUDA tensor.  Returns (scale, n_samples)."""
        p3 = np.ascontiguousarray(sparse_points, np.float64).reshape(-1, 3)
        p2 = np.ascontiguousarray(sparse_pts2d, np.float64).reshape(-1, 2)
        n = min(len(p3), len(p2))
        s = C.c_double(1.0)
        k = C.c_int64(0)
        check(self.lib.t3d_estimate_scale(self.handle, _ptr(depth), depth.shape[0], depth.shape[1], _np_ptr(p3),
                                          _np_ptr(p2), n, int(bool(gate)), int(min_input_points), C.byref(s),
                                          C.byref(k), _stream()))
        return float(s.value), int(k.value)

    def pack_pointcloud2(self, xyz, colors):
        """PointCloud2 records [N,4] f32 = x, y, z, rgb-as-float-bits (dp:744-758).
        colors: f32 in [0,1] (PointCloudGenerator.generate output) or u8 RGB."""
        torch = _torch()
        assert xyz.is_contiguous() and colors.is_contiguous() and xyz.dtype == torch.float32
        out = torch.empty((max(xyz.shape[0], 1), 4), dtype=torch.float32, device=xyz.device)
        check(self.lib.t3d_pack_pointcloud2(self.handle, _ptr(xyz), _ptr(colors), int(colors.dtype == torch.float32),
                                            xyz.shape[0], _ptr(out), _stream()))
        return out[: xyz.shape[0]]

    # ------------------------------------------------------------------ synth
    def synth_frame(self, scene, frame_index, H, W, fx, fy, cx, cy, seed=1234, noise_sigma=0.0,
                    depth=None, bgr=None, pose_only=False):
        """Returns (depth f32[H,W], bgr u8[H,W,3], T_cw f64[3,4])."""
        torch = _torch()
        T = np.zeros((3, 4), np.float64)
        if not pose_only:
            if depth is None:
                depth = torch.empty((H, W), dtype=torch.float32, device=self.device)
            if bgr is None:
                bgr = torch.empty((H, W, 3), dtype=torch.uint8, device=self.device)
        check(self.lib.t3d_synth_frame(self.handle, int(scene), int(frame_index), H, W, float(fx), float(fy),
                                       float(cx), float(cy), int(seed), float(noise_sigma),
                                       None if pose_only else _ptr(depth), None if pose_only else _ptr(bgr),
                                       _np_ptr(T), _stream()))
        return depth, bgr, T


class TSDFVolume:
    """Voxel-block TSDF volume (K4/K5/K6) on one GPU."""

    MAX_BATCH = 64

    def __init__(self, voxel_size=0.01, sdf_trunc=0.04, block_capacity=200_000, pixel_round=0,
                 ctx: Context | None = None):
        self.ctx = ctx or get_context()
        self.lib = self.ctx.lib
        p = TsdfParams(float(voxel_size), float(sdf_trunc), 8, int(pixel_round), int(block_capacity), 0)
        h = C.c_void_p()
        check(self.lib.t3d_tsdf_create(self.ctx.handle, C.byref(p), C.byref(h)))
        self.handle = h
        self.voxel_size = float(voxel_size)
        self.sdf_trunc = float(sdf_trunc)
        self.block_capacity = int(block_capacity)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.t3d_tsdf_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def reset(self):
        check(self.lib.t3d_tsdf_reset(self.handle, _stream()))

    @staticmethod
    def make_frame_views(depths, bgrs, Ks, T_cws, conf_masks=None):
        """Build the host-side frame descriptor array once (bench: outside the timed loop).
        conf_masks: optional per-frame (H, W) uint8 CUDA tensors (0 = pixel carries no measurement)."""
        n = len(depths)
        arr = (FrameView * n)()
        for i in range(n):
            cm = conf_masks[i] if conf_masks is not None else None
            _check_frame(depths[i], bgrs[i] if bgrs is not None else None, cm, f"TSDF frame {i}")
            arr[i].conf_mask = cm.data_ptr() if cm is not None else None
            if depths[i].shape != depths[0].shape or depths[i].dtype != depths[0].dtype:
                raise ValueError("TSDF integration: every frame of a batch must have the same size and dtype")
            arr[i].depth = depths[i].data_ptr()
            arr[i].bgr = bgrs[i].data_ptr() if bgrs is not None and bgrs[i] is not None else None
            K = np.asarray(Ks[i] if isinstance(Ks, (list, tuple)) else Ks, np.float32).reshape(4)
            T = np.asarray(T_cws[i], np.float64)[:3, :4].astype(np.float32).reshape(12)
            arr[i].K[:] = K.tolist()
            arr[i].T_cw[:] = T.tolist()
        return arr

    def integrate_views(self, views, start, count, H, W, depth_is_u16=False, depth_scale=1.0, depth_max=5.0):
        """Fuse views[start:start+count] (count <= 64) in one temporally blocked pass."""
        sub = C.cast(C.byref(views, start * C.sizeof(FrameView)), C.POINTER(FrameView))
        check(self.lib.t3d_tsdf_integrate(self.handle, sub, int(count), int(H), int(W), int(depth_is_u16),
                                          float(depth_scale), float(depth_max), _stream()))

    def integrate_sequence(self, views, n_frames, H, W, batch=64, depth_is_u16=False, depth_scale=1.0,
                           depth_max=5.0, start=0):
        """Fuse views[start:start+n_frames] in `batch`-frame passes with K4/K5 overlap."""
        sub = C.cast(C.byref(views, start * C.sizeof(FrameView)), C.POINTER(FrameView))
        check(self.lib.t3d_tsdf_integrate_sequence(self.handle, sub, int(n_frames), int(batch), int(H), int(W),
                                                   int(depth_is_u16), float(depth_scale), float(depth_max),
                                                   _stream()))

    def integrate_sequence_hooked(self, views, n_frames, H, W, batch, nblocks_dev, after_batch0, wait_event,
                                  depth_is_u16=False, depth_scale=1.0, depth_max=5.0, start=0, hook_batch=0):
        """integrate_sequence with the routing hooks of t3d_tsdf_integrate_sequence_hooked:
        nblocks_dev (int32[1] CUDA tensor) receives the block count after K4 of batch `hook_batch`,
        after_batch0(phase, event_a, event_b) is called on the host — phase b (every b >= hook_batch) once K5 of batch b
        is enqueued (event_a completes with K4 of that batch, event_b with its K5) and, only if wait_event is given,
        phase -1 once K4 of the last batch is enqueued (event_a completes with it; the callee must record wait_event
        before returning; K5 of the last batch waits for wait_event, a handle from t3d_event_create)."""
        sub = C.cast(C.byref(views, start * C.sizeof(FrameView)), C.POINTER(FrameView))
        err = []

        def _cb(_user, phase, ev_a, ev_b):
            try:
                after_batch0(int(phase), C.c_void_p(ev_a), C.c_void_p(ev_b))
            except BaseException as e:  # noqa: BLE001 - must not propagate through the C frame
                err.append(e)
        cb = _lib.SEQUENCE_HOOK(_cb)
        check(self.lib.t3d_tsdf_integrate_sequence_hooked(
            self.handle, sub, int(n_frames), int(batch), int(H), int(W), int(depth_is_u16), float(depth_scale),
            float(depth_max), _ptr(nblocks_dev), cb, None, wait_event, int(hook_batch), _stream()))
        if err:
            raise err[0]

    def integrate(self, depth, bgr, K, T_cw, depth_scale=1.0, depth_max=5.0, conf_mask=None):
        """Fuse one frame.  depth: (H,W) f32|u16 CUDA tensor; bgr: (H,W,3) u8 or None;
        K=(fx,fy,cx,cy); T_cw: world->camera 4x4 or 3x4; conf_mask: optional (H,W) u8, 0 = masked pixel."""
        torch = _torch()
        H, W = depth.shape
        views = self.make_frame_views([depth], [bgr], [K], [T_cw], None if conf_mask is None else [conf_mask])
        self.integrate_views(views, 0, 1, H, W, depth.dtype in (torch.uint16, torch.int16), depth_scale, depth_max)

    def integrate_batch(self, depths, bgrs, K, T_cws, depth_scale=1.0, depth_max=5.0):
        torch = _torch()
        H, W = depths[0].shape
        n = len(depths)
        views = self.make_frame_views(depths, bgrs, [K] * n, T_cws)
        self.integrate_sequence(views, n, H, W, self.MAX_BATCH, depths[0].dtype in (torch.uint16, torch.int16),
                                depth_scale, depth_max)

    def touch(self, depth, K, T_cw, depth_scale=1.0, depth_max=5.0, conf_mask=None):
        """K4 only: unique block keys (int32 [n,3], device) touched by one frame."""
        torch = _torch()
        H, W = depth.shape
        views = self.make_frame_views([depth], None, [K], [T_cw], None if conf_mask is None else [conf_mask])
        cap = (H // 4) * (W // 4) * 4 + 1
        keys = torch.empty((cap, 3), dtype=torch.int32, device=depth.device)
        n = torch.zeros(1, dtype=torch.int64, device=depth.device)
        check(self.lib.t3d_tsdf_touch(self.handle, views, H, W, int(depth.dtype in (torch.uint16, torch.int16)),
                                      float(depth_scale), float(depth_max), _ptr(keys), cap, _ptr(n), _stream()))
        return keys[: int(n.item())]

    @property
    def num_blocks(self) -> int:
        n = int(self.lib.t3d_tsdf_num_blocks(self.handle, _stream()))
        if n < 0:
            raise _lib.T3DError(n, _lib.last_error())
        return n

    def counters(self, detailed=False) -> dict:
        c = np.zeros(5, np.int64)
        check(self.lib.t3d_tsdf_counters(self.handle, _np_ptr(c), _stream()))
        out = dict(voxel_updates=int(c[0]), block_frames=int(c[1]), frames=int(c[2]))
        if detailed:
            out.update(voxels_changed_per_visit=int(c[3]), block_visits=int(c[4]))
        return out

    def set_profiling(self, enable=True):
        check(self.lib.t3d_tsdf_set_profiling(self.handle, int(bool(enable))))

    def get_profile(self) -> dict:
        p = np.zeros(3, np.float64)
        check(self.lib.t3d_tsdf_get_profile(self.handle, _np_ptr(p), _stream()))
        return dict(touch_ms=float(p[0]), integrate_ms=float(p[1]), calls=int(p[2]))

    def export_blocks(self):
        torch = _torch()
        nb = self.num_blocks
        dev = self.ctx.device
        keys = torch.empty((max(nb, 1), 3), dtype=torch.int32, device=dev)
        tsdf = torch.empty((max(nb, 1), 512), dtype=torch.float32, device=dev)
        w = torch.empty((max(nb, 1), 512), dtype=torch.float32, device=dev)
        rgb = torch.empty((max(nb, 1), 512, 3), dtype=torch.float32, device=dev)
        out_b = torch.zeros(1, dtype=torch.int64, device=dev)
        check(self.lib.t3d_tsdf_export_blocks(self.handle, _ptr(keys), _ptr(tsdf), _ptr(w), _ptr(rgb), max(nb, 1),
                                              _ptr(out_b), _stream()))
        return keys[:nb], tsdf[:nb], w[:nb], rgb[:nb]

    def count_blocks_range(self, axis, lo, hi) -> int:
        torch = _torch()
        out_b = torch.zeros(1, dtype=torch.int64, device=self.ctx.device)
        check(self.lib.t3d_tsdf_export_blocks_range(self.handle, int(axis), int(lo), int(hi), None, None, None, None,
                                                    0, _ptr(out_b), _stream()))
        return int(out_b.item())

    def export_blocks_range(self, axis, lo, hi):
        """Blocks with key[axis] in [lo, hi): (keys i32[n,3], tsdf[n,512], weight[n,512], rgb[n,512,3])."""
        torch = _torch()
        dev = self.ctx.device
        n = self.count_blocks_range(axis, lo, hi)
        keys = torch.empty((max(n, 1), 3), dtype=torch.int32, device=dev)
        tsdf = torch.empty((max(n, 1), 512), dtype=torch.float32, device=dev)
        w = torch.empty((max(n, 1), 512), dtype=torch.float32, device=dev)
        rgb = torch.empty((max(n, 1), 512, 3), dtype=torch.float32, device=dev)
        out_b = torch.zeros(1, dtype=torch.int64, device=dev)
        if n > 0:
            check(self.lib.t3d_tsdf_export_blocks_range(self.handle, int(axis), int(lo), int(hi), _ptr(keys), _ptr(tsdf),
                                                        _ptr(w), _ptr(rgb), n, _ptr(out_b), _stream()))
        return keys[:n], tsdf[:n], w[:n], rgb[:n]

    def export_blocks_outside(self, axis, lo, hi):
        """Blocks with key[axis] NOT in [lo, hi) — what this rank does not own — in one call.
        The buffers are sized from the previous call (one retry if the set grew)."""
        torch = _torch()
        dev = self.ctx.device
        cap = max(int(getattr(self, "_outside_cap", 0)), 1024)
        out_b = torch.zeros(1, dtype=torch.int64, device=dev)
        while True:
            keys = torch.empty((cap, 3), dtype=torch.int32, device=dev)
            tsdf = torch.empty((cap, 512), dtype=torch.float32, device=dev)
            w = torch.empty((cap, 512), dtype=torch.float32, device=dev)
            rgb = torch.empty((cap, 512, 3), dtype=torch.float32, device=dev)
            rc = self.lib.t3d_tsdf_export_blocks_outside(self.handle, int(axis), int(lo), int(hi), _ptr(keys), _ptr(tsdf),
                                                         _ptr(w), _ptr(rgb), cap, _ptr(out_b), _stream())
            n = int(out_b.item())
            if rc == _lib.T3D_E_CAPACITY and n > cap:
                cap = int(n * 1.25) + 64
                continue
            check(rc)
            self._outside_cap = max(cap, int(n * 1.25) + 64)
            return keys[:n], tsdf[:n], w[:n], rgb[:n]

    RECORD_WORDS = 4 + 5 * 512

    def route_counts(self, axis, slab_blocks, world, rank):
        """int32[world] (device): blocks owned by every other rank."""
        torch = _torch()
        counts = torch.zeros(world, dtype=torch.int32, device=self.ctx.device)
        check(self.lib.t3d_tsdf_route_counts(self.handle, int(axis), int(slab_blocks), int(world), int(rank),
                                             _ptr(counts), _stream()))
        return counts

    def route_export(self, axis, slab_blocks, world, rank, counts, total=None):
        """Non-owned blocks as wire records [n, 2564] f32, grouped by destination rank.
        total: sum(counts) if the caller already has it on the host (saves a sync)."""
        torch = _torch()
        base = (torch.cumsum(counts, 0, dtype=torch.int32) - counts).contiguous()
        n = int(counts.sum().item()) if total is None else int(total)
        rec = torch.empty((max(n, 1), self.RECORD_WORDS), dtype=torch.float32, device=self.ctx.device)
        fill = torch.zeros_like(counts)
        check(self.lib.t3d_tsdf_route_export(self.handle, int(axis), int(slab_blocks), int(world), int(rank),
                                             _ptr(base), _ptr(fill), _ptr(rec), _stream()))
        return rec[:n]

    def merge_records(self, records):
        assert records.is_contiguous()
        check(self.lib.t3d_tsdf_merge_records(self.handle, _ptr(records), records.shape[0], _stream()))

    def merge_blocks(self, keys, tsdf, weight, rgb):
        check(self.lib.t3d_tsdf_merge_blocks(self.handle, _ptr(keys), _ptr(tsdf), _ptr(weight), _ptr(rgb),
                                             keys.shape[0], _stream()))

    # ---- checkpoint / resume: the block list is the wire format
    def save(self, path):
        """Dump the volume (keys + per-voxel tsdf / weight / rgb of every block) to an .npz."""
        keys, tsdf, w, rgb = self.export_blocks()
        np.savez(path, keys=keys.cpu().numpy(), tsdf=tsdf.cpu().numpy(), weight=w.cpu().numpy(),
                 rgb=rgb.cpu().numpy(), voxel_size=np.float32(self.voxel_size), sdf_trunc=np.float32(self.sdf_trunc))

    @classmethod
    def load(cls, path, block_capacity=None, ctx=None):
        """Restore a volume saved by save(); fusion can continue on it (weights are kept)."""
        torch = _torch()
        z = np.load(path)
        n = len(z["keys"])
        vol = cls(float(z["voxel_size"]), float(z["sdf_trunc"]), int(block_capacity or max(2 * n, 1024)), ctx=ctx)
        dev = vol.ctx.device
        if n:
            vol.merge_blocks(torch.from_numpy(z["keys"]).to(dev), torch.from_numpy(z["tsdf"]).to(dev),
                             torch.from_numpy(z["weight"]).to(dev), torch.from_numpy(z["rgb"]).to(dev))
        return vol

    def extract_points(self, weight_threshold=3.0, with_normals=True, with_colors=True, capacity=None):
        torch = _torch()
        dev = self.ctx.device
        cap = int(capacity) if capacity else max(self.num_blocks * 96, 1024)
        while True:
            xyz = torch.empty((cap, 3), dtype=torch.float32, device=dev)
            nrm = torch.empty((cap, 3), dtype=torch.float32, device=dev) if with_normals else None
            rgb = torch.empty((cap, 3), dtype=torch.uint8, device=dev) if with_colors else None
            n = torch.zeros(1, dtype=torch.int64, device=dev)
            check(self.lib.t3d_tsdf_extract_points(self.handle, float(weight_threshold), _ptr(xyz), _ptr(nrm),
                                                   _ptr(rgb), cap, _ptr(n), _stream()))
            cnt = int(n.item())
            if cnt <= cap:
                return xyz[:cnt], (nrm[:cnt] if nrm is not None else None), (rgb[:cnt] if rgb is not None else None)
            cap = cnt


    def extract_points_range(self, axis, lo, hi, weight_threshold=3.0, with_normals=True, with_colors=True):
        """K6 over the blocks with key[axis] in [lo, hi) only (the owned slab of a sharded volume)."""
        torch = _torch()
        dev = self.ctx.device
        cap = max(self.count_blocks_range(axis, lo, hi) * 96, 1024)
        while True:
            xyz = torch.empty((cap, 3), dtype=torch.float32, device=dev)
            nrm = torch.empty((cap, 3), dtype=torch.float32, device=dev) if with_normals else None
            rgb = torch.empty((cap, 3), dtype=torch.uint8, device=dev) if with_colors else None
            n = torch.zeros(1, dtype=torch.int64, device=dev)
            check(self.lib.t3d_tsdf_extract_points_range(self.handle, int(axis), int(lo), int(hi),
                                                         float(weight_threshold), _ptr(xyz), _ptr(nrm), _ptr(rgb), cap,
                                                         _ptr(n), _stream()))
            cnt = int(n.item())
            if cnt <= cap:
                return xyz[:cnt], (nrm[:cnt] if nrm is not None else None), (rgb[:cnt] if rgb is not None else None)
            cap = cnt

    def extract_mesh(self, weight_threshold=3.0, with_normals=True, with_colors=True, capacity=None):
        """K10: marching-cubes triangle mesh of the fused surface (Open3D extract_triangle_mesh
        semantics).  Returns (vertices f32 Vx3, normals f32 Vx3|None, colours u8 Vx3|None,
        triangles i32 Tx3) on the device.  One pass into buffers sized from the block count
        (capacity = (vertices, triangles) overrides the guess); a second, exactly sized pass only if
        the guess was too small."""
        torch = _torch()
        dev = self.ctx.device
        nb = self.num_blocks
        vcap, tcap = capacity if capacity else (max(nb * 96, 1024), max(nb * 192, 1024))
        n = torch.zeros(2, dtype=torch.int64, device=dev)
        while True:
            vcap, tcap = min(int(vcap), 2**31 - 2), min(int(tcap), 2**31 - 2)
            xyz = torch.empty((vcap, 3), dtype=torch.float32, device=dev)
            nrm = torch.empty((vcap, 3), dtype=torch.float32, device=dev) if with_normals else None
            rgb = torch.empty((vcap, 3), dtype=torch.uint8, device=dev) if with_colors else None
            tri = torch.empty((tcap, 3), dtype=torch.int32, device=dev)
            check(self.lib.t3d_tsdf_extract_mesh(self.handle, float(weight_threshold), _ptr(xyz), _ptr(nrm), _ptr(rgb),
                                                 vcap, _ptr(tri), tcap, _ptr(n), _stream()))
            nv, nt = (int(x) for x in n.tolist())
            if nv >= 2**31 - 2 or nt >= 2**31 - 2:
                raise T3DError(-1, f"mesh too large for 32-bit indices: {nv} vertices, {nt} triangles")
            if nv <= vcap and nt <= tcap:
                return (xyz[:nv], nrm[:nv] if nrm is not None else None, rgb[:nv] if rgb is not None else None,
                        tri[:nt])
            vcap, tcap = max(nv, 1), max(nt, 1)

    def extract_points_view_async(self, K, T_cw, H, W, depth_max, weight_threshold, buffers, with_normals=True,
                                  with_colors=False):
        """K6 over the blocks visible from (K, T_cw) with NO host synchronisation: fills the capacity-sized
        tensors of `buffers` (created/grown by ensure_view_buffers) and leaves the point count in
        buffers["n"] (int64[1], device).  The caller checks buffers["n"] <= capacity when it next syncs."""
        arr = (FrameView * 1)()
        arr[0].depth = None
        arr[0].bgr = None
        arr[0].K[:] = np.asarray(K, np.float32).reshape(4).tolist()
        arr[0].T_cw[:] = np.asarray(T_cw, np.float64)[:3, :4].astype(np.float32).reshape(12).tolist()
        check(self.lib.t3d_tsdf_extract_points_view(
            self.handle, arr, int(H), int(W), float(depth_max), float(weight_threshold), _ptr(buffers["xyz"]),
            _ptr(buffers["nrm"]) if with_normals else None, _ptr(buffers["rgb"]) if with_colors else None,
            buffers["xyz"].shape[0], _ptr(buffers["n"]), None, _stream()))
        return buffers

    def ensure_view_buffers(self, buffers, cap):
        torch = _torch()
        dev = self.ctx.device
        if buffers.get("xyz") is None or buffers["xyz"].shape[0] < cap:
            buffers["xyz"] = torch.empty((cap, 3), dtype=torch.float32, device=dev)
            buffers["nrm"] = torch.empty((cap, 3), dtype=torch.float32, device=dev)
            buffers["rgb"] = torch.empty((cap, 3), dtype=torch.uint8, device=dev)
            buffers["n"] = torch.zeros(1, dtype=torch.int64, device=dev)
            buffers["cap"] = cap
        return buffers

    def extract_points_view(self, K, T_cw, H, W, depth_max=5.0, weight_threshold=1.0, with_normals=True,
                            with_colors=False, buffers=None):
        """K6 over the blocks visible from camera (K, T_cw) only — the frame-to-model ICP target.
        buffers: optional dict reused across calls (grown on demand).  Returns (xyz, nrm, rgb, n_blocks)."""
        torch = _torch()
        dev = self.ctx.device
        arr = (FrameView * 1)()
        arr[0].depth = None
        arr[0].bgr = None
        arr[0].K[:] = np.asarray(K, np.float32).reshape(4).tolist()
        arr[0].T_cw[:] = np.asarray(T_cw, np.float64)[:3, :4].astype(np.float32).reshape(12).tolist()
        buffers = buffers if buffers is not None else {}
        cap = int(buffers.get("cap", 1 << 20))
        while True:
            if buffers.get("xyz") is None or buffers["xyz"].shape[0] < cap:
                buffers["xyz"] = torch.empty((cap, 3), dtype=torch.float32, device=dev)
                buffers["nrm"] = torch.empty((cap, 3), dtype=torch.float32, device=dev)
                buffers["rgb"] = torch.empty((cap, 3), dtype=torch.uint8, device=dev)
                buffers["n"] = torch.zeros(1, dtype=torch.int64, device=dev)
                buffers["cap"] = cap
            nsel = C.c_int64(0)
            check(self.lib.t3d_tsdf_extract_points_view(
                self.handle, arr, int(H), int(W), float(depth_max), float(weight_threshold), _ptr(buffers["xyz"]),
                _ptr(buffers["nrm"]) if with_normals else None, _ptr(buffers["rgb"]) if with_colors else None,
                buffers["xyz"].shape[0], _ptr(buffers["n"]), C.byref(nsel), _stream()))
            cnt = int(buffers["n"].item())
            if cnt <= buffers["xyz"].shape[0]:
                return (buffers["xyz"][:cnt], buffers["nrm"][:cnt] if with_normals else None,
                        buffers["rgb"][:cnt] if with_colors else None, int(nsel.value))
            cap = int(cnt * 1.25) + 1024


def write_ply(path, points, colors=None, normals=None, layout=_lib.PLY_O3D_BINARY):
    """Host-side PLY writer through the C ABI (K9).  points: (N,3) f32|f64 NumPy."""
    lib = _lib.load()
    pts = np.ascontiguousarray(points)
    if pts.dtype not in (np.float32, np.float64):
        pts = pts.astype(np.float64)
    n = len(pts)
    cols = None
    if colors is not None:
        cols = np.ascontiguousarray(colors)
        if cols.dtype != np.uint8:
            cols = cols.astype(np.uint8)
    nr = None if normals is None else np.ascontiguousarray(normals, pts.dtype)
    check(lib.t3d_write_ply_h(str(path).encode(), _np_ptr(pts), int(pts.dtype == np.float64), _np_ptr(cols),
                              _np_ptr(nr), n, int(layout)))


def write_ply_mesh(path, vertices, triangles, colors=None, normals=None):
    """Triangle mesh -> .ply in Open3D's write_triangle_mesh layout (binary LE, double xyz [normals],
    uchar rgb, `list uchar uint vertex_indices`).  Accepts torch (any device) or NumPy arrays."""
    def host(a, dt):
        if a is None:
            return None
        if hasattr(a, "detach"):
            a = a.detach().cpu().numpy()
        return np.ascontiguousarray(a, dt)
    v = host(vertices, np.float32)
    t = host(triangles, np.int32)
    c = host(colors, np.uint8)
    nr = host(normals, np.float32)
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    check(_lib.load().t3d_write_ply_mesh_h(str(path).encode(), _np_ptr(v), _np_ptr(nr), _np_ptr(c), len(v),
                                           _np_ptr(t), len(t)))
