"""ctypes wrapper around oracle/libt3d_oracle.so (the C restatement, t3d_oracle.c).

TEST INFRASTRUCTURE ONLY — see the header of t3d_oracle.c.  Nothing in the
product package imports this.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_SO = _DIR / "libt3d_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    src = _DIR / "t3d_oracle.c"
    newest = max(src.stat().st_mtime, (_DIR / "mc_tables.h").stat().st_mtime)
    if force or not _SO.exists() or _SO.stat().st_mtime < newest:
        base = ["gcc", "-O2", "-ffp-contract=off", "-mfma", "-fno-fast-math", "-fPIC", "-shared",
                "-fvisibility=hidden", "-o", str(_SO), str(src), "-lm"]
        r = subprocess.run(base[:1] + ["-fopenmp"] + base[1:], capture_output=True, text=True)
        if r.returncode != 0:  # no libgomp on this box: serial oracle
            subprocess.run(base, check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_SO))
        _lib.o_backproject.restype = C.c_int64
        _lib.o_voxel_downsample.restype = C.c_int64
        _lib.o_statistical_outlier.restype = C.c_int64
        _lib.o_tsdf_create.restype = C.c_void_p
        _lib.o_tsdf_touch.restype = C.c_int64
        _lib.o_tsdf_num_blocks.restype = C.c_int64
        _lib.o_tsdf_export_flags.restype = C.c_int64
        _lib.o_tsdf_extract.restype = C.c_int64
        _lib.o_tsdf_extract_view.restype = C.c_int64
        _lib.o_icp_point_to_plane.restype = C.c_int
        _lib.o_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return lib().o_num_threads()


def set_num_threads(n: int):
    lib().o_set_num_threads(C.c_int(n))


def synth_frame(scene, i, H, W, fx, fy, cx, cy, seed=1234, noise_sigma=0.0):
    """(depth f32 [H,W], bgr u8 [H,W,3], T_cw f64 [3,4]) — C/OpenMP twin of the synthetic scene generator
    (same as textureless_3d_reconstruction_b200/synthetic.synth_frame, ~100x faster): the CPU arm of the
    benchmark builds its workload with it, so that arm never loads the product library."""
    if scene == 0:
        yaw = np.deg2rad(2.0) * np.sin(0.03 * i)
        o = np.array([0.1 * np.sin(0.05 * i), 0.0, 0.25 * i])
    else:
        yaw = np.deg2rad(0.5) * i
        o = np.array([0.05 * i, 0.0, 0.0])
    c, s = np.cos(yaw), np.sin(yaw)
    R = np.ascontiguousarray(np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], np.float64))
    T = np.zeros((3, 4))
    T[:, :3] = R.T
    T[:, 3] = -R.T @ o
    depth = np.empty((H, W), np.float32)
    bgr = np.empty((H, W, 3), np.uint8)
    seed32 = (seed ^ (seed >> 32)) & 0xFFFFFFFF
    o = np.ascontiguousarray(o, np.float64)
    lib().o_synth_frame(C.c_int(scene), C.c_int(i), C.c_int(H), C.c_int(W), C.c_double(fx), C.c_double(fy),
                        C.c_double(cx), C.c_double(cy), C.c_uint32(seed32), C.c_float(noise_sigma), _p(R), _p(o),
                        _p(depth), _p(bgr))
    return depth, bgr, T


def backproject(depth, bgr, fx, fy, cx, cy, scale=1.0, f64_mask=False, min_depth=0.1,
                max_depth=50.0, pose=None, subsample=1):
    depth = np.ascontiguousarray(depth, np.float32)
    H, W = depth.shape
    bgr = None if bgr is None else np.ascontiguousarray(bgr, np.uint8)
    Hs, Ws = -(-H // subsample), -(-W // subsample)
    xyz = np.empty((Hs * Ws, 3), np.float32)
    rgb = np.empty((Hs * Ws, 3), np.uint8)
    R = t = None
    if pose is not None:
        R = np.ascontiguousarray(pose[0], np.float64)
        t = np.ascontiguousarray(np.asarray(pose[1], np.float64).ravel())
    n = lib().o_backproject(_p(depth), _p(bgr), C.c_int(H), C.c_int(W), C.c_int(subsample),
                            C.c_double(fx), C.c_double(fy), C.c_double(cx), C.c_double(cy),
                            C.c_double(scale), C.c_int(1 if f64_mask else 0), C.c_double(min_depth),
                            C.c_double(max_depth), _p(R), _p(t), _p(xyz), _p(rgb))
    return xyz[:n], rgb[:n]


def voxel_downsample(xyz, rgb, voxel, min_bound=None):
    """R2.  Returns dict(points f64, colors_u8, colors_mean f64, count, idx, min_bound)."""
    xyz = np.ascontiguousarray(xyz, np.float64)
    n = len(xyz)
    rgb = None if rgb is None else np.ascontiguousarray(rgb, np.uint8)
    mb = None if min_bound is None else np.ascontiguousarray(min_bound, np.float64)
    o_xyz = np.empty((max(n, 1), 3), np.float64)
    o_rgb = np.empty((max(n, 1), 3), np.uint8)
    o_mean = np.empty((max(n, 1), 3), np.float64)
    o_cnt = np.empty(max(n, 1), np.uint32)
    o_idx = np.empty((max(n, 1), 3), np.int32)
    o_mb = np.zeros(3, np.float64)
    m = lib().o_voxel_downsample(_p(xyz), _p(rgb), C.c_int64(n), C.c_double(voxel), _p(mb), _p(o_xyz),
                                 _p(o_rgb), _p(o_mean), _p(o_cnt), _p(o_idx), _p(o_mb))
    if m == -1:
        raise ValueError("voxel_size <= 0.")
    if m == -2:
        raise ValueError("voxel_size is too small.")
    return dict(points=o_xyz[:m], colors_u8=o_rgb[:m] if rgb is not None else None,
                colors_mean=o_mean[:m] if rgb is not None else None, count=o_cnt[:m], idx=o_idx[:m],
                min_bound=o_mb)


def statistical_outlier(xyz, nb=20, std_ratio=2.0):
    """R3.  Returns (keep mask bool, mean_dist, (mu, sigma, thr))."""
    xyz = np.ascontiguousarray(xyz, np.float64)
    n = len(xyz)
    mean = np.empty(max(n, 1), np.float64)
    keep = np.zeros(max(n, 1), np.uint8)
    stats = np.zeros(3, np.float64)
    lib().o_statistical_outlier(_p(xyz), C.c_int64(n), C.c_int(nb), C.c_double(std_ratio), _p(mean),
                                _p(keep), _p(stats))
    return keep[:n].astype(bool), mean[:n], tuple(stats)


def knn_bruteforce(xyz, q, k):
    xyz = np.ascontiguousarray(xyz, np.float64)
    q = np.ascontiguousarray(q, np.float64)
    d2 = np.empty(k, np.float64)
    idx = np.empty(k, np.int64)
    lib().o_knn_bruteforce(_p(xyz), C.c_int64(len(xyz)), _p(q), C.c_int(k), _p(d2), _p(idx))
    return d2, idx


def estimate_normals(xyz, knn=30, orient_to=None):
    xyz = np.ascontiguousarray(xyz, np.float32)
    out = np.empty_like(xyz)
    o = None if orient_to is None else np.ascontiguousarray(orient_to, np.float64)
    lib().o_estimate_normals(_p(xyz), C.c_int64(len(xyz)), C.c_int(knn), _p(o), _p(out))
    return out


def icp_point_to_plane(src, tgt, tgt_nrm, max_corr, T0=None, max_iter=30, rel_fitness=1e-6,
                       rel_rmse=1e-6):
    """R8.  Returns dict(T, fitness, inlier_rmse, iterations, acc_first)."""
    src = np.ascontiguousarray(src, np.float32)
    tgt = np.ascontiguousarray(tgt, np.float32)
    tgt_nrm = np.ascontiguousarray(tgt_nrm, np.float32)
    T0a = None if T0 is None else np.ascontiguousarray(T0, np.float64)
    T = np.zeros((4, 4), np.float64)
    fr = np.zeros(2, np.float64)
    acc = np.zeros(29, np.float64)
    it = lib().o_icp_point_to_plane(_p(src), C.c_int64(len(src)), _p(tgt), _p(tgt_nrm), C.c_int64(len(tgt)),
                                    C.c_double(max_corr), _p(T0a), C.c_int(max_iter), C.c_double(rel_fitness),
                                    C.c_double(rel_rmse), _p(T), _p(fr), _p(acc))
    return dict(T=T, fitness=fr[0], inlier_rmse=fr[1], iterations=it, acc_first=acc)


def nearest_neighbor(q, ref, radius):
    q = np.ascontiguousarray(q, np.float32)
    ref = np.ascontiguousarray(ref, np.float32)
    idx = np.empty(len(q), np.int32)
    d2 = np.empty(len(q), np.float32)
    lib().o_nearest_neighbor(_p(q), C.c_int64(len(q)), _p(ref), C.c_int64(len(ref)), C.c_double(radius),
                             _p(idx), _p(d2))
    return idx, d2


class TSDFVolume:
    """R4-R6 oracle volume."""

    def __init__(self, voxel_size=0.01, sdf_trunc=0.04, pixel_round=0):
        self._h = C.c_void_p(lib().o_tsdf_create(C.c_float(voxel_size), C.c_float(sdf_trunc), C.c_int(pixel_round)))
        self.voxel_size = voxel_size

    def __del__(self):
        if getattr(self, "_h", None):
            lib().o_tsdf_destroy(self._h)
            self._h = None

    @staticmethod
    def _depth(depth):
        if depth.dtype == np.uint16:
            return np.ascontiguousarray(depth), 1
        return np.ascontiguousarray(depth, np.float32), 0

    @staticmethod
    def _conf(conf, shape):
        if conf is None:
            lib().o_tsdf_set_conf(None)
            return None
        conf = np.ascontiguousarray(conf, np.uint8)
        assert conf.shape == shape
        lib().o_tsdf_set_conf(_p(conf))
        return conf

    def touch(self, depth, K, T_cw, depth_scale=1.0, depth_max=5.0, conf_mask=None):
        depth, u16 = self._depth(depth)
        conf_mask = self._conf(conf_mask, depth.shape)
        H, W = depth.shape
        K = np.ascontiguousarray(K, np.float32)
        T = np.ascontiguousarray(np.asarray(T_cw, np.float32)[:3, :4])
        out = np.empty(((H // 4) * (W // 4) * 4 + 1, 3), np.int32)
        n = lib().o_tsdf_touch(self._h, _p(depth), C.c_int(u16), C.c_int(H), C.c_int(W), _p(K), _p(T),
                               C.c_float(depth_scale), C.c_float(depth_max), _p(out))
        lib().o_tsdf_set_conf(None)
        return out[:n].copy()

    def integrate(self, depth, bgr, K, T_cw, depth_scale=1.0, depth_max=5.0, keys=None, literal=False, conf_mask=None):
        """R4 + R5 for one frame.  literal=True: R5 exactly as SURVEY 8c writes it (true divisions, no FMA,
        no reciprocals; o_tsdf_integrate_literal) instead of the kernel-ordered arithmetic."""
        depth, u16 = self._depth(depth)
        H, W = depth.shape
        bgr = None if bgr is None else np.ascontiguousarray(bgr, np.uint8)
        K = np.ascontiguousarray(K, np.float32)
        T = np.ascontiguousarray(np.asarray(T_cw, np.float32)[:3, :4])
        if keys is None:
            keys = self.touch(depth, K, T, depth_scale, depth_max, conf_mask=conf_mask)
        keys = np.ascontiguousarray(keys, np.int32)
        conf_mask = self._conf(conf_mask, depth.shape)
        fn = lib().o_tsdf_integrate_literal if literal else lib().o_tsdf_integrate
        fn(self._h, _p(depth), C.c_int(u16), _p(bgr), C.c_int(H), C.c_int(W), _p(K), _p(T),
           C.c_float(depth_scale), C.c_float(depth_max), _p(keys), C.c_int64(len(keys)))
        lib().o_tsdf_set_conf(None)
        return keys

    def export_flags(self):
        """(flags u8 [nb, 512] in export() block order, n_sensitive voxel-frame pairs) of a volume fused
        with literal=True: bit 0 = the literal and the kernel-ordered R5 picked different pixels for some
        frame, bit 1 = same pixel but a different outcome of the zc / sdf tests."""
        flags = np.zeros((self.num_blocks, 512), np.uint8)
        n = lib().o_tsdf_export_flags(self._h, _p(flags))
        return flags, int(n)

    @property
    def num_blocks(self):
        return lib().o_tsdf_num_blocks(self._h)

    def counters(self):
        c = np.zeros(3, np.int64)
        lib().o_tsdf_counters(self._h, _p(c))
        return dict(voxel_updates=int(c[0]), block_frames=int(c[1]), frames=int(c[2]))

    def export(self):
        nb = self.num_blocks
        keys = np.empty((nb, 3), np.int32)
        tsdf = np.empty((nb, 512), np.float32)
        w = np.empty((nb, 512), np.float32)
        rgb = np.empty((nb, 512, 3), np.float32)
        lib().o_tsdf_export(self._h, _p(keys), _p(tsdf), _p(w), _p(rgb))
        return keys, tsdf, w, rgb

    def import_blocks(self, keys, tsdf, weight, rgb=None):
        keys = np.ascontiguousarray(keys, np.int32)
        tsdf = np.ascontiguousarray(tsdf, np.float32)
        weight = np.ascontiguousarray(weight, np.float32)
        rgb = None if rgb is None else np.ascontiguousarray(rgb, np.float32)
        lib().o_tsdf_import(self._h, _p(keys), _p(tsdf), _p(weight), _p(rgb), C.c_int64(len(keys)))

    def extract_points(self, weight_threshold=3.0, view=None):
        """view = (K(fx,fy,cx,cy), T_cw 3x4|4x4, H, W, depth_max): only blocks visible from that
        camera (the frame-to-model tracker's target, see o_block_in_view)."""
        cap = max(self.num_blocks * 96, 1024)
        q = None
        if view is not None:
            K, T, H, W, dmax = view
            T = np.asarray(T, np.float32)[:3, :4]
            q = np.ascontiguousarray(np.concatenate([np.asarray(K, np.float32).reshape(4),
                                                     np.array([W - 1, H - 1, dmax], np.float32),
                                                     T[:, :3].reshape(9), T[:, 3].reshape(3)]), np.float32)
        while True:
            xyz = np.empty((cap, 3), np.float32)
            nrm = np.empty((cap, 3), np.float32)
            rgb = np.empty((cap, 3), np.uint8)
            if q is None:
                n = lib().o_tsdf_extract(self._h, C.c_float(weight_threshold), _p(xyz), _p(nrm), _p(rgb), C.c_int64(cap))
            else:
                nsel = C.c_int64(0)
                n = lib().o_tsdf_extract_view(self._h, C.c_float(weight_threshold), _p(q), _p(xyz), _p(nrm), _p(rgb),
                                              C.c_int64(cap), C.byref(nsel))
                self.last_view_blocks = nsel.value
            if n <= cap:
                return xyz[:n].copy(), nrm[:n].copy(), rgb[:n].copy()
            cap = n

    def extract_mesh(self, weight_threshold=3.0):
        """R6m: (vertices f32 Vx3, normals f32 Vx3, colours u8 Vx3, triangles i32 Tx3)."""
        vcap, tcap = max(self.num_blocks * 96, 1024), max(self.num_blocks * 192, 1024)
        while True:
            xyz = np.empty((vcap, 3), np.float32)
            nrm = np.empty((vcap, 3), np.float32)
            rgb = np.empty((vcap, 3), np.uint8)
            tri = np.empty((tcap, 3), np.int32)
            n = np.zeros(2, np.int64)
            lib().o_tsdf_extract_mesh(self._h, C.c_float(weight_threshold), _p(xyz), _p(nrm), _p(rgb), C.c_int64(vcap),
                                      _p(tri), C.c_int64(tcap), _p(n))
            if n[0] <= vcap and n[1] <= tcap:
                return xyz[:n[0]].copy(), nrm[:n[0]].copy(), rgb[:n[0]].copy(), tri[:n[1]].copy()
            vcap, tcap = int(n[0]), int(n[1])
