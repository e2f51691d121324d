"""NumPy twin of csrc/synth.cu: deterministic synthetic RGB-D scenes.

Used by the CPU-side tests and to feed identical arrays to the oracle and the
GPU path.  Same formulas and the same integer hash as the device generator;
transcendental functions differ in the last ulp, so device and NumPy depth agree
to ~1e-6 m, not bit for bit (tests/test_synth.py states the tolerance).
"""
from __future__ import annotations

import numpy as np

CFG1_INTRINSICS = dict(fx=1719.0, fy=1719.0, cx=540.0, cy=960.0, W=1080, H=1920)


def scaled_intrinsics(H: int, W: int):
    """cfg-1 intrinsics scaled to a smaller frame with the same field of view."""
    s = W / 1080.0
    return dict(fx=1719.0 * s, fy=1719.0 * s, cx=540.0 * s, cy=960.0 * s, W=W, H=H)


def lowbias32(x):
    x = np.asarray(x, np.uint32).copy()
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def scene_pose(scene: int, i: int):
    """(R_wc camera->world, camera position o, T_cw world->camera 3x4)."""
    if scene == 0:
        yaw = np.deg2rad(2.0) * np.sin(0.03 * i)
        o = np.array([0.1 * np.sin(0.05 * i), 0.0, 0.25 * i])
    else:
        yaw = np.deg2rad(0.5) * i
        o = np.array([0.05 * i, 0.0, 0.0])
    c, s = np.cos(yaw), np.sin(yaw)
    R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    T = np.zeros((3, 4))
    T[:, :3] = R.T
    T[:, 3] = -R.T @ o
    return R, o, T


def _tunnel_radius(x, y, z):
    th = np.arctan2(y, x)
    return 1.5 + 0.08 * np.sin(2 * np.pi * z / 1.7) + 0.05 * np.cos(3.0 * th + 0.9 * z)


def synth_frame(scene: int, i: int, H: int, W: int, fx, fy, cx, cy, seed: int = 1234, noise_sigma: float = 0.0):
    """Returns (depth f32 [H,W], bgr u8 [H,W,3], T_cw f64 [3,4])."""
    R, o, T = scene_pose(scene, i)
    v, u = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    xc = (u - cx) / fx
    yc = (v - cy) / fy
    dx = R[0, 0] * xc + R[0, 1] * yc + R[0, 2]
    dy = R[1, 0] * xc + R[1, 1] * yc + R[1, 2]
    dz = R[2, 0] * xc + R[2, 1] * yc + R[2, 2]
    seed32 = np.uint32((seed ^ (seed >> 32)) & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        fh = lowbias32(np.uint32((i * 0x9E3779B9) & 0xFFFFFFFF) ^ seed32)
        h0 = lowbias32(np.arange(H * W, dtype=np.uint32).reshape(H, W) ^ fh)
    if scene == 0:
        a = dx * dx + dy * dy
        b = o[0] * dx + o[1] * dy
        c0 = o[0] ** 2 + o[1] ** 2
        with np.errstate(divide="ignore", invalid="ignore"):
            disc_in = b * b - a * (c0 - 1.37 ** 2)
            disc_out = b * b - a * (c0 - 1.63 ** 2)
            t_lo = np.where(disc_in > 0, (-b + np.sqrt(np.maximum(disc_in, 0))) / a, 0.0)
            t_hi = (-b + np.sqrt(np.maximum(disc_out, 0))) / a
        t_lo = np.maximum(t_lo, 0.0)
        ok = (a > 1e-12) & (t_hi > t_lo) & (t_lo < 12.0)
        t_lo = np.where(ok, t_lo, 0.0)
        t_hi = np.where(ok, t_hi, 1.0)

        def f(t):
            x, y, z = o[0] + t * dx, o[1] + t * dy, o[2] + t * dz
            return np.sqrt(x * x + y * y) - _tunnel_radius(x, y, z)

        lo, hi = t_lo.copy(), t_hi.copy()
        st = (t_hi - t_lo) / 8.0
        done = np.zeros_like(ok)
        for k in range(1, 9):
            t = t_lo + st * k
            hit = (f(t) >= 0) & ~done
            hi = np.where(hit, t, hi)
            lo = np.where(~done & ~hit, t, lo)
            done |= hit
        for _ in range(30):
            t = 0.5 * (lo + hi)
            pos = f(t) >= 0
            hi = np.where(pos, t, hi)
            lo = np.where(pos, lo, t)
        t_hit = np.where(ok, 0.5 * (lo + hi), 0.0)
        good = (t_hit > 0) & (t_hit < 12.0)
        dn = t_hit.copy()
        if noise_sigma > 0:
            with np.errstate(over="ignore"):
                h1 = lowbias32(h0 ^ np.uint32(0x68BC21EB))
                h2 = lowbias32(h0 ^ np.uint32(0x02E5BE93))
            u1 = (h1.astype(np.float64) + 1.0) / 4294967297.0
            u2 = h2.astype(np.float64) / 4294967296.0
            dn = dn + np.float64(np.float32(noise_sigma)) * np.sqrt(-2.0 * np.log(u1)) * np.cos(2 * np.pi * u2)
        depth = np.where(good, dn, 0.0).astype(np.float32)
        n0 = (h0 % np.uint32(5)).astype(np.int32) - 2
        n1 = ((h0 >> np.uint32(8)) % np.uint32(5)).astype(np.int32) - 2
        n2 = ((h0 >> np.uint32(16)) % np.uint32(5)).astype(np.int32) - 2
        Rr, G, B = 128 + n0, 128 + n1, 128 + n2
    else:
        t = (3.0 - o[2]) / dz
        for _ in range(8):
            x, y = o[0] + t * dx, o[1] + t * dy
            s = 3.0 + 0.25 * np.sin(2 * np.pi * x / 0.9) * np.cos(2 * np.pi * y / 1.3)
            t = (s - o[2]) / dz
        depth = t.astype(np.float32)
        sel = h0 % np.uint32(1000)
        depth = np.where(sel < 10, np.float32(0), depth)
        depth = np.where(sel == 10, np.float32(np.nan), depth)
        depth = np.where(sel == 11, np.float32(np.inf), depth)
        border = (u < 64) | (v < 64) | (u >= W - 64) | (v >= H - 64)
        depth = np.where(border, np.float32(60.0), depth).astype(np.float32)
        Rr = (u * 7 + v * 13 + i * 29) & 255
        G = (u ^ v) & 255
        B = (u + v + i) & 255
    bgr = np.stack([B, G, Rr], axis=-1).astype(np.uint8)
    return depth, bgr, T
