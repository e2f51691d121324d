"""Helpers shared by the mesh tests (CPU oracle KATs and GPU parity)."""
import numpy as np


def analytic_blocks(sdf, lo, hi, voxel, trunc, weight=5.0, rgb_fn=None):
    """Blocks with keys in [lo, hi)^3 filled with clamp(sdf(p)/trunc, -1, 1) at the voxel corners."""
    ks = np.arange(lo, hi)
    keys = np.stack(np.meshgrid(ks, ks, ks, indexing="ij"), -1).reshape(-1, 3).astype(np.int32)
    v = np.arange(512)
    loc = np.stack([v & 7, (v >> 3) & 7, v >> 6], -1)                      # x fastest
    g = keys[:, None, :] * 8 + loc[None, :, :]
    p = g.astype(np.float64) * voxel
    t = np.clip(sdf(p) / trunc, -1.0, 1.0).astype(np.float32)
    w = np.full(t.shape, weight, np.float32)
    rgb = (rgb_fn(p) if rgb_fn else np.stack([g[..., 0] % 256, g[..., 1] % 256, g[..., 2] % 256], -1)).astype(np.float32)
    return keys, t, w, rgb


def canonical_mesh(xyz, tri, *extra):
    """Order-independent form: vertices sorted by coordinate bits, triangles re-indexed, rotated so
    the smallest index comes first (winding kept) and sorted."""
    xyz = np.ascontiguousarray(xyz, np.float32)
    order = np.lexsort(xyz.T[::-1])
    inv = np.empty(len(order), np.int64)
    inv[order] = np.arange(len(order))
    t = inv[np.asarray(tri, np.int64)]
    k = t.argmin(1)
    t = np.stack([t[np.arange(len(t)), (k + j) % 3] for j in range(3)], 1)
    t = t[np.lexsort(t.T[::-1])]
    return (xyz[order], t) + tuple(np.asarray(e)[order] for e in extra)


def edge_manifold_stats(tri):
    """(#directed edges used more than once, #directed edges whose opposite is missing)."""
    t = np.asarray(tri, np.int64)
    t = t[(t[:, 0] != t[:, 1]) & (t[:, 1] != t[:, 2]) & (t[:, 0] != t[:, 2])]
    e = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]])
    n = int(t.max()) + 1 if len(t) else 1
    code = e[:, 0] * n + e[:, 1]
    rev = e[:, 1] * n + e[:, 0]
    u, c = np.unique(code, return_counts=True)
    dup = int((c > 1).sum())
    missing = int((~np.isin(rev, u)).sum())
    return dup, missing, len(t)
