"""K10 triangle-mesh extraction (north_star "surface/point extraction to .ply").

No reference code and no Open3D here => parity unpinned; the oracle (R6m in oracle/t3d_oracle.c)
is pinned by topology known-answers on analytic fields (closed oriented 2-manifold, Euler
characteristic, area) and the GPU path is compared with it as an order-independent mesh."""
import sys

import numpy as np
import pytest

from conftest import ROOT
from mesh_util import analytic_blocks, canonical_mesh, edge_manifold_stats

VOXEL, TRUNC = 0.01, 0.04
C0 = np.array([0.163, 0.157, 0.171])
R0 = 0.093


def sphere(p):
    return np.linalg.norm(p - C0, axis=-1) - R0


def gyroid(p):  # many ambiguous faces; open at the region border
    q = p * (2 * np.pi / 0.11)
    return 0.02 * (np.sin(q[..., 0]) * np.cos(q[..., 1]) + np.sin(q[..., 1]) * np.cos(q[..., 2]) +
                   np.sin(q[..., 2]) * np.cos(q[..., 0]))


def checker(p):  # sign alternates voxel to voxel: every cube is the 4-diagonal case
    g = np.rint(p / VOXEL).astype(np.int64)
    return np.where((g.sum(-1) & 1) == 0, 0.013, -0.017)


def oracle_volume(oracle, sdf, lo, hi, weight=5.0):
    keys, t, w, rgb = analytic_blocks(sdf, lo, hi, VOXEL, TRUNC, weight)
    ov = oracle.TSDFVolume(VOXEL, TRUNC)
    ov.import_blocks(keys, t, w, rgb)
    return ov, (keys, t, w, rgb)


def test_tables_are_reproducible(tmp_path):
    """The committed tables are exactly what the generator derives."""
    before = [(ROOT / p).read_text() for p in ("oracle/mc_tables.h",
                                               "textureless_3d_reconstruction_b200/csrc/mc_tables.cuh")]
    sys.path.insert(0, str(ROOT / "oracle"))
    import gen_mc_tables as g
    tris = g.build()
    assert max(len(t) for t in tris) == 5
    for case in range(256):  # complementary cases cut the same edges
        assert {e for t in tris[case] for e in t} == {e for t in tris[255 - case] for e in t}
    text = before[0]
    for case in (1, 3, 255 - 1):
        row = [e for t in tris[case] for e in t]
        row += [-1] * (15 - len(row))
        assert "{" + ", ".join(map(str, row)) + "}" in text


def test_oracle_sphere_is_closed_oriented_manifold(oracle):
    ov, _ = oracle_volume(oracle, sphere, -1, 5)
    xyz, nrm, col, tri = ov.extract_mesh(3.0)
    dup, missing, nt = edge_manifold_stats(tri)
    assert dup == 0 and missing == 0 and nt == len(tri) > 3000
    V, T = len(xyz), len(tri)
    assert V - (3 * T) // 2 + T == 2                                   # Euler characteristic of a sphere
    a, b, c = xyz[tri[:, 0]], xyz[tri[:, 1]], xyz[tri[:, 2]]
    n = np.cross(b - a, c - a)
    assert (np.einsum("ij,ij->i", n, (a + b + c) / 3 - C0) > 0).all()  # wound along the TSDF gradient
    assert abs(0.5 * np.linalg.norm(n, axis=1).sum() / (4 * np.pi * R0 * R0) - 1) < 0.01
    assert np.abs(np.linalg.norm(xyz - C0, axis=1) - R0).max() < 0.3 * VOXEL
    rad = (xyz - C0) / np.linalg.norm(xyz - C0, axis=1, keepdims=True)
    assert np.einsum("ij,ij->i", nrm, rad).min() > 0.999
    # vertex set == R6 surface points here (every sign-changing edge lies in a valid cube)
    p6, _, _ = ov.extract_points(3.0)
    assert np.array_equal(np.unique(xyz, axis=0), np.unique(p6, axis=0))


@pytest.mark.parametrize("field", [gyroid, checker])
def test_oracle_ambiguous_faces_stay_watertight(oracle, field):
    lo, hi = 0, 3
    ov, _ = oracle_volume(oracle, field, lo, hi)
    xyz, _, _, tri = ov.extract_mesh(3.0)
    assert len(tri) > 1000
    t = np.asarray(tri, np.int64)
    e = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]])
    code = e[:, 0] * len(xyz) + e[:, 1]
    u, cnt = np.unique(code, return_counts=True)
    assert (cnt == 1).all()                                            # no directed edge twice
    missing = ~np.isin(e[:, 1] * len(xyz) + e[:, 0], u)
    # an edge without its opposite may only lie on the border of the observed region
    border_lo, border_hi = lo * 8 * VOXEL, (hi * 8 - 1) * VOXEL
    pm = xyz[e[missing]]
    on_border = ((pm <= border_lo + 1e-6) | (pm >= border_hi - 1e-6)).any(-1).all(-1)
    assert on_border.all()


def test_oracle_weight_threshold_and_missing_blocks(oracle):
    keys, t, w, rgb = analytic_blocks(sphere, -1, 5, VOXEL, TRUNC)
    w[:] = 2.0
    ov = oracle.TSDFVolume(VOXEL, TRUNC)
    ov.import_blocks(keys, t, w, rgb)
    assert len(ov.extract_mesh(3.0)[3]) == 0                           # nothing observed often enough
    assert len(ov.extract_mesh(2.0)[3]) > 0
    # drop one block that the surface crosses: the mesh opens up there, nothing else changes
    hit = np.nonzero((t.min(1) < 0) & (t.max(1) > 0))[0]
    keep = np.ones(len(keys), bool)
    keep[hit[len(hit) // 2]] = False
    ov2 = oracle.TSDFVolume(VOXEL, TRUNC)
    ov2.import_blocks(keys[keep], t[keep], w[keep], rgb[keep])
    full = ov.extract_mesh(2.0)
    part = ov2.extract_mesh(2.0)
    assert 0 < len(part[3]) < len(full[3])
    fv = {tuple(r) for r in full[0].view(np.uint32).tolist()}
    assert all(tuple(r) in fv for r in part[0].view(np.uint32).tolist())


def test_ply_mesh_writer_layout(tmp_path):
    """Host-only: Open3D write_triangle_mesh layout (double xyz/normals, uchar rgb, list uchar uint)."""
    from textureless_3d_reconstruction_b200.runtime import write_ply_mesh
    rng = np.random.default_rng(3)
    v = rng.normal(size=(7, 3)).astype(np.float32)
    n = rng.normal(size=(7, 3)).astype(np.float32)
    c = rng.integers(0, 256, (7, 3), dtype=np.uint8)
    t = rng.integers(0, 7, (5, 3)).astype(np.int32)
    path = tmp_path / "sub" / "m.ply"
    write_ply_mesh(path, v, t, colors=c, normals=n)
    raw = path.read_bytes()
    head, body = raw.split(b"end_header\n", 1)
    assert head.decode().splitlines() == [
        "ply", "format binary_little_endian 1.0", "comment Created by Open3D", "element vertex 7",
        "property double x", "property double y", "property double z",
        "property double nx", "property double ny", "property double nz",
        "property uchar red", "property uchar green", "property uchar blue",
        "element face 5", "property list uchar uint vertex_indices"]
    vdt = np.dtype([("p", "<f8", 3), ("n", "<f8", 3), ("c", "u1", 3)])
    fdt = np.dtype([("k", "u1"), ("i", "<u4", 3)])
    assert len(body) == 7 * vdt.itemsize + 5 * fdt.itemsize
    vv = np.frombuffer(body[:7 * vdt.itemsize], vdt)
    ff = np.frombuffer(body[7 * vdt.itemsize:], fdt)
    assert np.array_equal(vv["p"], v.astype(np.float64)) and np.array_equal(vv["n"], n.astype(np.float64))
    assert np.array_equal(vv["c"], c) and (ff["k"] == 3).all() and np.array_equal(ff["i"], t.astype(np.uint32))


# ------------------------------------------------------------------ GPU parity (through the C ABI)
def gpu_volume(ctx, blocks, capacity=4096):
    import torch
    from textureless_3d_reconstruction_b200.runtime import TSDFVolume
    keys, t, w, rgb = blocks
    vol = TSDFVolume(VOXEL, TRUNC, block_capacity=capacity, ctx=ctx)
    vol.merge_blocks(torch.from_numpy(keys).cuda(), torch.from_numpy(t).cuda(), torch.from_numpy(w).cuda(),
                     torch.from_numpy(rgb).cuda())
    return vol


def assert_same_mesh(g, o):
    gx, gn, gc, gt = [x.cpu().numpy() for x in g]
    ox, on, oc, ot = o
    assert len(gx) == len(ox) and len(gt) == len(ot)                   # counts bit-exact
    if len(gx) == 0:
        return
    cg = canonical_mesh(gx, gt, gn, gc)
    co = canonical_mesh(ox, ot, on, oc)
    assert np.array_equal(cg[0].view(np.uint32), co[0].view(np.uint32))    # vertices bit-exact (<=1e-5 rel required)
    assert np.array_equal(cg[1], co[1])                                    # connectivity identical
    assert np.abs(cg[2] - co[2]).max() <= 1e-6 and np.array_equal(cg[3], co[3])


@pytest.mark.gpu
@pytest.mark.parametrize("field,lo,hi", [(sphere, -1, 5), (gyroid, 0, 3), (checker, 0, 2)])
def test_gpu_mesh_matches_oracle_on_analytic_fields(ctx, oracle, field, lo, hi):
    ov, blocks = oracle_volume(oracle, field, lo, hi)
    vol = gpu_volume(ctx, blocks)
    for thr in (3.0, 5.0, 6.0):
        assert_same_mesh(vol.extract_mesh(thr), ov.extract_mesh(thr))
    # a capacity guess that is too small triggers the exactly-sized second pass
    assert_same_mesh(vol.extract_mesh(3.0, capacity=(10, 7)), ov.extract_mesh(3.0))
    # without optional attributes
    x, n, c, t = vol.extract_mesh(3.0, with_normals=False, with_colors=False)
    assert n is None and c is None and len(t) == len(ov.extract_mesh(3.0)[3])


@pytest.mark.gpu
def test_gpu_mesh_of_fused_frames_matches_oracle(ctx, oracle, tmp_path):
    import torch
    from textureless_3d_reconstruction_b200 import synthetic as S
    from textureless_3d_reconstruction_b200.runtime import TSDFVolume, write_ply_mesh
    H, W = 240, 136
    it = S.scaled_intrinsics(H, W)
    K = (it["fx"], it["fy"], it["cx"], it["cy"])
    vol = TSDFVolume(VOXEL, TRUNC, block_capacity=60000, ctx=ctx)
    ov = oracle.TSDFVolume(VOXEL, TRUNC)
    for i in range(6):
        d, c, T = S.synth_frame(0, i, H, W, *K, noise_sigma=0.002)
        vol.integrate(torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda(), K, T, 1.0, 5.0)
        ov.integrate(d, c, K, T, 1.0, 5.0)
    g = vol.extract_mesh(3.0)
    o = ov.extract_mesh(3.0)
    assert len(o[3]) > 20000
    assert_same_mesh(g, o)
    # mesh vertices are a subset of the R6 surface points of the same volume
    p6 = vol.extract_points(3.0)[0].cpu().numpy()
    s6 = {tuple(r) for r in p6.view(np.uint32).tolist()}
    gx = g[0].cpu().numpy()
    extra = np.array([tuple(r) not in s6 for r in gx.view(np.uint32).tolist()])
    # ... except where a voxel's tsdf is exactly 0: marching cubes tests the sign BIT (0 counts as
    # positive, vertex on the voxel corner), R6 tests t_o * t_n < 0 and emits nothing there
    assert extra.sum() <= 1e-4 * len(gx)
    q = gx[extra] / np.float32(VOXEL)
    assert np.abs(q - np.rint(q)).max(initial=0.0) < 1e-3
    dup, _, _ = edge_manifold_stats(g[3].cpu().numpy())
    assert dup == 0
    write_ply_mesh(tmp_path / "mesh.ply", g[0], g[3], colors=g[2], normals=g[1])
    assert (tmp_path / "mesh.ply").stat().st_size > len(g[0]) * 51 + len(g[3]) * 13


@pytest.mark.gpu
def test_gpu_mesh_empty_and_truncated(ctx, oracle):
    import torch
    from textureless_3d_reconstruction_b200.runtime import TSDFVolume, _ptr, _stream, check
    vol = TSDFVolume(VOXEL, TRUNC, block_capacity=1024, ctx=ctx)
    x, n, c, t = vol.extract_mesh(3.0)
    assert len(x) == 0 and len(t) == 0
    ov, blocks = oracle_volume(oracle, sphere, -1, 5)
    vol = gpu_volume(ctx, blocks)
    full = ov.extract_mesh(3.0)
    # capacities smaller than the mesh: counts are still the true ones, nothing is written past the caps
    vcap, tcap = 100, 50
    xyz = torch.full((vcap + 8, 3), -7.0, device="cuda")
    tri = torch.full((tcap + 8, 3), -7, dtype=torch.int32, device="cuda")
    cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
    check(vol.lib.t3d_tsdf_extract_mesh(vol.handle, 3.0, _ptr(xyz), None, None, vcap, _ptr(tri), tcap, _ptr(cnt),
                                        _stream()))
    assert cnt.tolist() == [len(full[0]), len(full[3])]
    assert (xyz[vcap:] == -7.0).all() and (tri[tcap:] == -7).all()


def test_oracle_vertex_set_vs_numpy_restatement(oracle):
    """Independent NumPy restatement of which edges carry a vertex (sign-bit change on an edge of at
    least one cube whose 8 corners all have W >= thr), on a dense grid with a ragged observed region."""
    lo, hi = 0, 3
    keys, t, w, rgb = analytic_blocks(gyroid, lo, hi, VOXEL, TRUNC)
    rng = np.random.default_rng(8)
    w[rng.uniform(size=w.shape) < 0.08] = 1.0                     # unobserved-enough voxels punch holes
    ov = oracle.TSDFVolume(VOXEL, TRUNC)
    ov.import_blocks(keys, t, w, rgb)
    xyz, _, _, tri = ov.extract_mesh(3.0)
    n = (hi - lo) * 8
    T = np.zeros((n, n, n), np.float32)
    W = np.zeros((n, n, n), np.float32)
    v = np.arange(512)
    loc = np.stack([v & 7, (v >> 3) & 7, v >> 6], -1)
    for k, tt, ww in zip(keys, t, w):
        g = (k - lo) * 8 + loc
        T[g[:, 0], g[:, 1], g[:, 2]] = tt
        W[g[:, 0], g[:, 1], g[:, 2]] = ww
    ok = W >= 3.0
    cube = np.ones((n - 1, n - 1, n - 1), bool)                   # cube based at (x,y,z) valid
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                cube &= ok[dx:n - 1 + dx, dy:n - 1 + dy, dz:n - 1 + dz]
    neg = T < 0
    expected = 0
    for ax in range(3):
        a = [slice(0, n)] * 3
        b = [slice(0, n)] * 3
        a[ax], b[ax] = slice(0, n - 1), slice(1, n)
        change = neg[tuple(a)] != neg[tuple(b)]                   # edge (voxel, +ax), shape n-1 along ax
        # the edge belongs to the cubes based at voxel - {0,1} along the two other axes
        o1, o2 = [c for c in range(3) if c != ax]
        used = np.zeros_like(change)
        pad = np.zeros((n + 1, n + 1, n + 1), bool)
        pad[1:n, 1:n, 1:n] = cube                                 # pad[i+1] = cube[i]; index -1 and n-1.. are False
        for s1 in (0, 1):
            for s2 in (0, 1):
                idx = [None, None, None]
                idx[ax] = slice(1, n)
                idx[o1] = slice(1 - s1, n + 1 - s1)
                idx[o2] = slice(1 - s2, n + 1 - s2)
                used |= pad[tuple(idx)]
        expected += int((change & used).sum())
    assert len(xyz) == expected
    cases = np.zeros((n - 1, n - 1, n - 1), np.int32)
    corner = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
    for i, (dx, dy, dz) in enumerate(corner):
        cases |= neg[dx:n - 1 + dx, dy:n - 1 + dy, dz:n - 1 + dz].astype(np.int32) << i
    crossed = cube & (cases != 0) & (cases != 255)
    assert crossed.sum() > 1000 and len(tri) >= crossed.sum()     # every crossed valid cube emits >= 1 triangle
