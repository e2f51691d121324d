"""Independent cross-checks of the unpinned part of the oracle (K2, K3, K7, K8, K5).

Open3D itself is unavailable, so oracle/t3d_oracle.c is additionally checked
against *independent* implementations of the same published algorithms built from SciPy /
NumPy primitives (cKDTree, numpy.unique, numpy.linalg.eigh, lstsq).  Agreement here does not
pin Open3D parity — it removes "the oracle has a private bug" as an explanation."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

from textureless_3d_reconstruction_b200 import synthetic as S


def _cloud(n=6000, seed=3):
    rng = np.random.default_rng(seed)
    th = rng.uniform(0, 2 * np.pi, n)
    z = rng.uniform(0, 3, n)
    r = 1.5 + 0.05 * np.cos(3 * th + 0.9 * z)
    return np.column_stack([r * np.cos(th), r * np.sin(th), z]) + rng.normal(0, 0.002, (n, 3))


def test_r2_vs_numpy_unique(oracle):
    rng = np.random.default_rng(0)
    pts = _cloud(20000).astype(np.float32).astype(np.float64)
    cols = rng.integers(0, 256, (len(pts), 3), dtype=np.uint8)
    v = 0.05
    o = oracle.voxel_downsample(pts, cols, v)
    minb = pts.min(0) - 0.5 * v
    idx = np.floor((pts - minb) / v).astype(np.int64)
    uniq, inv, cnt = np.unique(idx, axis=0, return_inverse=True, return_counts=True)
    inv = inv.reshape(-1)
    sums = np.zeros((len(uniq), 3))
    np.add.at(sums, inv, pts)
    csum = np.zeros((len(uniq), 3))
    np.add.at(csum, inv, cols / 255.0)
    order = np.lexsort(o["idx"].T[::-1])
    assert np.array_equal(o["idx"][order], uniq)                     # voxel index set: bit-exact
    assert np.array_equal(o["count"][order], cnt)
    assert np.allclose(o["points"][order], sums / cnt[:, None], rtol=1e-12)
    assert np.abs(o["colors_mean"][order] - csum / cnt[:, None]).max() < 1e-12


def test_r3_vs_ckdtree(oracle):
    pts = _cloud(5000)
    pts[:20] += 0.4                                                   # a few outliers
    keep, mean, (mu, sigma, thr) = oracle.statistical_outlier(pts, nb=20, std_ratio=2.0)
    d, _ = cKDTree(pts).query(pts, k=20)                              # includes self at distance 0
    ref_mean = d.mean(1)
    assert np.allclose(mean, ref_mean, rtol=1e-12)
    ref_mu, ref_sigma = ref_mean.mean(), ref_mean.std(ddof=1)
    assert np.isclose(mu, ref_mu, rtol=1e-12) and np.isclose(sigma, ref_sigma, rtol=1e-10)
    ref_keep = (ref_mean > 0) & (ref_mean < ref_mu + 2.0 * ref_sigma)
    boundary = np.abs(ref_mean - thr) < 1e-9 * thr
    assert np.array_equal(keep[~boundary], ref_keep[~boundary]) and 0 < (~keep).sum() < 400


def test_r7_vs_eigh(oracle):
    pts = _cloud(3000).astype(np.float32)
    nrm = oracle.estimate_normals(pts, knn=30).astype(np.float64)
    p64 = pts.astype(np.float64)
    _, nn = cKDTree(p64).query(p64, k=30)
    worst = 1.0
    for i in range(0, len(pts), 7):
        q = p64[nn[i]]
        cov = np.cov(q.T, bias=True)
        w, vec = np.linalg.eigh(cov)
        if w[1] - w[0] < 1e-3 * w[2]:
            continue                                                   # ill-separated: direction not unique
        worst = min(worst, abs(float(vec[:, 0] @ nrm[i])))
    assert worst > 1 - 1e-6


def test_r8_vs_numpy_gauss_newton(oracle):
    """One point-to-plane linearisation + the full registration against a NumPy restatement."""
    rng = np.random.default_rng(5)
    xy = rng.uniform(-1, 1, (8000, 2))                                 # bumpy height field: all 6 DoF observable
    zz = 0.3 * np.sin(2 * xy[:, 0]) * np.cos(1.5 * xy[:, 1]) + 0.1 * xy[:, 0]
    tgt = np.column_stack([xy, zz]).astype(np.float32)
    c = tgt.astype(np.float64)
    gx = 0.6 * np.cos(2 * c[:, 0]) * np.cos(1.5 * c[:, 1]) + 0.1
    gy = -0.45 * np.sin(2 * c[:, 0]) * np.sin(1.5 * c[:, 1])
    tn = np.column_stack([-gx, -gy, np.ones(len(c))])
    tn = (tn / np.linalg.norm(tn, axis=1, keepdims=True)).astype(np.float32)
    ang = np.deg2rad(0.4)
    R = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    t = np.array([0.004, -0.003, 0.006])
    src = ((c[::3] - t) @ R).astype(np.float32)                        # src = R^T (tgt - t)  =>  T* = [R | t]
    o = oracle.icp_point_to_plane(src, tgt, tn, 0.05, T0=np.eye(4), max_iter=30)
    # independent Gauss-Newton with cKDTree correspondences
    T = np.eye(4)
    tree = cKDTree(c)
    prev = None
    for it in range(31):
        s = src.astype(np.float64) @ T[:3, :3].T + T[:3, 3]
        d, j = tree.query(s, k=1, distance_upper_bound=0.05)
        m = np.isfinite(d)
        fit, rmse = m.mean(), np.sqrt((d[m] ** 2).mean())
        if it == 0:
            a = o["acc_first"]
            n = tn[j[m]].astype(np.float64)
            J = np.column_stack([np.cross(s[m], n), n])
            r = ((s[m] - c[j[m]]) * n).sum(1)
            JtJ = J.T @ J
            assert np.allclose([a[k] for k in range(6)], JtJ[0], rtol=1e-9)
            assert np.allclose(a[21:27], J.T @ r, rtol=1e-7, atol=1e-12) and a[28] == m.sum()
        if prev is not None and abs(prev[0] - fit) < 1e-6 and abs(prev[1] - rmse) < 1e-6:
            break
        prev = (fit, rmse)
        if it == 30:
            break
        n = tn[j[m]].astype(np.float64)
        J = np.column_stack([np.cross(s[m], n), n])
        r = ((s[m] - c[j[m]]) * n).sum(1)
        x = np.linalg.solve(J.T @ J, -J.T @ r)
        ca, sa, cb, sb, cg, sg = np.cos(x[0]), np.sin(x[0]), np.cos(x[1]), np.sin(x[1]), np.cos(x[2]), np.sin(x[2])
        U = np.eye(4)
        U[:3, :3] = [[cg * cb, cg * sb * sa - sg * ca, cg * sb * ca + sg * sa],
                     [sg * cb, sg * sb * sa + cg * ca, sg * sb * ca - cg * sa], [-sb, cb * sa, cb * ca]]
        U[:3, 3] = x[3:]
        T = U @ T
    assert np.abs(o["T"] - T).max() < 1e-9
    assert abs(o["fitness"] - fit) < 1e-12 and abs(o["inlier_rmse"] - rmse) < 1e-9
    # and it recovers the ground-truth motion (source = a subset of the target samples)
    assert np.abs(o["T"][:3, 3] - t).max() < 1e-4 and np.abs(o["T"][:3, :3] - R).max() < 1e-4


def test_r5_vs_numpy_projective_tsdf(oracle):
    """R5 with a NumPy restatement in float64 (weights/occupancy exact, tsdf to f32 rounding)."""
    H, W = 48, 64
    it = S.scaled_intrinsics(H, W)
    K = (it["fx"], it["fy"], it["cx"], it["cy"])
    d, c, T = S.synth_frame(0, 2, H, W, *K, noise_sigma=0.0)
    vol = oracle.TSDFVolume(0.02, 0.08)
    keys = vol.integrate(d, c, K, T, 1.0, 5.0)
    okeys, ot, ow, oc = vol.export()
    T32 = np.asarray(T, np.float32).astype(np.float64)
    K32 = np.asarray(K, np.float32).astype(np.float64)
    vs, tr = np.float64(np.float32(0.02)), np.float64(np.float32(0.08))
    bad = checked = 0
    for b in range(0, len(okeys), max(1, len(okeys) // 40)):
        g = np.stack(np.meshgrid(np.arange(8), np.arange(8), np.arange(8), indexing="ij"), -1).reshape(-1, 3)
        vi = g[:, 0] + 8 * g[:, 1] + 64 * g[:, 2]
        p = (okeys[b].astype(np.float64) * 8 + g) * vs
        pc = p @ T32[:, :3].T + T32[:, 3]
        u = K32[0] * pc[:, 0] / pc[:, 2] + K32[2]
        v = K32[1] * pc[:, 1] / pc[:, 2] + K32[3]
        inside = (u >= 0) & (v >= 0) & (u <= W - 1) & (v <= H - 1) & (pc[:, 2] > 0)
        ui = np.floor(np.where(inside, u, 0) + 0.5).astype(int)
        vv = np.floor(np.where(inside, v, 0) + 0.5).astype(int)
        dd = d[vv, ui].astype(np.float64)
        sdf = dd - pc[:, 2]
        upd = inside & (dd > 0) & (dd <= 5.0) & (sdf >= -tr)
        # voxels within f32 rounding of a decision boundary are excluded from the exact check
        edge = (np.abs(sdf + tr) < 1e-5) | (np.abs(u - np.round(u)) > 0.4999) | (np.abs(v - np.round(v)) > 0.4999) \
            | (np.minimum(u, v) < 1e-4) | (u > W - 1 - 1e-4) | (v > H - 1 - 1e-4)
        w_ref = upd.astype(np.float32)
        sel = ~edge
        bad += int((ow[b][vi][sel] != w_ref[sel]).sum())
        ts = np.minimum(sdf, tr) / tr
        m = sel & upd
        checked += int(m.sum())
        if m.any():
            assert np.abs(ot[b][vi][m] - ts[m]).max() < 1e-4
    assert bad == 0 and checked > 2000


def test_r4_vs_numpy_block_keys(oracle):
    """R4 (block allocation): a vectorised NumPy restatement over the whole stride-4 grid.
    (i) the same f32 operation sequence => the key SET must be identical;
    (ii) evaluated in float64 => every key of a sample that is not within 1e-4 block of a block face
         must be in the oracle's set, and every oracle key must come from some sample's f64 key or from a
         face-boundary sample (no key the geometry does not imply)."""
    H, W = 120, 68
    it = S.scaled_intrinsics(H, W)
    K = (it["fx"], it["fy"], it["cx"], it["cy"])
    f32 = np.float32
    for frame, voxel, trunc, dmax in ((0, 0.01, 0.04, 5.0), (7, 0.02, 0.08, 3.5), (3, 0.005, 0.02, 5.0)):
        d, c, T = S.synth_frame(0, frame, H, W, *K, noise_sigma=0.002)
        vol = oracle.TSDFVolume(voxel, trunc)
        okeys = vol.touch(d, K, T, 1.0, dmax)
        oset = set(map(tuple, okeys.tolist()))
        assert len(oset) == len(okeys) > 20                              # unique, first-touch order
        T32 = np.asarray(T, f32)[:3, :4]
        R, t = T32[:, :3], T32[:, 3]
        Rwc = R.T.copy()
        o = (-(R.astype(np.float64).T @ t.astype(np.float64))).astype(f32)
        fx, fy, cx, cy = (f32(x) for x in K)
        bs = f32(voxel) * f32(8.0)
        tr, dm = f32(trunc), f32(dmax)
        py, px = np.meshgrid(np.arange(H // 4) * 4, np.arange(W // 4) * 4, indexing="ij")
        dd = d[py, px].astype(f32) / f32(1.0)
        valid = (dd > 0) & (dd < dm)
        xc = (px.astype(f32) - cx) / fx
        yc = (py.astype(f32) - cy) / fy
        g = [((Rwc[i, 0] * xc + Rwc[i, 1] * yc) + Rwc[i, 2] * f32(1.0)) + o[i] for i in range(3)]
        dirv = [g[i] - o[i] for i in range(3)]
        t_min = np.maximum(dd - tr, f32(0.0))
        t_max = np.minimum(dd + tr, dm)
        step = (t_max - t_min) / f32(3.0)
        assert all(a.dtype == f32 for a in (xc, yc, g[0], dirv[0], step))
        tt = t_min.copy()
        got32, must64, may64 = set(), set(), set()
        for s in range(4):
            k32 = [np.floor((o[i] + tt * dirv[i]) / bs).astype(np.int64) for i in range(3)]
            got32 |= set(zip(*(k[valid].tolist() for k in k32)))
            # float64 evaluation of the same geometry
            q = [(np.float64(o[i]) + tt.astype(np.float64) * dirv[i].astype(np.float64)) / np.float64(bs) for i in range(3)]
            near = np.zeros_like(valid)
            for qi in q:
                near |= np.abs(qi - np.round(qi)) < 1e-4
            k64 = [np.floor(qi).astype(np.int64) for qi in q]
            must64 |= set(zip(*(k[valid & ~near].tolist() for k in k64)))
            may64 |= set(zip(*(k[valid].tolist() for k in k64)))
            for sh in ((1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)):   # face-boundary samples
                may64 |= set(zip(*((k64[i][valid & near] + sh[i]).tolist() for i in range(3))))
            tt = tt + step
        assert got32 == oset                                              # (i) bit-exact key set
        assert must64 <= oset <= may64                                    # (ii) geometry, independent of f32 order
        assert len(must64) > 0.95 * len(oset)


def test_r6_vs_numpy_dense_grid(oracle):
    """R6 (surface points): rebuild a dense (tsdf, weight) grid from the exported blocks and restate the
    extraction with shifted-array NumPy operations; positions, normals and colours must be the same
    multiset, bit for bit (same f32 formulas, independent traversal / neighbour lookup)."""
    H, W = 120, 68
    it = S.scaled_intrinsics(H, W)
    K = (it["fx"], it["fy"], it["cx"], it["cy"])
    f32 = np.float32
    voxel = 0.02
    vol = oracle.TSDFVolume(voxel, 0.08)
    for i in range(5):
        d, c, T = S.synth_frame(0, i, H, W, *K, noise_sigma=0.002)
        vol.integrate(d, c, K, T, 1.0, 5.0)
    thr = 2.0
    keys, tsdf, w, rgb = vol.export()
    lo = keys.min(0) - 1                                   # one empty block of margin on every side
    dims = (keys.max(0) - lo + 2) * 8
    Tg = np.zeros(dims, f32)                               # missing block: tsdf reads as 0 (gradients) ...
    Wg = np.full(dims, -1.0, f32)                          # ... and weight as "absent"
    Cg = np.zeros(tuple(dims) + (3,), f32)
    for b, k in enumerate(keys):
        x0, y0, z0 = (k - lo) * 8
        # block layout: voxel index = x + 8 y + 64 z
        Tg[x0:x0 + 8, y0:y0 + 8, z0:z0 + 8] = tsdf[b].reshape(8, 8, 8).transpose(2, 1, 0)
        Wg[x0:x0 + 8, y0:y0 + 8, z0:z0 + 8] = w[b].reshape(8, 8, 8).transpose(2, 1, 0)
        Cg[x0:x0 + 8, y0:y0 + 8, z0:z0 + 8] = rgb[b].reshape(8, 8, 8, 3).transpose(2, 1, 0, 3)
    core = tuple(slice(8, n - 8) for n in dims)            # voxels that can own a point

    def sh(a, dx, dy, dz):                                 # a[x+dx, y+dy, z+dz] on the core region
        return a[8 + dx:dims[0] - 8 + dx, 8 + dy:dims[1] - 8 + dy, 8 + dz:dims[2] - 8 + dz]

    gx, gy, gz = np.meshgrid(*(np.arange(8, n - 8) + l * 8 for n, l in zip(dims, lo)), indexing="ij")
    rows = []
    for ax in range(3):
        e = [int(ax == 0), int(ax == 1), int(ax == 2)]
        t_o, w_o, t_i, w_i = Tg[core], Wg[core], sh(Tg, *e), sh(Wg, *e)
        m = (w_o >= thr) & (w_i >= thr) & (t_o * t_i < 0)
        with np.errstate(divide="ignore", invalid="ignore"):
            ratio = (f32(0) - t_o) / (t_i - t_o)
        r = ratio[m]
        P = [f32(voxel) * ((g[m].astype(f32) + r) if e[i] else g[m].astype(f32)) for i, g in enumerate((gx, gy, gz))]
        om = f32(1) - r
        N = []
        for c_ in range(3):
            ce = [int(c_ == 0), int(c_ == 1), int(c_ == 2)]
            go = sh(Tg, *ce)[m] - sh(Tg, *[-x for x in ce])[m]
            gi = sh(Tg, *[a + b for a, b in zip(e, ce)])[m] - sh(Tg, *[a - b for a, b in zip(e, ce)])[m]
            N.append(om * go + r * gi)
        ln = np.sqrt((N[0] * N[0] + N[1] * N[1]) + N[2] * N[2])
        N = [np.where(ln > 0, n / np.where(ln > 0, ln, 1), n).astype(f32) for n in N]
        co, ci = Cg[core][m], sh(Cg, *e)[m]
        col = np.clip(om[:, None] * co + r[:, None] * ci, 0, 255)
        col = np.floor(col + f32(0.5)).astype(np.uint32)                # roundf for non-negative values
        rows.append(np.column_stack([np.column_stack(P).view(np.uint32), np.column_stack(N).view(np.uint32), col]))
    mine = np.concatenate(rows)
    op, on, oc = vol.extract_points(thr)
    theirs = np.column_stack([op.view(np.uint32), on.view(np.uint32), oc.astype(np.uint32)])
    assert len(mine) == len(theirs) > 500
    mine = mine[np.lexsort(mine.T[::-1])]
    theirs = theirs[np.lexsort(theirs.T[::-1])]
    assert np.array_equal(mine[:, :3], theirs[:, :3])                   # positions: bit-exact
    assert np.array_equal(mine[:, 3:6], theirs[:, 3:6])                 # normals: bit-exact
    assert np.array_equal(mine[:, 6:], theirs[:, 6:])                   # colours


@pytest.mark.parametrize("pixel_round", [0, 1])
def test_r5_literal_vs_kernel_ordered_formulation(oracle, pixel_round):
    """o_tsdf_integrate (the arithmetic order the CUDA kernel uses) against o_tsdf_integrate_literal (SURVEY 8c
    R5 term by term: true divisions, roundf, no FMA, no reciprocals): identical weights except on the voxels the
    literal oracle marks as formulation-sensitive (pixel-rounding / image-border / threshold boundaries), tsdf
    within north_star's 1e-4 elsewhere.  The full-size version (1080x1920, 2160x3840) runs on the GPU box
    (tests/test_tsdf_fullsize_gpu.py) with the CUDA kernel as the third party."""
    H, W = 240, 136
    it = S.scaled_intrinsics(H, W)
    K = (it["fx"], it["fy"], it["cx"], it["cy"])
    a, b = oracle.TSDFVolume(0.01, 0.04, pixel_round), oracle.TSDFVolume(0.01, 0.04, pixel_round)
    for i in range(6):
        d, c, T = S.synth_frame(0, i, H, W, *K, noise_sigma=0.002)
        keys = a.integrate(d, c, K, T, 1.0, 5.0)
        b.integrate(d, c, K, T, 1.0, 5.0, keys=keys, literal=True)
    ka, ta, wa, ca = a.export()
    kb, tb, wb, cb = b.export()
    flags, n_pairs = b.export_flags()
    assert np.array_equal(ka, kb)
    sens = flags != 0
    upd = a.counters()["voxel_updates"]
    assert upd > 500_000 and 0 < n_pairs < 2e-3 * upd
    assert not ((wa != wb) & ~sens).any()
    assert np.abs(ta - tb)[~sens].max() <= 1e-4
    assert np.abs(ca - cb)[~sens].max() <= 1e-2
    assert abs(a.counters()["voxel_updates"] - b.counters()["voxel_updates"]) <= n_pairs
