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
