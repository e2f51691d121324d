"""K2 voxel downsample, K3 outlier removal, K7 normals, K8 ICP, K9 PLY — CUDA vs oracle.
(Open3D semantics restated in oracle/t3d_oracle.c; parity unpinned by the reference.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from textureless_3d_reconstruction_b200 import synthetic as S


def cloud(n_frames=3, H=240, W=136, sub=1):
    """fused T1 cloud through the oracle's back-projection (f32 xyz, u8 rgb)."""
    from oracle import capi
    it = S.scaled_intrinsics(H, W)
    P, C = [], []
    for i in range(n_frames):
        d, c, T = S.synth_frame(0, i, H, W, it["fx"], it["fy"], it["cx"], it["cy"], noise_sigma=0.002)
        c = (np.arange(H * W * 3, dtype=np.uint32).reshape(H, W, 3) * 2654435761 >> 13).astype(np.uint8)
        p, col = capi.backproject(d, c, it["fx"], it["fy"], it["cx"], it["cy"], max_depth=5.0,
                                  pose=(T[:, :3].copy(), T[:, 3:4].copy()), subsample=sub)
        P.append(p), C.append(col)
    return np.vstack(P), np.vstack(C)


def sort_by_idx(idx, *arrs):
    order = np.lexsort((idx[:, 2], idx[:, 1], idx[:, 0]))
    return (idx[order],) + tuple(a[order] for a in arrs)


@pytest.mark.parametrize("voxel", [0.005, 0.02, 0.1])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_voxel_downsample_vs_oracle(ctx, oracle, voxel, dtype):
    import torch
    pts, cols = cloud()
    o = oracle.voxel_downsample(pts.astype(np.float64), cols, voxel)
    g = ctx.voxel_downsample(torch.from_numpy(pts.astype(dtype)).cuda(), torch.from_numpy(cols).cuda(), voxel)
    assert g["m"] == len(o["points"])                                   # post-downsample count bit-exact
    gi = g["idx"].cpu().numpy()
    assert np.array_equal(gi, o["idx"])                                 # sorted output == oracle's key order
    assert np.array_equal(g["count"].cpu().numpy().astype(np.uint32), o["count"])
    assert np.allclose(g["points"].cpu().numpy(), o["points"], rtol=1e-5, atol=1e-12)
    assert np.array_equal(g["min_bound"], o["min_bound"])
    gc, oc = g["colors"].cpu().numpy().astype(int), o["colors_u8"].astype(int)
    assert np.abs(gc - oc).max() <= 1                                   # +-1 LSB
    single = o["count"] == 1
    assert np.array_equal(gc[single], oc[single])                       # single-point voxels exact
    # unsorted mode: same set
    g2 = ctx.voxel_downsample(torch.from_numpy(pts.astype(dtype)).cuda(), torch.from_numpy(cols).cuda(), voxel,
                              sorted_output=False)
    i2, p2 = sort_by_idx(g2["idx"].cpu().numpy(), g2["points"].cpu().numpy())
    assert np.array_equal(i2, o["idx"]) and np.allclose(p2, o["points"], rtol=1e-5, atol=1e-12)


def test_voxel_downsample_adversarial(ctx, oracle):
    """points exactly on voxel faces, duplicates, large coordinates, single point, errors."""
    import torch
    from textureless_3d_reconstruction_b200._lib import T3DError
    v = 0.01
    rng = np.random.default_rng(3)
    k = rng.integers(-2000, 2000, size=(5000, 3))
    faces = (k * v).astype(np.float32)                         # on voxel faces (after f32 rounding)
    dup = np.repeat(rng.uniform(-1, 1, size=(500, 3)).astype(np.float32), 7, axis=0)
    far = (rng.uniform(-1, 1, size=(300, 3)) * 300.0).astype(np.float32)
    pts = np.vstack([faces, dup, far])
    cols = rng.integers(0, 256, size=(len(pts), 3), dtype=np.uint8)
    o = oracle.voxel_downsample(pts.astype(np.float64), cols, v)
    g = ctx.voxel_downsample(torch.from_numpy(pts).cuda(), torch.from_numpy(cols).cuda(), v)
    assert g["m"] == len(o["points"])
    assert np.array_equal(g["idx"].cpu().numpy(), o["idx"])
    assert np.array_equal(g["count"].cpu().numpy().astype(np.uint32), o["count"])
    one = ctx.voxel_downsample(torch.tensor([[1.0, 2.0, 3.0]], device="cuda"), None, 0.5)
    assert one["m"] == 1 and np.allclose(one["points"].cpu().numpy(), [[1, 2, 3]])
    with pytest.raises(T3DError):
        ctx.voxel_downsample(torch.from_numpy(pts).cuda(), None, 0.0)
    with pytest.raises(T3DError):
        ctx.voxel_downsample(torch.from_numpy(pts).cuda(), None, 1e-9)  # > 2^21 voxels per axis
    empty = ctx.voxel_downsample(torch.empty((0, 3), device="cuda"), None, 0.1)
    assert empty["m"] == 0


def test_statistical_outlier_vs_oracle(ctx, oracle):
    import torch
    pts, cols = cloud(2)
    ds = oracle.voxel_downsample(pts.astype(np.float64), cols, 0.02)["points"]
    rng = np.random.default_rng(0)
    ds = np.vstack([ds, rng.uniform(-3, 3, size=(200, 3)) + [0, 0, 3], ds[:5]])   # outliers + exact duplicates
    keep_o, mean_o, st_o = oracle.statistical_outlier(ds, 20, 2.0)
    keep, mean, st, kept = ctx.statistical_outlier(torch.from_numpy(ds).cuda(), 20, 2.0)
    mean = mean.cpu().numpy()
    assert np.allclose(mean, mean_o, rtol=1e-12, atol=0)
    assert np.allclose(st, st_o, rtol=1e-10)
    boundary = np.abs(mean_o - st_o[2]) < 1e-9 * st_o[2]
    k = keep.cpu().numpy().astype(bool)
    assert np.array_equal(k[~boundary], keep_o[~boundary])
    assert kept == int(k.sum()) and 0 < kept < len(ds)
    out = ctx.compact_rows(torch.from_numpy(ds).cuda(), keep)
    assert np.array_equal(out.cpu().numpy(), ds[k])                      # input order preserved (R3)
    c8 = torch.from_numpy(rng.integers(0, 256, size=(len(ds), 3), dtype=np.uint8)).cuda()
    assert np.array_equal(ctx.compact_rows(c8, keep).cpu().numpy(), c8.cpu().numpy()[k])


def test_knn_grid_vs_bruteforce(ctx, oracle):
    """the oracle's own grid KNN is checked against brute force, then the GPU mean distances."""
    import torch
    rng = np.random.default_rng(11)
    pts = np.vstack([rng.normal(size=(3000, 3)), rng.normal(size=(50, 3)) * 20.0])
    _, mean_o, _ = oracle.statistical_outlier(pts, 20, 2.0)
    for i in rng.integers(0, len(pts), size=40):
        d2, _ = oracle.knn_bruteforce(pts, pts[i], 20)
        assert abs(np.sqrt(d2).mean() - mean_o[i]) <= 1e-12 * max(1.0, mean_o[i])
    _, mean, _, _ = ctx.statistical_outlier(torch.from_numpy(pts).cuda(), 20, 2.0)
    assert np.allclose(mean.cpu().numpy(), mean_o, rtol=1e-12)


def test_normals_vs_oracle(ctx, oracle):
    import torch
    pts, cols = cloud(2)
    ds = oracle.voxel_downsample(pts.astype(np.float64), cols, 0.02)["points"].astype(np.float32)
    cam = np.array([0.0, 0.0, 0.0])
    n_o = oracle.estimate_normals(ds, 30, orient_to=cam)
    n_g = ctx.estimate_normals(torch.from_numpy(ds).cuda(), 30, orient_to=cam).cpu().numpy()
    dots = np.abs((n_o * n_g).sum(1))
    assert (dots >= 1 - 1e-6).mean() > 0.999 and dots.min() > 0.99       # |n.n_ref| >= 1-1e-6
    # analytic check: a plane z = 0.3x + 0.1y has normal ~ (-0.3,-0.1,1)
    g = np.stack(np.meshgrid(np.linspace(0, 1, 60), np.linspace(0, 1, 60)), -1).reshape(-1, 2)
    plane = np.column_stack([g, 0.3 * g[:, 0] + 0.1 * g[:, 1]]).astype(np.float32)
    n = ctx.estimate_normals(torch.from_numpy(plane).cuda(), 30).cpu().numpy()
    ref = np.array([-0.3, -0.1, 1.0]) / np.linalg.norm([-0.3, -0.1, 1.0])
    assert np.abs(np.abs(n @ ref) - 1).max() < 1e-5
    tiny = ctx.estimate_normals(torch.tensor([[0., 0., 0.], [1., 0., 0.]], device="cuda"), 30).cpu().numpy()
    assert np.array_equal(tiny, [[0, 0, 1], [0, 0, 1]])                  # < 3 neighbours -> (0,0,1)


def test_icp_vs_oracle(ctx, oracle):
    import torch
    pts, cols = cloud(2)
    tgt = oracle.voxel_downsample(pts.astype(np.float64), cols, 0.02)["points"].astype(np.float32)
    nrm = oracle.estimate_normals(tgt, 30, orient_to=np.zeros(3))
    a = 0.01
    Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    T_true = np.eye(4); T_true[:3, :3] = Rz; T_true[:3, 3] = [0.01, -0.008, 0.012]
    src = ((tgt[::3] - T_true[:3, 3]) @ Rz).astype(np.float32)            # src = T_true^-1 tgt
    o = oracle.icp_point_to_plane(src, tgt, nrm, 0.05, max_iter=30)
    g = ctx.icp_point_to_plane(torch.from_numpy(src).cuda(), torch.from_numpy(tgt).cuda(),
                               torch.from_numpy(nrm).cuda(), 0.05, max_iter=30)
    # first linearisation: 27 normal-equation terms
    a27, sd2, cnt = ctx.icp_linearize(torch.from_numpy(src).cuda(), torch.from_numpy(tgt).cuda(),
                                      torch.from_numpy(nrm).cuda(), 0.05, np.eye(4))
    assert cnt == o["acc_first"][28]                                      # correspondence count exact
    assert np.allclose(a27, o["acc_first"][:27], rtol=1e-9, atol=1e-12)
    assert np.isclose(sd2, o["acc_first"][27], rtol=1e-9)
    assert g.iterations == o["iterations"]
    dT = g.transformation @ np.linalg.inv(o["T"])
    ang = np.arccos(np.clip((np.trace(dT[:3, :3]) - 1) / 2, -1, 1))
    assert ang <= 1e-4 and np.linalg.norm(dT[:3, 3]) <= 1e-4              # north_star ICP tolerance
    assert abs(g.fitness - o["fitness"]) < 1e-9 and abs(g.inlier_rmse - o["inlier_rmse"]) < 1e-9
    # and it actually recovers the motion
    dT = g.transformation @ np.linalg.inv(T_true)
    assert np.linalg.norm(dT[:3, 3]) < 2e-3
    idx, d2 = ctx.nearest_neighbor(torch.from_numpy(src[:500]).cuda(), torch.from_numpy(tgt).cuda(), 0.05)
    oi, od = oracle.nearest_neighbor(src[:500], tgt, 0.05)
    assert np.array_equal(idx.cpu().numpy(), oi)


def _brute_nn(src, tgt, T, r):
    """Exact reference of the registration's search: q = T p accumulated left to right in f64, candidates from a
    KD-tree (8 nearest), d2 = (dx^2 + dy^2) + dz^2 in f64, smallest (d2, index) within r."""
    from scipy.spatial import cKDTree
    p = src.astype(np.float64)
    q = np.stack([((T[k, 0] * p[:, 0] + T[k, 1] * p[:, 1]) + T[k, 2] * p[:, 2]) + T[k, 3] for k in range(3)], 1)
    t64 = tgt.astype(np.float64)
    _, cand = cKDTree(t64).query(q, k=8)
    d = t64[cand] - q[:, None, :]
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    order = np.lexsort((cand, d2), axis=1)[:, 0]
    rows = np.arange(len(q))
    idx, best = cand[rows, order], d2[rows, order]
    ok = best <= r * r
    return np.where(ok, idx, -1).astype(np.int32), np.where(ok, best, r * r)


@pytest.mark.parametrize("offset", [0.0, 31.0, 900.0])
def test_icp_search_equals_brute_force(ctx, offset):
    """icp_nn_kernel (f32 rejection tests, f64 decision) against an exact f64 reference: a noisy surface with
    coordinates up to `offset` metres (the f32 rounding of the query grows with them), cold and warm-seeded."""
    import torch
    rng = np.random.default_rng(5)
    g = np.stack(np.meshgrid(np.arange(420), np.arange(400)), -1).reshape(-1, 2) * 0.01
    tgt = np.column_stack([g[:, 0], g[:, 1], 0.3 * np.sin(3 * g[:, 0]) + 0.2 * np.cos(2 * g[:, 1])])
    tgt = (tgt + rng.normal(0, 0.002, tgt.shape) + offset).astype(np.float32)
    tgt[1000:1200] = tgt[3000:3200]                                      # exact duplicates: ties -> lowest index
    nrm = np.tile(np.array([0, 0, 1], np.float32), (len(tgt), 1))
    sel = rng.choice(len(tgt), 40000, replace=False)
    src = (tgt[sel] + rng.normal(0, 0.004, (len(sel), 3))).astype(np.float32)
    src[:300] += 0.2                                                      # some queries without a correspondence
    src[300:400] = tgt[sel[300:400]]                                      # zero-distance queries
    a = 0.002
    T0 = np.eye(4); T0[:3, :3] = [[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]
    c = tgt.mean(0).astype(np.float64)
    T0[:3, 3] = c - T0[:3, :3] @ c + [0.004, -0.003, 0.002]               # small motion about the cloud centre
    T1 = T0.copy(); T1[:3, 3] += [-0.003, 0.002, 0.001]
    S, Tg, Ng = (torch.from_numpy(x).cuda() for x in (src, tgt, nrm))
    for Ta, Tb in ((T0, None), (T0, T1)):
        idx, d2 = ctx.icp_correspondences(S, Tg, Ng, 0.05, Ta, Tb)
        ri, rd = _brute_nn(src, tgt, Ta if Tb is None else Tb, 0.05)
        idx, d2 = idx.cpu().numpy(), d2.cpu().numpy()
        assert np.array_equal(idx, ri), (offset, Tb is not None, int((idx != ri).sum()))
        assert np.array_equal(d2[ri >= 0], rd[ri >= 0])
    assert (ri >= 0).mean() > 0.9 and (ri < 0).sum() >= 200


def test_ply_writers(tmp_path):
    """K9 host writers: reference ASCII layout byte-for-byte vs the reference's own file;
    Open3D binary layout structurally (R9)."""
    from conftest import GOLDEN
    from textureless_3d_reconstruction_b200 import _lib
    from textureless_3d_reconstruction_b200.runtime import write_ply
    z = np.load(GOLDEN / "k9_ply_ascii.npz")
    f = tmp_path / "a.ply"
    write_ply(f, z["points"], z["colors"], layout=_lib.PLY_REF_ASCII)
    assert f.read_bytes() == bytes(z["ply_f32"])
    write_ply(f, z["points64"], z["colors"][:50], layout=_lib.PLY_REF_ASCII)
    assert f.read_bytes() == bytes(z["ply_f64"])
    write_ply(f, z["points"], z["colors"], layout=_lib.PLY_O3D_BINARY)
    raw = f.read_bytes()
    head, body = raw.split(b"end_header\n", 1)
    assert head.decode().splitlines() == [
        "ply", "format binary_little_endian 1.0", "comment Created by Open3D", "element vertex 200",
        "property double x", "property double y", "property double z",
        "property uchar red", "property uchar green", "property uchar blue"]
    rec = np.frombuffer(body, dtype=np.dtype([("p", "<f8", 3), ("c", "u1", 3)]))
    assert np.array_equal(rec["p"], z["points"].astype(np.float64)) and np.array_equal(rec["c"], z["colors"])


def test_device_synth_matches_numpy_twin(ctx):
    H, W = 240, 136
    it = S.scaled_intrinsics(H, W)
    for scene in (0, 1):
        d, c, T = ctx.synth_frame(scene, 5, H, W, it["fx"], it["fy"], it["cx"], it["cy"], noise_sigma=0.002)
        dn, cn, Tn = S.synth_frame(scene, 5, H, W, it["fx"], it["fy"], it["cx"], it["cy"], noise_sigma=0.002)
        d = d.cpu().numpy()
        assert np.allclose(T, Tn, atol=1e-12)
        assert np.array_equal(c.cpu().numpy(), cn)
        assert np.array_equal(np.isnan(d), np.isnan(dn)) and np.array_equal(np.isinf(d), np.isinf(dn))
        fin = np.isfinite(dn)
        assert np.abs(d[fin] - dn[fin]).max() < 2e-5                     # libm vs CUDA transcendentals


def test_voxel_downsample_full_size_properties(ctx, oracle):
    """cfg-4 style stress at 20 M points (fused 1080x1920 tunnel clouds): count conservation,
    idempotence of the voxel index set, agreement of a 2 M prefix with the oracle (bit-exact indices)."""
    import torch
    H, W = 1920, 1080
    K = (1719.0, 1719.0, 540.0, 960.0)
    nf = 10
    ds, cs, poses = [], [], []
    for i in range(nf):
        d, c, T = ctx.synth_frame(0, i, H, W, *K, noise_sigma=0.002)
        ds.append(d)
        cs.append(c)
        poses.append((T[:, :3].copy(), T[:, 3:4].copy()))
    fr = ctx.make_backproject_frames(ds, cs, poses)
    xyz, rgb, offs = ctx.backproject_batch(fr, nf, H, W, fx=K[0], fy=K[1], cx=K[2], cy=K[3], max_depth=5.0)
    n = int(offs[-1].item())
    assert n > 10_000_000
    p, c = xyz[:n].contiguous(), rgb[:n].contiguous()
    for v in (0.005, 0.02):
        r = ctx.voxel_downsample(p, c, v, sorted_output=True, want_idx=True)
        assert int(r["count"].to(torch.int64).sum().item()) == n                      # nothing lost or doubled
        assert int(r["rgb_sum"].to(torch.int64).sum().item()) == int(c.to(torch.int64).sum().item())
        idx = r["idx"]
        assert torch.equal(torch.unique(idx, dim=0), idx)                               # unique + (x,y,z)-sorted
        centres = ((idx.to(torch.float64) + 0.5) * v + torch.from_numpy(r["min_bound"]).cuda()).contiguous()
        r2 = ctx.voxel_downsample(centres, None, v, min_bound=r["min_bound"], sorted_output=True, want_idx=True)
        assert r2["m"] == r["m"] and torch.equal(r2["idx"], idx)                        # idempotent
        lo = p.min(0).values.double().cpu().numpy()
        assert np.all(r["points"].min(0).values.cpu().numpy() >= lo - 1e-9)
    ns = 2_000_000
    g = ctx.voxel_downsample(p[:ns].contiguous(), c[:ns].contiguous(), 0.01, sorted_output=True, want_idx=True)
    o = oracle.voxel_downsample(p[:ns].cpu().numpy().astype(np.float64), c[:ns].cpu().numpy(), 0.01)
    oo = np.lexsort(o["idx"].T[::-1])
    assert np.array_equal(g["idx"].cpu().numpy(), o["idx"][oo])
    assert np.array_equal(g["count"].cpu().numpy().astype(np.uint32), o["count"][oo])
    assert np.allclose(g["points"].cpu().numpy(), o["points"][oo], rtol=1e-12, atol=0)
    # colours: trunc(mean(c/255)*255) can flip by 1 LSB at exact integers (f64 summation order)
    dc = np.abs(g["colors"].cpu().numpy().astype(np.int16) - o["colors_u8"][oo].astype(np.int16))
    assert dc.max() <= 1 and (dc != 0).mean() < 0.02


@pytest.mark.parametrize("world", [1, 3, 8])
def test_voxel_partials_then_merge_equals_downsample(ctx, world):
    """Sharded K2 building blocks on one GPU: split the cloud in `world` shards, take each shard's
    per-voxel partial sums (records grouped by owner), hand every owner its records, merge — the
    union of the owners' voxels is bit-identical to one voxel_downsample over all points."""
    import torch
    rng = np.random.default_rng(21)
    n = 200_000
    p = (rng.uniform(-2.0, 2.0, (n, 3)) * np.array([1.0, 0.6, 2.5])).astype(np.float32)
    p[::11] = p[5]
    c = rng.integers(0, 256, (n, 3), dtype=np.uint8)
    xyz, rgb = torch.from_numpy(p).cuda(), torch.from_numpy(c).cuda()
    voxel = 0.03
    ref = ctx.voxel_downsample(xyz, rgb, voxel, sorted_output=True)
    mn, mx = ctx.bounds(xyz)
    minb = mn - voxel * 0.5
    assert np.array_equal(minb, ref["min_bound"])
    cuts = np.linspace(0, n, world + 1).astype(int)
    inbox = [[] for _ in range(world)]
    sent = 0
    for r in range(world):
        rec, counts = ctx.voxel_partials(xyz[cuts[r]:cuts[r + 1]].contiguous(), rgb[cuts[r]:cuts[r + 1]].contiguous(),
                                         voxel, minb, mx, world)
        cl = counts.cpu().tolist()
        assert sum(cl) == rec.shape[0] and rec.shape[1] == 56
        off = 0
        for d in range(world):
            inbox[d].append(rec[off:off + cl[d]])
            off += cl[d]
        sent += rec.shape[0]
    outs = [ctx.voxel_merge_partials(torch.cat(inbox[d]).contiguous(), True, voxel, minb, mx) for d in range(world)]
    assert sum(o["m"] for o in outs) == ref["m"] and sent >= ref["m"]
    idx = torch.cat([o["idx"] for o in outs]).cpu().numpy()
    pts = torch.cat([o["points"] for o in outs]).cpu().numpy()
    col = torch.cat([o["colors"] for o in outs]).cpu().numpy()
    cnt = torch.cat([o["count"] for o in outs]).cpu().numpy()
    order = np.lexsort(idx[:, ::-1].T)
    assert np.array_equal(idx[order], ref["idx"].cpu().numpy())                       # ascending (x,y,z) like sorted K2
    assert np.array_equal(pts[order].view(np.uint64), ref["points"].cpu().numpy().view(np.uint64))   # bit-identical means
    assert np.array_equal(col[order], ref["colors"].cpu().numpy())
    assert np.array_equal(cnt[order], ref["count"].cpu().numpy())
    # every owner holds only its own voxels, each voxel exactly once
    assert len(np.unique(idx, axis=0)) == len(idx)


@pytest.mark.parametrize("parts", [1, 2, 5])
def test_sor_parts_tile_the_cloud_and_equal_the_single_call(ctx, parts):
    """Sharded K3 building blocks on one GPU: the `parts` shares of the grid-sorted cloud cover every
    point exactly once, and mean distances / mu / sigma / mask equal t3d_statistical_outlier's."""
    import torch
    pts, _ = cloud(2)
    ds = ctx.voxel_downsample(torch.from_numpy(pts).cuda(), None, 0.01)["points"].contiguous()
    n = ds.shape[0]
    keep0, mean0, stats0, kept0 = ctx.statistical_outlier(ds, 20, 2.0)
    acc = torch.full((n,), float("-inf"), dtype=torch.float64, device=ds.device)
    covered = torch.zeros(n, dtype=torch.int32, device=ds.device)
    for r in range(parts):
        part = torch.full((n,), float("-inf"), dtype=torch.float64, device=ds.device)
        ctx.sor_mean_distances_part(ds, 20, r, parts, part)
        covered += torch.isfinite(part).to(torch.int32)
        acc = torch.maximum(acc, part)                              # what all_reduce(MAX) does
    assert int(covered.min()) == 1 and int(covered.max()) == 1
    assert torch.equal(acc.view(torch.int64), mean0.view(torch.int64))       # bit-identical mean distances
    keep, stats, kept = ctx.sor_from_mean_distances(acc, 2.0)
    assert stats == stats0 and kept == kept0 and torch.equal(keep, keep0)
