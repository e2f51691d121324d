"""Hand-computed known-answer tests for the oracle's restatement of the Open3D stages
(SURVEY.md §8c R2..R8).  The reference holds no vectors for these, so these KATs are
the only pin ("parity unpinned" against Open3D itself)."""
import numpy as np
import pytest


# ------------------------------------------------------------------ R2 voxel_down_sample
def test_r2_known_answer(oracle):
    v = 1.0
    pts = np.array([[0.0, 0.0, 0.0],      # min point -> minb = -0.5 -> idx (0,0,0)
                    [0.4, 0.4, 0.4],      # (0.9,0.9,0.9) -> idx 0   (same voxel)
                    [0.5, 0.0, 0.0],      # (1.0,..) exactly on a face -> idx (1,0,0)
                    [2.6, 0.0, 0.0],      # 3.1 -> idx (3,0,0)
                    [2.7, 0.2, 0.1]])     # 3.2 -> idx (3,0,0)
    cols = np.array([[10, 20, 30], [30, 40, 50], [255, 0, 7], [1, 1, 1], [2, 3, 4]], np.uint8)
    o = oracle.voxel_downsample(pts, cols, v)
    assert np.array_equal(o["min_bound"], [-0.5, -0.5, -0.5])
    assert np.array_equal(o["idx"], [[0, 0, 0], [1, 0, 0], [3, 0, 0]])
    assert np.array_equal(o["count"], [2, 1, 2])
    assert np.allclose(o["points"], [[0.2, 0.2, 0.2], [0.5, 0, 0], [2.65, 0.1, 0.05]], rtol=0, atol=1e-15)
    # colours: mean of c/255 then (x*255).astype(u8) (truncation, d2r:418)
    assert np.array_equal(o["colors_u8"][1], [255, 0, 7])            # single point round-trips
    assert np.array_equal(o["colors_u8"][0], [20, 30, 40])
    assert np.array_equal(o["colors_u8"][2], [1, 2, 2])              # (1+2)/2=1.5->1, (1+3)/2=2, (1+4)/2=2.5->2
    with pytest.raises(ValueError):
        oracle.voxel_downsample(pts, cols, 0.0)
    with pytest.raises(ValueError):
        oracle.voxel_downsample(np.array([[0, 0, 0], [1e6, 0, 0.0]]), None, 1e-5)   # v*INT_MAX < extent
    assert len(oracle.voxel_downsample(np.zeros((0, 3)), None, 1.0)["points"]) == 0


def test_r2_all_u8_values_round_trip(oracle):
    pts = np.column_stack([np.arange(256) * 10.0, np.zeros(256), np.zeros(256)])
    cols = np.repeat(np.arange(256, dtype=np.uint8)[:, None], 3, axis=1)
    o = oracle.voxel_downsample(pts, cols, 1.0)
    assert np.array_equal(o["colors_u8"], cols)                       # exact for all 256 values


# ------------------------------------------------------------------ R3 statistical outlier
def test_r3_known_answer(oracle):
    # 5 collinear points at x = 0,1,2,3 and an outlier at 100; nb=3 (self + 2 nearest)
    pts = np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [3, 0, 0], [100, 0, 0]], np.float64)
    keep, mean, (mu, sigma, thr) = oracle.statistical_outlier(pts, nb=3, std_ratio=1.0)
    exp = np.array([(0 + 1 + 2) / 3, (0 + 1 + 1) / 3, (0 + 1 + 1) / 3, (0 + 1 + 2) / 3, (0 + 97 + 98) / 3])
    assert np.allclose(mean, exp, rtol=1e-15)
    assert np.isclose(mu, exp.mean()) and np.isclose(sigma, exp.std(ddof=1))     # Bessel
    assert np.isclose(thr, exp.mean() + exp.std(ddof=1))
    assert keep.tolist() == [True, True, True, True, False]
    # exact duplicates: mean distance 0 -> dropped (mean_i > 0 required)
    dup = np.array([[0, 0, 0], [0, 0, 0], [5, 0, 0], [5, 0, 0]], np.float64)
    keep, mean, _ = oracle.statistical_outlier(dup, nb=2, std_ratio=2.0)
    assert np.array_equal(mean, [0, 0, 0, 0]) and not keep.any()


# ------------------------------------------------------------------ R4/R5 TSDF
def _front_plane_frame(H=16, W=16, z=1.0):
    depth = np.full((H, W), z, np.float32)
    bgr = np.zeros((H, W, 3), np.uint8)
    bgr[..., 0], bgr[..., 1], bgr[..., 2] = 30, 20, 10               # B, G, R
    K = (16.0, 16.0, 8.0, 8.0)
    T = np.eye(4)[:3]
    return depth, bgr, K, T


def test_r4_touch_known_answer(oracle):
    depth, bgr, K, T = _front_plane_frame()
    vol = oracle.TSDFVolume(voxel_size=0.05, sdf_trunc=0.2)          # block = 0.4 m
    keys = vol.touch(depth, K, T, 1.0, 3.0)
    # rays: x = (u-8)/16 * t, u in {0,4,8,12} -> x/t in {-.5,-.25,0,.25}; t in {0.8,0.9333,1.0667,1.2}
    exp = set()
    for u in (0, 4, 8, 12):
        for v in (0, 4, 8, 12):
            t = np.float32(0.8)
            step = np.float32((np.float32(1.2) - np.float32(0.8)) / np.float32(3))
            for _ in range(4):
                p = np.array([(u - 8) / 16 * t, (v - 8) / 16 * t, t], np.float32)
                exp.add(tuple(np.floor(p / np.float32(0.4)).astype(int)))
                t = np.float32(t + step)
    assert set(map(tuple, keys.tolist())) == exp
    assert len(keys) == len(exp)
    # depth 0, NaN, beyond depth_max: nothing touched
    for bad in (0.0, np.nan, 3.0, 10.0, -1.0):
        assert len(vol.touch(np.full((16, 16), bad, np.float32), K, T, 1.0, 3.0)) == 0


def test_r5_integrate_known_answer(oracle):
    depth, bgr, K, T = _front_plane_frame()
    vol = oracle.TSDFVolume(voxel_size=0.05, sdf_trunc=0.2)
    vol.integrate(depth, bgr, K, T, 1.0, 3.0)
    vol.integrate(depth, bgr, K, T, 1.0, 3.0)
    keys, tsdf, w, rgb = vol.export()
    kmap = {tuple(k): i for i, k in enumerate(keys.tolist())}
    b = kmap[(0, 0, 2)]                                               # voxels z = 0.8 .. 1.15
    def vox(x, y, z):
        return x + 8 * y + 64 * z
    # voxel (0,0,16) -> z_c = 0.8, projects to (8,8): sdf = 1 - 0.8 = 0.2 -> tsdf 1, weight 2
    i = vox(0, 0, 0)
    assert w[b, i] == 2.0 and np.isclose(tsdf[b, i], 1.0)
    assert np.allclose(rgb[b, i], [10, 20, 30])                       # stored R,G,B from BGR (30,20,10)
    # z_c = 0.95 -> sdf 0.05 -> 0.25 ; z_c = 1.1 -> sdf -0.1 -> -0.5
    assert np.isclose(tsdf[b, vox(0, 0, 3)], 0.25, atol=1e-6)
    assert np.isclose(tsdf[b, vox(0, 0, 6)], -0.5, atol=1e-6)
    # z_c = 1.2 is exactly -trunc behind the surface (not < -trunc up to rounding); 1.25 is skipped
    b3 = kmap.get((0, 0, 3))
    if b3 is not None:
        assert w[b3, vox(0, 0, 1)] == 0.0                             # z_c = 1.25 -> sdf -0.25 < -0.2
    # voxels that project outside the 16x16 image are untouched: x_c = 0.35 -> u = 16*0.35/0.8+8 = 15.0 ok,
    # x_c = 0.4 (block 1) would give u = 16 > W-1
    assert w[b, vox(7, 0, 0)] == 2.0
    c = vol.counters()
    assert c["frames"] == 2 and c["voxel_updates"] == int(w.sum())
    # running mean: a third frame at depth 1.1 moves tsdf of z_c=0.95 to (2*0.25 + 0.75)/3
    vol.integrate(np.full((16, 16), 1.1, np.float32), bgr, K, T, 1.0, 3.0)
    keys, tsdf, w, rgb = vol.export()
    b = {tuple(k): i for i, k in enumerate(keys.tolist())}[(0, 0, 2)]
    assert w[b, vox(0, 0, 3)] == 3.0
    assert np.isclose(tsdf[b, vox(0, 0, 3)], (2 * 0.25 + 0.75) / 3, atol=1e-6)


def test_r5_pixel_rounding_modes(oracle):
    depth = np.zeros((8, 8), np.float32)
    depth[4, 5] = 1.0                                                 # only pixel (u=5, v=4) is valid
    K = (8.0, 8.0, 4.0, 4.0)
    T = np.eye(4)[:3]
    # voxel at x = 0.1: u = 8*0.1/0.95+4 = 4.84 -> round 5 (hit), trunc 4 (miss)
    hits = []
    for mode in (0, 1):
        vol = oracle.TSDFVolume(0.05, 0.2, pixel_round=mode)
        vol.integrate(depth, None, K, T, 1.0, 3.0, keys=np.array([[0, 0, 2]], np.int32))
        _, _, w, _ = vol.export()
        hits.append(w[0, 2 + 8 * 0 + 64 * 3])                         # voxel (2,0,3): x=0.1, z=0.95
    assert hits == [1.0, 0.0]


# ------------------------------------------------------------------ R6 surface extraction
def test_r6_extract_known_answer(oracle):
    depth, bgr, K, T = _front_plane_frame(z=1.02)
    vol = oracle.TSDFVolume(voxel_size=0.05, sdf_trunc=0.2)
    for _ in range(3):
        vol.integrate(depth, bgr, K, T, 1.0, 3.0)
    xyz, nrm, rgb = vol.extract_points(3.0)
    assert len(xyz) > 20
    assert np.allclose(xyz[:, 2], 1.02, atol=1e-6)                    # zero crossing between z=1.0 and 1.05
    assert np.allclose(np.abs(nrm[:, 2]), 1.0, atol=1e-5) and (nrm[:, 2] < 0).all()   # gradient points to camera
    assert np.array_equal(np.unique(rgb, axis=0), [[10, 20, 30]])
    assert len(vol.extract_points(4.0)[0]) == 0                       # weight threshold (W >= thr)


# ------------------------------------------------------------------ R7 normals
def test_r7_known_answer(oracle):
    g = np.stack(np.meshgrid(np.arange(12.0), np.arange(12.0)), -1).reshape(-1, 2)
    plane = np.column_stack([g[:, 0], g[:, 1], 2 * g[:, 0] + 1]).astype(np.float32)      # z = 2x + 1
    n = oracle.estimate_normals(plane, knn=9)
    ref = np.array([-2, 0, 1]) / np.sqrt(5)
    assert np.allclose(np.abs(n @ ref), 1.0, atol=1e-6)
    n2 = oracle.estimate_normals(plane, knn=9, orient_to=np.array([0.0, 0.0, 100.0]))
    assert (n2 @ np.array([0, 0, 1.0]) > 0).all()
    assert np.array_equal(oracle.estimate_normals(plane[:2], knn=9), [[0, 0, 1], [0, 0, 1]])


# ------------------------------------------------------------------ R8 ICP
def test_r8_known_answer(oracle):
    rng = np.random.default_rng(1)
    g = np.stack(np.meshgrid(np.linspace(-1, 1, 50), np.linspace(-1, 1, 50)), -1).reshape(-1, 2)
    z = 0.3 * np.sin(2 * g[:, 0]) * np.cos(1.5 * g[:, 1])
    tgt = np.column_stack([g, z]).astype(np.float32)
    dzdx = 0.6 * np.cos(2 * g[:, 0]) * np.cos(1.5 * g[:, 1])
    dzdy = -0.45 * np.sin(2 * g[:, 0]) * np.sin(1.5 * g[:, 1])
    n = np.column_stack([-dzdx, -dzdy, np.ones(len(g))])
    nrm = (n / np.linalg.norm(n, axis=1, keepdims=True)).astype(np.float32)
    t_true = np.array([0.004, -0.006, 0.008])
    src = (tgt[::2] - t_true.astype(np.float32)).astype(np.float32)
    r = oracle.icp_point_to_plane(src, tgt, nrm, 0.05, max_iter=30)
    assert np.allclose(r["T"][:3, 3], t_true, atol=2e-3) and np.allclose(r["T"][:3, :3], np.eye(3), atol=2e-3)
    assert r["fitness"] > 0.95 and 1 <= r["iterations"] <= 30
    # no correspondences (clouds 10 m apart) -> singular system -> identity, fitness 0
    far = oracle.icp_point_to_plane(src + 10.0, tgt, nrm, 0.05, max_iter=5)
    assert np.array_equal(far["T"], np.eye(4)) and far["fitness"] == 0.0
    # Jacobian layout: one correspondence with s=(1,2,3), t=(1,2,2.9), n=(0,0,1)
    one = oracle.icp_point_to_plane(np.array([[1, 2, 3.0]], np.float32), np.array([[1, 2, 2.9]], np.float32),
                                    np.array([[0, 0, 1.0]], np.float32), 0.5, max_iter=0)["acc_first"]
    J = np.array([2.0, -1.0, 0.0, 0.0, 0.0, 1.0])                      # [s x n ; n]
    JtJ = np.outer(J, J)[np.triu_indices(6)]
    assert np.allclose(one[:21], JtJ) and np.allclose(one[21:27], J * np.float32(3.0 - np.float32(2.9)), atol=1e-7)
    assert one[28] == 1.0 and np.isclose(one[27], 0.01, atol=1e-6)


def test_tsdf_confidence_mask_equals_zeroed_depth(oracle):
    """Confidence masking (north_star; no reference code): masked pixels behave exactly like pixels without depth."""
    from textureless_3d_reconstruction_b200 import synthetic as S
    H, W = 120, 68
    it = S.scaled_intrinsics(H, W)
    K = (it["fx"], it["fy"], it["cx"], it["cy"])
    rng = np.random.default_rng(3)
    a, b = oracle.TSDFVolume(0.02, 0.08), oracle.TSDFVolume(0.02, 0.08)
    for i in range(3):
        d, c, T = S.synth_frame(0, i, H, W, *K, noise_sigma=0.002)
        m = (rng.random((H, W)) > 0.4).astype(np.uint8)
        a.integrate(d, c, K, T, 1.0, 5.0, conf_mask=m)
        dz = d.copy()
        dz[m == 0] = 0
        b.integrate(dz, c, K, T, 1.0, 5.0)
    assert a.counters() == b.counters() and a.counters()["voxel_updates"] > 1000
    for x, y in zip(a.export(), b.export()):
        assert np.array_equal(x, y)
