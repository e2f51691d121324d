"""K1 parity: CUDA back-projection vs vectors produced by the unmodified reference
(tests/golden/k1_backproject.npz) and vs the oracle at full 1080x1920 size."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SCALES = {"pyfloat1": 1.0, "pyfloat": 1.37, "npf64": np.float64(1.37), "npf64_1": np.float64(1.0)}
REL_TOL = 1e-5          # north_star: <= 1e-5 relative on point coordinates


def poses_of(z):
    return {"none": None, "identity": (np.eye(3), np.zeros((3, 1))),
            "rotated": (z["pose_rotated_R"], z["pose_rotated_t"])}


def close(a, b):
    return np.allclose(a, b, rtol=REL_TOL, atol=1e-12)


def test_golden_all_reference_copies(k1_golden, ctx):
    from textureless_3d_reconstruction_b200 import depth_enhanced_reconstruction as der
    from textureless_3d_reconstruction_b200 import depth_processor as dp
    from textureless_3d_reconstruction_b200 import depth_to_reconstruction as d2r
    z, meta = k1_golden
    fx, fy, cx, cy = z["intrinsics"]
    H, W = z["depth"].shape
    poses = poses_of(z)
    dense = d2r.DenseReconstructor(d2r.ReconstructionConfig(fx=fx, fy=fy, cx=cx, cy=cy))
    gen = der.DensePointCloudGenerator(der.CameraIntrinsics(fx=fx, fy=fy, cx=cx, cy=cy, width=W, height=H))
    exact = total = 0
    for m in meta:
        k = m["id"]
        if m["kind"] in ("d2r", "d2r_zero"):
            depth, color = (z["depth"], z["color"]) if m["kind"] == "d2r" else (
                np.zeros((8, 12), np.float32), np.zeros((8, 12, 3), np.uint8))
            pts, cols = dense.depth_to_pointcloud(depth, color, pose=poses[m["pose"]], scale=SCALES[m["scale"]],
                                                  subsample=m["subsample"])
            g_pts, g_cols = z[f"d2r_{k}_pts"], z[f"d2r_{k}_cols"]
        elif m["kind"] in ("der", "der_b"):
            depth, color = (z["depth"], z["color"]) if m["kind"] == "der" else (z["depth_b"], z["color_b"])
            if m["depth"] == "f64":
                depth = depth * np.float64(0.83)
            pts, cols = gen.depth_to_pointcloud(depth, color, pose=poses[m["pose"]], subsample=m["subsample"])
            g_pts, g_cols = z[f"der_{k}_pts"], z[f"der_{k}_cols"]
        else:
            g = dp.PointCloudGenerator(dp.CameraIntrinsics(fx=fx, fy=fy, cx=cx, cy=cy, width=W, height=H),
                                       downsample_factor=m["downsample"])
            pts, cols = g.generate(z["depth"], z["color"] if m["rgb"] else None, max_depth=20.0, min_depth=0.1)
            g_pts = z[f"dp_{k}_pts"]
            g_cols = z[f"dp_{k}_cols"] if m["rgb"] else None
        assert pts.dtype == np.float32 and pts.shape == g_pts.shape, m   # mask + count bit-exact
        assert close(pts, g_pts), m
        if g_cols is None:
            assert cols is None
        else:
            assert cols.dtype == g_cols.dtype and np.array_equal(cols, g_cols), m   # colours + order exact
        exact += int((pts.view(np.uint32) == g_pts.view(np.uint32)).sum())
        total += pts.size
    assert exact / max(total, 1) > 0.999, exact / total


@pytest.mark.parametrize("subsample", [1, 2, 4])
def test_full_size_vs_oracle(ctx, oracle, subsample):
    """cfg-1 frame (S1, W=1080 H=1920, NaN/inf/zero/border pixels) on device vs the C oracle."""
    import torch
    H, W = 1920, 1080
    fx, fy, cx, cy = 1719.0, 1719.0, 540.0, 960.0
    depth, bgr, T = ctx.synth_frame(1, 7, H, W, fx, fy, cx, cy)
    pose = (T[:, :3].copy(), T[:, 3:4].copy())
    xyz, rgb, n = ctx.backproject(depth, bgr, fx=fx, fy=fy, cx=cx, cy=cy, subsample=subsample, pose=pose)
    n = int(n.item())
    o_xyz, o_rgb = oracle.backproject(depth.cpu().numpy(), bgr.cpu().numpy(), fx, fy, cx, cy, pose=pose,
                                      subsample=subsample)
    assert n == len(o_xyz)
    got = xyz[:n].cpu().numpy()
    assert np.array_equal(rgb[:n].cpu().numpy(), o_rgb)
    assert close(got, o_xyz)
    assert (got.view(np.uint32) == o_xyz.view(np.uint32)).mean() > 0.9999
    # size-independent property: every output point re-projects onto its own pixel
    Pc = (pose[0] @ got.astype(np.float64).T + pose[1]).T
    u = fx * Pc[:, 0] / Pc[:, 2] + cx
    v = fy * Pc[:, 1] / Pc[:, 2] + cy
    assert np.abs(u - np.round(u)).max() < 1e-2 and np.abs(v - np.round(v)).max() < 1e-2
    lin = np.round(v).astype(np.int64) * W + np.round(u).astype(np.int64)
    assert np.all(np.diff(lin) > 0)                     # row-major order, no duplicates


def test_edge_shapes_and_conf_mask(ctx, oracle):
    import torch
    rng = np.random.default_rng(5)
    for (H, W, s) in [(1, 1, 1), (5, 3, 2), (64, 32, 1), (33, 2049, 1), (7, 4100, 3), (300, 17, 7)]:
        depth = rng.uniform(0.0, 3.0, size=(H, W)).astype(np.float32)
        bgr = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        d, c = torch.from_numpy(depth).cuda(), torch.from_numpy(bgr).cuda()
        xyz, rgb, n = ctx.backproject(d, c, fx=50.0, fy=51.0, cx=W / 2, cy=H / 2, subsample=s, min_depth=0.5,
                                      max_depth=2.5)
        n = int(n.item())
        o_xyz, o_rgb = oracle.backproject(depth, bgr, 50.0, 51.0, W / 2, H / 2, min_depth=0.5, max_depth=2.5,
                                          subsample=s)
        assert n == len(o_xyz), (H, W, s)
        assert np.array_equal(xyz[:n].cpu().numpy().view(np.uint32), o_xyz.view(np.uint32))
        assert np.array_equal(rgb[:n].cpu().numpy(), o_rgb)
    # confidence mask extension: identical to zeroing the depth of masked pixels
    H, W = 40, 50
    depth = rng.uniform(0.5, 3.0, size=(H, W)).astype(np.float32)
    bgr = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    conf = (rng.uniform(size=(H, W)) > 0.3).astype(np.uint8)
    xyz, rgb, n = ctx.backproject(torch.from_numpy(depth).cuda(), torch.from_numpy(bgr).cuda(), fx=50., fy=50.,
                                  cx=25., cy=20., conf_mask=torch.from_numpy(conf).cuda())
    o_xyz, _ = oracle.backproject(np.where(conf > 0, depth, 0).astype(np.float32), bgr, 50., 50., 25., 20.)
    assert int(n.item()) == len(o_xyz)
    assert np.array_equal(xyz[: len(o_xyz)].cpu().numpy(), o_xyz)


@pytest.mark.parametrize("n_frames,H,W,s", [(1, 33, 47, 1), (5, 120, 68, 2), (40, 64, 48, 1), (70, 31, 29, 3)])
def test_batch_equals_per_frame_vstack(ctx, oracle, n_frames, H, W, s):
    """t3d_backproject_batch == the reference's per-frame loop + np.vstack (d2r:566-581, 401-402):
    bit-identical rows, frame-ordered, offsets = per-frame counts; > 32 frames spans several launches."""
    import torch
    from textureless_3d_reconstruction_b200 import synthetic as S
    it = S.scaled_intrinsics(H, W)
    K = (it["fx"], it["fy"], it["cx"], it["cy"])
    depths, bgrs, poses = [], [], []
    rng = np.random.default_rng(n_frames)
    for i in range(n_frames):
        d, c, T = S.synth_frame(1, i, H, W, *K)
        d = d.copy()
        d[rng.uniform(size=d.shape) < 0.2] = 0.0            # ragged validity per frame
        if i == 1:
            d[:] = 0.0                                        # an empty frame in the middle
        depths.append(torch.from_numpy(d).cuda())
        bgrs.append(torch.from_numpy(c).cuda())
        poses.append((T[:, :3].copy(), T[:, 3:4].copy()))
    frames = ctx.make_backproject_frames(depths, bgrs, poses)
    xyz, rgb, offs = ctx.backproject_batch(frames, n_frames, H, W, fx=K[0], fy=K[1], cx=K[2], cy=K[3], subsample=s,
                                           max_depth=50.0)
    offs = offs.cpu().numpy()
    assert offs[0] == 0 and np.all(np.diff(offs) >= 0)
    for i in range(n_frames):
        o_xyz, o_rgb = oracle.backproject(depths[i].cpu().numpy(), bgrs[i].cpu().numpy(), *K, pose=poses[i],
                                          max_depth=50.0, subsample=s)
        a, b = int(offs[i]), int(offs[i + 1])
        assert b - a == len(o_xyz), i
        assert np.array_equal(xyz[a:b].cpu().numpy().view(np.uint32), o_xyz.view(np.uint32)), i
        assert np.array_equal(rgb[a:b].cpu().numpy(), o_rgb), i
    if n_frames > 1:
        assert offs[2] == offs[1]                             # the empty frame owns no rows


def test_capacity_and_argument_errors(ctx):
    import torch
    from textureless_3d_reconstruction_b200._lib import T3DError
    d = torch.ones((8, 8), device="cuda")
    small = torch.empty((3, 3), device="cuda")
    with pytest.raises(T3DError):
        ctx.backproject(d, None, fx=1., fy=1., cx=0., cy=0., out_xyz=small)
    with pytest.raises(T3DError):
        ctx.backproject(d, None, fx=1., fy=1., cx=0., cy=0., subsample=0)


def _misaligned(t):
    """Same values at an address that is NOT 16-byte aligned: forces the generic K1 kernel."""
    import torch
    flat = torch.empty(t.numel() + 16, dtype=t.dtype, device=t.device)
    off = 1 if t.element_size() >= 4 else 3
    view = flat[off:off + t.numel()].view(t.shape)
    view.copy_(t)
    assert view.data_ptr() % 16 != 0 and view.is_contiguous()
    return view


@pytest.mark.parametrize("mode", ["f32", "f64scale", "f64depth"])
@pytest.mark.parametrize("rgb_f32,with_pose,with_conf", [(False, True, False), (True, False, True), (False, False, True)])
def test_streaming_kernel_equals_generic_kernel(ctx, oracle, mode, rgb_f32, with_pose, with_conf):
    """The s = 1 TMA-staged kernel (16-byte aligned frames) and the generic kernel (any alignment)
    must agree bit for bit in every mode, on sizes with many full tiles + a ragged last tile."""
    import torch
    H, W = 397, 1083                                    # 429 951 px = 209 full tiles + 1919 px, rows wrap inside threads
    fx, fy, cx, cy = 1719.0, 1719.0, W / 2, H / 2
    depth, bgr, T = ctx.synth_frame(1, 3, H, W, fx, fy, cx, cy)
    rng = np.random.default_rng(11)
    conf = torch.from_numpy((rng.uniform(size=(H, W)) > 0.25).astype(np.uint8)).cuda() if with_conf else None
    pose = (T[:, :3].copy(), T[:, 3:4].copy()) if with_pose else None
    kw = dict(fx=fx, fy=fy, cx=cx, cy=cy, pose=pose, rgb_out_f32=rgb_f32, scale=1.37, min_depth=0.2, max_depth=40.0,
              scale_is_f64=(mode != "f32"))
    d = depth.double() if mode == "f64depth" else depth
    assert d.data_ptr() % 16 == 0 and bgr.data_ptr() % 16 == 0
    a = ctx.backproject(d, bgr, conf_mask=conf, **kw)
    b = ctx.backproject(_misaligned(d), _misaligned(bgr), conf_mask=None if conf is None else _misaligned(conf), **kw)
    n = int(a[2].item())
    assert n == int(b[2].item()) and 1000 < n < H * W
    assert torch.equal(a[0][:n].view(torch.int32), b[0][:n].view(torch.int32))
    assert torch.equal(a[1][:n], b[1][:n])
    if mode != "f64depth" and not rgb_f32:
        dm = depth.cpu().numpy()
        if conf is not None:
            dm = np.where(conf.cpu().numpy() > 0, dm, 0).astype(np.float32)
        o_xyz, o_rgb = oracle.backproject(dm, bgr.cpu().numpy(), fx, fy, cx, cy, scale=1.37, f64_mask=(mode != "f32"),
                                          min_depth=0.2, max_depth=40.0, pose=pose)
        assert n == len(o_xyz)
        assert np.array_equal(a[1][:n].cpu().numpy(), o_rgb)
        assert close(a[0][:n].cpu().numpy(), o_xyz)


def test_streaming_batch_full_size(ctx, oracle):
    """3 x 1080p frames through t3d_backproject_batch (streaming kernel, tickets across frames)."""
    import torch
    H, W = 1920, 1080
    fx, fy, cx, cy = 1719.0, 1719.0, 540.0, 960.0
    depths, bgrs, poses = [], [], []
    for i in range(3):
        d, c, T = ctx.synth_frame(1, i, H, W, fx, fy, cx, cy)
        if i == 1:
            d[: H // 2] = 0.0
        depths.append(d); bgrs.append(c); poses.append((T[:, :3].copy(), T[:, 3:4].copy()))
    frames = ctx.make_backproject_frames(depths, bgrs, poses)
    xyz, rgb, offs = ctx.backproject_batch(frames, 3, H, W, fx=fx, fy=fy, cx=cx, cy=cy, subsample=1)
    offs = offs.cpu().numpy()
    for i in range(3):
        o_xyz, o_rgb = oracle.backproject(depths[i].cpu().numpy(), bgrs[i].cpu().numpy(), fx, fy, cx, cy, pose=poses[i])
        a, b = int(offs[i]), int(offs[i + 1])
        assert b - a == len(o_xyz)
        assert np.array_equal(rgb[a:b].cpu().numpy(), o_rgb)
        got = xyz[a:b].cpu().numpy()
        assert close(got, o_xyz) and (got.view(np.uint32) == o_xyz.view(np.uint32)).mean() > 0.9999


def test_streaming_kernel_random_shapes_and_repeatability(ctx):
    """Randomised geometry / validity patterns: the TMA-staged kernel must equal the generic kernel bit
    for bit (counts, order, points, colours), also for batches whose frames differ in validity, and a
    repeated launch must reproduce itself exactly (the rings, look-back and staging are race-free)."""
    import torch
    rng = np.random.default_rng(2024)
    shapes = [(3, 2048), (2, 4096), (1, 2049), (7, 293), (64, 32), (211, 1024), (97, 2051), (5, 4), (600, 700)]
    for (H, W) in shapes:
        P = H * W
        depth = rng.uniform(0.05, 6.0, size=(H, W)).astype(np.float32)
        mode = rng.integers(0, 4)
        if mode == 0:
            depth[rng.uniform(size=(H, W)) < 0.9] = 0.0                 # sparse
        elif mode == 1:
            depth[:, : W // 2] = np.nan                                 # half of every row invalid
        elif mode == 2:
            depth.reshape(-1)[: (P // 2048) * 2048 // 2] = np.inf       # whole tiles invalid
        bgr = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        d, c = torch.from_numpy(depth).cuda(), torch.from_numpy(bgr).cuda()
        R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        pose = (R, rng.normal(size=(3, 1)))
        kw = dict(fx=300.0, fy=310.0, cx=W / 2.0, cy=H / 2.0, pose=pose, min_depth=0.1, max_depth=5.0)
        a = ctx.backproject(d, c, **kw)
        b = ctx.backproject(_misaligned(d), _misaligned(c), **kw)
        n = int(a[2].item())
        assert n == int(b[2].item()), (H, W)
        assert torch.equal(a[0][:n].view(torch.int32), b[0][:n].view(torch.int32)), (H, W)
        assert torch.equal(a[1][:n], b[1][:n]), (H, W)
    # batch of 9 frames with different validity, 5 repetitions
    H, W = 230, 1000
    depths, bgrs, poses = [], [], []
    for i in range(9):
        dd = rng.uniform(0.2, 4.0, size=(H, W)).astype(np.float32)
        dd[rng.uniform(size=(H, W)) < 0.1 * i] = 0.0
        depths.append(torch.from_numpy(dd).cuda())
        bgrs.append(torch.from_numpy(rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)).cuda())
        poses.append((np.linalg.qr(rng.normal(size=(3, 3)))[0], rng.normal(size=(3, 1))))
    fa = ctx.make_backproject_frames(depths, bgrs, poses)
    kw = dict(fx=500.0, fy=500.0, cx=W / 2.0, cy=H / 2.0, subsample=1, min_depth=0.1, max_depth=5.0)
    mis_d = [_misaligned(x) for x in depths]
    mis_c = [_misaligned(x) for x in bgrs]
    fb = ctx.make_backproject_frames(mis_d, mis_c, poses)
    ref = ctx.backproject_batch(fb, 9, H, W, **kw)
    ro = ref[2].cpu().numpy()
    for rep in range(5):
        out = ctx.backproject_batch(fa, 9, H, W, **kw)
        assert np.array_equal(out[2].cpu().numpy(), ro)
        n = int(ro[-1])
        assert torch.equal(out[0][:n].view(torch.int32), ref[0][:n].view(torch.int32))
        assert torch.equal(out[1][:n], ref[1][:n])
