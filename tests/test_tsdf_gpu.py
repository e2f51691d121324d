"""K4/K5/K6 parity vs the oracle (parity unpinned by the reference: no TSDF code there)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from textureless_3d_reconstruction_b200 import synthetic as S


def frames(n, H, W, noise=0.002, scene=0, step=1):
    it = S.scaled_intrinsics(H, W)
    out = []
    for i in range(n):
        d, c, T = S.synth_frame(scene, i * step, H, W, it["fx"], it["fy"], it["cx"], it["cy"], noise_sigma=noise)
        out.append((d, c, T))
    return out, (it["fx"], it["fy"], it["cx"], it["cy"])


def key_rows(a):
    a = np.ascontiguousarray(a, np.int32)
    return set(map(tuple, a.tolist()))


def by_key(keys, *arrs):
    order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
    return (keys[order],) + tuple(a[order] for a in arrs)


@pytest.mark.parametrize("pixel_round", [0, 1])
def test_touch_and_integrate_bit_exact(ctx, oracle, pixel_round):
    import torch
    from textureless_3d_reconstruction_b200.runtime import TSDFVolume
    H, W = 240, 136
    fr, K = frames(5, H, W)
    vol = TSDFVolume(0.01, 0.04, block_capacity=60000, pixel_round=pixel_round, ctx=ctx)
    ov = oracle.TSDFVolume(0.01, 0.04, pixel_round)
    for d, c, T in fr:
        dd, cc = torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda()
        gk = vol.touch(dd, K, T, 1.0, 5.0).cpu().numpy()
        ok = ov.touch(d, K, T, 1.0, 5.0)
        assert key_rows(gk) == key_rows(ok) and len(gk) == len(ok)      # R4: block key set bit-exact
        vol.integrate(dd, cc, K, T, 1.0, 5.0)
        ov.integrate(d, c, K, T, 1.0, 5.0)
    assert vol.num_blocks == ov.num_blocks
    assert vol.counters() == ov.counters()
    gk, gt, gw, gc = [x.cpu().numpy() for x in vol.export_blocks()]
    okk, ot, ow, oc = ov.export()
    gk, gt, gw, gc = by_key(gk, gt, gw, gc)
    okk, ot, ow, oc = by_key(okk, ot, ow, oc)
    assert np.array_equal(gk, okk)
    assert np.array_equal(gw, ow)                                        # occupancy + integer weights
    assert np.array_equal(gt.view(np.uint32), ot.view(np.uint32))        # tsdf bit-exact (<=1e-4 required)
    assert np.array_equal(gc.view(np.uint32), oc.view(np.uint32))
    # R6 surface points: same multiset
    gp, gn, gcol = [x.cpu().numpy() for x in vol.extract_points(3.0)]
    op, on, ocol = ov.extract_points(3.0)
    assert len(gp) == len(op) and len(gp) > 1000
    def rows(p, n, c):
        a = np.concatenate([p.view(np.uint32), n.view(np.uint32), c.astype(np.uint32)], axis=1)
        return a[np.lexsort(a.T[::-1])]
    assert np.array_equal(rows(gp, gn, gcol), rows(op, on, ocol))


def test_batched_equals_sequential(ctx):
    """Temporal blocking must not change a single bit: 1 x 7 frames == 7 x 1 frame."""
    import torch
    from textureless_3d_reconstruction_b200.runtime import TSDFVolume
    H, W = 240, 136
    fr, K = frames(7, H, W)
    ds = [torch.from_numpy(f[0]).cuda() for f in fr]
    cs = [torch.from_numpy(f[1]).cuda() for f in fr]
    Ts = [f[2] for f in fr]
    a = TSDFVolume(0.01, 0.04, block_capacity=60000, ctx=ctx)
    b = TSDFVolume(0.01, 0.04, block_capacity=60000, ctx=ctx)
    a.integrate_batch(ds, cs, K, Ts, 1.0, 5.0)
    for d, c, T in zip(ds, cs, Ts):
        b.integrate(d, c, K, T, 1.0, 5.0)
    assert a.counters() == b.counters() and a.num_blocks == b.num_blocks
    ea = by_key(*[x.cpu().numpy() for x in a.export_blocks()])
    eb = by_key(*[x.cpu().numpy() for x in b.export_blocks()])
    for x, y in zip(ea, eb):
        assert np.array_equal(x.view(np.uint32) if x.dtype == np.float32 else x,
                              y.view(np.uint32) if y.dtype == np.float32 else y)
    # pipelined sequence (K4 of batch b+1 under K5 of batch b): 4 batches of 2,2,2,1 frames
    c = TSDFVolume(0.01, 0.04, block_capacity=60000, ctx=ctx)
    views = c.make_frame_views(ds, cs, [K] * 7, Ts)
    for rep in range(3):                       # repeat: exercises ping-pong parity carry-over + reset
        c.reset()
        c.integrate_sequence(views, 7, H, W, batch=2, depth_scale=1.0, depth_max=5.0)
        assert c.counters() == b.counters() and c.num_blocks == b.num_blocks
        ec = by_key(*[x.cpu().numpy() for x in c.export_blocks()])
        for x, y in zip(ec, eb):
            assert np.array_equal(x.view(np.uint32) if x.dtype == np.float32 else x,
                                  y.view(np.uint32) if y.dtype == np.float32 else y)
    # reset really forgets
    a.reset()
    assert a.num_blocks == 0 and a.counters()["voxel_updates"] == 0


def test_u16_depth_no_color_and_merge(ctx, oracle):
    import torch
    from textureless_3d_reconstruction_b200.runtime import TSDFVolume
    H, W = 160, 92
    fr, K = frames(4, H, W, noise=0.0)
    vol = TSDFVolume(0.02, 0.08, block_capacity=20000, ctx=ctx)
    ov = oracle.TSDFVolume(0.02, 0.08)
    for d, c, T in fr:
        mm = np.clip(d * 1000.0, 0, 65535).astype(np.uint16)            # dp:919-921 depth PNG format
        vol.integrate(torch.from_numpy(mm).cuda(), None, K, T, 1000.0, 5.0)
        ov.integrate(mm, None, K, T, 1000.0, 5.0)
    g = by_key(*[x.cpu().numpy() for x in vol.export_blocks()])
    o = by_key(*ov.export())
    assert np.array_equal(g[0], o[0]) and np.array_equal(g[2], o[2])
    assert np.array_equal(g[1].view(np.uint32), o[1].view(np.uint32))
    # merge of two half-volumes == weights add, tsdf = weighted mean
    A = TSDFVolume(0.02, 0.08, block_capacity=20000, ctx=ctx)
    B = TSDFVolume(0.02, 0.08, block_capacity=20000, ctx=ctx)
    for i, (d, c, T) in enumerate(fr):
        (A if i % 2 == 0 else B).integrate(torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda(), K, T, 1.0, 5.0)
    full = TSDFVolume(0.02, 0.08, block_capacity=20000, ctx=ctx)
    for d, c, T in fr:
        full.integrate(torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda(), K, T, 1.0, 5.0)
    kb, tb, wb, cb = B.export_blocks()
    A.merge_blocks(kb.contiguous(), tb.contiguous(), wb.contiguous(), cb.contiguous())
    m = by_key(*[x.cpu().numpy() for x in A.export_blocks()])
    f = by_key(*[x.cpu().numpy() for x in full.export_blocks()])
    assert np.array_equal(m[0], f[0]) and np.array_equal(m[2], f[2])     # keys + weights exact
    assert np.abs(m[1] - f[1]).max() <= 1e-4                             # north_star tsdf tolerance
    assert np.abs(m[3] - f[3]).max() <= 1e-2


def test_capacity_overflow_is_reported(ctx):
    import torch
    from textureless_3d_reconstruction_b200._lib import T3DError
    from textureless_3d_reconstruction_b200.runtime import TSDFVolume
    fr, K = frames(1, 160, 92, noise=0.0)
    vol = TSDFVolume(0.01, 0.04, block_capacity=16, ctx=ctx)
    vol.integrate(torch.from_numpy(fr[0][0]).cuda(), None, K, fr[0][2], 1.0, 5.0)
    with pytest.raises(T3DError):
        _ = vol.num_blocks


def test_full_size_properties(ctx):
    """cfg-2 sized frames (1080x1920, device-generated): size-independent invariants."""
    import torch
    from textureless_3d_reconstruction_b200.runtime import TSDFVolume
    H, W = 1920, 1080
    K = (1719.0, 1719.0, 540.0, 960.0)
    n = 6
    fr = [ctx.synth_frame(0, i, H, W, *K, noise_sigma=0.002) for i in range(n)]
    ds, cs, Ts = [f[0] for f in fr], [f[1] for f in fr], [f[2] for f in fr]
    a = TSDFVolume(0.01, 0.04, block_capacity=120000, ctx=ctx)
    a.integrate_batch(ds, cs, K, Ts, 1.0, 5.0)
    cnt = a.counters()
    keys, tsdf, w, rgb = a.export_blocks()
    assert cnt["frames"] == n and cnt["voxel_updates"] > 1_000_000
    assert int(w.sum().item()) == cnt["voxel_updates"]          # checksum: sum of weights == updates applied
    assert float(w.max().item()) <= n and float(tsdf.abs().max().item()) <= 1.0
    assert torch.isfinite(tsdf).all() and torch.isfinite(rgb).all()
    assert len(key_rows(keys.cpu().numpy())) == a.num_blocks    # no duplicate blocks in the hash
    touched = set()
    for d, T in zip(ds, Ts):
        touched |= key_rows(a.touch(d, K, T, 1.0, 5.0).cpu().numpy())
    assert touched == key_rows(keys.cpu().numpy())
    # grey albedo 128+-2 -> fused colours stay in [126,130] wherever weight > 0
    m = w > 0
    assert float(rgb[m].min().item()) >= 126.0 and float(rgb[m].max().item()) <= 130.0
    # idempotent order: batch of 6 == 6 singles (bit-exact) at full size too
    b = TSDFVolume(0.01, 0.04, block_capacity=120000, ctx=ctx)
    for d, c, T in zip(ds, cs, Ts):
        b.integrate(d, c, K, T, 1.0, 5.0)
    assert b.counters() == cnt
    kb, tb, wb, _ = b.export_blocks()
    ka = by_key(keys.cpu().numpy(), tsdf.cpu().numpy(), w.cpu().numpy())
    kbb = by_key(kb.cpu().numpy(), tb.cpu().numpy(), wb.cpu().numpy())
    assert np.array_equal(ka[0], kbb[0]) and np.array_equal(ka[2], kbb[2])
    assert np.array_equal(ka[1].view(np.uint32), kbb[1].view(np.uint32))
    # surface points lie on the tunnel wall: radius within relief bounds +- a voxel
    p, nrm, _ = a.extract_points(3.0)
    r = torch.sqrt(p[:, 0] ** 2 + p[:, 1] ** 2)
    assert p.shape[0] > 50_000
    assert float(r.min().item()) > 1.37 - 0.03 and float(r.max().item()) < 1.63 + 0.03
    assert torch.allclose(nrm.norm(dim=1), torch.ones_like(r), atol=1e-4)


def test_extract_view_matches_oracle(ctx, oracle):
    """View-restricted K6 (the tracker's ICP target): same visible-block count and the same
    multiset of surface points/normals as the oracle's rule (o_block_in_view)."""
    import torch
    from textureless_3d_reconstruction_b200.runtime import TSDFVolume
    H, W = 240, 136
    fr, K = frames(6, H, W, step=3)
    vol = TSDFVolume(0.01, 0.04, block_capacity=80000, ctx=ctx)
    ov = oracle.TSDFVolume(0.01, 0.04)
    for d, c, T in fr:
        vol.integrate(torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda(), K, T, 1.0, 5.0)
        ov.integrate(d, c, K, T, 1.0, 5.0)
    for T, dmax in ((fr[-1][2], 5.0), (fr[0][2], 3.5), (fr[3][2], 5.0)):
        gp, gn, _, nsel = vol.extract_points_view(K, T, H, W, dmax, 1.0)
        op, on, _ = ov.extract_points(1.0, view=(K, T, H, W, dmax))
        assert nsel == ov.last_view_blocks and 0 < nsel < vol.num_blocks
        assert len(gp) == len(op) and len(gp) > 1000
        a = np.concatenate([gp.cpu().numpy().view(np.uint32), gn.cpu().numpy().view(np.uint32)], axis=1)
        b = np.concatenate([op.view(np.uint32), on.view(np.uint32)], axis=1)
        assert np.array_equal(a[np.lexsort(a.T[::-1])], b[np.lexsort(b.T[::-1])])
    # a camera looking away sees nothing
    Taway = np.array(fr[0][2], np.float64).copy()
    Taway[:3, :3] = np.diag([1.0, -1.0, -1.0]) @ Taway[:3, :3]
    Taway[:3, 3] = np.diag([1.0, -1.0, -1.0]) @ Taway[:3, 3] + np.array([0, 0, -50.0])
    gp, _, _, nsel = vol.extract_points_view(K, Taway, H, W, 5.0, 1.0)
    assert nsel == 0 and len(gp) == 0


def test_frame_to_model_tracker_vs_oracle(ctx, oracle):
    """cfg-3 loop (ICP + TSDF): GPU tracker vs the CPU restatement (oracle/ref_tracker.py) on the
    same frames.  north_star tolerance on ICP poses: <= 1e-4 rad / 1e-4 m."""
    import torch
    from oracle import ref_tracker
    from textureless_3d_reconstruction_b200.tracking import FrameToModelTracker
    H, W = 240, 136
    fr, K = frames(7, H, W)
    kw = dict(voxel_size=0.01, sdf_trunc=0.04, depth_max=5.0, icp_subsample=2, icp_max_corr=0.05)
    gt = FrameToModelTracker(K, H, W, block_capacity=80000, ctx=ctx, **kw)
    ot = ref_tracker.FrameToModelTracker(K, H, W, **kw)
    for i, (d, c, T) in enumerate(fr):
        Tg = gt.add_frame(torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda(), known_pose=T if i == 0 else None)
        To = ot.add_frame(d, c, known_pose=T if i == 0 else None)
        E = Tg @ np.linalg.inv(To)
        ang = np.arccos(np.clip((np.trace(E[:3, :3]) - 1.0) / 2.0, -1.0, 1.0))
        assert np.linalg.norm(E[:3, 3]) <= 1e-4 and ang <= 1e-4, (i, E)
        if i > 0:
            assert gt.icp_log[i].iterations == ot.icp_log[i]["iterations"]
            assert abs(gt.icp_log[i].fitness - ot.icp_log[i]["fitness"]) < 1e-6
    # tracked poses stay near the ground truth (textureless tunnel, coarse 240x136 frames)
    Tgt = np.eye(4)
    Tgt[:3, :4] = fr[-1][2]
    assert np.linalg.norm((gt.poses[-1] @ np.linalg.inv(Tgt))[:3, 3]) < 0.05
    assert gt.volume.num_blocks == ot.volume.num_blocks


def test_route_records_round_trip(ctx):
    """Wire-record routing (t3d_tsdf_route_counts / route_export / merge_records) on one GPU:
    records grouped by owner carry exactly the non-owned blocks, and merging them equals the
    export_blocks_range + merge_blocks path bit for bit (same weighted-mean rule)."""
    import torch
    from textureless_3d_reconstruction_b200.distributed import block_owner_range, owner_of
    from textureless_3d_reconstruction_b200.runtime import TSDFVolume
    H, W = 240, 136
    fr, K = frames(6, H, W, step=2)
    src = TSDFVolume(0.01, 0.04, block_capacity=60000, ctx=ctx)
    for d, c, T in fr:
        src.integrate(torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda(), K, T, 1.0, 5.0)
    keys_all = src.export_blocks()[0].cpu().numpy()
    world, rank, axis = 3, 1, 2
    slab = max(1, (int(keys_all[:, axis].max()) - int(keys_all[:, axis].min())) // 3)
    counts = src.route_counts(axis, slab, world, rank)
    own = owner_of(keys_all[:, axis].astype(np.int64), world, slab)
    expect = np.bincount(own[own != rank], minlength=world)
    assert counts.cpu().numpy().tolist() == expect.tolist() and expect.sum() > 100
    rec = src.route_export(axis, slab, world, rank, counts)
    assert rec.shape == (int(expect.sum()), 2564)
    rk = rec[:, :4].contiguous().view(torch.int32).cpu().numpy()
    assert np.array_equal(rk[:, 3], np.sort(rk[:, 3]))                       # grouped by destination
    assert np.array_equal(rk[:, 3], owner_of(rk[:, axis].astype(np.int64), world, slab))
    lo, hi = block_owner_range(rank, world, slab)
    assert key_rows(rk[:, :3]) == key_rows(keys_all[(keys_all[:, axis] < lo) | (keys_all[:, axis] >= hi)])
    # merging the records into a volume that already holds overlapping data == merge_blocks
    a = TSDFVolume(0.01, 0.04, block_capacity=60000, ctx=ctx)
    b = TSDFVolume(0.01, 0.04, block_capacity=60000, ctx=ctx)
    for vol in (a, b):
        for d, c, T in fr[:2]:
            vol.integrate(torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda(), K, T, 1.0, 5.0)
    a.merge_records(rec.contiguous())
    ok, ot, ow, oc = src.export_blocks_outside(axis, lo, hi)
    b.merge_blocks(ok.contiguous(), ot.contiguous(), ow.contiguous(), oc.contiguous())
    ga = by_key(*[x.cpu().numpy() for x in a.export_blocks()])
    gb = by_key(*[x.cpu().numpy() for x in b.export_blocks()])
    assert a.num_blocks == b.num_blocks
    for x, y in zip(ga, gb):
        assert np.array_equal(x.view(np.uint32) if x.dtype == np.float32 else x,
                              y.view(np.uint32) if y.dtype == np.float32 else y)


def test_checkpoint_resume_is_bit_exact(ctx, tmp_path):
    """save() after k frames + load() + the remaining frames == uninterrupted fusion, bit for bit."""
    import torch
    from textureless_3d_reconstruction_b200.runtime import TSDFVolume
    H, W = 240, 136
    fr, K = frames(6, H, W)
    full = TSDFVolume(0.01, 0.04, block_capacity=60000, ctx=ctx)
    part = TSDFVolume(0.01, 0.04, block_capacity=60000, ctx=ctx)
    for i, (d, c, T) in enumerate(fr):
        full.integrate(torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda(), K, T, 1.0, 5.0)
        if i < 3:
            part.integrate(torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda(), K, T, 1.0, 5.0)
    part.save(tmp_path / "vol.npz")
    res = TSDFVolume.load(tmp_path / "vol.npz", block_capacity=60000, ctx=ctx)
    assert res.num_blocks == part.num_blocks
    for d, c, T in fr[3:]:
        res.integrate(torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda(), K, T, 1.0, 5.0)
    a = by_key(*[x.cpu().numpy() for x in full.export_blocks()])
    b = by_key(*[x.cpu().numpy() for x in res.export_blocks()])
    for x, y in zip(a, b):
        assert np.array_equal(x.view(np.uint32) if x.dtype == np.float32 else x,
                              y.view(np.uint32) if y.dtype == np.float32 else y)


def test_multi_gpu_routers_agree():
    """Needs >= 2 GPUs (skipped on the 1-GPU tier): tests/mgpu_route_check.py under torchrun."""
    import subprocess
    import sys
    import torch
    from pathlib import Path
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(root / "tests" / "mgpu_route_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "p2p==nccl True owned==serial True overlapped==serial True" in r.stdout, \
        r.stdout[-2000:] + r.stderr[-2000:]


def test_two_contexts_in_one_process(oracle):
    """One process, one Context per GPU (needs >= 2 GPUs): per-device one-time initialisation (ICP neighbour
    offsets in constant memory, the >48 KB shared-memory opt-in of the streaming back-projection kernel) must
    happen on EACH device, and every entry point must select its context's device."""
    import torch
    from textureless_3d_reconstruction_b200.runtime import TSDFVolume, get_context
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    H, W = 240, 136
    fr, K = frames(3, H, W)
    ref = None
    for dev in (0, 1, 0):
        ctx = get_context(dev)
        with torch.cuda.device(1 - dev):                 # the CURRENT device is deliberately the other one
            d = torch.from_numpy(fr[0][0]).to(ctx.device)
            c = torch.from_numpy(fr[0][1]).to(ctx.device)
            pose = (fr[0][2][:, :3].copy(), fr[0][2][:, 3:4].copy())
            # K1 streaming kernel (s = 1: needs the dynamic shared-memory attribute on this device)
            xyz, rgb, n = ctx.backproject(d, c, fx=K[0], fy=K[1], cx=K[2], cy=K[3], pose=pose, max_depth=5.0)
            n = int(n.item())
            o_xyz, o_rgb = oracle.backproject(fr[0][0], fr[0][1], *K, pose=pose, max_depth=5.0)
            assert n == len(o_xyz) and np.array_equal(rgb[:n].cpu().numpy(), o_rgb)
            # TSDF + surface + ICP of the frame's own cloud against that surface (needs c_ofs on this device)
            vol = TSDFVolume(0.01, 0.04, block_capacity=40000, ctx=ctx)
            for dd, cc, T in fr:
                vol.integrate(torch.from_numpy(dd).to(ctx.device), torch.from_numpy(cc).to(ctx.device), K, T, 1.0, 5.0)
            p, nrm, _ = vol.extract_points(2.0)
            src = xyz[:n:7].contiguous()
            T0 = np.eye(4)
            T0[0, 3] = 0.004
            res = ctx.icp_point_to_plane(src, p.contiguous(), nrm.contiguous(), 0.05, init=T0, max_iter=30)
            out = (vol.counters(), int(p.shape[0]), res.iterations, np.round(res.transformation, 9).tolist())
        if ref is None:
            ref = out
        assert out == ref, (dev, out, ref)
    assert ref[0]["voxel_updates"] > 10000 and abs(ref[3][0][3]) < 2e-3       # ICP pulled the offset back


def test_full_size_tracking_properties(ctx):
    """cfg-3 at BASELINE frame size (1080x1920, 0.25 m/frame): the tracker stays on the ground-truth
    trajectory, its poses are rigid, and with known poses it reproduces plain integration."""
    import torch
    from textureless_3d_reconstruction_b200.tracking import FrameToModelTracker
    H, W = 1920, 1080
    K = (1719.0, 1719.0, 540.0, 960.0)
    n = 16
    fr = [ctx.synth_frame(0, i, H, W, *K, noise_sigma=0.002) for i in range(n)]
    trk = FrameToModelTracker(K, H, W, block_capacity=120000, icp_subsample=4, icp_max_corr=0.05, ctx=ctx)
    for i, (d, c, T) in enumerate(fr):
        trk.add_frame(d, c, known_pose=T if i == 0 else None)
    for i in range(n):
        P = trk.poses[i]
        assert np.allclose(P[:3, :3] @ P[:3, :3].T, np.eye(3), atol=1e-9) and abs(np.linalg.det(P[:3, :3]) - 1) < 1e-9
        G = np.eye(4)
        G[:3, :4] = fr[i][2]
        E = P @ np.linalg.inv(G)
        assert np.linalg.norm(E[:3, 3]) < 0.02, (i, E[:3, 3])              # < 2 cm after up to 3.75 m of travel
        assert np.arccos(np.clip((np.trace(E[:3, :3]) - 1) / 2, -1, 1)) < 5e-3
    log = [r for r in trk.icp_log if r is not None]
    assert len(log) == n - 1 and all(r.fitness > 0.8 and r.inlier_rmse < 0.02 and r.iterations <= 30 for r in log)
    # known poses through the same class == TSDFVolume.integrate
    from textureless_3d_reconstruction_b200.runtime import TSDFVolume
    a = FrameToModelTracker(K, H, W, block_capacity=60000, ctx=ctx)
    b = TSDFVolume(0.01, 0.04, block_capacity=60000, ctx=ctx)
    for d, c, T in fr[:3]:
        a.add_frame(d, c, known_pose=T)
        b.integrate(d, c, K, T, 1.0, 5.0)
    assert a.volume.counters() == b.counters() and a.volume.num_blocks == b.num_blocks


def test_confidence_mask(ctx, oracle):
    """north_star "confidence masking" in TSDF fusion (t3d_frame_view.conf_mask): a pixel whose mask byte is 0
    carries no measurement.  (i) GPU == oracle with the same mask, bit for bit; (ii) independent of any oracle:
    fusing with a mask == fusing the same frames with the depth of the masked pixels set to 0."""
    import torch
    from textureless_3d_reconstruction_b200.runtime import TSDFVolume
    H, W = 240, 136
    fr, K = frames(4, H, W)
    rng = np.random.default_rng(7)
    masks = []
    for i in range(4):
        m = (rng.random((H, W)) > 0.3).astype(np.uint8) * rng.integers(1, 256, (H, W)).astype(np.uint8)
        m[40:90, 20:70] = 0                                   # a solid masked rectangle as well as salt-and-pepper
        masks.append(m)
    a = TSDFVolume(0.01, 0.04, block_capacity=60000, ctx=ctx)     # masked
    b = TSDFVolume(0.01, 0.04, block_capacity=60000, ctx=ctx)     # depth zeroed instead
    full = TSDFVolume(0.01, 0.04, block_capacity=60000, ctx=ctx)  # no mask
    ov = oracle.TSDFVolume(0.01, 0.04)
    for (d, c, T), m in zip(fr, masks):
        dd, cc, mm = torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda(), torch.from_numpy(m).cuda()
        gk = a.touch(dd, K, T, 1.0, 5.0, conf_mask=mm).cpu().numpy()
        ok = ov.touch(d, K, T, 1.0, 5.0, conf_mask=m)
        assert key_rows(gk) == key_rows(ok) and len(gk) == len(ok)
        a.integrate(dd, cc, K, T, 1.0, 5.0, conf_mask=mm)
        ov.integrate(d, c, K, T, 1.0, 5.0, conf_mask=m)
        dz = d.copy()
        dz[m == 0] = 0.0
        b.integrate(torch.from_numpy(dz).cuda(), cc, K, T, 1.0, 5.0)
        full.integrate(dd, cc, K, T, 1.0, 5.0)
    assert a.counters() == ov.counters() == b.counters()
    assert a.counters()["voxel_updates"] < 0.8 * full.counters()["voxel_updates"]
    ea = by_key(*[x.cpu().numpy() for x in a.export_blocks()])
    eb = by_key(*[x.cpu().numpy() for x in b.export_blocks()])
    eo = by_key(*ov.export())
    for x, y, z in zip(ea, eb, eo):
        bits = (lambda q: q.view(np.uint32)) if x.dtype == np.float32 else (lambda q: q)
        assert np.array_equal(bits(x), bits(y)) and np.array_equal(bits(x), bits(z))
    # a batch with a mask on some frames only
    c2 = TSDFVolume(0.01, 0.04, block_capacity=60000, ctx=ctx)
    ds = [torch.from_numpy(f[0]).cuda() for f in fr]
    cs = [torch.from_numpy(f[1]).cuda() for f in fr]
    views = c2.make_frame_views(ds, cs, [K] * 4, [f[2] for f in fr], [torch.from_numpy(m).cuda() for m in masks])
    c2.integrate_sequence(views, 4, H, W, batch=4, depth_scale=1.0, depth_max=5.0)
    ec = by_key(*[x.cpu().numpy() for x in c2.export_blocks()])
    for x, y in zip(ea, ec):
        bits = (lambda q: q.view(np.uint32)) if x.dtype == np.float32 else (lambda q: q)
        assert np.array_equal(bits(x), bits(y))
    with pytest.raises(ValueError):
        a.integrate(ds[0], cs[0], K, fr[0][2], 1.0, 5.0, conf_mask=torch.zeros((H, W + 1), dtype=torch.uint8, device="cuda"))
