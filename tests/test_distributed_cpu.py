"""N>1 host logic under gloo on CPU (world_size 2): block ownership, the variable-size
all_to_all routing + weighted merge, and the ICP normal-equation all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from textureless_3d_reconstruction_b200 import distributed as D


class FakeVolume:
    """CPU stand-in with the two methods BlockRouter needs (same merge rule as
    t3d_tsdf_merge_blocks)."""

    def __init__(self):
        self.blocks = {}

    def add(self, key, tsdf, weight, rgb):
        self.blocks[tuple(int(k) for k in key)] = [tsdf.copy(), weight.copy(), rgb.copy()]

    def export_blocks_range(self, axis, lo, hi):
        sel = [k for k in sorted(self.blocks) if lo <= k[axis] < hi]
        n = len(sel)
        keys = torch.tensor(sel, dtype=torch.int32).reshape(n, 3)
        t = torch.from_numpy(np.stack([self.blocks[k][0] for k in sel]) if n else np.zeros((0, 512), np.float32))
        w = torch.from_numpy(np.stack([self.blocks[k][1] for k in sel]) if n else np.zeros((0, 512), np.float32))
        c = torch.from_numpy(np.stack([self.blocks[k][2] for k in sel]) if n else np.zeros((0, 512, 3), np.float32))
        return keys, t, w, c

    def merge_blocks(self, keys, tsdf, weight, rgb):
        for k, t, w, c in zip(keys.numpy(), tsdf.numpy(), weight.numpy(), rgb.numpy()):
            k = tuple(int(x) for x in k)
            if k not in self.blocks:
                self.blocks[k] = [np.zeros(512, np.float32), np.zeros(512, np.float32), np.zeros((512, 3), np.float32)]
            ta, wa, ca = self.blocks[k]
            ws = wa + w
            with np.errstate(invalid="ignore", divide="ignore"):
                tn = np.where(ws > 0, (wa * ta + w * t) / ws, 0).astype(np.float32)
                cn = np.where(ws[:, None] > 0, (wa[:, None] * ca + w[:, None] * c) / ws[:, None], 0).astype(np.float32)
            self.blocks[k] = [tn, ws.astype(np.float32), cn]


def make_rank_blocks(rank, slab_blocks):
    rng = np.random.default_rng(100 + rank)
    vol = FakeVolume()
    # own slab, the neighbour's slab (look-ahead) and one shared boundary block
    zs = list(range(rank * slab_blocks, rank * slab_blocks + 6)) + list(range((1 - rank) * slab_blocks, (1 - rank) * slab_blocks + 3))
    for z in zs:
        for x in (-1, 0):
            w = rng.integers(0, 5, 512).astype(np.float32)
            vol.add((x, 2, z), rng.uniform(-1, 1, 512).astype(np.float32), w,
                    rng.uniform(0, 255, (512, 3)).astype(np.float32))
    return vol


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, slab_blocks, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        vol = make_rank_blocks(rank, slab_blocks)
        router = D.BlockRouter(vol, rank, world, slab_frames=slab_blocks, frame_advance=1.0, block_size=1.0)
        assert router.slab_blocks == slab_blocks
        got = router.route()
        lo, hi = D.block_owner_range(rank, world, slab_blocks)
        owned = {k: [a.copy() for a in v] for k, v in vol.blocks.items() if lo <= k[2] < hi}
        # ICP all-reduce: per-rank partial sums -> identical totals on every rank
        part = np.arange(27, dtype=np.float64) * (rank + 1)
        a27, sd2, cnt = D.allreduce_normal_equations(part, 0.5 * (rank + 1), 10.0 * (rank + 1))
        q.put((rank, got, router.last_sent, owned, a27, sd2, cnt))
    finally:
        dist.destroy_process_group()


def test_block_routing_world2():
    world, slab = 2, 10
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, slab, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r = q.get(timeout=120)
        res[r[0]] = r
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # expectation: merge everything in one process
    vols = [make_rank_blocks(r, slab) for r in range(world)]
    for r in range(world):
        lo, hi = D.block_owner_range(r, world, slab)
        ref = FakeVolume()
        ref.blocks = {k: [a.copy() for a in v] for k, v in vols[r].blocks.items()}
        other = vols[1 - r]
        keys, t, w, c = other.export_blocks_range(2, lo, hi)
        ref.merge_blocks(keys, t, w, c)
        exp = {k: v for k, v in ref.blocks.items() if lo <= k[2] < hi}
        got = res[r][3]
        assert set(got) == set(exp)
        for k in exp:
            assert np.array_equal(got[k][1], exp[k][1])                    # weights exact
            assert np.allclose(got[k][0], exp[k][0], atol=1e-6)
            assert np.allclose(got[k][2], exp[k][2], atol=1e-3)
        assert res[r][1] == 6 and res[r][2] == 6                           # 3 z x 2 x blocks each way
        assert np.allclose(res[r][4], np.arange(27) * 3.0) and res[r][5] == 1.5 and res[r][6] == 30.0


def test_owner_ranges_partition_all_keys():
    for world in (1, 2, 4, 8):
        slab = 937
        z = np.arange(-3000, 9000)
        own = D.owner_of(z, world, slab)
        for r in range(world):
            lo, hi = D.block_owner_range(r, world, slab)
            assert np.array_equal((z >= lo) & (z < hi), own == r)
        zt = torch.from_numpy(z)
        assert np.array_equal(D.owner_of(zt, world, slab).numpy(), own)


def test_solve_icp_update_matches_oracle(oracle):
    """allreduce + host solve == the oracle's first ICP iteration."""
    rng = np.random.default_rng(4)
    g = np.stack(np.meshgrid(np.linspace(-1, 1, 40), np.linspace(-1, 1, 40)), -1).reshape(-1, 2)
    z = 0.2 * np.sin(2 * g[:, 0]) + 0.1 * g[:, 1] ** 2
    tgt = np.column_stack([g, z]).astype(np.float32)
    n = np.column_stack([-0.4 * np.cos(2 * g[:, 0]), -0.2 * g[:, 1], np.ones(len(g))])
    nrm = (n / np.linalg.norm(n, axis=1, keepdims=True)).astype(np.float32)
    src = (tgt + np.array([0.004, -0.003, 0.006], np.float32)).astype(np.float32)
    full = oracle.icp_point_to_plane(src, tgt, nrm, 0.05, max_iter=0)["acc_first"]
    halves = [oracle.icp_point_to_plane(src[i::2], tgt, nrm, 0.05, max_iter=0)["acc_first"] for i in range(2)]
    assert np.allclose(halves[0] + halves[1], full, rtol=1e-9, atol=1e-9)
    one = oracle.icp_point_to_plane(src, tgt, nrm, 0.05, max_iter=1, rel_fitness=0.0, rel_rmse=0.0)
    assert np.allclose(D.solve_icp_update(full[:27]), one["T"], atol=1e-10)
    # ill-posed system -> identity update (R8 det guard)
    assert np.array_equal(D.solve_icp_update(np.zeros(27)), np.eye(4))


# ------------------------------------------------------------------ sharded K2 + gather
VREC = np.dtype([("ix", "<i4"), ("iy", "<i4"), ("iz", "<i4"), ("cnt", "<u4"), ("sx", "<f8"), ("sy", "<f8"),
                 ("sz", "<f8"), ("r", "<u4"), ("g", "<u4"), ("b", "<u4"), ("pad", "<u4")])
assert VREC.itemsize == 56


class NumpyCloudOps:
    """CPU stand-in for runtime.Context in sharded_voxel_downsample: the same wire record, R2's
    index rule (floor((p - minb) / v) in f64, true division) and sum-then-divide means."""

    @staticmethod
    def bounds(xyz):
        a = xyz.numpy().astype(np.float64)
        return a.min(0), a.max(0)

    @staticmethod
    def voxel_partials(xyz, rgb, v, minb, gmax, world):
        p = xyz.numpy().astype(np.float64)
        idx = np.floor((p - minb) / v).astype(np.int64)
        u, inv = np.unique(idx, axis=0, return_inverse=True)
        inv = inv.reshape(-1)
        rec = np.zeros(len(u), VREC)
        rec["ix"], rec["iy"], rec["iz"] = u[:, 0], u[:, 1], u[:, 2]
        rec["cnt"] = np.bincount(inv, minlength=len(u))
        for c, f in enumerate(("sx", "sy", "sz")):
            rec[f] = np.bincount(inv, weights=p[:, c], minlength=len(u))
        if rgb is not None:
            col = rgb.numpy().astype(np.float64)
            for c, f in enumerate(("r", "g", "b")):
                rec[f] = np.bincount(inv, weights=col[:, c], minlength=len(u)).astype(np.uint32)
        owner = (u[:, 0] * 73856093 ^ u[:, 1] * 19349663 ^ u[:, 2] * 83492791) % world
        order = np.argsort(owner, kind="stable")
        counts = np.bincount(owner, minlength=world).astype(np.int32)
        raw = np.frombuffer(rec[order].tobytes(), np.uint8).reshape(-1, 56).copy()
        return torch.from_numpy(raw), torch.from_numpy(counts)

    @staticmethod
    def voxel_merge_partials(records, has_rgb, v, minb, gmax, sorted_output=True):
        rec = np.frombuffer(records.numpy().tobytes(), VREC)
        idx = np.stack([rec["ix"], rec["iy"], rec["iz"]], 1).astype(np.int64)
        u, inv = np.unique(idx, axis=0, return_inverse=True)       # np.unique sorts lexicographically
        inv = inv.reshape(-1)
        cnt = np.bincount(inv, weights=rec["cnt"], minlength=len(u))
        pts = np.stack([np.bincount(inv, weights=rec[f], minlength=len(u)) for f in ("sx", "sy", "sz")], 1) / cnt[:, None]
        out = dict(points=torch.from_numpy(pts), idx=torch.from_numpy(u.astype(np.int32)),
                   count=torch.from_numpy(cnt.astype(np.int32)), colors=None, rgb_sum=None, m=len(u))
        if has_rgb:
            s = np.stack([np.bincount(inv, weights=rec[f], minlength=len(u)) for f in ("r", "g", "b")], 1)
            out["rgb_sum"] = torch.from_numpy(s.astype(np.int32))
            out["colors"] = torch.from_numpy((s / 255.0 / cnt[:, None] * 255.0).astype(np.uint8))
        return out


def _cloud(rank_count=2, n=4000):
    rng = np.random.default_rng(77)
    p = rng.uniform(-1.0, 1.0, (n, 3)).astype(np.float32)
    p[::7] = p[3]                                       # duplicates that land on both ranks
    c = rng.integers(0, 256, (n, 3), dtype=np.uint8)
    cut = [0, n // 3, n] if rank_count == 2 else [0, n]
    return p, c, cut


def _k2_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p, c, cut = _cloud(world)
        xs, cs = torch.from_numpy(p[cut[rank]:cut[rank + 1]]), torch.from_numpy(c[cut[rank]:cut[rank + 1]])
        out = D.sharded_voxel_downsample(NumpyCloudOps, xs, cs, 0.05)
        rows = torch.cat([out["idx"].double(), out["points"], out["count"].double().unsqueeze(1),
                          out["rgb_sum"].double()], 1)
        allrows = D.gather_rows(rows, dst=0)
        # a rank without points takes part in every collective
        empty = D.sharded_voxel_downsample(NumpyCloudOps, xs[:0] if rank == 1 else xs, cs[:0] if rank == 1 else cs, 0.05)
        q.put((rank, None if allrows is None else allrows.numpy(), out["m"], empty["m"], out["records_sent"]))
    finally:
        dist.destroy_process_group()


def test_sharded_voxel_downsample_world2(oracle):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_k2_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = {}
    for _ in range(world):
        r = q.get(timeout=120)
        res[r[0]] = r
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p, c, cut = _cloud(world)
    ref = oracle.voxel_downsample(p.astype(np.float64), c, 0.05)
    rows = res[0][1]
    assert res[1][1] is None and rows.shape[0] == len(ref["points"]) == res[0][2] + res[1][2]
    order = np.lexsort(rows[:, [2, 1, 0]].T)
    rows = rows[order]
    ro = np.lexsort(ref["idx"][:, ::-1].T)
    assert np.array_equal(rows[:, :3].astype(np.int32), ref["idx"][ro])             # voxel index set bit-exact
    assert np.array_equal(rows[:, 6].astype(np.uint32), ref["count"][ro])
    assert np.allclose(rows[:, 3:6], ref["points"][ro], rtol=1e-12, atol=1e-12)
    # rank 0 alone (rank 1 contributed nothing) == oracle on rank 0's shard
    ref0 = oracle.voxel_downsample(p[cut[0]:cut[1]].astype(np.float64), c[cut[0]:cut[1]], 0.05)
    assert res[0][3] + res[1][3] == len(ref0["points"])
    assert res[0][4] > 0


class NumpySorOps:
    """CPU stand-in for the two sharded-K3 calls: brute-force kNN for this part of the (index-ordered)
    cloud; R3's mu / sigma / threshold from the complete vector."""

    @staticmethod
    def sor_mean_distances_part(xyz, nb, part, parts, out_mean):
        p = xyz.numpy()
        n = len(p)
        lo, hi = n * part // parts, n * (part + 1) // parts
        d = np.sqrt(((p[lo:hi, None, :] - p[None, :, :]) ** 2).sum(-1))
        d.sort(axis=1)
        out_mean[lo:hi] = torch.from_numpy(d[:, :nb].mean(1))
        return out_mean

    @staticmethod
    def sor_from_mean_distances(mean, std_ratio):
        m = mean.numpy()
        mu = m.sum() / len(m)
        sigma = np.sqrt(((m - mu) ** 2).sum() / (len(m) - 1))
        thr = mu + std_ratio * sigma
        keep = (m > 0) & (m < thr)
        return torch.from_numpy(keep.astype(np.uint8)), (mu, sigma, thr), int(keep.sum())


def _sor_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(9)
        pts = rng.uniform(0, 1, (600, 3))
        pts[:5] += 3.0                                     # outliers
        keep, mean, stats, kept = D.sharded_statistical_outlier(NumpySorOps, torch.from_numpy(pts), 20, 2.0)
        q.put((rank, keep.numpy(), mean.numpy(), stats, kept))
    finally:
        dist.destroy_process_group()


def test_sharded_statistical_outlier_world2(oracle):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sor_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = {}
    for _ in range(world):
        r = q.get(timeout=120)
        res[r[0]] = r
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    rng = np.random.default_rng(9)
    pts = rng.uniform(0, 1, (600, 3))
    pts[:5] += 3.0
    ref = oracle.statistical_outlier(pts, 20, 2.0)
    for r in range(world):
        assert np.isfinite(res[r][2]).all()                                  # every query was owned by some rank
        assert np.array_equal(res[r][1], res[0][1]) and res[r][4] == res[0][4]
    ref_keep = ref["keep"] if isinstance(ref, dict) else ref[0]
    assert np.array_equal(res[0][1].astype(bool), np.asarray(ref_keep).astype(bool))
    assert not res[0][1][:5].any()


def test_overlap_order_puts_the_tail_first_and_the_head_last():
    """Frame order of routing-underneath-fusion: batches descending, frames ascending inside a batch;
    the first batch holds every frame that can reach past the slab, the last batch the slab's head."""
    for n, b in [(300, 64), (300, 32), (64, 64), (100, 24), (7, 3)]:
        order = D.P2PBlockRouter.overlap_order(n, b)
        assert sorted(order) == list(range(n))
        assert order[:min(b, n)] == list(range(max(0, n - b), n))            # tail batch first, in frame order
        nb = -(-n // b)
        last = order[(nb - 1) * b:] if nb > 1 else order
        assert last[0] == 0 and last == list(range(len(last)))                # head batch last
        for k in range(0, n, b):                                              # ascending inside every batch
            chunk = order[k:k + b]
            assert chunk == sorted(chunk)


def test_tail_head_order_puts_the_tail_first_and_the_head_second():
    """Frame order of the copy-engine router: the tail batch (the frames whose blocks travel) first, the head
    batch (the frames that meet incoming blocks) second, the middle batches after them; every frame exactly once,
    ascending inside a batch, and batch boundaries of the reordered sequence never split tail or head."""
    for n, b in [(300, 64), (300, 32), (625, 64), (192, 64), (130, 64), (64, 64), (7, 3)]:
        order = D.CopyEngineBlockRouter.tail_head_order(n, b)
        assert sorted(order) == list(range(n))
        tail = order[:min(b, n)]
        assert tail == list(range(n - len(tail), n)) or n <= b                # the last frames of the slab, in order
        if n >= 2 * b:
            assert order[b:2 * b] == list(range(b))                           # the slab's first frames second
            assert order[2 * b:] == list(range(b, n - b))                     # then the middle, in frame order
        for k in range(0, n, b):
            chunk = order[k:k + b]
            assert chunk == sorted(chunk)
    # the frames that reach past the slab (depth_max / advance + 2 = 22 at cfg 2) are all in the first batch
    order = D.CopyEngineBlockRouter.tail_head_order(300, 64)
    assert set(range(300 - 22, 300)) <= set(order[:64])
