"""Multi-GPU sharding of the fused path: one process per GPU,
torch.distributed (NCCL over NVLink/NVSwitch) only where data really has to move.

* frame stream: rank r fuses frames [r*F, (r+1)*F) — no communication;
* TSDF volume: split spatially by voxel-block ownership.  The camera advances along
  +z, so ownership is the z-slab of the block key: owner(zb) = clamp(zb // slab_blocks).
  After a rank has fused its frames, blocks it touched but does not own (the ~5 m of
  look-ahead past its last camera) are exported once, packed into 10 KiB records sorted
  by owner, routed with one variable-size all_to_all and merged there as weighted means (running averages with unit weights
  are mergeable: w = wa + wb, tsdf = (wa*ta + wb*tb) / w);
* ICP: source points sharded, per-iteration all_reduce of the 29 normal-equation sums.

Everything is written against `torch.distributed` collectives on whatever device the
volume's tensors live on, so the same code runs under gloo on CPU tensors in the
world_size-2 tests (tests/test_distributed_cpu.py) and under NCCL on the GPUs.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def block_owner_range(rank: int, world: int, slab_blocks: int):
    """[lo, hi) of block z-keys owned by `rank`; the first/last slabs are open-ended."""
    lo = -(1 << 30) if rank == 0 else rank * slab_blocks
    hi = (1 << 30) if rank == world - 1 else (rank + 1) * slab_blocks
    return lo, hi


def owner_of(zb, world: int, slab_blocks: int):
    """Vectorised owner of block z-keys (NumPy or torch integer array)."""
    o = zb // slab_blocks
    if hasattr(o, "clamp"):
        return o.clamp(0, world - 1)
    return np.clip(o, 0, world - 1)


class BlockRouter:
    """Routes non-owned TSDF blocks to their owner rank and merges them there.

    `vol` needs: export_blocks_range(axis, lo, hi) -> (keys[n,3] i32, tsdf[n,512] f32,
    weight[n,512] f32, rgb[n,512,3] f32) and merge_blocks(keys, tsdf, weight, rgb).
    """

    AXIS = 2

    def __init__(self, vol, rank: int, world: int, slab_frames: int, frame_advance: float, block_size: float,
                 group=None):
        self.vol, self.rank, self.world, self.group = vol, rank, world, group
        self.slab_blocks = max(1, int(round(slab_frames * frame_advance / block_size)))
        self.last_sent = 0
        self.last_received = 0

    RECORD = 4 + 512 + 512 + 1536   # f32 words per block on the wire: key(3)+pad | tsdf | weight | rgb

    def _export_non_owned(self):
        lo, hi = block_owner_range(self.rank, self.world, self.slab_blocks)
        if hasattr(self.vol, "export_blocks_outside"):
            return self.vol.export_blocks_outside(self.AXIS, lo, hi)
        import torch
        below = self.vol.export_blocks_range(self.AXIS, -(1 << 30), lo)
        above = self.vol.export_blocks_range(self.AXIS, hi, 1 << 30)
        return tuple(torch.cat([a, b]) for a, b in zip(below, above))

    def route(self):
        """One export of everything this rank does not own, one packed record per block
        (10 256 B), sorted by owner; one all_to_all for the counts and one for the
        records; owners merge what they receive, one source rank at a time (a block can
        arrive from several ranks, and merge_blocks needs unique keys per call)."""
        import torch
        import torch.distributed as dist
        world, rank = self.world, self.rank
        if world == 1:
            return 0
        if hasattr(self.vol, "route_export"):
            return self._route_records()
        keys, tsdf, weight, rgb = self._export_non_owned()
        dev = keys.device
        n = keys.shape[0]
        owner = owner_of(keys[:, self.AXIS].to(torch.int64), world, self.slab_blocks)
        order = torch.argsort(owner, stable=True)
        send_counts = torch.bincount(owner, minlength=world)
        rec = torch.empty((n, self.RECORD), dtype=torch.float32, device=dev)
        if n:
            rec[:, :3] = keys.view(torch.float32)
            rec[:, 3] = 0
            rec[:, 4:516] = tsdf
            rec[:, 516:1028] = weight
            rec[:, 1028:] = rgb.reshape(n, 1536)
            rec = rec.index_select(0, order)
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=self.group)
        sc, rc = send_counts.tolist(), recv_counts.tolist()
        self.last_sent, self.last_received = int(sum(sc)), int(sum(rc))
        if self.last_sent == 0 and self.last_received == 0:
            return 0
        recv = torch.empty((sum(rc), self.RECORD), dtype=torch.float32, device=dev)
        dist.all_to_all_single(recv, rec.contiguous(), output_split_sizes=rc, input_split_sizes=sc, group=self.group)
        off = 0
        for cnt in rc:
            if cnt > 0:
                seg = recv[off:off + cnt]
                self.vol.merge_blocks(seg[:, :3].contiguous().view(torch.int32), seg[:, 4:516].contiguous(),
                                      seg[:, 516:1028].contiguous(),
                                      seg[:, 1028:].contiguous().reshape(cnt, 512, 3))
            off += cnt
        return self.last_received


def _route_records_impl(self):
    """Device path: the library counts, packs (grouped by destination) and merges wire
    records itself; torch.distributed only moves them."""
    import torch
    import torch.distributed as dist
    vol, world, rank = self.vol, self.world, self.rank
    send_counts = vol.route_counts(self.AXIS, self.slab_blocks, world, rank)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=self.group)
    both = torch.stack([send_counts, recv_counts]).tolist()      # the route's only host sync
    sc, rc = both[0], both[1]
    self.last_sent, self.last_received = int(sum(sc)), int(sum(rc))
    if self.last_sent == 0 and self.last_received == 0:
        return 0
    rec = vol.route_export(self.AXIS, self.slab_blocks, world, rank, send_counts, total=self.last_sent)
    recv = torch.empty((sum(rc), vol.RECORD_WORDS), dtype=torch.float32, device=rec.device)
    dist.all_to_all_single(recv, rec, output_split_sizes=rc, input_split_sizes=sc, group=self.group)
    off = 0
    for cnt in rc:          # one merge per source rank: keys are unique per source
        if cnt > 0:
            vol.merge_records(recv[off:off + cnt])
        off += cnt
    return self.last_received


BlockRouter._route_records = _route_records_impl


class P2PBlockRouter:
    """Block routing over peer memory (one box, NVLink/NVSwitch): the export kernel stores every
    non-owned block directly into its owner's receive region (CUDA IPC mapping), a tiny
    all_reduce orders "all exports done" before "merge", and every rank merges the records
    that landed in its own memory.  No staging buffer, no data-path collective, no host sync.

    Receive buffer of a rank (two of them, used on alternating steps so that a fast neighbour
    can never overwrite records that are still being merged):
        [int32 count_from[world] | pad to 256 B][region 0][region 1] ... [region world-1]
    region s holds up to `region_records` 10 256-byte records coming from rank s."""

    AXIS = 2
    HEADER = 256

    def __init__(self, vol, rank, world, slab_frames, frame_advance, block_size, region_records=16384, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _lib
        self.vol, self.rank, self.world, self.group = vol, rank, world, group
        self.lib, self.ctx = vol.lib, vol.ctx
        self.slab_blocks = max(1, int(round(slab_frames * frame_advance / block_size)))
        self.region_records = int(region_records)
        self.region_bytes = self.region_records * vol.RECORD_WORDS * 4
        total = self.HEADER + world * self.region_bytes
        self.local, self.peers = [], []
        for parity in range(2):
            ptr = C.c_void_p()
            handle = (C.c_uint8 * 64)()
            _lib.check(self.lib.t3d_ipc_alloc(self.ctx.handle, total, C.byref(ptr), handle))
            handles = [None] * world
            dist.all_gather_object(handles, bytes(handle), group=group)
            bases = []
            for d in range(world):
                if d == rank:
                    bases.append(ptr.value)
                    continue
                q = C.c_void_p()
                hb = (C.c_uint8 * 64).from_buffer_copy(handles[d])
                _lib.check(self.lib.t3d_ipc_open(self.ctx.handle, hb, C.byref(q)))
                bases.append(q.value)
            self.local.append(ptr.value)
            self.peers.append(bases)
        self.fill = torch.zeros(world + 2, dtype=torch.int32, device=self.ctx.device)
        self.token = torch.zeros(1, dtype=torch.int32, device=self.ctx.device)
        self.step = 0
        self._side = self._done = self._snap = None
        self._regions = [(C.c_void_p * world)() for _ in range(2)]
        self._counts = [(C.c_void_p * world)() for _ in range(2)]
        for parity in range(2):
            for d in range(world):
                self._regions[parity][d] = self.peers[parity][d] + self.HEADER + rank * self.region_bytes
                self._counts[parity][d] = self.peers[parity][d] + 4 * rank

    def _export(self, n_blocks_dev=None):
        """export kernel (peer stores) + fence, on the current stream; returns the receive-buffer parity."""
        import torch.distributed as dist
        from . import _lib
        from .runtime import _ptr, _stream
        par = self.step & 1
        self.step += 1
        _lib.check(self.lib.t3d_tsdf_route_export_p2p(self.vol.handle, self.AXIS, self.slab_blocks, self.world, self.rank,
                                                      self._regions[par], self._counts[par], self.region_records,
                                                      _ptr(self.fill), _ptr(n_blocks_dev), _stream()))
        dist.all_reduce(self.token, group=self.group)     # every rank's export precedes every rank's merge
        return par

    def _merge(self, par):
        import ctypes as C
        from . import _lib
        from .runtime import _stream
        base = self.local[par]
        for s in range(self.world):
            if s != self.rank:
                _lib.check(self.lib.t3d_tsdf_merge_records_dev(
                    self.vol.handle, C.c_void_p(base + self.HEADER + s * self.region_bytes), C.c_void_p(base + 4 * s),
                    self.region_records, _stream()))
        _lib.check(self.lib.t3d_memset_async(C.c_void_p(base), 0, self.HEADER, _stream()))

    def route(self, n_blocks_dev=None):
        if self.world == 1:
            return
        self._merge(self._export(n_blocks_dev))

    # ---- routing that overlaps fusion -------------------------------------------------------------
    # The frames whose blocks must travel (the last ceil(depth_max / frame_advance) + 2 of a rank's
    # slab: only they reach past its end) are fused FIRST, then export -> fence -> merge run on the
    # router's own stream underneath the fusion of the other batches; the batch that meets the
    # incoming blocks (the slab's first frames) is fused LAST, after the merge.  Unit-weight running
    # means do not depend on the order in which frames arrive (weights exact, tsdf to rounding).
    @staticmethod
    def tail_head_order(n_frames, batch):
        """Frame order of CopyEngineBlockRouter.fuse_overlapped: the TAIL batch (the only frames whose blocks travel)
        first, the HEAD batch (the slab's first frames, the only ones that meet incoming blocks) second, the middle
        batches after them; frames ascending inside a batch."""
        head = list(range(0, min(batch, n_frames)))
        tail_lo = max(len(head), n_frames - batch)
        tail = list(range(tail_lo, n_frames))
        middle = list(range(len(head), tail_lo))
        return tail + head + middle

    @staticmethod
    def overlap_order(n_frames, batch):
        """Frame order for fuse_overlapped: batches in descending order, frames ascending inside."""
        order = []
        hi = n_frames
        while hi > 0:
            lo = max(0, hi - batch)
            order.extend(range(lo, hi))
            hi = lo
        return order

    def fuse_overlapped(self, views_reordered, n_frames, H, W, batch, depth_is_u16=False, depth_scale=1.0,
                        depth_max=5.0, frame_advance=0.25):
        """Fuse `views_reordered` (frame views in overlap_order) and route in the same pass.  The
        current stream ends ordered after fusion AND routing."""
        import torch
        from . import _lib
        from .runtime import _stream
        nb = -(-n_frames // batch)
        reach = int(np.ceil(depth_max / frame_advance)) + 2
        if nb < 3 or batch < reach:
            raise ValueError(f"overlapped routing needs >= 3 batches of >= {reach} frames (got {nb} x {batch})")
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.ctx.device)
            self._done = C.c_void_p()
            _lib.check(self.lib.t3d_event_create(C.byref(self._done)))
            self._snap = torch.zeros(1, dtype=torch.int32, device=self.ctx.device)
        side = self._side

        state = {}

        def after_batch0(phase, ev_a, ev_b):
            h = C.c_void_p(side.cuda_stream)
            with torch.cuda.stream(side):
                if phase == 0:      # behind K5 of batch 0: export + fence, underneath the middle batches
                    _lib.check(self.lib.t3d_stream_wait_event(h, ev_b))
                    state["par"] = self._export(self._snap)
                elif phase == -1:   # behind K4 of the last batch (both allocate blocks): merge, then K5 of the last batch
                    _lib.check(self.lib.t3d_stream_wait_event(h, ev_a))
                    self._merge(state["par"])
                    _lib.check(self.lib.t3d_event_record(self._done, h))

        self.vol.integrate_sequence_hooked(views_reordered, n_frames, H, W, batch, self._snap, after_batch0,
                                           self._done, depth_is_u16, depth_scale, depth_max)
        _lib.check(self.lib.t3d_stream_wait_event(_stream(), self._done))

    def stats(self):
        """(records sent per destination, records that did not fit) of the last route — synchronises, and
        raises if the export kernel had to drop records (the counters live on the device: this transport
        never syncs on its own, so callers check here, at their next host sync)."""
        from . import _lib
        f = self.fill.cpu().tolist()
        sent = [f[d] if d != self.rank else 0 for d in range(self.world)]
        if f[self.world + 1] != 0:
            raise _lib.T3DError(_lib.T3D_E_CAPACITY,
                                f"P2PBlockRouter: {f[self.world + 1]} block records did not fit the receive regions "
                                f"(region_records={self.region_records}); the fused volume is incomplete")
        return sent, f[self.world + 1]

    def close(self):
        if self._done is not None:
            self.lib.t3d_event_destroy(self._done)
            self._done = None
        for parity in range(2):
            for d in range(self.world):
                if d != self.rank:
                    self.lib.t3d_ipc_close(self.ctx.handle, self.peers[parity][d])
            self.lib.t3d_ipc_free(self.ctx.handle, self.local[parity])
        self.peers, self.local = [], []

class CopyEngineBlockRouter(P2PBlockRouter):
    """Block routing with the transfer on the COPY ENGINE (default on one NVLink box).  K5's persistent CTAs
    own every SM while a step runs, so an export kernel that stores 114 MB per rank over NVLink holds SMs for
    as long as the link takes; here the SMs only pack the records into a local send buffer (an HBM-speed
    copy), each destination's group then travels with ONE peer cudaMemcpyAsync (NVLink through the copy
    engine, concurrent with fusion) plus a 4-byte copy of its count, a 4-byte NCCL all_reduce orders
    "all copies done" before "merge", and ONE fused launch merges every source's region.

    The per-destination counts are read on the host (the copies need their sizes): in the overlapped mode
    that read waits only for K4 of batch 0 — all blocks that will travel exist from then on — i.e. it returns
    while K5 of batch 0 is still running, so fusion never waits for it.  Knowing the counts on the host also
    means a receive region that is too small raises immediately instead of dropping records."""

    def __init__(self, vol, rank, world, slab_frames, frame_advance, block_size, region_records=16384, group=None):
        import torch
        super().__init__(vol, rank, world, slab_frames, frame_advance, block_size, region_records, group)
        dev = self.ctx.device
        self.counts_dev = torch.zeros(world, dtype=torch.int32, device=dev)
        self.base_dev = torch.zeros(world, dtype=torch.int32, device=dev)
        self.fill_dev = torch.zeros(world, dtype=torch.int32, device=dev)
        self.counts_pin = torch.zeros(world, dtype=torch.int32).pin_memory()
        self.base_pin = torch.zeros(world, dtype=torch.int32).pin_memory()
        self.send = None
        self.last_counts = [0] * world
        self.route_events = []          # (start, end) CUDA events of every route, for bench.py

    def _enqueue_counts(self, n_blocks_dev):
        """Enqueue the per-owner count and its copy to pinned host memory on the current stream (no host sync)."""
        from . import _lib
        from .runtime import _ptr, _stream
        _lib.check(self.lib.t3d_tsdf_route_counts_upto(self.vol.handle, self.AXIS, self.slab_blocks, self.world, self.rank,
                                                       _ptr(n_blocks_dev), _ptr(self.counts_dev), _stream()))
        self.counts_pin.copy_(self.counts_dev, non_blocking=True)

    def _counts_on_host(self):
        """The counts of the last _enqueue_counts (the caller has synchronised with it); raises if a region is too small."""
        from . import _lib
        counts = self.counts_pin.tolist()
        worst = max(counts)
        if worst > self.region_records:
            raise _lib.T3DError(_lib.T3D_E_CAPACITY,
                                f"CopyEngineBlockRouter: {worst} block records for one owner exceed the receive region "
                                f"(region_records={self.region_records}); nothing was sent — raise region_records")
        return counts

    def _read_counts(self, n_blocks_dev):
        """Enqueue the per-owner count on the current stream and read it on the host."""
        import torch
        self._enqueue_counts(n_blocks_dev)
        torch.cuda.current_stream().synchronize()
        return self._counts_on_host()

    def _send(self, counts, n_blocks_dev):
        """pack -> per-destination peer copies -> fence, all on the current stream; returns the buffer parity."""
        import torch
        import torch.distributed as dist
        from . import _lib
        from .runtime import _ptr, _stream
        par = self.step & 1
        self.step += 1
        total = sum(counts)
        rec_bytes = self.vol.RECORD_WORDS * 4
        if total > 0:
            if self.send is None or self.send.shape[0] < total:
                self.send = torch.empty((int(total * 1.25) + 16, self.vol.RECORD_WORDS), dtype=torch.float32,
                                        device=self.ctx.device)
            base, acc = [], 0
            for c in counts:
                base.append(acc)
                acc += c
            self.base_pin.copy_(torch.tensor(base, dtype=torch.int32))
            self.base_dev.copy_(self.base_pin, non_blocking=True)
            _lib.check(self.lib.t3d_tsdf_route_export_upto(self.vol.handle, self.AXIS, self.slab_blocks, self.world,
                                                           self.rank, _ptr(n_blocks_dev), _ptr(self.base_dev),
                                                           _ptr(self.fill_dev), _ptr(self.send), _stream()))
            for d, c in enumerate(counts):
                if d == self.rank or c == 0:
                    continue
                _lib.check(self.lib.t3d_memcpy_async(self._regions[par][d], C.c_void_p(self.send.data_ptr() + base[d] * rec_bytes),
                                                     c * rec_bytes, _stream()))
                _lib.check(self.lib.t3d_memcpy_async(self._counts[par][d], C.c_void_p(self.counts_dev.data_ptr() + 4 * d),
                                                     4, _stream()))
        dist.all_reduce(self.token, group=self.group)     # every rank's copies precede every rank's merge
        self.last_counts = counts
        return par

    def _merge_all(self, par):
        """ONE fused insert + merge launch pair for every source's region, then the header is cleared."""
        from . import _lib
        from .runtime import _stream
        local = self.local[par]
        _lib.check(self.lib.t3d_tsdf_merge_records_multi(self.vol.handle, C.c_void_p(local), self.HEADER, self.region_bytes,
                                                         self.world, self.rank, self.region_records, _stream()))
        _lib.check(self.lib.t3d_memset_async(C.c_void_p(local), 0, self.HEADER, _stream()))

    def route(self, n_blocks_dev=None):
        import torch
        if self.world == 1:
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        self._merge_all(self._send(self._read_counts(n_blocks_dev), n_blocks_dev))
        e1.record()
        self.route_events.append((e0, e1))

    def fuse_overlapped(self, views_reordered, n_frames, H, W, batch, depth_is_u16=False, depth_scale=1.0,
                        depth_max=5.0, frame_advance=0.25):
        """Fuse `views_reordered` (frame views in tail_head_order) and route in the same pass.  Batch 0 is the slab's
        tail (the only frames whose blocks travel), batch 1 its head (the only frames that meet incoming blocks):
        count / pack / copies / fence are enqueued on the router's stream behind K5 of batch 0 and run underneath the
        other batches; the merge follows K5 of batch 1.  NOTHING in the fusion waits for the routing — no later batch
        reaches an incoming block — only the end of the step does, so the route has the whole step to hide in.
        The current stream ends ordered after fusion AND routing."""
        import torch
        from . import _lib
        from .runtime import _stream
        nb = -(-n_frames // batch)
        reach = int(np.ceil(depth_max / frame_advance)) + 2
        if nb < 3 or batch < reach:
            raise ValueError(f"overlapped routing needs >= 3 batches of >= {reach} frames (got {nb} x {batch})")
        if self._side is None:
            # high priority: at a K5 launch boundary the routing kernels are scheduled before the next K5's CTAs
            self._side = torch.cuda.Stream(device=self.ctx.device, priority=-1)
            self._done = C.c_void_p()
            _lib.check(self.lib.t3d_event_create(C.byref(self._done)))
            self._snap = torch.zeros(1, dtype=torch.int32, device=self.ctx.device)
        side = self._side
        state = {}

        def hook(phase, ev_a, ev_b):
            h = C.c_void_p(side.cuda_stream)
            if phase == 0:
                # the blocks that will travel exist once K4 of the tail batch is done: count them on the device.  The
                # host does NOT wait here: K5's persistent CTAs own every SM, so the count kernel only runs when K5 of
                # the tail batch retires, and a host that waits for it cannot enqueue the next batch meanwhile (the
                # GPU then idles for the count, the host's pack / copy / fence calls and a whole K4).
                with torch.cuda.stream(side):
                    _lib.check(self.lib.t3d_stream_wait_event(h, ev_a))
                    self._enqueue_counts(self._snap)
                    state["counted"] = torch.cuda.Event()
                    state["counted"].record()
                state["k5_tail"] = ev_b
            elif phase == 1:
                # K5 of the head batch is enqueued: the GPU is busy for two more batches.  Now read the counts (they
                # arrive when K5 of the tail batch ends) and enqueue pack / copies / fence behind that K5 and the
                # merge behind the head batch's K5 — the head batch is the only one that touches incoming blocks.
                state["counted"].synchronize()
                counts = self._counts_on_host()
                with torch.cuda.stream(side):
                    _lib.check(self.lib.t3d_stream_wait_event(h, state["k5_tail"]))
                    e0 = torch.cuda.Event(enable_timing=True)
                    e0.record()
                    par = self._send(counts, self._snap)
                    _lib.check(self.lib.t3d_stream_wait_event(h, ev_b))
                    self._merge_all(par)
                    e1 = torch.cuda.Event(enable_timing=True)
                    e1.record()
                    self.route_events.append((e0, e1))
                _lib.check(self.lib.t3d_event_record(self._done, h))

        self.vol.integrate_sequence_hooked(views_reordered, n_frames, H, W, batch, self._snap, hook, None,
                                           depth_is_u16, depth_scale, depth_max, hook_batch=0)
        _lib.check(self.lib.t3d_stream_wait_event(_stream(), self._done))

    def stats(self):
        """(records sent per destination, records dropped = always 0: an overflow raises before anything is sent)."""
        return [c if d != self.rank else 0 for d, c in enumerate(self.last_counts)], 0



def allreduce_normal_equations(acc27, sum_d2, count, device=None, group=None):
    """ICP across ranks: every rank linearises its shard of source points
    (Context.icp_linearize), then the 29 sums are all-reduced and every rank solves the
    same 6x6 system."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(acc27) + [sum_d2, count], dtype=torch.float64, device=device)
    dist.all_reduce(t, group=group)
    out = t.cpu().numpy()
    return out[:27], float(out[27]), float(out[28])


def solve_icp_update(acc27):
    """6x6 solve of the point-to-plane normal equations (host, f64) — same rule as the
    library: |det| < 1e-6 or non-finite -> identity (R8)."""
    A = np.zeros((6, 6))
    q = 0
    for a in range(6):
        for b in range(a, 6):
            A[a, b] = A[b, a] = acc27[q]
            q += 1
    b = -np.asarray(acc27[21:27], np.float64)
    det = np.linalg.det(A)
    U = np.eye(4)
    if not np.isfinite(det) or abs(det) < 1e-6:
        return U
    x = np.linalg.solve(A, b)
    if not np.all(np.isfinite(x)):
        return U
    ca, sa, cb, sb, cg, sg = np.cos(x[0]), np.sin(x[0]), np.cos(x[1]), np.sin(x[1]), np.cos(x[2]), np.sin(x[2])
    U[:3, :3] = [[cg * cb, cg * sb * sa - sg * ca, cg * sb * ca + sg * sa],
                 [sg * cb, sg * sb * sa + cg * ca, sg * sb * ca - cg * sa],
                 [-sb, cb * sa, cb * ca]]
    U[:3, 3] = x[3:]
    return U


def sharded_icp(ctx, src_shard, tgt, tgt_nrm, max_corr_dist, n_src_total, init=None, max_iter=30,
                relative_fitness=1e-6, relative_rmse=1e-6, group=None):
    """Point-to-plane ICP with the source cloud sharded across ranks.  Returns
    (T, fitness, rmse, iterations) — identical on every rank."""
    T = np.eye(4) if init is None else np.array(init, np.float64)

    def lin(Tc):
        a27, sd2, cnt = ctx.icp_linearize(src_shard, tgt, tgt_nrm, max_corr_dist, Tc)
        return allreduce_normal_equations(a27, sd2, cnt, device=src_shard.device, group=group)

    a27, sd2, cnt = lin(T)
    fit = cnt / n_src_total
    rmse = np.sqrt(sd2 / cnt) if cnt > 0 else 0.0
    it = 0
    while it < max_iter:
        T = solve_icp_update(a27) @ T
        a27, sd2, cnt = lin(T)
        f2 = cnt / n_src_total
        r2 = np.sqrt(sd2 / cnt) if cnt > 0 else 0.0
        stop = abs(fit - f2) < relative_fitness and abs(rmse - r2) < relative_rmse
        fit, rmse = f2, r2
        it += 1
        if stop:
            break
    return T, fit, rmse, it


# ---------------------------------------------------------------------------------------------------
# K2 sharded: every rank holds a shard of the points; owner(voxel) = hash(index) mod G.
# ---------------------------------------------------------------------------------------------------
def sharded_voxel_downsample(ops, xyz, rgb, voxel_size, group=None, sorted_output=True):
    """Voxel-grid downsample of the union of all ranks' points.  `ops` provides bounds(),
    voxel_partials() and voxel_merge_partials() (runtime.Context on the GPU; a NumPy stand-in in the
    gloo tests).  Steps: exact all_reduce of the bounds -> per-rank partial sums per voxel, grouped by
    owner -> one all_to_all for the counts and one for the 56-byte records -> the owner adds the
    partials and divides once.  Returns this rank's OWNED voxels (dict as voxel_downsample); the
    union over ranks equals the single-GPU result bit for bit for float32 points."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = xyz.device
    big = np.full(3, np.inf)
    mn, mx = ops.bounds(xyz) if xyz.shape[0] else (big, -big)
    b = torch.tensor(np.concatenate([mn, -np.asarray(mx)]), dtype=torch.float64, device=dev)
    dist.all_reduce(b, op=dist.ReduceOp.MIN, group=group)
    b = b.cpu().numpy()
    if not np.isfinite(b).all():                      # no rank has points
        return dict(points=torch.empty((0, 3), dtype=torch.float64, device=dev), colors=None, rgb_sum=None,
                    count=torch.empty(0, dtype=torch.int32, device=dev),
                    idx=torch.empty((0, 3), dtype=torch.int32, device=dev), min_bound=None, m=0)
    gmin, gmax = b[:3], -b[3:]
    minb = gmin - voxel_size * 0.5                    # R2: min_bound = min(p) - 0.5 * v, one global grid
    rec, counts = ops.voxel_partials(xyz, rgb, voxel_size, minb, gmax, world)
    send_counts = counts.to(torch.int64)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    sc, rc = send_counts.cpu().tolist(), recv_counts.cpu().tolist()
    recv = torch.empty((sum(rc), rec.shape[1]), dtype=rec.dtype, device=dev)
    dist.all_to_all_single(recv, rec.contiguous(), output_split_sizes=rc, input_split_sizes=sc, group=group)
    out = ops.voxel_merge_partials(recv, rgb is not None, voxel_size, minb, gmax, sorted_output)
    out["records_sent"], out["records_received"] = int(sum(sc)), int(sum(rc))
    return out


def sharded_statistical_outlier(ops, xyz, nb_neighbors=20, std_ratio=2.0, group=None):
    """remove_statistical_outlier (R3) with the kNN queries sharded over the ranks.  `xyz`
    is the FULL cloud, replicated on every rank (it is the downsampled one: small); rank r computes
    the mean neighbour distance for its share of the grid-sorted points, one all_reduce(MAX) over a
    -inf-initialised vector assembles them, and every rank derives the same mu / sigma / mask.
    Returns (keep u8[N], mean f64[N], (mu, sigma, thr), kept) — identical to the single-GPU call."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mean = torch.full((xyz.shape[0],), float("-inf"), dtype=torch.float64, device=xyz.device)
    ops.sor_mean_distances_part(xyz, nb_neighbors, rank, world, mean)
    dist.all_reduce(mean, op=dist.ReduceOp.MAX, group=group)
    keep, stats, kept = ops.sor_from_mean_distances(mean, std_ratio)
    return keep, mean, stats, kept


def gather_rows(rows, dst=0, group=None):
    """Concatenate every rank's (n_r, ...) tensor on rank `dst` in rank order (None elsewhere):
    all_gather of the row counts + one all_to_all in which only `dst` receives."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n, group=group)
    ns = [int(x.item()) for x in ns]
    flat = rows.reshape(rows.shape[0], -1).contiguous()
    sc = [flat.shape[0] if d == dst else 0 for d in range(world)]
    rc = ns if rank == dst else [0] * world
    recv = torch.empty((sum(rc), flat.shape[1]), dtype=flat.dtype, device=flat.device)
    dist.all_to_all_single(recv, flat, output_split_sizes=rc, input_split_sizes=sc, group=group)
    return recv.reshape((-1,) + tuple(rows.shape[1:])) if rank == dst else None


# ---------------------------------------------------------------------------------------------------
# K6 / K9 sharded: each owner extracts the surface of its z-slab; the +z neighbour tests
# and gradients at the slab's upper face need the first block layer(s) of the next slab (a halo).
# ---------------------------------------------------------------------------------------------------
def extract_owned_surface(vol, rank, world, slab_blocks, weight_threshold=3.0, axis=2, halo_blocks=1, group=None):
    """Surface points (R6) of the blocks this rank owns, exactly as a single volume holding every
    rank's blocks would emit them: the halo layers [hi, hi + halo_blocks) of the upper neighbour and
    [lo - halo_blocks, lo) of the lower one are fetched first (block records, one all_to_all), merged
    as read-only context, and extraction is restricted to the owned range.
    Returns (xyz f32, normals f32, rgb u8) device tensors."""
    import torch
    import torch.distributed as dist
    lo, hi = block_owner_range(rank, world, slab_blocks)
    lo_c, hi_c = max(lo, -(1 << 20)), min(hi, 1 << 20)
    # what the neighbours need from me: my lowest layers go down, my highest layers go up
    send = []
    for d in range(world):
        if d == rank - 1:
            send.append(vol.export_blocks_range(axis, lo_c, lo_c + halo_blocks))
        elif d == rank + 1:
            send.append(vol.export_blocks_range(axis, hi_c - halo_blocks, hi_c))
        else:
            send.append(None)
    dev = vol.ctx.device if hasattr(vol, "ctx") else "cpu"
    sc = torch.tensor([0 if s_ is None else s_[0].shape[0] for s_ in send], dtype=torch.int64, device=dev)
    rcv = torch.empty_like(sc)
    dist.all_to_all_single(rcv, sc, group=group)
    scl, rcl = sc.cpu().tolist(), rcv.cpu().tolist()

    def cat(i, shape, dtype):
        parts = [s_[i].reshape((s_[i].shape[0],) + shape) for s_ in send if s_ is not None and s_[0].shape[0]]
        return torch.cat(parts).contiguous() if parts else torch.empty((0,) + shape, dtype=dtype, device=dev)
    out = []
    for i, (shape, dtype) in enumerate([((3,), torch.int32), ((512,), torch.float32), ((512,), torch.float32),
                                        ((512, 3), torch.float32)]):
        snd = cat(i, shape, dtype)
        flat = snd.reshape(snd.shape[0], -1)
        r = torch.empty((sum(rcl), flat.shape[1]), dtype=dtype, device=dev)
        dist.all_to_all_single(r, flat, output_split_sizes=rcl, input_split_sizes=scl, group=group)
        out.append(r.reshape((-1,) + shape))
    # a scratch volume = exact copy of the owned slab + the neighbours' halo layers: the fused volume
    # is not modified, and this rank's own partial copies of non-owned blocks stay out of the picture
    own = vol.export_blocks_range(axis, lo_c, hi_c)
    tmp = type(vol)(vol.voxel_size, vol.sdf_trunc, block_capacity=max(2 * (own[0].shape[0] + out[0].shape[0]), 1024),
                    ctx=vol.ctx)
    if own[0].shape[0]:
        tmp.merge_blocks(*[t.contiguous() for t in own])
    if out[0].shape[0]:
        tmp.merge_blocks(out[0].contiguous(), out[1].contiguous(), out[2].contiguous(), out[3].contiguous())
    res = tmp.extract_points_range(axis, lo_c, hi_c, weight_threshold)
    tmp.close()
    return res
