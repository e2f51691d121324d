"""Timing probe of BlockRouter.route() stages on N GPUs (torchrun). Debug aid."""
import os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch, torch.distributed as dist
from textureless_3d_reconstruction_b200.runtime import TSDFVolume, get_context
from textureless_3d_reconstruction_b200 import distributed as D
H, W = 1920, 1080; K = (1719.0, 1719.0, 540.0, 960.0)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = get_context(local); F = 300
depth = torch.empty((F, H, W), dtype=torch.float32, device=ctx.device); bgr = torch.empty((F, H, W, 3), dtype=torch.uint8, device=ctx.device)
poses = []
for i in range(F):
    _, _, T = ctx.synth_frame(0, rank * F + i, H, W, *K, seed=1234, noise_sigma=0.002, depth=depth[i], bgr=bgr[i]); poses.append(T)
vol = TSDFVolume(0.01, 0.04, block_capacity=600000, ctx=ctx)
views = vol.make_frame_views([depth[i] for i in range(F)], [bgr[i] for i in range(F)], [K] * F, poses)
router = D.BlockRouter(vol, rank, world, slab_frames=F, frame_advance=0.25, block_size=0.08)
def T(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(4):
    vol.reset(); vol.integrate_sequence(views, F, H, W, 32, False, 1.0, 5.0)
    t0 = T(); keys, tsdf, w, rgb = router._export_non_owned(); t1 = T()
    n = keys.shape[0]
    owner = D.owner_of(keys[:, 2].to(torch.int64), world, router.slab_blocks); order = torch.argsort(owner, stable=True); sc = torch.bincount(owner, minlength=world); t2 = T()
    rec = torch.empty((n, router.RECORD), dtype=torch.float32, device=keys.device)
    rec[:, :3] = keys.view(torch.float32); rec[:, 3] = 0; rec[:, 4:516] = tsdf; rec[:, 516:1028] = w; rec[:, 1028:] = rgb.reshape(n, 1536); rec = rec.index_select(0, order); t3 = T()
    rcnt = torch.empty_like(sc); dist.all_to_all_single(rcnt, sc); scl, rcl = sc.tolist(), rcnt.tolist(); t4 = T()
    recv = torch.empty((sum(rcl), router.RECORD), dtype=torch.float32, device=keys.device)
    dist.all_to_all_single(recv, rec.contiguous(), output_split_sizes=rcl, input_split_sizes=scl); t5 = T()
    off = 0
    for cnt in rcl:
        if cnt: 
            seg = recv[off:off + cnt]
            vol.merge_blocks(seg[:, :3].contiguous().view(torch.int32), seg[:, 4:516].contiguous(), seg[:, 516:1028].contiguous(), seg[:, 1028:].contiguous().reshape(cnt, 512, 3))
        off += cnt
    t6 = T()
    print(f"rank {rank} it {it}: n={n} send={scl} recv={rcl} export={1e3*(t1-t0):.3f} owner={1e3*(t2-t1):.3f} pack={1e3*(t3-t2):.3f} a2a_cnt={1e3*(t4-t3):.3f} a2a={1e3*(t5-t4):.3f} merge={1e3*(t6-t5):.3f} total={1e3*(t6-t0):.3f} ms", flush=True)
dist.destroy_process_group()
