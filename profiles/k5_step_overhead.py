import sys, json
sys.path.insert(0, '/root/repo')
import torch, numpy as np
import bench as B
from textureless_3d_reconstruction_b200.runtime import TSDFVolume, get_context
ctx = get_context(0); dev = ctx.device
F, H, W = 300, B.H, B.W
depth = torch.empty((F, H, W), dtype=torch.float32, device=dev); bgr = torch.empty((F, H, W, 3), dtype=torch.uint8, device=dev); poses = []
for i in range(F):
    _, _, T = ctx.synth_frame(0, i, H, W, *B.KINTR, seed=B.SEED, noise_sigma=B.NOISE, depth=depth[i], bgr=bgr[i]); poses.append(T)
vol = TSDFVolume(B.VOXEL, B.TRUNC, block_capacity=600_000, ctx=ctx)
views = vol.make_frame_views([depth[i] for i in range(F)], [bgr[i] for i in range(F)], [B.KINTR] * F, poses)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
res = {}
res['reset'] = t(lambda: vol.reset())
for bsz in (32, 16, 8):
    def step(bsz=bsz):
        vol.reset(); vol.integrate_sequence(views, F, H, W, bsz, False, 1.0, B.DEPTH_MAX)
    res[f'step_b{bsz}'] = t(step)
def serial():
    vol.reset()
    for s in range(0, F, 32): vol.integrate_views(views, s, min(32, F - s), H, W, False, 1.0, B.DEPTH_MAX)
res['step_serial_b32'] = t(serial)
vol.set_profiling(True); serial(); p = vol.get_profile(); vol.set_profiling(False)
res['serial_profile'] = p
print(json.dumps(res))
