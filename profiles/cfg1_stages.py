"""cfg 1 (depth_to_reconstruction dense path) stage times on the GPU: K1 batch, K2, K3, compaction, D2H."""
import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from textureless_3d_reconstruction_b200.runtime import get_context, to_host
from textureless_3d_reconstruction_b200 import depth_to_reconstruction as d2r
ctx = get_context(0); dev = ctx.device
H, W = 1920, 1080; K4 = (1719.0, 1719.0, 540.0, 960.0); n1 = 30
d1 = torch.empty((n1, H, W), dtype=torch.float32, device=dev); c1 = torch.empty((n1, H, W, 3), dtype=torch.uint8, device=dev); p1 = []
for i in range(n1):
    _, _, T = ctx.synth_frame(1, i, H, W, *K4, seed=1234, depth=d1[i], bgr=c1[i]); p1.append((T[:, :3].copy(), T[:, 3:4].copy()))
fr1 = ctx.make_backproject_frames([d1[i] for i in range(n1)], [c1[i] for i in range(n1)], p1)
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t = {}
    t0 = sync()
    xyz, rgb, offs = ctx.backproject_batch(fr1, n1, H, W, fx=K4[0], fy=K4[1], cx=K4[2], cy=K4[3], subsample=2, min_depth=0.1, max_depth=50.0)
    n = int(offs[-1].item()); t1 = sync(); t['k1'] = t1 - t0
    ds = ctx.voxel_downsample(xyz[:n], rgb[:n], 0.005, sorted_output=True, want_idx=False)
    pts, cols = ds["points"].contiguous(), ds["colors"].contiguous(); t2 = sync(); t['k2'] = t2 - t1
    keep, _, _, _ = ctx.statistical_outlier(pts, 20, 2.0); t3 = sync(); t['k3'] = t3 - t2
    pts2 = ctx.compact_rows(pts, keep); cols2 = ctx.compact_rows(cols, keep); t4 = sync(); t['compact'] = t4 - t3
    ph, ch = to_host(pts2), to_host(cols2); t5 = sync(); t['d2h_pinned'] = t5 - t4
    ph2, ch2 = pts2.cpu().numpy(), cols2.cpu().numpy(); t6 = sync(); t['d2h_pageable'] = t6 - t5
    assert np.array_equal(ph, ph2) and np.array_equal(ch, ch2)
    print(json.dumps({k: round(v * 1e3, 3) for k, v in t.items()}), n, len(ph))
