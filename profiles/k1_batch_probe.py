import sys, torch
sys.path.insert(0, '.')
from textureless_3d_reconstruction_b200.runtime import get_context
ctx = get_context(0); dev = ctx.device
H, W, K4, NF = 1920, 1080, (1719.0, 1719.0, 540.0, 960.0), 24
depth = torch.empty((NF, H, W), dtype=torch.float32, device=dev); bgr = torch.empty((NF, H, W, 3), dtype=torch.uint8, device=dev); poses = []
for i in range(NF):
    _, _, T = ctx.synth_frame(0, i, H, W, *K4, seed=1234, noise_sigma=0.002, depth=depth[i], bgr=bgr[i]); poses.append(T)
frames = ctx.make_backproject_frames([depth[i] for i in range(NF)], [bgr[i] for i in range(NF)], [(poses[i][:, :3], poses[i][:, 3:4]) for i in range(NF)])
for s in (2, 4):
    P = (-(-H // s)) * (-(-W // s))
    o_xyz = torch.empty((P * NF, 3), dtype=torch.float32, device=dev); o_rgb = torch.empty((P * NF, 3), dtype=torch.uint8, device=dev); offs = torch.zeros(NF + 1, dtype=torch.int64, device=dev)
    for _ in range(4):
        ctx.backproject_batch(frames, NF, H, W, fx=K4[0], fy=K4[1], cx=K4[2], cy=K4[3], subsample=s, min_depth=0.1, max_depth=50.0, out_xyz=o_xyz, out_rgb=o_rgb, out_offsets=offs)
    torch.cuda.synchronize()
print("ok")
