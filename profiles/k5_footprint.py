"""K5: bytes a TMA-staged (block, frame) image tile would move vs the sectors the per-voxel gathers touch.

north_star proposes "TMA-staged shared-memory tiles for the projected voxel-block frustum" for TSDF integration.
For every (touched block, frame) pair of a few cfg-2 frames this script computes, on the CPU (oracle frames, oracle
block keys, the kernel's own projection formula in float64):
  * tile bytes  : the image rectangle that bounds the block's 512 projected voxels, rows rounded out to 16 B
                  (cp.async.bulk granularity), 4 B/pixel depth (+ 3 B/pixel colour for the `with colour` figure);
  * gather bytes: the DISTINCT 32-byte sectors the 512 depth gathers touch (+ the sectors of the colour words of
                  the voxels that pass the R5 tests) — what the L1 misses of the gather formulation can cost at most.
Printed per frame and in total; profiles/r2_k5_footprint.txt is its output.  (CPU only: python profiles/k5_footprint.py)"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import capi  # noqa: E402

H, W, K = 1920, 1080, (1719.0, 1719.0, 540.0, 960.0)
VOXEL, TRUNC, DMAX = 0.01, 0.04, 5.0
FRAMES = [0, 150, 299] if len(sys.argv) < 2 else [int(a) for a in sys.argv[1:]]

g = np.stack(np.meshgrid(np.arange(8), np.arange(8), np.arange(8), indexing="ij"), -1).reshape(-1, 3)   # x, y, z
tot = np.zeros(6)
print("# frame  pairs  | per (block, frame): tile px  tile B depth  tile B depth+bgr | gather B depth  gather B depth+bgr | ratio depth  ratio all")
for fi in FRAMES:
    d, c, T = capi.synth_frame(0, fi, H, W, *K, seed=1234, noise_sigma=0.002)
    vol = capi.TSDFVolume(VOXEL, TRUNC)
    keys = vol.touch(d, K, T, 1.0, DMAX)
    T64 = np.asarray(T, np.float64)
    R, t = T64[:, :3], T64[:, 3]
    nb = len(keys)
    vox = (keys[:, None, :].astype(np.float64) * 8 + g[None]) * VOXEL            # [nb, 512, 3] world metres
    pc = vox @ R.T + t
    zc = pc[..., 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        u = K[0] * pc[..., 0] / zc + K[2]
        v = K[1] * pc[..., 1] / zc + K[3]
    inimg = (u >= 0) & (v >= 0) & (u <= W - 1) & (v <= H - 1) & (zc > 0)
    ui = np.where(inimg, np.floor(u + 0.5), 0).astype(np.int64)
    vi = np.where(inimg, np.floor(v + 0.5), 0).astype(np.int64)
    dd = d[vi, ui]
    ok = inimg & (dd > 0) & (dd <= DMAX) & (dd - zc >= -TRUNC)
    tile_px = tile_d = tile_all = gat_d = gat_all = 0.0
    for b in range(nb):
        m = inimg[b]
        if not m.any():
            continue
        u0, u1, v0, v1 = ui[b][m].min(), ui[b][m].max(), vi[b][m].min(), vi[b][m].max()
        rows = v1 - v0 + 1
        ua, ub = (u0 // 4) * 4, -(-(u1 + 1) // 4) * 4                            # 16-byte aligned f32 row segments
        ca, cb = (3 * u0 // 16) * 16, -(-(3 * (u1 + 1)) // 16) * 16               # 16-byte aligned colour row segments
        tile_px += rows * (u1 - u0 + 1)
        tile_d += rows * (ub - ua) * 4
        tile_all += rows * ((ub - ua) * 4 + (cb - ca))
        pix = vi[b][m] * W + ui[b][m]
        gd = len(np.unique(pix * 4 // 32)) * 32
        mo = ok[b]
        po = (vi[b][mo] * W + ui[b][mo]) * 3
        gc = len(np.unique(np.concatenate([po // 32, (po + 2) // 32]))) * 32 if mo.any() else 0
        gat_d += gd
        gat_all += gd + gc
    print(f"{fi:7d} {nb:6d} | {tile_px / nb:10.0f} {tile_d / nb:13.0f} {tile_all / nb:16.0f} | {gat_d / nb:14.0f} {gat_all / nb:18.0f} |"
          f" {tile_d / gat_d:11.2f} {tile_all / gat_all:9.2f}")
    tot += [nb, tile_px, tile_d, tile_all, gat_d, gat_all]
nb = tot[0]
print(f"#   all {int(nb):6d} | {tot[1] / nb:10.0f} {tot[2] / nb:13.0f} {tot[3] / nb:16.0f} | {tot[4] / nb:14.0f} {tot[5] / nb:18.0f} |"
      f" {tot[2] / tot[4]:11.2f} {tot[3] / tot[5]:9.2f}")
print("# a staged tile moves `ratio` times the bytes the gathers' distinct sectors hold; ncu (profiles/r2_k5_full.txt) measures the")
print("# gathers' real L2->L1 traffic at 2.40 GB per 64-frame launch = 9.1 KB per (block, frame) pair, 54 % of it L1 misses of sectors")
print("# a neighbouring warp had already fetched; W*3 = 3240 B rows are only 8-byte aligned, so colour rows cannot be bulk-copied at all.")
