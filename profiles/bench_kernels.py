#!/usr/bin/env python
"""Per-kernel roofline + CPU baseline for every row of SURVEY §8a besides the
bench.py headline (K5).  Run on the GPU box:

    python profiles/bench_kernels.py [--quick] > gpurun_out/kernels_rN.jsonl

One JSON line per kernel: device time (CUDA events on the launching stream, after
warm-up, inputs cycled so the working set exceeds the 126 MB L2), SURVEY §8d
algorithmic bytes, achieved GB/s against MEASURED_PEAKS.json, and the CPU oracle
(oracle/t3d_oracle.c, OpenMP, all host cores — test infrastructure, used here only
as the timed baseline) on a bounded sample of the same input.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

H, W = 1920, 1080
K4 = (1719.0, 1719.0, 540.0, 960.0)


def peak():
    try:
        return float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]), "measured"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback"


def gpu_time(fn, iters, warm=3):
    import torch
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(warm + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def emit(name, unit_name, units, ms, alg_bytes, cpu=None, **extra):
    pk, src = peak()
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    line = {"kernel": name, "unit": unit_name, "units_per_call": units, "ms_per_call": ms,
            "units_per_s": units / (ms * 1e-3), "algorithmic_bytes": alg_bytes, "achieved_GBs": gbs,
            "peak_GBs": pk, "peak_source": src, "frac": gbs / pk, "cpu_baseline": cpu}
    line.update(extra)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="", help="comma list of kernels to run, e.g. K1,K2 (default all)")
    args = ap.parse_args()
    only = set(x for x in args.only.split(",") if x)

    def want(k):
        return not only or k in only
    import torch
    from oracle import capi
    from textureless_3d_reconstruction_b200.runtime import TSDFVolume, get_context, to_host

    ctx = get_context(0)
    dev = ctx.device
    cores = capi.num_threads()
    NF = 8 if args.quick else 24           # 24 x 14.5 MB = 348 MB of frames > L2
    depth = torch.empty((NF, H, W), dtype=torch.float32, device=dev)
    bgr = torch.empty((NF, H, W, 3), dtype=torch.uint8, device=dev)
    poses = []
    for i in range(NF):
        _, _, T = ctx.synth_frame(0, i, H, W, *K4, seed=1234, noise_sigma=0.002, depth=depth[i], bgr=bgr[i])
        poses.append(T)
    torch.cuda.synchronize()
    h_depth0, h_bgr0 = depth[0].cpu().numpy(), bgr[0].cpu().numpy()

    # ------------------------------------------------------------------ K1
    for s in ((1, 2, 4) if want("K1") else ()):
        Hs, Ws = -(-H // s), -(-W // s)
        P = Hs * Ws
        o_xyz = torch.empty((P, 3), dtype=torch.float32, device=dev)
        o_rgb = torch.empty((P, 3), dtype=torch.uint8, device=dev)
        o_n = torch.zeros(1, dtype=torch.int64, device=dev)

        def k1(i, s=s):
            j = i % NF
            ctx.backproject(depth[j], bgr[j], fx=K4[0], fy=K4[1], cx=K4[2], cy=K4[3], subsample=s,
                            min_depth=0.1, max_depth=50.0, pose=(poses[j][:, :3], poses[j][:, 3:4]),
                            out_xyz=o_xyz, out_rgb=o_rgb, out_n=o_n)
        ms = gpu_time(k1, 50)
        nvalid = int(o_n.item())
        pose0 = (poses[0][:, :3].copy(), poses[0][:, 3:4].copy())
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            capi.backproject(h_depth0, h_bgr0, *K4, pose=pose0, max_depth=50.0, subsample=s)
        cpu_c = (time.perf_counter() - t0) / reps
        from oracle import ref_numpy
        t0 = time.perf_counter()
        ref_numpy.d2r_depth_to_pointcloud(h_depth0, h_bgr0, *K4, pose=pose0, scale=1.0, subsample=s,
                                          min_depth=0.1, max_depth=50.0)
        cpu_np = time.perf_counter() - t0
        emit(f"K1 backproject s={s}", "frames", 1, ms, 7 * P + 15 * nvalid,
             cpu={"frames_per_s_c_openmp": 1 / cpu_c, "cores": cores,
                  "frames_per_s_numpy_restatement_1core": (1 / cpu_np) if cpu_np > 1e-4 else None},
             valid_points=nvalid, sampled_pixels=P)

    if want("K1"):
        # batched entry point: all NF frames, one launch per 32 frames, vstack-ordered output
        for s in (1, 2, 4):
            Hs, Ws = -(-H // s), -(-W // s)
            P = Hs * Ws
            frames = ctx.make_backproject_frames([depth[i] for i in range(NF)], [bgr[i] for i in range(NF)],
                                                 [(poses[i][:, :3], poses[i][:, 3:4]) for i in range(NF)])
            o_xyz = torch.empty((P * NF, 3), dtype=torch.float32, device=dev)
            o_rgb = torch.empty((P * NF, 3), dtype=torch.uint8, device=dev)
            offs = torch.zeros(NF + 1, dtype=torch.int64, device=dev)

            def k1b(i, s=s):
                ctx.backproject_batch(frames, NF, H, W, fx=K4[0], fy=K4[1], cx=K4[2], cy=K4[3], subsample=s,
                                      min_depth=0.1, max_depth=50.0, out_xyz=o_xyz, out_rgb=o_rgb, out_offsets=offs)
            ms = gpu_time(k1b, 20)
            nvalid = int(offs[-1].item())
            emit(f"K1 backproject_batch s={s} ({NF} frames/call)", "frames", NF, ms, 7 * P * NF + 15 * nvalid,
                 valid_points=nvalid, sampled_pixels=P * NF)
            del o_xyz, o_rgb

    if "CFG1" in only:
        # BASELINE configs[0]: depth_to_reconstruction's dense path on 30 x 1080x1920 frames (scene S1, given
        # poses, subsample 2 = CLI default, voxel 0.005, outlier removal) from HBM-resident frames to a .ply
        from oracle import ref_numpy
        from textureless_3d_reconstruction_b200 import depth_to_reconstruction as d2r
        from textureless_3d_reconstruction_b200.runtime import write_ply
        n1 = 30
        d1 = torch.empty((n1, H, W), dtype=torch.float32, device=dev)
        c1 = torch.empty((n1, H, W, 3), dtype=torch.uint8, device=dev)
        p1 = []
        for i in range(n1):
            _, _, T = ctx.synth_frame(1, i, H, W, *K4, seed=1234, depth=d1[i], bgr=c1[i])
            p1.append((T[:, :3].copy(), T[:, 3:4].copy()))
        dense = d2r.DenseReconstructor(d2r.ReconstructionConfig())
        fr1 = ctx.make_backproject_frames([d1[i] for i in range(n1)], [c1[i] for i in range(n1)], p1)
        out = {}

        def cfg1(i):
            xyz, rgb, offs = ctx.backproject_batch(fr1, n1, H, W, fx=K4[0], fy=K4[1], cx=K4[2], cy=K4[3], subsample=2,
                                                   min_depth=0.1, max_depth=50.0)
            n = int(offs[-1].item())
            pts, cols = dense.merge_pointclouds_device(xyz[:n], rgb[:n], 0.005)
            out["r"] = (n, to_host(pts), to_host(cols))
        ms = gpu_time(cfg1, 6, warm=3)   # steady state: the pinned result buffers alternate between two cached blocks
        n_in, pts_h, cols_h = out["r"]
        t0 = time.perf_counter()
        write_ply("/tmp/_t3d_cfg1.ply", pts_h, cols_h, layout=0)
        ply_s = time.perf_counter() - t0
        # CPU: the reference's NumPy arithmetic for K1 + the oracle for the two Open3D calls
        hd, hc = d1.cpu().numpy(), c1.cpu().numpy()
        t0 = time.perf_counter()
        clouds = [ref_numpy.d2r_depth_to_pointcloud(hd[i], hc[i], *K4, pose=p1[i], scale=1.0, subsample=2) for i in range(n1)]
        t_k1 = time.perf_counter() - t0
        P = np.vstack([a for a, _ in clouds]).astype(np.float64)
        Cc = np.vstack([b for _, b in clouds])
        t0 = time.perf_counter()
        o = capi.voxel_downsample(P, Cc, 0.005)
        t_k2 = time.perf_counter() - t0
        t0 = time.perf_counter()
        keep, _, _ = capi.statistical_outlier(o["points"], 20, 2.0)
        t_k3 = time.perf_counter() - t0
        print(json.dumps({"kernel": "CFG1 d2r dense path: 30 frames 1080x1920, s=2, voxel 5 mm, SOR (HBM -> host arrays)",
                          "frames": n1, "ms": ms, "frames_per_s": n1 / (ms * 1e-3), "points_in": n_in,
                          "points_out": int(len(pts_h)), "ply_write_s": ply_s,
                          "cpu_baseline": {"seconds": t_k1 + t_k2 + t_k3, "frames_per_s": n1 / (t_k1 + t_k2 + t_k3),
                                           "k1_numpy_reference_restatement_s": t_k1, "k2_oracle_serial_s": t_k2,
                                           "k3_oracle_openmp_s": t_k3, "cores": cores, "points_out": int(keep.sum())},
                          "same_point_count_as_cpu": bool(abs(int(keep.sum()) - len(pts_h)) <= 2)}), flush=True)
        del d1, c1
        return

    # ------------------------------------------------------------------ fused cloud for K2/K3/K7
    NC = 4 if args.quick else 16
    clouds, cols = [], []
    for j in range(NC):
        x, c, n = ctx.backproject(depth[j], bgr[j], fx=K4[0], fy=K4[1], cx=K4[2], cy=K4[3], subsample=1,
                                  min_depth=0.1, max_depth=5.0, pose=(poses[j][:, :3], poses[j][:, 3:4]))
        k = int(n.item())
        clouds.append(x[:k].clone())
        cols.append(c[:k].clone())
    pts = torch.cat(clouds).contiguous()
    rgb = torch.cat(cols).contiguous()
    del clouds, cols
    N = pts.shape[0]
    for voxel in ((0.005, 0.01, 0.02) if want("K2") else ()):
        res = {}

        def k2(i, voxel=voxel):
            res["r"] = ctx.voxel_downsample(pts, rgb, voxel, sorted_output=False, want_idx=False)
        walls = []
        for _ in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            k2(0)
            torch.cuda.synchronize()
            walls.append(round((time.perf_counter() - t0) * 1e3, 3))
        print(f"K2 v={voxel} host wall per call (ms): {walls}", file=sys.stderr, flush=True)
        ms = gpu_time(k2, 5, warm=1)
        M = res["r"]["m"]
        ns = min(N, 4_000_000)
        hp, hc = pts[:ns].cpu().numpy().astype(np.float64), rgb[:ns].cpu().numpy()
        t0 = time.perf_counter()
        o = capi.voxel_downsample(hp, hc, voxel)
        cpu_s = time.perf_counter() - t0
        emit(f"K2 voxel_downsample v={voxel}", "points", N, ms, 15 * N + 27 * M,
             cpu={"points_per_s": ns / cpu_s, "cores": 1, "sample": f"first {ns} points, serial hash (Open3D's is serial)",
                  "voxels": int(len(o["points"]))},
             voxels=M, note="wall of the whole call incl. bounds, insert, collect, finalise and 2 host syncs")

    # K2 sorted output + K3 SOR on the v=0.01 result
    ds = ctx.voxel_downsample(pts, rgb, 0.01, sorted_output=True, want_idx=False)
    P64 = ds["points"].contiguous()
    M = P64.shape[0]

    if want("K3"):
        def k3(i):
            ctx.statistical_outlier(P64, 20, 2.0)
        ms = gpu_time(k3, 3, warm=1)
        ns = min(M, 300_000)
        hp = P64[:ns].cpu().numpy()
        t0 = time.perf_counter()
        capi.statistical_outlier(hp, 20, 2.0)
        cpu_s = time.perf_counter() - t0
        emit("K3 statistical_outlier nb=20", "points", M, ms, 24 * M + M,
             cpu={"points_per_s": ns / cpu_s, "cores": cores, "sample": f"first {ns} of the downsampled points"},
             uncached_upper_bound_bytes=24 * M * 21)

    # K7 normals on the same cloud (f32)
    if want("K7"):
        P32 = P64.to(torch.float32).contiguous()

        def k7(i):
            ctx.estimate_normals(P32, 30)
        ms = gpu_time(k7, 3, warm=1)
        ns = min(M, 300_000)
        hp = P32[:ns].cpu().numpy()
        t0 = time.perf_counter()
        capi.estimate_normals(hp, 30)
        cpu_s = time.perf_counter() - t0
        emit("K7 estimate_normals knn=30", "points", M, ms, 24 * M,
             cpu={"points_per_s": ns / cpu_s, "cores": cores, "sample": f"first {ns} points"})
    del pts, rgb
    if only and not (only & {"K4", "K6", "K8", "K9", "K10"}):
        return

    # ------------------------------------------------------------------ K4 / K6 on a fused volume
    vol = TSDFVolume(0.01, 0.04, block_capacity=120_000, ctx=ctx)
    views = vol.make_frame_views([depth[i] for i in range(NF)], [bgr[i] for i in range(NF)], [K4] * NF, poses)
    vol.integrate_sequence(views, NF, H, W, 32, False, 1.0, 5.0)
    nb = vol.num_blocks
    cap = nb * 96
    xyz = torch.empty((cap, 3), dtype=torch.float32, device=dev)
    nrm = torch.empty((cap, 3), dtype=torch.float32, device=dev)
    crgb = torch.empty((cap, 3), dtype=torch.uint8, device=dev)
    import ctypes as C
    from textureless_3d_reconstruction_b200.runtime import _ptr, _stream
    from textureless_3d_reconstruction_b200._lib import check
    n_out = torch.zeros(1, dtype=torch.int64, device=dev)

    def k6(i):
        check(vol.lib.t3d_tsdf_extract_points(vol.handle, 1.0, _ptr(xyz), _ptr(nrm), _ptr(crgb), cap, _ptr(n_out),
                                              _stream()))
    ms = gpu_time(k6, 10)
    npts = int(n_out.item())
    assert npts <= cap
    ov = capi.TSDFVolume(0.01, 0.04)
    nf_cpu = 4
    for i in range(nf_cpu):
        ov.integrate(depth[i].cpu().numpy(), bgr[i].cpu().numpy(), K4, poses[i], 1.0, 5.0)
    t0 = time.perf_counter()
    op, _, _ = ov.extract_points(1.0)
    cpu_s = time.perf_counter() - t0
    emit("K6 extract_points thr=1", "blocks", nb, ms, 20 * 512 * nb + 27 * npts,
         cpu={"blocks_per_s": ov.num_blocks / cpu_s, "cores": cores, "sample": f"{ov.num_blocks} blocks ({nf_cpu} frames fused)"},
         surface_points=npts)

    # ------------------------------------------------------------------ K10 triangle mesh
    if want("K10"):
        cnt = torch.zeros(2, dtype=torch.int64, device=dev)
        check(vol.lib.t3d_tsdf_extract_mesh(vol.handle, 1.0, None, None, None, 0, None, 0, _ptr(cnt), _stream()))
        nv, nt = cnt.tolist()
        mxyz = torch.empty((nv, 3), dtype=torch.float32, device=dev)
        mnrm = torch.empty((nv, 3), dtype=torch.float32, device=dev)
        mrgb = torch.empty((nv, 3), dtype=torch.uint8, device=dev)
        mtri = torch.empty((nt, 3), dtype=torch.int32, device=dev)

        def k10(i):
            check(vol.lib.t3d_tsdf_extract_mesh(vol.handle, 1.0, _ptr(mxyz), _ptr(mnrm), _ptr(mrgb), nv, _ptr(mtri), nt,
                                                _ptr(cnt), _stream()))
        ms = gpu_time(k10, 10)
        t0 = time.perf_counter()
        om = ov.extract_mesh(1.0)
        cpu_s = time.perf_counter() - t0
        emit("K10 extract_mesh thr=1 (vertex + triangle kernels, incl. block-count D2H)", "blocks", nb, ms,
             8 * 512 * nb + 27 * nv + 12 * nt + 2 * 4 * 512 * nb,
             cpu={"blocks_per_s": ov.num_blocks / cpu_s, "cores": 1, "sample": f"{ov.num_blocks} blocks ({nf_cpu} frames fused), "
                  f"{len(om[0])} vertices, {len(om[3])} triangles"},
             vertices=nv, triangles=nt)
        del mxyz, mnrm, mrgb, mtri

    def k4(i):
        vol.touch(depth[i % NF], K4, poses[i % NF], 1.0, 5.0)
    ms = gpu_time(k4, 20)
    t0 = time.perf_counter()
    keys = ov.touch(h_depth0, K4, poses[0], 1.0, 5.0)
    cpu_s = time.perf_counter() - t0
    emit("K4 touch (1 frame, export variant incl. count D2H)", "frames", 1, ms, 4 * (H // 4) * (W // 4) + 12 * len(keys),
         cpu={"frames_per_s": 1 / cpu_s, "cores": cores})

    # ------------------------------------------------------------------ K8 ICP (frame-to-model)
    tgt, tgt_n = xyz[:npts].contiguous(), nrm[:npts].contiguous()
    j = NF - 1
    src, _, n = ctx.backproject(depth[j], None, fx=K4[0], fy=K4[1], cx=K4[2], cy=K4[3], subsample=2, min_depth=0.1,
                                max_depth=5.0, pose=None)
    src = src[: int(n.item())].contiguous()
    T_wc = np.eye(4)
    T_wc[:3, :4] = poses[j]
    T_init = np.linalg.inv(T_wc)
    d = np.eye(4)
    d[:3, 3] = [0.004, -0.003, 0.005]
    res = {}

    def k8(i):
        res["r"] = ctx.icp_point_to_plane(src, tgt, tgt_n, 0.05, init=d @ T_init, max_iter=30)
    ms = gpu_time(k8, 3, warm=1)
    r = res["r"]
    its = max(r.iterations, 1)
    ns = min(src.shape[0], 130_000)
    t0 = time.perf_counter()
    o = capi.icp_point_to_plane(src[:ns].cpu().numpy(), tgt.cpu().numpy(), tgt_n.cpu().numpy(), 0.05, T0=d @ T_init,
                                max_iter=30)
    cpu_s = time.perf_counter() - t0
    emit("K8 icp_point_to_plane (whole registration)", "source points x iterations", src.shape[0] * (its + 1), ms,
         36 * r.correspondences * (its + 1) + 12 * src.shape[0] * (its + 1),
         cpu={"src_points_x_iters_per_s": ns * (o["iterations"] + 1) / cpu_s, "cores": cores,
              "sample": f"first {ns} source points, {o['iterations']} iterations, same target ({npts} pts)"},
         iterations=r.iterations, fitness=r.fitness, rmse=r.inlier_rmse, n_src=int(src.shape[0]), n_tgt=npts,
         pose_err_m=float(np.linalg.norm((r.transformation @ T_wc)[:3, 3])),
         note="includes target grid build once + per-iteration 29-double D2H and host 6x6 solve")

    # ------------------------------------------------------------------ K9 PLY
    from textureless_3d_reconstruction_b200.runtime import write_ply
    hp = tgt.cpu().numpy().astype(np.float64)
    hc = crgb[:npts].cpu().numpy()
    for layout, nm, n_w in ((0, "o3d-binary", npts), (1, "ref-ascii", min(npts, 1_000_000))):
        t0 = time.perf_counter()
        write_ply("/tmp/_t3d_bench.ply", hp[:n_w], hc[:n_w], layout=layout)
        s = time.perf_counter() - t0
        sz = Path("/tmp/_t3d_bench.ply").stat().st_size
        print(json.dumps({"kernel": f"K9 write_ply {nm} (host)", "points": n_w, "seconds": s, "MB_per_s": sz / s / 1e6,
                          "points_per_s": n_w / s}), flush=True)


if __name__ == "__main__":
    main()
