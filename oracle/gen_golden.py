"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

Run here (where /root/reference exists):   python oracle/gen_golden.py
The GPU box has no /root/reference; tests only read the committed fixtures.

Covers SURVEY.md §8c (i): K1 over {subsample 1,2,4} x {python-float scale, np.float64 scale}
x {pose None, identity, rotated} with NaN/inf/<=min/>=max/threshold-exact pixels, for the three
reference copies (d2r, der, dp); the ASCII PLY fallback writer; intrinsics JSON handling;
the depth-file loader.
"""
from __future__ import annotations

import io
import json
import sys
import tempfile
from contextlib import redirect_stdout
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def make_inputs(H, W, seed):
    rng = np.random.default_rng(seed)
    depth = rng.uniform(0.05, 60.0, size=(H, W)).astype(np.float32)
    flat = depth.reshape(-1)
    n = flat.size
    special = [0.0, np.nan, np.inf, -np.inf, -1.0, 0.1, np.float32(0.1), np.nextafter(np.float32(0.1), np.float32(1)),
               np.nextafter(np.float32(0.1), np.float32(0)), 50.0, np.nextafter(np.float32(50), np.float32(0)),
               np.nextafter(np.float32(50), np.float32(100)), 100.0, np.nextafter(np.float32(100), np.float32(0)),
               1e-30, 3.4e38, 0.099999994, 0.10000001, 49.999996, 99.99999]
    pos = rng.choice(n, size=len(special) * 3, replace=False)
    for k, p in enumerate(pos):
        flat[p] = np.float32(special[k % len(special)])
    color = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
    return depth, color


def rot(yaw, pitch, roll):
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
    Rz = np.array([[cr, -sr, 0], [sr, cr, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def main():
    sys.path.insert(0, str(REF))
    with redirect_stdout(io.StringIO()):
        import depth_to_reconstruction as d2r
        import depth_enhanced_reconstruction as der
        # depth_processor.py references rclpy/sensor_msgs names (Node, Header, ...) at class-definition
        # time even when ROS 2 is absent; give each missing name a stub and retry.  The reference file
        # itself is not modified.
        import builtins
        import re
        stubs = []
        for _ in range(32):
            try:
                import depth_processor as dp
                break
            except NameError as e:
                name = re.search(r"name '(\w+)' is not defined", str(e)).group(1)
                setattr(builtins, name, type(name, (), {}))
                stubs.append(name)
                sys.modules.pop("depth_processor", None)
        for name in stubs:
            delattr(builtins, name)
    OUT.mkdir(parents=True, exist_ok=True)

    poses = {
        "none": None,
        "identity": (np.eye(3), np.zeros((3, 1))),
        "rotated": (rot(0.3, -0.2, 0.1), np.array([[0.4], [-1.3], [2.2]])),
    }
    scales = {"pyfloat1": 1.0, "pyfloat": 1.37, "npf64": np.float64(1.37), "npf64_1": np.float64(1.0)}
    cases = {}
    meta = []
    H, W = 37, 53  # odd sizes: ragged subsampled grids
    fx, fy, cx, cy = 61.5, 59.25, 26.5, 18.0
    depth, color = make_inputs(H, W, 1234)
    cases["depth"], cases["color"] = depth, color
    cfg = d2r.ReconstructionConfig(fx=fx, fy=fy, cx=cx, cy=cy)
    dense = d2r.DenseReconstructor(cfg)
    intr = der.CameraIntrinsics(fx=fx, fy=fy, cx=cx, cy=cy, width=W, height=H)
    gen = der.DensePointCloudGenerator(intr)
    k = 0
    for sub in (1, 2, 4):
        for sname, sc in scales.items():
            for pname, pose in poses.items():
                pts, cols = dense.depth_to_pointcloud(depth, color, pose=pose, scale=sc, subsample=sub)
                cases[f"d2r_{k}_pts"], cases[f"d2r_{k}_cols"] = pts, cols
                meta.append(dict(kind="d2r", id=k, subsample=sub, scale=sname, pose=pname))
                k += 1
    for sub in (1, 2, 4):
        for pname, pose in poses.items():
            for dt in ("f32", "f64"):
                dd = depth if dt == "f32" else depth * np.float64(0.83)  # der:1135 passes an f64 array
                pts, cols = gen.depth_to_pointcloud(dd, color, pose=pose, subsample=sub)
                cases[f"der_{k}_pts"], cases[f"der_{k}_cols"] = pts, cols
                meta.append(dict(kind="der", id=k, subsample=sub, pose=pname, depth=dt))
                k += 1
    # der with a depth whose shape differs from the intrinsics (recomputes maps, der:575-578)
    d2, c2 = make_inputs(20, 31, 99)
    cases["depth_b"], cases["color_b"] = d2, c2
    pts, cols = gen.depth_to_pointcloud(d2, c2, pose=poses["rotated"], subsample=2)
    cases[f"der_{k}_pts"], cases[f"der_{k}_cols"] = pts, cols
    meta.append(dict(kind="der_b", id=k, subsample=2, pose="rotated", depth="f32"))
    k += 1
    for ds in (1, 2, 3):
        ii = dp.CameraIntrinsics(fx=fx, fy=fy, cx=cx, cy=cy, width=W, height=H)
        g = dp.PointCloudGenerator(ii, downsample_factor=ds)
        for with_rgb in (True, False):
            pts, cols = g.generate(depth, color if with_rgb else None, max_depth=20.0, min_depth=0.1)
            cases[f"dp_{k}_pts"] = pts
            if cols is not None:
                cases[f"dp_{k}_cols"] = cols
            meta.append(dict(kind="dp", id=k, downsample=ds, rgb=with_rgb))
            k += 1
    # all-invalid and empty-ish frames
    dz = np.zeros((8, 12), np.float32)
    cz = np.zeros((8, 12, 3), np.uint8)
    pts, cols = dense.depth_to_pointcloud(dz, cz, pose=poses["rotated"], scale=1.0, subsample=1)
    cases[f"d2r_{k}_pts"], cases[f"d2r_{k}_cols"] = pts, cols
    meta.append(dict(kind="d2r_zero", id=k, subsample=1, scale="pyfloat1", pose="rotated"))
    k += 1
    cases["pose_rotated_R"], cases["pose_rotated_t"] = poses["rotated"]
    cases["intrinsics"] = np.array([fx, fy, cx, cy])
    cases["meta_json"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(OUT / "k1_backproject.npz", **cases)

    # merge (Open3D absent -> vstack) + ASCII PLY fallback writer
    pipe = d2r.DepthToReconstructionPipeline(cfg)
    p1, c1 = dense.depth_to_pointcloud(depth, color, pose=poses["rotated"], scale=1.0, subsample=4)
    p2, c2_ = dense.depth_to_pointcloud(depth, color, pose=poses["identity"], scale=1.0, subsample=4)
    with redirect_stdout(io.StringIO()):
        mp, mc = dense.merge_pointclouds([(p1, c1), (np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint8)), (p2, c2_)])
        with tempfile.TemporaryDirectory() as td:
            f = Path(td) / "sub" / "x.ply"
            pipe.save_reconstruction(mp[:200], mc[:200], str(f))
            ply_f32 = f.read_bytes()
            pipe.save_reconstruction(mp[:50].astype(np.float64) * 1e-5, mc[:50], str(f))
            ply_f64 = f.read_bytes()
    np.savez_compressed(OUT / "k9_ply_ascii.npz", points=mp[:200], colors=mc[:200],
                        ply_f32=np.frombuffer(ply_f32, np.uint8), points64=mp[:50].astype(np.float64) * 1e-5,
                        ply_f64=np.frombuffer(ply_f64, np.uint8), merged_n=np.array([len(mp)]))

    # intrinsics JSON (dp:89-102) + config defaults (d2r:45-73)
    jcases = [
        dict(width=640, height=480),
        dict(width=640, height=480, fx=500.0, fy=501.0, cx=320.5, cy=240.5, depth_scale=0.001),
        dict(width=1080, height=1920, focal_length_x=1719.0, focal_length_y=1718.0),
        dict(width=100, height=50, principal_point_x=49.0, principal_point_y=26.0, fx=80.0),
    ]
    jout = []
    with tempfile.TemporaryDirectory() as td:
        for j in jcases:
            p = Path(td) / "i.json"
            p.write_text(json.dumps(j))
            ci = dp.CameraIntrinsics.from_json(str(p))
            jout.append(dict(input=j, fx=ci.fx, fy=ci.fy, cx=ci.cx, cy=ci.cy, width=ci.width, height=ci.height,
                             depth_scale=ci.depth_scale, K=ci.to_matrix().tolist()))
    dflt = dp.CameraIntrinsics.default(800, 600)
    rs = dp.CameraIntrinsics.realsense_d455()
    c0 = d2r.ReconstructionConfig()
    golden = dict(
        from_json=jout,
        default=dict(fx=dflt.fx, fy=dflt.fy, cx=dflt.cx, cy=dflt.cy, width=dflt.width, height=dflt.height),
        realsense=dict(fx=rs.fx, fy=rs.fy, cx=rs.cx, cy=rs.cy, width=rs.width, height=rs.height, depth_scale=rs.depth_scale),
        config=dict(fx=c0.fx, fy=c0.fy, cx=c0.cx, cy=c0.cy, depth_scale=c0.depth_scale, min_depth=c0.min_depth,
                    max_depth=c0.max_depth, match_ratio=c0.match_ratio, ransac_threshold=c0.ransac_threshold,
                    voxel_size=c0.voxel_size, subsample_factor=c0.subsample_factor, K=c0.K.tolist()),
        depth_name_patterns=["{stem}_depth.npy", "{stem}_depth.png", "{stem}.npy", "{stem}.png",
                             "depth_{stem}.npy", "depth_{stem}.png"],
    )
    (OUT / "intrinsics_config.json").write_text(json.dumps(golden, indent=1))
    gen_formats(d2r, der)
    print("wrote", sorted(p.name for p in OUT.iterdir()))


def gen_formats(d2r, der):
    """tests/golden/formats.npz — SURVEY 8f rows: depth file formats (dp:905-921 writer, d2r:80-97
    reader, run through cv2 / NumPy exactly as the reference does), depth->RGB resize (cv2.resize
    INTER_LINEAR, d2r:465-467), depth-scale estimation (the reference's own two functions), and the
    PointCloud2 colour packing (statements of dp:750-756 executed verbatim; the enclosing class needs ROS)."""
    import cv2
    rng = np.random.default_rng(77)
    out = {}
    # 16-bit millimetre depth: writer then reader, through real PNG files
    depth = rng.uniform(0.0, 70.0, size=(41, 57)).astype(np.float32)
    flat = depth.reshape(-1)
    for k, v in enumerate([0.0, 65.535, 65.5351, 65.536, 65.5365, 131.072, 1e-4, 0.0005, 0.001, 0.0019999, np.nan,
                           np.inf, -np.inf, -0.001, -1.0, -65.536, 3e9, -3e9, 2147483.5, 2147483.9]):
        flat[k * 7] = np.float32(v)
    with np.errstate(invalid="ignore"):
        depth_mm = (depth * 1000).astype(np.uint16)                                    # dp:920
    with tempfile.TemporaryDirectory() as td:
        f = Path(td) / "a_depth.png"
        cv2.imwrite(str(f), depth_mm)                                                    # dp:921
        back = d2r.DepthImageLoader.load_depth(f)                                        # d2r:85-90
        np.save(Path(td) / "a_depth.npy", depth)                                         # dp:908
        back_npy = d2r.DepthImageLoader.load_depth(Path(td) / "a_depth.npy")
    out.update(depth=depth, depth_mm=depth_mm, depth_back=back, depth_back_npy=back_npy)
    # cv2.resize INTER_LINEAR (d2r:465-467): depth -> RGB size
    src = rng.uniform(0.2, 9.0, size=(37, 53)).astype(np.float32)
    src[5, 7] = 0.0
    rs = []
    for i, (dh, dw) in enumerate([(74, 106), (100, 91), (37, 53), (18, 26), (19, 27), (120, 160), (1, 1), (3, 200)]):
        out[f"resize_{i}"] = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
        rs.append([dh, dw])
    src2 = rng.uniform(0.2, 9.0, size=(120, 160)).astype(np.float32)
    out["resize_src2_half"] = cv2.resize(src2, (80, 60), interpolation=cv2.INTER_LINEAR)  # exact 2x: cv2's area fast path
    out.update(resize_src=src, resize_shapes=np.array(rs), resize_src2=src2)
    # depth-scale estimation: the reference's two functions
    dm = rng.uniform(0.0, 4.0, size=(48, 64)).astype(np.float32)
    dm[rng.uniform(size=dm.shape) < 0.1] = 0.0
    cfg = d2r.ReconstructionConfig()
    dense = d2r.DenseReconstructor(cfg)
    sc = []
    for i, n in enumerate([0, 2, 3, 4, 5, 40, 41, 300]):
        p3 = rng.uniform(-2.0, 6.0, size=(n, 3))
        p2 = np.column_stack([rng.uniform(-6.0, 70.0, n), rng.uniform(-6.0, 54.0, n)])
        if n >= 40:
            p3[0, 2] = 1e5                                                               # ratio > 1000: gated in d2r only
            p3[1, 2] = 1e-6
            p2[2] = [-0.5, -0.7]                                                         # int() truncates toward 0 -> pixel (0,0)
        with redirect_stdout(io.StringIO()):
            s_d2r = dense.estimate_scale(p3, p2, dm)
            s_der = der.DepthScaleEstimator.estimate_scale(p3, p2, dm, cfg.K)
        out[f"scale_{i}_p3"], out[f"scale_{i}_p2"] = p3, p2
        sc.append([float(s_d2r), float(s_der)])
    out.update(scale_depth=dm, scale_expected=np.array(sc))
    # PointCloud2 colour packing, dp:750-756
    pts = rng.uniform(-3, 3, size=(300, 3)).astype(np.float32)
    cols = (rng.integers(0, 256, size=(300, 3)).astype(np.float32) / 255.0)[:, ::-1].copy()   # dp:417-420 output format
    rgb_packed = np.zeros(len(pts), dtype=np.float32)
    for i in range(len(pts)):
        r, g, b = (cols[i] * 255).astype(np.uint8)
        rgb_packed[i] = np.frombuffer(np.array([b, g, r, 0], dtype=np.uint8).tobytes(), dtype=np.float32)[0]
    out.update(pc2_points=pts, pc2_colors=cols, pc2_cloud=np.column_stack([pts, rgb_packed]).view(np.uint32))
    np.savez_compressed(OUT / "formats.npz", **out)


if __name__ == "__main__":
    main()
