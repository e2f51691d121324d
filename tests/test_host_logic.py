"""Host-side mirror of the reference interface (no GPU): config defaults, intrinsics JSON,
depth-file naming/format rules, NumPy-promotion rule, CLI flags — against values the
reference itself produced (tests/golden/intrinsics_config.json)."""
import json

import numpy as np
import pytest

from conftest import GOLDEN
from textureless_3d_reconstruction_b200 import depth_enhanced_reconstruction as der
from textureless_3d_reconstruction_b200 import depth_processor as dp
from textureless_3d_reconstruction_b200 import depth_to_reconstruction as d2r

G = json.loads((GOLDEN / "intrinsics_config.json").read_text())


def test_reconstruction_config_defaults_match_reference():
    c = d2r.ReconstructionConfig()
    for k, v in G["config"].items():
        if k == "K":
            assert np.array_equal(c.K, np.array(v)) and c.K.dtype == np.float64
        else:
            assert getattr(c, k) == v, k


def test_intrinsics_from_json_matches_reference(tmp_path):
    for case in G["from_json"]:
        p = tmp_path / "i.json"
        p.write_text(json.dumps(case["input"]))
        ci = dp.CameraIntrinsics.from_json(str(p))
        for k in ("fx", "fy", "cx", "cy", "width", "height", "depth_scale"):
            assert getattr(ci, k) == case[k], (case["input"], k)
        assert np.array_equal(ci.to_matrix(), np.array(case["K"]))
    p = tmp_path / "bad.json"
    p.write_text(json.dumps(dict(fx=1.0, fy=1.0, cx=2.0, cy=2.0)))
    with pytest.raises(KeyError):                       # width/height are required (dp:98-101)
        dp.CameraIntrinsics.from_json(str(p))
    d = dp.CameraIntrinsics.default(800, 600)
    assert {k: getattr(d, k) for k in G["default"]} == G["default"]
    r = dp.CameraIntrinsics.realsense_d455()
    assert {k: getattr(r, k) for k in G["realsense"]} == G["realsense"]
    K = np.array([[10.0, 0, 3], [0, 11, 4], [0, 0, 1]])
    ci = der.CameraIntrinsics.from_matrix(K, 7, 9)
    assert (ci.fx, ci.fy, ci.cx, ci.cy, ci.width, ci.height) == (10.0, 11.0, 3.0, 4.0, 7, 9)
    assert np.array_equal(ci.to_matrix(), K)


def test_depth_loader_rules(tmp_path):
    depth = np.linspace(0.2, 4.0, 12, dtype=np.float64).reshape(3, 4)
    np.save(tmp_path / "a_depth.npy", depth)
    np.save(tmp_path / "a.npy", depth * 2)
    np.save(tmp_path / "depth_b.npy", depth * 3)
    # search order d2r:105-112 — {stem}_depth.npy wins over {stem}.npy
    assert d2r.DepthImageLoader.find_matching_depth("a.jpg", tmp_path).name == "a_depth.npy"
    assert d2r.DepthImageLoader.find_matching_depth("b.png", tmp_path).name == "depth_b.npy"
    assert d2r.DepthImageLoader.find_matching_depth("zzz.png", tmp_path) is None
    got = d2r.DepthImageLoader.load_depth(tmp_path / "a_depth.npy")
    assert got.dtype == np.float32 and np.array_equal(got, depth.astype(np.float32))
    assert d2r.DepthImageLoader.load_depth(tmp_path / "x.tiff") is None        # unknown suffix (d2r:97)
    cv2 = pytest.importorskip("cv2")
    mm = (depth * 1000).astype(np.uint16)                                       # dp:919-921 writer format
    cv2.imwrite(str(tmp_path / "c_depth.png"), mm)
    got = d2r.DepthImageLoader.load_depth(tmp_path / "c_depth.png")
    assert got.dtype == np.float32 and np.array_equal(got, mm.astype(np.float32) / 1000.0)
    assert [p.format(stem="s") for p in G["depth_name_patterns"]] == [
        "s_depth.npy", "s_depth.png", "s.npy", "s.png", "depth_s.npy", "depth_s.png"]


def test_scale_promotion_rule():
    d = np.ones((2, 2), np.float32)
    for scale in (1.0, 2, 1.37, np.float32(1.37), np.float64(1.37), np.median([1.0, 2.0]), np.array(1.5)):
        assert d2r._scale_is_f64(scale) == ((d * scale).dtype == np.float64), repr(scale)


def test_cli_flags_match_reference(monkeypatch, capsys):
    import argparse
    seen = {}

    class Stop(Exception):
        pass

    def fake_parse(self, argv=None):
        seen["opts"] = {a.option_strings[0]: a.default for a in self._actions if a.option_strings}
        raise Stop

    monkeypatch.setattr(argparse.ArgumentParser, "parse_args", fake_parse)
    with pytest.raises(Stop):
        d2r.main([])
    o = seen["opts"]
    ref = {"--rgb-folder": None, "--depth-folder": None, "--output": "./output/reconstruction.ply", "--fx": 1719.0,
           "--fy": 1719.0, "--cx": 540.0, "--cy": 960.0, "--voxel-size": 0.005, "--subsample": 2, "--no-vis": False}
    for k, v in ref.items():                                                     # d2r:771-784
        assert o[k] == v, k
    with pytest.raises(Stop):
        der.main([])
    o = seen["opts"]
    ref = {"--input": "./input_folder/buddha_images", "--output": "./output", "--fx": 1719.0, "--fy": 1719.0,
           "--cx": 540.0, "--cy": 960.0, "--no-depth": False, "--no-hybrid": False}
    for k, v in ref.items():                                                     # der:1421-1431
        assert o[k] == v, k


def test_load_poses_and_empty_paths(tmp_path, capsys):
    T = np.tile(np.eye(4), (3, 1, 1))
    T[1, :3, 3] = [1, 2, 3]
    np.save(tmp_path / "p.npy", T)
    poses = d2r.load_poses(tmp_path / "p.npy")
    assert len(poses) == 3 and poses[1][1].shape == (3, 1) and np.array_equal(poses[1][1].ravel(), [1, 2, 3])
    (tmp_path / "p.json").write_text(json.dumps(T.tolist()))
    assert np.array_equal(d2r.load_poses(tmp_path / "p.json")[1][1].ravel(), [1, 2, 3])
    pipe = d2r.DepthToReconstructionPipeline()
    assert pipe.reconstruct() == (None, None, None)                              # d2r:484-486
    assert "Need at least 2 images" in capsys.readouterr().out
    pipe.save_reconstruction(np.zeros((0, 3)), np.zeros((0, 3)), str(tmp_path / "x.ply"))
    assert "No points to save" in capsys.readouterr().out and not (tmp_path / "x.ply").exists()
    dense = d2r.DenseReconstructor(d2r.ReconstructionConfig())
    p, c = dense.merge_pointclouds([(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint8))])
    assert p.shape == (0,) and c.shape == (0,)                                   # d2r:399
    assert dense.estimate_scale(np.zeros((0, 3)), np.zeros((0, 2)), np.ones((4, 4))) == 1.0


def test_ply_writers_host_only(tmp_path):
    """K9 is host code in the C ABI: runs without a GPU; ASCII layout byte-exact vs the reference's file."""
    from textureless_3d_reconstruction_b200 import _lib
    from textureless_3d_reconstruction_b200.runtime import write_ply
    z = np.load(GOLDEN / "k9_ply_ascii.npz")
    f = tmp_path / "a.ply"
    write_ply(f, z["points"], z["colors"], layout=_lib.PLY_REF_ASCII)
    assert f.read_bytes() == bytes(z["ply_f32"])
    write_ply(f, z["points64"], z["colors"][:50], layout=_lib.PLY_REF_ASCII)
    assert f.read_bytes() == bytes(z["ply_f64"])
    vals = np.array([[0.0, -0.0, 1e-5], [1e16, 123456789012345678.0, 0.0001], [9.999e15, 1.5e16, 1e-7],
                     [np.float32(0.1), np.float32(16777216.0), -2.5e-323], [1e22, 1e23, 5e-324]], np.float64)
    cols = np.zeros((5, 3), np.uint8)
    write_ply(f, vals, cols, layout=_lib.PLY_REF_ASCII)
    body = f.read_text().split("end_header\n")[1].splitlines()
    for row, line in zip(vals, body):
        assert line == f"{row[0]} {row[1]} {row[2]} 0 0 0"                       # Python repr rules
    with pytest.raises(_lib.T3DError):
        write_ply(tmp_path / "no_such_dir" / "x.ply", vals, cols)


def test_ply_ascii_writer_threaded_path_is_byte_exact(tmp_path):
    """Above 200 k points the ASCII writer formats chunks on all host threads: the file must still be
    the reference's line-by-line output (`f"{x} {y} {z} {r} {g} {b}\\n"`, d2r:700-701), in order."""
    import time
    from textureless_3d_reconstruction_b200 import _lib
    from textureless_3d_reconstruction_b200.runtime import write_ply
    rng = np.random.default_rng(12)
    n = 450_001
    pts = (rng.normal(size=(n, 3)) * np.array([3.0, 1e-3, 250.0])).astype(np.float32)
    pts[::1000] = 0.0
    cols = rng.integers(0, 256, (n, 3), dtype=np.uint8)
    f = tmp_path / "big.ply"
    t0 = time.perf_counter()
    write_ply(f, pts, cols, layout=_lib.PLY_REF_ASCII)
    dt = time.perf_counter() - t0
    head, body = f.read_text().split("end_header\n")
    assert f"element vertex {n}" in head
    lines = body.split("\n")
    assert lines[-1] == "" and len(lines) == n + 1
    p64 = pts.astype(np.float64)
    for i in list(range(0, n, 997)) + [n - 1, 224_999, 225_000, 225_001]:
        x, y, z = (float(v) for v in p64[i])
        assert lines[i] == f"{x} {y} {z} {cols[i, 0]} {cols[i, 1]} {cols[i, 2]}", i
    # single-chunk reference: the same points written in two halves through the small-n path
    g = tmp_path / "half.ply"
    write_ply(g, pts[:150_000], cols[:150_000], layout=_lib.PLY_REF_ASCII)
    assert g.read_text().split("end_header\n")[1] == "\n".join(lines[:150_000]) + "\n"
    assert dt < 30
