"""Pin the oracle: oracle/ref_numpy.py and the C back-projection against vectors the
UNMODIFIED reference produced (oracle/gen_golden.py, tests/golden/*)."""
import json

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import ref_numpy as R

SCALES = {"pyfloat1": 1.0, "pyfloat": 1.37, "npf64": np.float64(1.37), "npf64_1": np.float64(1.0)}


def poses_of(z):
    return {"none": None, "identity": (np.eye(3), np.zeros((3, 1))),
            "rotated": (z["pose_rotated_R"], z["pose_rotated_t"])}


def test_ref_numpy_matches_reference_bit_for_bit(k1_golden):
    z, meta = k1_golden
    fx, fy, cx, cy = z["intrinsics"]
    poses = poses_of(z)
    n = 0
    for m in meta:
        k = m["id"]
        if m["kind"] in ("d2r", "d2r_zero"):
            depth, color = (z["depth"], z["color"]) if m["kind"] == "d2r" else (np.zeros((8, 12), np.float32), np.zeros((8, 12, 3), np.uint8))
            pts, cols = R.d2r_depth_to_pointcloud(depth, color, fx, fy, cx, cy, pose=poses[m["pose"]],
                                                  scale=SCALES[m["scale"]], subsample=m["subsample"])
            g_pts, g_cols = z[f"d2r_{k}_pts"], z[f"d2r_{k}_cols"]
        elif m["kind"] in ("der", "der_b"):
            depth, color = (z["depth"], z["color"]) if m["kind"] == "der" else (z["depth_b"], z["color_b"])
            if m["depth"] == "f64":
                depth = depth * np.float64(0.83)
            pts, cols = R.der_depth_to_pointcloud(depth, color, fx, fy, cx, cy, pose=poses[m["pose"]],
                                                  subsample=m["subsample"])
            g_pts, g_cols = z[f"der_{k}_pts"], z[f"der_{k}_cols"]
        else:
            pts, cols = R.dp_generate(z["depth"], z["color"] if m["rgb"] else None, fx, fy, cx, cy,
                                      downsample=m["downsample"], max_depth=20.0, min_depth=0.1)
            g_pts = z[f"dp_{k}_pts"]
            g_cols = z[f"dp_{k}_cols"] if m["rgb"] else None
        assert pts.dtype == g_pts.dtype and pts.shape == g_pts.shape, m
        assert np.array_equal(pts.view(np.uint32), g_pts.view(np.uint32)), m
        if g_cols is None:
            assert cols is None
        else:
            assert cols.dtype == g_cols.dtype and np.array_equal(cols, g_cols), m
        n += 1
    assert n == len(meta) and n > 50


def test_c_backproject_matches_reference(k1_golden, oracle):
    """The C restatement (CPU-baseline leg) against the reference vectors: identical masks
    and counts; coordinates within 1e-5 relative (its FMA chain is the dgemm order, so in
    practice bit-equal)."""
    z, meta = k1_golden
    fx, fy, cx, cy = z["intrinsics"]
    poses = poses_of(z)
    exact = total = 0
    for m in meta:
        if m["kind"] != "d2r":
            continue
        sc = SCALES[m["scale"]]
        pts, cols = oracle.backproject(z["depth"], z["color"], fx, fy, cx, cy, scale=float(sc),
                                       f64_mask=isinstance(sc, np.floating), pose=poses[m["pose"]],
                                       subsample=m["subsample"])
        g_pts, g_cols = z[f"d2r_{m['id']}_pts"], z[f"d2r_{m['id']}_cols"]
        assert pts.shape == g_pts.shape, m                      # mask / count exact
        assert np.array_equal(cols, g_cols), m                  # colours + order exact
        assert np.allclose(pts, g_pts, rtol=1e-5, atol=1e-12), m
        exact += int((pts.view(np.uint32) == g_pts.view(np.uint32)).sum())
        total += pts.size
    assert exact / total > 0.999


def test_ascii_ply_writer_restatement():
    z = np.load(GOLDEN / "k9_ply_ascii.npz")
    body = "".join(R.ascii_ply_lines(z["points"], z["colors"]))
    assert (R.ASCII_PLY_HEADER.format(n=200) + body).encode() == bytes(z["ply_f32"])
    body = "".join(R.ascii_ply_lines(z["points64"], z["colors"][:50]))
    assert (R.ASCII_PLY_HEADER.format(n=50) + body).encode() == bytes(z["ply_f64"])


def test_intrinsics_json_restatement():
    g = json.loads((GOLDEN / "intrinsics_config.json").read_text())
    for case in g["from_json"]:
        out = R.intrinsics_from_json_dict(case["input"])
        for key in ("fx", "fy", "cx", "cy", "width", "height", "depth_scale"):
            assert out[key] == case[key]
    with pytest.raises(KeyError):
        R.intrinsics_from_json_dict(dict(fx=1.0, fy=1.0, cx=1.0, cy=1.0))  # width/height required (dp:98)
