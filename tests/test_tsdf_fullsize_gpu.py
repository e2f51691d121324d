"""K4/K5 parity at the BASELINE sizes (cfg 2: 1080x1920 / 1 cm / trunc 4 cm; cfg 5: 2160x3840 / 5 mm /
trunc 2 cm) against BOTH oracle formulations of R5:

* ``o_tsdf_integrate``          - the kernel-ordered arithmetic (scaled rotation, fmaf chains, reciprocals):
                                  the GPU must agree bit for bit (keys, integer weights, tsdf, colours, counters);
* ``o_tsdf_integrate_literal``  - SURVEY 8c R5 transcribed term by term (true divisions, roundf, no FMA):
                                  the independent yardstick.  The two formulations can only decide differently
                                  where u / v sits on a rounding or image-border boundary (or zc / sdf on a
                                  test threshold); the literal oracle marks exactly those voxels.  Everywhere
                                  else weights must be identical and tsdf within north_star's 1e-4.

The deviation numbers are printed and written to gpurun_out/parity_fullsize.json (copied into DESIGN.md 3).
"""
import json
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent


def by_key(keys, *arrs):
    order = np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))
    return (keys[order],) + tuple(a[order] for a in arrs)


def record(name, obj):
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    p = out / "parity_fullsize.json"
    try:
        d = json.loads(p.read_text())
    except Exception:  # noqa: BLE001
        d = {}
    d[name] = obj
    p.write_text(json.dumps(d, indent=1))


def run_case(ctx, oracle, name, H, W, K, voxel, trunc, frame_ids, capacity, tsdf_tol):
    import torch
    from textureless_3d_reconstruction_b200.runtime import TSDFVolume
    fr = [ctx.synth_frame(0, i, H, W, *K, seed=1234, noise_sigma=0.002) for i in frame_ids]
    ds, cs, Ts = [f[0] for f in fr], [f[1] for f in fr], [f[2] for f in fr]
    vol = TSDFVolume(voxel, trunc, block_capacity=capacity, ctx=ctx)
    vol.integrate_batch(ds, cs, K, Ts, 1.0, 5.0)
    torch.cuda.synchronize()
    ko = oracle.TSDFVolume(voxel, trunc)          # kernel-ordered R5
    lo = oracle.TSDFVolume(voxel, trunc)          # literal R5
    for d, c, T in fr:
        dn, cn = d.cpu().numpy(), c.cpu().numpy()
        keys = ko.integrate(dn, cn, K, T, 1.0, 5.0)
        lo.integrate(dn, cn, K, T, 1.0, 5.0, keys=keys, literal=True)
    # ---- GPU vs kernel-ordered oracle: bit for bit
    assert vol.num_blocks == ko.num_blocks == lo.num_blocks
    assert vol.counters() == ko.counters()
    g = by_key(*[x.cpu().numpy() for x in vol.export_blocks()])
    k = by_key(*ko.export())
    assert np.array_equal(g[0], k[0])                                   # block key set
    assert np.array_equal(g[2], k[2])                                   # occupancy + integer weights
    assert np.array_equal(g[1].view(np.uint32), k[1].view(np.uint32))   # tsdf
    assert np.array_equal(g[3].view(np.uint32), k[3].view(np.uint32))   # colours
    # ---- GPU vs literal oracle: identical except the formulation-sensitive voxels it marks
    lk, lt, lw, lc = lo.export()
    flags, n_pairs = lo.export_flags()
    lk, lt, lw, lc, flags = by_key(lk, lt, lw, lc, flags)
    assert np.array_equal(g[0], lk)
    sens = flags != 0
    dw = g[2] != lw
    assert not (dw & ~sens).any(), "weights differ on a voxel the literal oracle did not mark as boundary"
    occ = (g[2] > 0) != (lw > 0)
    clean = ~sens
    dt = np.abs(g[1] - lt)
    dc = np.abs(g[3] - lc)
    upd = int(ko.counters()["voxel_updates"])
    res = {
        "frames": list(map(int, frame_ids)), "H": H, "W": W, "voxel": voxel, "trunc": trunc,
        "blocks": int(vol.num_blocks), "voxel_updates": upd,
        "voxel_updates_literal": int(lo.counters()["voxel_updates"]),
        "sensitive_voxel_frame_pairs": int(n_pairs), "sensitive_voxels": int(sens.sum()),
        "sensitive_by_pixel": int(((flags & 1) != 0).sum()), "sensitive_by_threshold": int(((flags & 2) != 0).sum()),
        "sensitive_fraction_of_updates": n_pairs / max(upd, 1),
        "weights_differ": int(dw.sum()), "occupancy_differs": int(occ.sum()),
        "tsdf_maxdiff_clean_voxels": float(dt[clean].max()), "tsdf_maxdiff_sensitive_voxels": float(dt[sens].max()) if sens.any() else 0.0,
        "rgb_maxdiff_clean_voxels": float(dc[clean[..., None].repeat(3, -1)].max()),
        "gpu_equals_kernel_ordered_oracle": "bit-exact (keys, weights, tsdf, rgb, counters)",
    }
    print(name, json.dumps(res))
    record(name, res)
    assert res["sensitive_fraction_of_updates"] < 2e-3
    assert res["tsdf_maxdiff_clean_voxels"] <= tsdf_tol
    assert res["rgb_maxdiff_clean_voxels"] <= 1e-2
    return res


def test_cfg2_full_size_vs_kernel_ordered_and_literal_oracles(ctx, oracle):
    """8 full 1080x1920 cfg-2 frames (the first 8 of the bench workload)."""
    run_case(ctx, oracle, "cfg2_frames_0_7", 1920, 1080, (1719.0, 1719.0, 540.0, 960.0), 0.01, 0.04,
             range(8), 120_000, tsdf_tol=1e-4)


def test_cfg2_far_from_origin_deviation(ctx, oracle):
    """The last 8 frames of the 300-frame workload sit 73-75 m from the origin: an f32 rigid transform
    cancels ~75 m against ~75 m, so ANY two f32 formulations of R5 differ by a few ulp(75 m) = 7.6e-6 m
    in zc, i.e. ~2e-4 in sdf / 0.04.  Measured and bounded here (not a property of the kernel: the
    GPU still equals the kernel-ordered oracle bit for bit)."""
    run_case(ctx, oracle, "cfg2_frames_292_299", 1920, 1080, (1719.0, 1719.0, 540.0, 960.0), 0.01, 0.04,
             range(292, 300), 120_000, tsdf_tol=1e-3)


def test_cfg5_full_size_vs_kernel_ordered_and_literal_oracles(ctx, oracle):
    """2 full 2160x3840 cfg-5 frames, 5 mm voxels, trunc 2 cm."""
    run_case(ctx, oracle, "cfg5_frames_0_1", 3840, 2160, (3438.0, 3438.0, 1080.0, 1920.0), 0.005, 0.02,
             range(2), 400_000, tsdf_tol=1e-4)
