"""Turn ncu outputs into the small text summaries committed under profiles/.

  python profiles/summarize_ncu.py launches gpurun_out/launches_r1.csv      > profiles/r1_launches.txt
  python profiles/summarize_ncu.py full     gpurun_out/prof_k4k5_r1.ncu-rep > profiles/r1_k4k5_full.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        # what a gather-bound kernel is bound BY: L1 wavefronts / sector lookups, L2 sector lookups, L2->L1 bytes
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum",
        "l1tex__t_sectors.sum", "l1tex__t_sectors_lookup_miss.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "lts__t_sectors.sum", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else v * 1e3 if r[ui] in ("ms", "msecond") else v
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({path}); cold-cache, serialised: compare shares")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / tot:6.3f} share  n={v[0]:4d}  avg={v[1] / v[0]:9.1f} us  total={v[1]:11.1f} us  {k[:110]}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none  ({path})")
    for r in rows[2:]:
        print("kernel:", r[h.index("Kernel Name")][:140])
        for k in KEYS:
            cols = [i for i, name in enumerate(h) if name == k or name.endswith("." + k)]   # some live under a section prefix
            cols = [i for i in cols if r[i] != ""]
            if cols:
                print(f"  {k:68s} {r[cols[0]]:>16s} {units[cols[0]]}")
        st = [(float(r[i] or 0), h[i]) for i in range(len(h)) if "pcsamp_warps_issue_stalled" in h[i] and not h[i].endswith("_not_issued")]
        tot = sum(v for v, _ in st) or 1.0
        for v, c in sorted(st, reverse=True)[:7]:
            print(f"  stall {v / tot:6.3f}  {c.replace('smsp__pcsamp_warps_issue_stalled_', '')}")


def traffic(path, kernel="integrate_kernel"):
    """JSON for bench.py's roofline.traffic: mean dram read+write bytes per launch of `kernel`, tagged with the
    sha256 of csrc/tsdf.cu as it is NOW (run this right after the capture): bench.py refuses to replay the
    number once the kernel source has changed."""
    import hashlib
    import json
    import pathlib
    src = pathlib.Path(__file__).resolve().parent.parent / "textureless_3d_reconstruction_b200" / "csrc" / "tsdf.cu"
    sha = hashlib.sha256(src.read_bytes()).hexdigest()
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = []
    for r in rows[2:]:
        if kernel not in r[h.index("Kernel Name")]:
            continue
        rd = float(r[h.index("dram__bytes_read.sum")]) * mult[units[h.index("dram__bytes_read.sum")]]
        wr = float(r[h.index("dram__bytes_write.sum")]) * mult[units[h.index("dram__bytes_write.sum")]]
        us = float(r[h.index("gpu__time_duration.sum")])
        per.append({"read": rd, "write": wr, "duration_" + units[h.index("gpu__time_duration.sum")]: us})
    n = max(len(per), 1)
    print(json.dumps({"kernel": kernel, "launches": len(per), "kernel_source_sha256": sha,
                      "dram_bytes_per_launch": sum(p["read"] + p["write"] for p in per) / n,
                      "dram_read_per_launch": sum(p["read"] for p in per) / n,
                      "dram_write_per_launch": sum(p["write"] for p in per) / n,
                      "source": f"ncu --set full --clock-control none, {path.split('/')[-1]}, mean of {len(per)} launches "
                                "of one 300-frame step (batches of <= 64 frames)",
                      "per_launch": per}, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])
