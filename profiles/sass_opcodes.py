"""Per-kernel counts of the SASS opcodes that prove (or disprove) Blackwell-specific paths in libt3d.so:
UBLKCP / UTMALDG / UTMASTG (TMA bulk / tensor copies), SYNCS (mbarrier), LDGSTS (cp.async), UTC*MMA / LDTM / STTM
(tcgen05 — none expected: nothing here is a dense contraction), HMMA (legacy tensor path — none expected),
FFMA2 / FADD2 / FMUL2 (packed f32x2), MUFU, ATOMG / REDG.   python profiles/sass_opcodes.py > profiles/sass_opcodes.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

so = Path(__file__).resolve().parent.parent / "textureless_3d_reconstruction_b200" / "libt3d.so"
out = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
pat = re.compile(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)")
WATCH = ["UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "HMMA", "FFMA2", "FADD2",
         "FMUL2", "MUFU", "ATOMG", "REDG", "ATOMS", "MATCH", "LDG", "STG", "LDS", "STS"]
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    if "Function :" in line:
        cur = line.split("Function :")[1].strip()
        counts[cur] = collections.Counter()
        continue
    m = pat.match(line)
    if m and cur:
        counts[cur][m.group(1)] += 1
        counts[cur]["_total"] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {so.name}: static instruction counts per kernel (sm_100a)")
print("# " + " ".join(f"{w:>7s}" for w in ["total"] + WATCH) + "  kernel")
agg = collections.Counter()
for (name, c), dn in zip(counts.items(), demangle):
    short = re.sub(r"\(anonymous namespace\)::", "", dn)
    short = re.sub(r"\(.*", "", short)[:70]
    print("  " + " ".join(f"{c[w]:7d}" for w in ["_total"] + WATCH) + "  " + short)
    agg.update(c)
print("# " + " ".join(f"{agg[w]:7d}" for w in ["_total"] + WATCH) + "  ALL KERNELS")
