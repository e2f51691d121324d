"""Per-source-line executed-instruction / stall-sample shares of one kernel launch.

  python profiles/sass_lines.py <nvdisasm -g -c dump of the kernel> <ncu --page source --csv of the launch> <source.cu>

ncu's source page exports per-SASS-instruction counters but no line numbers in CSV; `nvdisasm -g` of the same
cubin has the line table.  Both list the kernel's instructions in address order, so they are joined by position
(opcodes are cross-checked)."""
import collections
import csv
import re
import sys


def main(dis_path, csv_path, src_path, top=40):
    out, cur, started = [], None, False
    for ln in open(dis_path).read().splitlines():
        if re.match(r'\s*\.section\s', ln):
            if started:
                break
            started = True
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m:
            out.append((cur, m.group(2)))
    rows = list(csv.reader(open(csv_path)))
    hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
    H = rows[hi]
    ie, sm = H.index('Instructions Executed'), H.index('# Samples')
    sass = []
    for r in rows[hi + 1:]:
        if r and r[0] == 'Kernel Name':
            break
        try:
            sass.append((r[1].strip(), int(r[ie]), int(r[sm])))
        except (ValueError, IndexError):
            pass
    n = min(len(out), len(sass))
    mism = sum(1 for (loc, txt), (s, e, sp) in zip(out[:n], sass[:n]) if txt.split()[0].split('.')[0] not in s)
    agg, samp = collections.Counter(), collections.Counter()
    for (loc, txt), (s, e, sp) in zip(out[:n], sass[:n]):
        agg[loc] += e
        samp[loc] += sp
    tot, st = sum(agg.values()) or 1, sum(samp.values()) or 1
    print(f"# {len(out)} disassembled / {len(sass)} profiled instructions, {mism} opcode mismatches; {tot} warp instructions executed")
    print("# share of executed instructions | share of stall samples | file:line | source")
    src = open(src_path).read().splitlines()
    base = src_path.split('/')[-1]
    for loc, e in agg.most_common(top):
        f, l = loc if loc else ('?', 0)
        text = src[l - 1].strip()[:100] if (f == base or f.startswith('_icp')) and 0 < l <= len(src) else ''
        print(f"{e / tot:6.3f} {samp[loc] / st:6.3f} {f}:{l:<4d} {text}")


if __name__ == "__main__":
    main(*sys.argv[1:4])
