"""Print the key numbers of bench.py JSON lines found in the given log files."""
import json
import sys

for path in sys.argv[1:]:
    for line in open(path):
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        r = d.get("roofline") or {}
        tb = r.get("temporal_blocking", {})
        print(f"{path}: value={d['value']:.0f} {d['unit']} ms/step={d['ms_per_step']:.3f} "
              f"e2e={(d.get('e2e') or {}).get('value')} cpu={(d.get('cpu_baseline') or {}).get('value')} "
              f"K5_ms={r.get('avg_launch_ms')} achieved={r.get('achieved')} frac={r.get('frac')} "
              f"minbytes_GBs={tb.get('achieved_on_batched_min_bytes')} K4_ms={r.get('k4_touch_avg_launch_ms')} "
              f"clocks={d.get('clocks')}")
