# K5 timing decomposition (run under gpurun): T3D_K5_DEBUG bit 0 = depth gathers read pixel 0, bit 1 = colour
# gathers read pixel 0, bit 2 = block state neither loaded nor stored (results are wrong on purpose)
for dbg in 0 1 2 3 4 7; do
  T3D_K5_DEBUG=$dbg python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2e_bench_dbg$dbg.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2e_bench_dbg$dbg.log").read().strip().splitlines()[-1])
    print("debug $dbg: %.0f frames/s  step %.3f ms  k5 launch %.4f ms" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"]))
except Exception as e:
    print("debug $dbg: FAILED", e)
PY
done
