# K5 tuning sweep (run under gpurun): T3D_K5_MODE / T3D_K5_VPT / T3D_K5_MINB variants of the default bench
T3D_K5_MODE=6 T3D_K5_MINB=7 python -m pytest tests/test_tsdf_gpu.py tests/test_tsdf_fullsize_gpu.py -m gpu -x -q > gpurun_out/r2f_tests_m6.log 2>&1; echo "mode 6 tests rc=$? $(tail -1 gpurun_out/r2f_tests_m6.log)"
for cfg in "0 4 8" "6 4 8" "6 4 7" "6 4 6"; do
  set -- $cfg
  T3D_K5_MODE=$1 T3D_K5_VPT=$2 T3D_K5_MINB=$3 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2f_bench_m$1_v$2_b$3.log 2>&1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2f_bench_m$1_v$2_b$3.log").read().strip().splitlines()[-1])
    print("mode $1 vpt $2 minb $3: %.0f frames/s  step %.3f ms  k5 launch %.4f ms" % (d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"]))
except Exception as e:
    print("mode $1 vpt $2 minb $3: FAILED", e)
PY
done
