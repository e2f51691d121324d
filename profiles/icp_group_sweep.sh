#!/bin/bash
# ICP correspondence search variants (tuning): T3D_ICP_FUSE=1 linearisation fused into the search, 0 = two launches
python -m pytest tests -m gpu -x -q -k "icp_search" 2>&1 | tail -1
for f in ${FUSE_TO_RUN:-1 0}; do
  T3D_ICP_FUSE=$f python -m pytest tests -m gpu -x -q -k "icp or track" 2>&1 | tail -1
  T3D_ICP_FUSE=$f python bench.py --workload cfg3 --frames 120 > gpurun_out/icp_sweep_f$f.json 2> gpurun_out/icp_sweep_f$f.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/icp_sweep_f$f.json").read().strip().splitlines()[-1])
print("fuse=$f", round(d["value"]), d["ms_per_frame"], d["stage_ms_per_frame_synchronised"]["icp"], d["tracking"]["final_translation_error_m"], d["cpu_baseline"]["max_abs_pose_difference_gpu_vs_cpu"])
PY
done
