#!/bin/bash
# ncu launch list of the tracked-frame loop (cfg 3): which kernels a frame is made of.
#   usage: bash profiles/capture_cfg3.sh <tag>
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --workload cfg3 --frames 24 --no-cpu"
$CMD > $OUT/plain_cfg3_${TAG}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file $OUT/launches_cfg3_${TAG}.csv $CMD > $OUT/ncu_launches_cfg3_${TAG}.log 2>&1
