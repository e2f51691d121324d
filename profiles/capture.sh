#!/bin/bash
# Run on the GPU box (under gpurun) to refresh the evidence in profiles/.
# Each ncu pass runs only after the identical plain command exited 0.
#   usage: bash profiles/capture.sh <tag>        e.g. r1
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
# 1) launch list of the default bench command (device-resident leg only; shares, not absolutes)
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-secondary"
$CMD > $OUT/plain_${TAG}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"touch|integrate" -c 200 --csv \
    --log-file $OUT/launches_${TAG}.csv $CMD > $OUT/ncu_launches_${TAG}.log 2>&1
# 2) full-set capture of the dominant kernel (K5): the 5 launches (64-frame batches) of the first
#    timed step (launches 0-4 are the warm-up step), plus two K4 launches
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"integrate_kernel" -s 5 -c 5 \
    -o $OUT/prof_k5_${TAG} $CMD > $OUT/ncu_full_k5_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"touch_kernel" -s 6 -c 2 \
    -o $OUT/prof_k4_${TAG} $CMD > $OUT/ncu_full_k4_${TAG}.log 2>&1
