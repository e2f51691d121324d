#!/bin/bash
# ncu --set full captures (one launch each) of every kernel besides K4/K5 (capture.sh does those).
#   usage: bash profiles/capture_kernels.sh <tag>
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
K="python profiles/bench_kernels.py --quick"
NCU="ncu --set full --clock-control none --import-source on"
$K --only K1,K2,K3,K6,K7,K8 > $OUT/plain_kernels_${TAG}.jsonl 2> $OUT/plain_kernels_${TAG}.err || exit 1
timeout 600 $NCU -k regex:"backproject_kernel" -s 60 -c 1 -o $OUT/prof_k1_${TAG} $K --only K1 > /dev/null 2>&1
timeout 600 $NCU -k regex:"voxel_keys|voxel_accum" -s 4 -c 2 -o $OUT/prof_k2_${TAG} $K --only K2 > /dev/null 2>&1
timeout 600 $NCU -k regex:"sor_mean" -s 1 -c 1 -o $OUT/prof_k3_${TAG} $K --only K3 > /dev/null 2>&1
timeout 600 $NCU -k regex:"normals_kernel" -s 1 -c 1 -o $OUT/prof_k7_${TAG} $K --only K7 > /dev/null 2>&1
timeout 600 $NCU -k regex:"extract_kernel" -s 2 -c 1 -o $OUT/prof_k6_${TAG} $K --only K6 > /dev/null 2>&1
timeout 600 $NCU -k regex:"icp_nn|icp_acc" -s 2 -c 2 -o $OUT/prof_k8_${TAG} $K --only K8 > /dev/null 2>&1
