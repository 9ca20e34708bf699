#!/bin/bash
# one ncu --set full capture of a kernel at the bench size. usage: gpurun -- 'bash tools/gpu_ncu_k.sh <kernel regex> <tag> [skip]'
K=$1; TAG=$2; SKIP=${3:-3}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o gpurun_out/${TAG} \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/${TAG}.ncu-rep
