#!/bin/bash
# ncu --set full captures of the three hot kernels at the bench size (one launch each, after warm-up)
mkdir -p gpurun_out
for K in mg_cycle_kernel g2p_advect_kernel p2g_gather_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/full_$K \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/full_$K.log 2>&1
  echo "$K ncu exit $?"
done
ls -la gpurun_out/*.ncu-rep
