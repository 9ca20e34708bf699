#!/bin/bash
# A/B of p2g_xrow_kernel staging capacities: rebuilds p2g.o with -DFB_XR_CAP=<n> on the GPU box. usage: gpurun -- 'bash tools/gpu_ab_p2g.sh 1024 1152 1344'
cd zeno_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Xcompiler -O2 --expt-relaxed-constexpr"
for CAP in "$@"; do
  nvcc $FLAGS -DFB_XR_CAP=$CAP $FB_EXTRA -c p2g.cu -o _build/p2g.o 2>/dev/null && nvcc -shared -o ../libflipb200.so _build/*.o -lcudart -ldl
  ( cd ../.. && timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().splitlines()[-1]); k=d['kernels']['p2g_gather']
print('CAP $CAP: p2g_gather %.1f us  step %.3f ms' % (k['avg_us'], d['ms_per_step']))" )
done
