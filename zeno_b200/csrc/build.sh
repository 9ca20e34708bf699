#!/bin/bash
# Builds zeno_b200/libflipb200.so for sm_100a (cross-compiles without a GPU).
# -fmad=false: the transfer kernels must execute the oracle's un-contracted op sequence
# (explicit __fmaf_rn is used where the reference has FMA intrinsics).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libflipb200.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -Xcompiler -O2 --expt-relaxed-constexpr -Xptxas -v"
mkdir -p _build
pids=()
for f in scan topo particles p2g g2p stencils poisson abi comm dd reseed; do
  stale=0
  [ -f _build/$f.o ] || stale=1
  for dep in $f.cu *.cuh ../../include/flipb200.h build.sh; do [ $dep -nt _build/$f.o ] && stale=1; done
  if [ $stale = 1 ]; then
    ( $NVCC $FLAGS -c $f.cu -o _build/$f.o > _build/$f.log 2>&1 || { cat _build/$f.log; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o $OUT _build/*.o -lcudart -ldl
echo "built $OUT"
