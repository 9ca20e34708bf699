#!/bin/bash
# A/B of the coarse-coefficient all-reduce (one staged message vs four fresh arrays) under the host-phase trace. usage: gpurun --gpus N -- 'bash tools/gpu_ar_ab.sh N'
N=${1:-4}
for S in 1 0; do
  FLIPB200_DD_AR_STAGE=$S FLIPB200_PHASE_TRACE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29730+S)) bench.py --gpus $N --steps 2 --warmup 3 --no-cpu --no-e2e --no-check > gpurun_out/ar${S}_n$N.json 2> gpurun_out/ar${S}_n$N.err
  echo "stage=$S rc=$?"; grep -E "all-reduce coefficients|ppe dd_build_coarse|S4 solve_ppe" gpurun_out/ar${S}_n$N.err | head -6
done
