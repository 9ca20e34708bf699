#!/bin/bash
# bench.py at N = 1, 2, 4, 8 ranks the way the driver launches it. usage: gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_scale.sh tag "1 2 4 8"'
TAG=${1:-scale}; NS=${2:-"1 2 4 8"}
mkdir -p gpurun_out
for N in $NS; do
  if [ "$N" = "1" ]; then
    timeout 400 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
  fi
  echo "N=$N rc=$?"; tail -c 400 gpurun_out/${TAG}_n$N.err | grep -v "^\*\|OMP_NUM" 
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_n$N.json").read().splitlines()[-1])
    print("N=%d value %.4g  ms/step %.3f  e2e %.4g (%.2f ms) launches %d iters %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["config"].get("pcg_iterations")))
    print("  stage_ms", {k: round(v,3) for k,v in d["stage_ms"].items()})
except Exception as e: print("no json", e)
PY
done
