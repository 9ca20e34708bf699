#!/bin/bash
# A/B of the ghost exchange route at N ranks: peer memory (default) vs NCCL send/recv (FLIPB200_DD_P2P=0). usage: gpurun --gpus N -- 'bash tools/gpu_p2p_ab.sh N'
N=${1:-2}
mkdir -p gpurun_out
for P in 1 0; do
  OUT=gpurun_out/p2p${P}_n$N
  FLIPB200_DD_P2P=$P timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N+P)) bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-e2e > $OUT.json 2> $OUT.err
  echo "P2P=$P rc=$?"
  python - <<PY
import json
d=json.loads(open("$OUT.json").read().splitlines()[-1])
print("  value %.4g ms/step %.3f" % (d["value"], d["ms_per_step"]), {k: round(v,3) for k,v in d["stage_ms"].items()})
for k,v in d["kernels"].items():
    if k.startswith("dd_") or k.startswith("comm"): print("   %-22s %6.1f launches/step %7.2f us avg %7.3f ms/step" % (k, v["launches_per_step"], v["avg_us"], v["ms_per_step"]))
PY
done
