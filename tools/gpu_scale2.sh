#!/bin/bash
# default bench (no e2e, no cpu leg) and the C4 solver line at the given rank counts, each under its own timeout.
# usage: gpurun --gpus 8 --timeout 1200 -- 'bash tools/gpu_scale2.sh tag "2 4 8"'
TAG=${1:-scale}; NS=${2:-"2 4 8"}
mkdir -p gpurun_out
for N in $NS; do
  for WL in c2 c4; do
    OUT=gpurun_out/${TAG}_${WL}_n$N
    EXTRA="--no-cpu --no-e2e"; [ "$WL" = "c4" ] && EXTRA="--workload c4"
    if [ "$N" = "1" ]; then
      timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 $EXTRA > $OUT.json 2> $OUT.err
    else
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 5 --warmup 3 $EXTRA > $OUT.json 2> $OUT.err
    fi
    echo "N=$N $WL rc=$?"
    python - <<PY
import json
try:
    d=json.loads(open("$OUT.json").read().splitlines()[-1])
    print("  N=%d %s value %.4g %s  ms/step %.3f" % (d["n_gpus"], "$WL", d["value"], d["unit"], d["ms_per_step"]))
    if "stage_ms" in d: print("  stage_ms", {k: round(v,3) for k,v in d["stage_ms"].items()}, d["config"].get("dd_result_check"))
    if "solve_1e-6" in d: print("  1e-6", d["solve_1e-6"], "\n  5e-5", d["solve_5e-5"])
except Exception as e: print("  no json", e); import subprocess; print(subprocess.run("tail -c 600 $OUT.err", shell=True, capture_output=True, text=True).stdout)
PY
  done
done
