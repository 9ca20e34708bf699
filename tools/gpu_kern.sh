#!/bin/bash
# quick per-kernel timing of one bench configuration. usage: gpurun -- 'ENVVAR=.. bash tools/gpu_kern.sh tag'
TAG=${1:-k}
mkdir -p gpurun_out
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err || tail -5 gpurun_out/${TAG}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}.json").read().splitlines()[-1])
print("${TAG}: value %.4g  ms/step %.3f launches %d iters %s" % (d["value"], d["ms_per_step"], d["gpu_launches"], d["config"].get("pcg_iterations")))
print("  stage_ms", {k: round(v,3) for k,v in d["stage_ms"].items()})
for k,v in list(d["kernels"].items())[:6]: print("  %-24s %8.3f ms/step  %7.1f launches  %8.2f us avg" % (k, v["ms_per_step"], v["launches_per_step"], v["avg_us"]))
PY
