#!/bin/bash
# run selected GPU tests + optional bench. usage: gpurun -- 'bash tools/gpu_t.sh tag "<pytest args>" [bench args...]'
TAG=${1:-t}; PYARGS=${2:-tests -m gpu}; shift; shift
mkdir -p gpurun_out
timeout 900 python -m pytest $PYARGS -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -30 gpurun_out/${TAG}_pytest.log
if [ $# -gt 0 ]; then
timeout 600 python bench.py "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench.json"))
    print("value %.4g  ms/step %.3f  e2e %.4g  launches %d iters %s" % (d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value", 0), d["gpu_launches"], d["config"].get("pcg_iterations")))
    print("stage_ms", {k: round(v,3) for k,v in d["stage_ms"].items()})
    for k,v in d["kernels"].items(): print("  %-24s %8.3f ms/step  %7.1f launches  %8.2f us avg  %7.1f GB/s" % (k, v["ms_per_step"], v["launches_per_step"], v["avg_us"], v["algorithmic_GBps"]))
except Exception as e: print("no bench json", e)
PY
fi
