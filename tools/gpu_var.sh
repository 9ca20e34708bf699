#!/bin/bash
# two back-to-back quick benches on one box (host-gap variance) + host CPU info
bash tools/gpu_kern.sh ${1:-v}a | head -3; bash tools/gpu_kern.sh ${1:-v}b | head -3
python - <<PY
import json
for t in ("${1:-v}a","${1:-v}b"):
    d=json.loads(open("gpurun_out/%s.json"%t).read().splitlines()[-1]); print(t, "syncs/step", d["host_syncs_per_step"], "kernel ms sum", round(d["kernel_ms_sum_per_step"],3), "host step ms", d["config"]["step_ms_host"])
PY
lscpu | grep -i "model name\|MHz" | head -4
