for P in 0.02 5.0; do
FLIPB200_CLOCK_PERIOD=$P timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/p$P.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/p$P.json").read().splitlines()[-1]); print("period $P", d["ms_per_step"], d["config"]["step_ms_host"], d["clocks"]["samples"])
PY
done
