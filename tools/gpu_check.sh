#!/bin/bash
# One gpurun call: GPU parity tests, a bench line, and the ncu launch list of one bench step.
# usage (from the repo root): gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag]'
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"
tail -c 1500 gpurun_out/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu \
    > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "ncu exit $?"
wc -l gpurun_out/${TAG}_launches.csv
