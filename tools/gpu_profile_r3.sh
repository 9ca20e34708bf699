#!/bin/bash
# round-2 (second session) profile set on one B200: the default bench line (with e2e and the reference arm), the launch list of two
# substeps, one ncu --set full pass over every hot kernel of one substep, the C3 microbench and the C4 solver line.
# usage: gpurun -- 'bash tools/gpu_profile_r3.sh <tag>'
TAG=${1:-r03}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench exit $?"; cut -c1-400 gpurun_out/${TAG}_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_launches.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:mg_cluster_kernel|mg_cycle_kernel|g2p_tile_kernel|p2g_xrow_kernel' -s 66 -c 22 -f -o gpurun_out/${TAG}_hot \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_hot.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/${TAG}_hot.ncu-rep
timeout 600 python bench.py --workload c3 > gpurun_out/${TAG}_c3.json 2> gpurun_out/${TAG}_c3.err; echo "c3 exit $?"; tail -2 gpurun_out/${TAG}_c3.err; cut -c1-600 gpurun_out/${TAG}_c3.json
timeout 600 python bench.py --workload c4 > gpurun_out/${TAG}_c4.json 2> gpurun_out/${TAG}_c4.err; echo "c4 exit $?"; tail -2 gpurun_out/${TAG}_c4.err; cut -c1-600 gpurun_out/${TAG}_c4.json
