#!/bin/bash
# ncu launch list of the C1 bench incl. the streamed pass (per-kernel times of a latency-bound step; cold-cache, serialised)
set -u
mkdir -p gpurun_out
export KUCD_COOP=0
CMD="python bench.py --workload c1 --steps 20 --warmup 3 --no-cpu-baseline"
timeout 40 $CMD > gpurun_out/r02x_plain.log 2>&1 && timeout 45 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02x_c1_launches.csv $CMD > gpurun_out/r02x_ncu.log 2>&1
echo "rc=$?"; wc -l gpurun_out/r02x_c1_launches.csv
