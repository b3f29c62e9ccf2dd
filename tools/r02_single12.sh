#!/bin/bash
# ncu --set full of the mid chain variant (a 512-row shard of C3: what bounds the strong-scaled step?)
set -u
mkdir -p gpurun_out
P=gpurun_out/r02u
export KUCD_COOP=0
CMD="python bench.py --workload c3 --batch 512 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-f32-grade"
$CMD > ${P}_plain.log 2>&1 && echo "plain ok" &&
ncu --set full --clock-control none --import-source on -k "regex:chain_kernel" -s 2 -c 1 -o ${P}_prof_mid_chain $CMD > ${P}_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i ${P}_prof_mid_chain.ncu-rep --page raw --csv > ${P}_mid_chain_raw.csv 2>/dev/null; wc -c ${P}_mid_chain_raw.csv
