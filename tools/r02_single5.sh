#!/bin/bash
set -u
mkdir -p gpurun_out
P=gpurun_out/r02l
timeout 300 python tools/diag/stream_vs_resident.py 60 > ${P}_diag_stream.log 2>&1; echo "diag rc=$?"; grep -c trial ${P}_diag_stream.log; grep -A8 "c\[" ${P}_diag_stream.log | cut -c1-600 | head -60
timeout 300 python tools/diag/stepwise_divergence.py 150 > ${P}_stepwise.log 2>&1; echo "stepwise rc=$?"; cut -c1-500 ${P}_stepwise.log | head -60
