#!/bin/bash
# How fast do the SMs actually run during back-to-back GEMMs?  (power cap vs kernel efficiency)
mkdir -p gpurun_out
LOG=gpurun_out/probe_clock.log
: > $LOG
P=build/probe_gemm
for cg in 1 2; do
  export KUCD_CG=$cg
  for it in 5 20 200 1000; do
    echo "== cg=$cg iters=$it" >> $LOG
    timeout 120 $P 0 1 4096 4096 4096 256 1 0 $it >> $LOG 2>&1
  done
  echo "== cg=$cg 8192^3 iters=50" >> $LOG
  timeout 120 $P 0 1 8192 8192 8192 256 1 0 50 >> $LOG 2>&1
done
grep "==\|CLOCK\|TIMING" $LOG
