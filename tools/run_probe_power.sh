#!/bin/bash
# Is the MMA rate limited by power?  Same kernel, operands all zero / small integers / full-mantissa random.
mkdir -p gpurun_out
LOG=gpurun_out/probe_power.log
: > $LOG
P=build/probe_gemm
for cg in 1 2; do
  for fill in zero int rand; do
    for fl in 0 1; do
      export KUCD_CG=$cg KUCD_DBG_FLAGS=$fl KUCD_PROBE_FILL=$fill
      echo "== cg=$cg fill=$fill flags=$fl  M=18944 N=256 K=16384 (one tile per SM)" >> $LOG
      timeout 120 $P 0 1 $((128*148)) 256 16384 256 1 0 200 2>&1 | grep "CLOCK\|TIMING" >> $LOG
    done
    export KUCD_DBG_FLAGS=0
    echo "== cg=$cg fill=$fill flags=0  8192^3" >> $LOG
    timeout 120 $P 0 1 8192 8192 8192 256 1 0 30 2>&1 | grep "CLOCK\|TIMING" >> $LOG
  done
done
cat $LOG
