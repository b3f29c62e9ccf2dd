#!/bin/bash
# Which side paces the mainloop?  flags 0 = real kernel, 1 = MMA alone (no TMA), 2 = TMA alone (no MMA)
mkdir -p gpurun_out
LOG=gpurun_out/probe_pacing.log
: > $LOG
P=build/probe_gemm
for cg in 1 2; do
  for fl in 0 1 2; do
    export KUCD_CG=$cg KUCD_DBG_FLAGS=$fl
    echo "== cg=$cg flags=$fl   4096^3" >> $LOG
    timeout 120 $P 0 1 4096 4096 4096 256 1 0 50 2>&1 | grep "CLOCK\|TIMING" >> $LOG
    echo "== cg=$cg flags=$fl   one tile per SM-unit, K=16384 (M=$((128*148)) N=256)" >> $LOG
    timeout 120 $P 0 1 $((128*148)) 256 16384 256 1 0 50 2>&1 | grep "CLOCK\|TIMING" >> $LOG
  done
done
cat $LOG
