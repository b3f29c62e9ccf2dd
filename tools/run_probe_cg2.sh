#!/bin/bash
# cta_group::2 bring-up: the same probe as run_probe.sh with KUCD_CG=2 (256 x 256 tiles on CTA pairs).
mkdir -p gpurun_out
LOG=gpurun_out/probe_cg2.log
: > $LOG
P=build/probe_gemm
export KUCD_CG=2
run() { timeout 120 $P "$@" >> $LOG 2>&1; echo "exit=$? args: $*" >> $LOG; }
for maj in "0 0" "0 1" "1 1" "1 0"; do
  run $maj 256 256 64 256 1 0
  run $maj 256 512 256 256 1 0
  run $maj 512 512 1024 256 1 0
  run $maj 300 500 200 256 1 0
  run $maj 1000 777 136 256 1 0
  run $maj 512 512 128 256 2 2
  if grep -q "FAIL\|ERROR\|timed out" $LOG; then break; fi
done
echo "---- timing ----" >> $LOG
for maj in "0 0" "0 1" "1 1"; do
  run $maj 4096 4096 4096 256 1 0 20
done
run 1 1 4096 4096 4096 256 2 2 20
run 0 1 8192 8192 8192 256 1 0 10
export KUCD_CG=1
echo "---- cg=1 reference timing ----" >> $LOG
run 0 1 4096 4096 4096 256 1 0 20
grep -c PASS $LOG; grep "FAIL\|ERROR\|timed out\|TIMING" $LOG | head -40
