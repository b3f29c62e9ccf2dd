#!/bin/bash
# Runs the tcgen05 GEMM bring-up probe over operand-major combinations, tile widths and ragged
# shapes.  Meant for `gpurun -- bash tools/run_probe.sh`; the log comes back in gpurun_out/.
mkdir -p gpurun_out
LOG=gpurun_out/probe.log
: > $LOG
P=build/probe_gemm
run() { timeout 150 $P "$@" >> $LOG 2>&1; echo "exit=$? args: $*" >> $LOG; }

nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1

for maj in "0 0" "0 1" "1 1" "1 0"; do
  run $maj 128 256 64 256 1 0
  run $maj 128 256 256 256 1 0
  run $maj 256 512 1024 256 1 0
  run $maj 300 500 200 128 1 0
  run $maj 300 500 200 64 1 0
  run $maj 256 512 128 256 2 2
done

# alternates for MN-major descriptors, only informative if the defaults above fail:
#   swapped LBO/SBO roles
if grep -q "FAIL\|ERROR" $LOG; then
  echo "---- alternates ----" >> $LOG
  run 0 1 128 256 64 256 1 0 0   0 0 0     1024 8192 2048
  run 0 1 128 256 64 256 1 0 0   0 0 0     8192 1024 1024
  run 0 1 128 256 64 256 1 0 0   0 0 0     1024 8192 1024
  run 1 0 128 256 64 256 1 0 0   1024 8192 2048   0 0 0
  run 1 0 128 256 64 256 1 0 0   8192 1024 1024   0 0 0
  run 1 0 128 256 64 256 1 0 0   1024 8192 1024   0 0 0
fi

echo "---- timing ----" >> $LOG
for maj in "0 0" "0 1" "1 1"; do
  run $maj 4096 4096 4096 256 1 0 20
  run $maj 4096 4096 4096 128 1 0 20
done
run 1 1 4096 4096 4096 256 2 2 20
run 0 1 8192 8192 8192 256 1 0 10
tail -60 $LOG
