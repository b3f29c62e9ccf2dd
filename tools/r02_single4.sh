#!/bin/bash
# single-GPU diagnostics: where a rare small bf16 fit diverges; float32-grade piece length (CH = 4 / 8 / 16)
set -u
mkdir -p gpurun_out
P=gpurun_out/r02k
timeout 400 python tools/diag/stepwise_divergence.py 150 > ${P}_stepwise.log 2>&1; echo "stepwise rc=$?"; cat ${P}_stepwise.log | cut -c1-400 | head -60
for lib in keras_unsupervised_b200/libkucd.so build/libkucd_ch8.so build/libkucd_ch16.so; do
  timeout 200 python tools/diag/ch_experiment.py $lib 2>&1 | tail -4
done | tee ${P}_ch_experiment.log
