#!/bin/bash
# First GPU calls of the next round: the opt-in paths that were written without a GPU at the end of round 1
# (all compile, none has run).  Each is checked for parity first, then measured on / off inside one call.
#
#   KUCD_STREAM_CHUNK=C   chunked streaming of latency-bound minibatches (kucd.cu: fit_host_chunked)
#   KUCD_PLANE_POOL=1     free list of data-set plane buffers (no cudaMalloc / cudaFree per transform_dataset)
#   kucd_rbm_delta_rule   DBN.fine_tune's building block (tests/test_fine_tune.py, KUCD_TEST_UNVERIFIED=1)
#   KUCD_WIRE_BF16=1      bf16 partial sums of dW on the wire: bf16 slots in the fused exchange, a bf16 ncclAllReduce
#                         otherwise (gemm.cuh: kEpiRawPush16)
#   KUCD_AR_SLABS=S       dW contracted, all-reduced (NCCL, second stream) and applied in S row slabs that overlap
#
#   gpurun --timeout 900 -- 'bash tools/run_round2_switches.sh single'          (one GPU)
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/run_round2_switches.sh multi check'
#   gpurun --gpus 8 --timeout 600 -- 'bash tools/run_round2_switches.sh multi bench [c4|c3]'   (C4 is quoted on 8 GPUs)
set -u
mkdir -p gpurun_out
MODE=${1:-single}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"

if [ "$MODE" = single ]; then
  LOG=gpurun_out/r02_stream_chunk.log
  : > $LOG
  # parity: every test that streams from host arrays, with the chunked path forced (4 minibatches per chunk leaves a
  # partial last chunk and a remainder minibatch at the tests' sizes)
  echo "== parity, KUCD_STREAM_CHUNK=4" >> $LOG
  KUCD_STREAM_CHUNK=4 timeout 600 python -m pytest tests -m gpu -x -q \
      -k "stream or fit_host or packed or one_epoch or keras or shuffl" >> $LOG 2>&1
  echo "rc=$?" >> $LOG
  echo "== kucd_rbm_delta_rule / DBN.fine_tune against the oracle (written without a GPU)" >> $LOG
  KUCD_TEST_UNVERIFIED=1 timeout 600 python -m pytest tests/test_fine_tune.py -m gpu -q >> $LOG 2>&1
  echo "rc=$?" >> $LOG
  # measurement: one epoch of the C1 shape through kucd_rbm_fit_host, per chunk size (0 = per-minibatch stream)
  for c in 0 4 8 16 32 64; do
    echo "== bench c1, KUCD_STREAM_CHUNK=$c" >> $LOG
    KUCD_STREAM_CHUNK=$c timeout 300 python bench.py --workload c1 --steps 469 --warmup 5 --no-cpu-baseline \
        > gpurun_out/r02_bench_c1_chunk$c.json 2>> $LOG
    python - "$c" >> $LOG <<'EOF'
import json, sys
try:
    d = json.loads(open("gpurun_out/r02_bench_c1_chunk%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print("chunk=%s value=%.0f e2e=%.0f samples/s (e2e %.1f us per step)" % (
        sys.argv[1], d["value"], d["e2e"]["value"], 1e3 * d["e2e"].get("ms_per_step", float("nan"))))
except Exception as e:  # noqa: BLE001
    print("chunk=%s: no line (%s)" % (sys.argv[1], e))
EOF
  done
  # KUCD_AR_SLABS on one GPU: the update of dW slab i (second stream, HBM-bound) overlaps the contraction of slab i+1
  echo "== parity, KUCD_AR_SLABS=3 forced at test sizes (training tests: parameters must not change)" >> $LOG
  KUCD_AR_SLABS=3 KUCD_AR_SLABS_MIN_ELEMS=1 timeout 900 python -m pytest tests -m gpu -x -q \
      -k "cd_step or fit or chain or persistent or momentum or dbn or stream" >> $LOG 2>&1
  echo "rc=$?" >> $LOG
  for sl in 1 2 4 8; do
    for w in c4 c3; do
      echo "== bench $w, one GPU, KUCD_AR_SLABS=$sl" >> $LOG
      KUCD_AR_SLABS=$sl timeout 300 python bench.py --workload $w --steps 60 --warmup 5 --no-cpu-baseline --no-e2e \
          > gpurun_out/r02_bench_${w}_slabs$sl.json 2>> $LOG
      python - "$w" "$sl" >> $LOG <<'EOF'
import json, sys
try:
    d = json.loads(open("gpurun_out/r02_bench_%s_slabs%s.json" % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
    print("%s slabs=%s: %.4f ms per step, %.3f M samples/s" % (sys.argv[1], sys.argv[2], d["ms_per_step"], d["value"] / 1e6))
except Exception as e:  # noqa: BLE001
    print("%s slabs=%s: no line (%s)" % (sys.argv[1], sys.argv[2], e))
EOF
    done
  done
  # where a latency-bound step spends its time: kernel durations of the C1 graph replay (3 launches per step); what is
  # left of the 43 us per step is launch gaps - the case for programmatic dependent launch (what comes next)
  echo "== ncu launch list, bench c1 (cold-cache, serialised: shares only)" >> $LOG
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
      --log-file gpurun_out/r02_c1_launches.csv python bench.py --workload c1 --steps 40 --warmup 5 --no-cpu-baseline \
      --no-e2e > gpurun_out/r02_c1_ncu.log 2>&1
  echo "rc=$?" >> $LOG
  # KUCD_PLANE_POOL=1: free list of data-set plane buffers (small-batch inference is cudaMalloc-bound without it)
  echo "== parity, KUCD_PLANE_POOL=1 (every GPU test that creates data sets)" >> $LOG
  KUCD_PLANE_POOL=1 timeout 900 python -m pytest tests -m gpu -x -q >> $LOG 2>&1
  echo "rc=$?" >> $LOG
  for v in 0 1; do
    echo "== bench c5, KUCD_PLANE_POOL=$v" >> $LOG
    KUCD_PLANE_POOL=$v timeout 600 python bench.py --workload c5 > gpurun_out/r02_bench_c5_pool$v.json 2>> $LOG
    python - "$v" >> $LOG <<'EOF'
import json, sys
try:
    d = json.loads(open("gpurun_out/r02_bench_c5_pool%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print("pool=%s " % sys.argv[1] + "  ".join("n=%d: %.3f ms" % (r["rows"], r["transform_ms"]) for r in d["sweep"]))
except Exception as e:  # noqa: BLE001
    print("pool=%s: no line (%s)" % (sys.argv[1], e))
EOF
  done
  cat $LOG
else
  # multi check : the four data-parallel parity runs (2 GPUs are enough: gpurun --gpus 2, charged twice the box time)
  # multi bench : C4 on every GPU of the box, six exchange variants, ~35 s each (gpurun --gpus 8: ~30 GPU-minutes)
  WHAT=${2:-check}
  N=$(python -c 'import torch; print(torch.cuda.device_count())')
  LOG=gpurun_out/r02_exchange_$WHAT.log
  : > $LOG
  if [ "$WHAT" = check ]; then
    echo "== dp_check, $N ranks, KUCD_WIRE_BF16=1, fused exchange (oracle models the rounded partial sums)" >> $LOG
    KUCD_WIRE_BF16=1 timeout 600 $TR --nproc-per-node $N --master-port 29531 tests/dp_check.py >> $LOG 2>&1
    echo "rc=$?" >> $LOG
    echo "== dp_check, $N ranks, KUCD_WIRE_BF16=1, bf16 ncclAllReduce (exact model at 2 ranks, tolerance beyond)" >> $LOG
    KUCD_WIRE_BF16=1 KUCD_FUSED_REDUCE=0 timeout 600 $TR --nproc-per-node $N --master-port 29534 tests/dp_check.py >> $LOG 2>&1
    echo "rc=$?" >> $LOG
    echo "== dp_check, $N ranks, KUCD_AR_SLABS=2, fp32 ncclAllReduce in row slabs (same bar as the default)" >> $LOG
    KUCD_AR_SLABS=2 KUCD_AR_SLABS_MIN_ELEMS=1 KUCD_FUSED_REDUCE=0 timeout 600 $TR --nproc-per-node $N --master-port 29535 tests/dp_check.py >> $LOG 2>&1
    echo "rc=$?" >> $LOG
    echo "== dp_check, $N ranks, default (must stay bit-identical to one GPU)" >> $LOG
    timeout 600 $TR --nproc-per-node $N --master-port 29532 tests/dp_check.py >> $LOG 2>&1
    echo "rc=$?" >> $LOG
  else
    w=${3:-c4}
    for v in "nccl32 KUCD_FUSED_REDUCE=0 KUCD_WIRE_BF16=0" "nccl16 KUCD_FUSED_REDUCE=0 KUCD_WIRE_BF16=1" \
             "fused16 KUCD_FUSED_MIN_ROWS=1 KUCD_WIRE_BF16=1" \
             "nccl32s4 KUCD_FUSED_REDUCE=0 KUCD_AR_SLABS=4" \
             "nccl16s4 KUCD_FUSED_REDUCE=0 KUCD_AR_SLABS=4 KUCD_WIRE_BF16=1" \
             "nccl16s8 KUCD_FUSED_REDUCE=0 KUCD_AR_SLABS=8 KUCD_WIRE_BF16=1" \
             "nccl16s4r32 KUCD_FUSED_REDUCE=0 KUCD_AR_SLABS=4 KUCD_WIRE_BF16=1 KUCD_AR_RESERVE_SMS=32"; do
      set -- $v
      echo "== bench $w at $N GPUs, $1 ($2 $3 ${4:-} ${5:-})" >> $LOG
      env $2 $3 ${4:-X_UNUSED=0} ${5:-X_UNUSED2=0} timeout 300 $TR --nproc-per-node $N --master-port 29533 bench.py --gpus $N --workload $w \
          --steps 60 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_${w}_n${N}_$1.json 2>> $LOG
      python - "$w" "$N" "$1" >> $LOG <<'EOF'
import json, sys
w, n, v = sys.argv[1:4]
try:
    d = json.loads(open("gpurun_out/r02_bench_%s_n%s_%s.json" % (w, n, v)).read().strip().splitlines()[-1])
    print("%s n=%s %s: %.3f M samples/s, %.4f ms per step" % (w, n, v, d["value"] / 1e6, d["ms_per_step"]))
except Exception as e:  # noqa: BLE001
    print("%s n=%s %s: no line (%s)" % (w, n, v, e))
EOF
    done
  fi
  cat $LOG
fi
