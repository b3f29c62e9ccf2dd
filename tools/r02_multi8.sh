#!/bin/bash
# 8-GPU call of round 2: parity at 8 ranks, then C4 and C3 (weak + strong) under the exchange variants
set -u
mkdir -p gpurun_out
N=$(python -c 'import torch; print(torch.cuda.device_count())')
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
timeout 600 $TR --master-port 29532 tests/dp_check.py > gpurun_out/r02c_dp_check_n$N.log 2>&1
echo "dp_check rc=$?" >> gpurun_out/r02c_dp_check_n$N.log
grep -E "dp_check\]|rc=" gpurun_out/r02c_dp_check_n$N.log
run() {  # name, env..., -- bench args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 300 $TR --master-port 29541 bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline "$@" \
      > gpurun_out/r02c_$name.json 2> gpurun_out/r02c_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r02c_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    s = d.get("strong_scaling")
    e = d.get("e2e")
    print("%s: %.4f ms/step %.3f M samples/s [%s]%s%s" % (sys.argv[1], d["ms_per_step"], d["value"] / 1e6, d["config"].get("exchange", "")[:40],
          "  strong: %.4f ms %.3f M [%s]" % (s["ms_per_step"], s["value"] / 1e6, s["exchange"][:30]) if s else "",
          "  e2e %.3f M (u8 %.3f, bits %.3f)" % (e["value"] / 1e6, e.get("uint8_input", {}).get("value", 0) / 1e6,
                                                  e.get("packed_bits_input", {}).get("value", 0) / 1e6) if e else ""))
except Exception as e:
    print(sys.argv[1], "no line", e)
    print(open("gpurun_out/r02c_%s.err" % sys.argv[1]).read()[-1500:])
PY
}
run c4_units X=1 -- --workload c4 --no-e2e
run c4_dp_nccl KUCD_EXCHANGE=dp -- --workload c4 --no-e2e
run c4_dp_nccl16s4 KUCD_EXCHANGE=dp KUCD_FUSED_REDUCE=0 KUCD_AR_SLABS=4 KUCD_WIRE_BF16=1 -- --workload c4 --no-e2e
run c3_weak X=1 -- --workload c3
run c3_strong_fused KUCD_FUSED_MIN_ROWS=1 KUCD_EXCHANGE=dp -- --workload c3 --scaling strong --no-e2e
run c3_strong_fused16 KUCD_FUSED_MIN_ROWS=1 KUCD_EXCHANGE=dp KUCD_WIRE_BF16=1 -- --workload c3 --scaling strong --no-e2e
run c3_strong_units KUCD_EXCHANGE=units -- --workload c3 --scaling strong --no-e2e
