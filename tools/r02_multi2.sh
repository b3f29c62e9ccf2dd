#!/bin/bash
# 2-GPU call of round 2: parity of every exchange variant (dp_check, incl. the unit-sharded step), then first timings
set -u
mkdir -p gpurun_out
N=$(python -c 'import torch; print(torch.cuda.device_count())')
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
bash tools/run_round2_switches.sh multi check > gpurun_out/r02b_check.out 2>&1
grep -E "^==|dp_check\]|rc=" gpurun_out/r02_exchange_check.log
run() {  # name, env..., -- bench args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 300 $TR --master-port 29541 bench.py --gpus $N --steps 30 --warmup 3 --no-cpu-baseline --no-e2e "$@" \
      > gpurun_out/r02b_$name.json 2> gpurun_out/r02b_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r02b_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    s = d.get("strong_scaling")
    print("%s: %.4f ms/step %.3f M samples/s [%s]%s" % (sys.argv[1], d["ms_per_step"], d["value"] / 1e6, d["config"].get("exchange", "")[:40],
          "  strong: %.4f ms %.3f M [%s]" % (s["ms_per_step"], s["value"] / 1e6, s["exchange"][:30]) if s else ""))
except Exception as e:
    print(sys.argv[1], "no line", e)
    print(open("gpurun_out/r02b_%s.err" % sys.argv[1]).read()[-1500:])
PY
}
run c4_auto X=1 -- --workload c4
run c4_dp KUCD_EXCHANGE=dp -- --workload c4
run c3_weak X=1 -- --workload c3
run c3_strong_units KUCD_EXCHANGE=units -- --workload c3 --scaling strong
run c3_strong_fused KUCD_FUSED_MIN_ROWS=1 KUCD_EXCHANGE=dp -- --workload c3 --scaling strong
