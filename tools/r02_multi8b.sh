#!/bin/bash
# final 8-GPU call of round 2 (kept short: eight GPUs are charged eight times)
set -u
mkdir -p gpurun_out
N=$(python -c 'import torch; print(torch.cuda.device_count())')
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
timeout 400 $TR --master-port 29532 tests/dp_check.py > gpurun_out/r02h_dp_check_n$N.log 2>&1
echo "dp_check rc=$?" >> gpurun_out/r02h_dp_check_n$N.log
grep -E "dp_check\]|rc=" gpurun_out/r02h_dp_check_n$N.log
run() {  # name, env..., -- bench args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 300 $TR --master-port 29541 bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline --no-e2e "$@" \
      > gpurun_out/r02h_$name.json 2> gpurun_out/r02h_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r02h_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    s = d.get("strong_scaling")
    r = d.get("roofline") or {}
    print("%s: %.4f ms/step %.3f M samples/s [%s]%s" % (sys.argv[1], d["ms_per_step"], d["value"] / 1e6, d["config"].get("exchange", "")[:40],
          "  strong: %.4f ms %.3f M [%s]" % (s["ms_per_step"], s["value"] / 1e6, s["exchange"][:30]) if s else ""))
    if r.get("step_breakdown_ms"): print("   breakdown", {k: v for k, v in r["step_breakdown_ms"].items() if k != "what"}, "launch_ms", r.get("launch_ms"), "dw", r.get("dw_launch_ms"))
except Exception as e:
    print(sys.argv[1], "no line", e)
    print(open("gpurun_out/r02h_%s.err" % sys.argv[1]).read()[-1500:])
PY
}
run c4_units X=1 -- --workload c4
run c4_units_inorder KUCD_UNITS_OVERLAP=0 -- --workload c4
run c3_weak X=1 -- --workload c3
run c3_strong_fused KUCD_FUSED_MIN_ROWS=1 -- --workload c3 --scaling strong
