#!/bin/bash
# the 8-GPU call of the re-entry (eight GPUs are charged eight times: three programs only)
set -u
mkdir -p gpurun_out
P=gpurun_out/r02o
N=$(python -c 'import torch; print(torch.cuda.device_count())')
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
timeout 400 $TR --master-port 29532 tests/dp_check.py > ${P}_dp_check_n$N.log 2>&1
echo "dp_check rc=$?" >> ${P}_dp_check_n$N.log
grep -E "dp_check\]|rc=" ${P}_dp_check_n$N.log | cut -c1-260
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 $TR --master-port 29541 bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline "$@" \
      > ${P}_$name.json 2> ${P}_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r02o_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    s = d.get("strong_scaling")
    r = d.get("roofline") or {}
    e = d.get("e2e") or {}
    print("%s: %.4f ms/step %.3f M samples/s [%s] e2e %.3f M%s" % (sys.argv[1], d["ms_per_step"], d["value"] / 1e6, d["config"].get("exchange", "")[:40], (e.get("value") or 0) / 1e6,
          "  strong: %.4f ms %.3f M [%s]" % (s["ms_per_step"], s["value"] / 1e6, s["exchange"][:30]) if s else ""))
    if r.get("step_breakdown_ms"): print("   breakdown", {k: v for k, v in r["step_breakdown_ms"].items() if k != "what"}, "launch_ms", r.get("launch_ms"), "dw", r.get("dw_launch_ms"))
    print("   clocks", d.get("clocks"))
except Exception as e:
    print(sys.argv[1], "no line", e)
    print(open("gpurun_out/r02o_%s.err" % sys.argv[1]).read()[-1500:])
PY
}
run c4_units X=1 -- --workload c4 --no-e2e
run c3_weak X=1 -- --workload c3
run c3_strong X=1 -- --workload c3 --scaling strong --no-e2e
