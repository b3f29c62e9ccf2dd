#!/bin/bash
# 2-GPU call after the re-entry of round 2: (1) is test_one_epoch_fit_streamed_or_resident flaky?  (2) dp_check incl. the
# 16-rows-per-rank shapes, (3) the bench commands of the final 8-GPU call at 2 ranks (de-risk)
set -u
mkdir -p gpurun_out
P=gpurun_out/r02j
N=$(python -c 'import torch; print(torch.cuda.device_count())')
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/diag/stream_vs_resident.py 12 > ${P}_diag_stream.log 2>&1; echo "diag rc=$?"; grep -c trial ${P}_diag_stream.log; grep "c\[" ${P}_diag_stream.log | head -8
timeout 600 $TR --master-port 29531 tests/dp_check.py > ${P}_dp_check_n$N.log 2>&1
echo "dp_check rc=$?" >> ${P}_dp_check_n$N.log
grep -E "dp_check\]|rc=" ${P}_dp_check_n$N.log | cut -c1-260
show() {
  python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r02j_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    s = d.get("strong_scaling")
    print("%s: %.4f ms/step %.3f M samples/s [%s] launch_ms %s dw %s%s" % (sys.argv[1], d["ms_per_step"], d["value"] / 1e6,
          d["config"].get("exchange", "")[:30], r.get("launch_ms"), r.get("dw_launch_ms"),
          "  strong: %.4f ms %.3f M [%s]" % (s["ms_per_step"], s["value"] / 1e6, s["exchange"][:30]) if s else ""))
    if r.get("step_breakdown_ms"): print("   breakdown", {k: v for k, v in r["step_breakdown_ms"].items() if k != "what"})
except Exception as ex:
    print(sys.argv[1], "no line", ex); print(open("gpurun_out/r02j_%s.err" % sys.argv[1]).read()[-1500:])
PY
}
multi() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 $TR --master-port 29541 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-e2e "$@" > ${P}_$name.json 2> ${P}_$name.err; show $name; }
multi c4_units X=1 -- --workload c4
multi c4_units_inorder KUCD_UNITS_OVERLAP=0 -- --workload c4
multi c3_weak X=1 -- --workload c3
multi c3_strong_auto X=1 -- --workload c3 --scaling strong
multi c3_strong_fused KUCD_FUSED_MIN_ROWS=1 -- --workload c3 --scaling strong
