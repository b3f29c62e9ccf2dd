#!/bin/bash
set -u
mkdir -p gpurun_out
N=$(python -c 'import torch; print(torch.cuda.device_count())')
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02e_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest_gpu.log
tail -15 gpurun_out/r02e_pytest_gpu.log
timeout 600 $TR --master-port 29532 tests/dp_check.py > gpurun_out/r02e_dp_check_n$N.log 2>&1
echo "dp_check rc=$?" >> gpurun_out/r02e_dp_check_n$N.log
grep -E "dp_check\]|rc=" gpurun_out/r02e_dp_check_n$N.log
show() {
  python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r02e_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    r = d.get("roofline") or {}
    print("%s: %.4f ms/step %.3f M samples/s [%s] e2e %.3f M  launch_ms %s dw %s" % (sys.argv[1], d["ms_per_step"], d["value"] / 1e6,
          d["config"].get("exchange", "")[:30], e.get("value", 0) / 1e6, r.get("launch_ms"), r.get("dw_launch_ms")))
    if r.get("step_breakdown_ms"): print("   breakdown", {k: v for k, v in r["step_breakdown_ms"].items() if k != "what"})
except Exception as ex:
    print(sys.argv[1], "no line", ex); print(open("gpurun_out/r02e_%s.err" % sys.argv[1]).read()[-1500:])
PY
}
multi() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 $TR --master-port 29541 bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r02e_$name.json 2> gpurun_out/r02e_$name.err; show $name; }
single() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env CUDA_VISIBLE_DEVICES=0 "${envs[@]}" timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r02e_$name.json 2> gpurun_out/r02e_$name.err; show $name; }
multi c4_units X=1 -- --workload c4
multi c3_b512_nccl X=1 -- --workload c3 --batch 512
multi c3_b512_fused KUCD_FUSED_MIN_ROWS=1 -- --workload c3 --batch 512
multi c3_b512_nochain KUCD_FUSED_MIN_ROWS=1 KUCD_MID_CHAIN=0 -- --workload c3 --batch 512
multi c3_b1024_fused KUCD_FUSED_MIN_ROWS=1 -- --workload c3 --batch 1024
single c3_b512_1gpu X=1 -- --workload c3 --batch 512 --no-e2e
single c3_b512_1gpu_nomid KUCD_MID_CHAIN=0 -- --workload c3 --batch 512 --no-e2e
single c1 X=1 -- --workload c1 --steps 469
single c1f32 X=1 -- --workload c1f32 --steps 469
