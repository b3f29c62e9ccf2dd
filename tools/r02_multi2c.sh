#!/bin/bash
set -u
mkdir -p gpurun_out
N=$(python -c 'import torch; print(torch.cuda.device_count())')
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02f_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest_gpu.log
tail -8 gpurun_out/r02f_pytest_gpu.log
timeout 600 $TR --master-port 29532 tests/dp_check.py > gpurun_out/r02f_dp_check_n$N.log 2>&1
echo "dp_check rc=$?" >> gpurun_out/r02f_dp_check_n$N.log
grep -E "dp_check\]|rc=" gpurun_out/r02f_dp_check_n$N.log
show() {
  python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r02f_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print("%s: %.4f ms/step %.3f M samples/s [%s] launch_ms %s dw %s" % (sys.argv[1], d["ms_per_step"], d["value"] / 1e6,
          d["config"].get("exchange", "")[:30], r.get("launch_ms"), r.get("dw_launch_ms")))
    if r.get("step_breakdown_ms"): print("   breakdown", {k: v for k, v in r["step_breakdown_ms"].items() if k != "what"})
except Exception as ex:
    print(sys.argv[1], "no line", ex); print(open("gpurun_out/r02f_%s.err" % sys.argv[1]).read()[-1500:])
PY
}
multi() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 $TR --master-port 29541 bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r02f_$name.json 2> gpurun_out/r02f_$name.err; show $name; }
single() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env CUDA_VISIBLE_DEVICES=0 "${envs[@]}" timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r02f_$name.json 2> gpurun_out/r02f_$name.err; show $name; }
multi c4_units_overlap X=1 -- --workload c4
multi c4_units_inorder KUCD_UNITS_OVERLAP=0 -- --workload c4
multi c4_units_overlap2 X=1 -- --workload c4
multi c3_b512_fused_bn128 KUCD_FUSED_MIN_ROWS=1 KUCD_MID_BN=128 -- --workload c3 --batch 512
multi c3_b512_fused_bn256 KUCD_FUSED_MIN_ROWS=1 KUCD_MID_BN=256 -- --workload c3 --batch 512
single c3_b512_bn128 KUCD_MID_BN=128 -- --workload c3 --batch 512
single c3_b512_bn256 KUCD_MID_BN=256 -- --workload c3 --batch 512
single c3_b1024_bn128 KUCD_MID_BN=128 -- --workload c3 --batch 1024
single c3_b1024_bn256 KUCD_MID_BN=256 -- --workload c3 --batch 1024
# ---- ncu (one GPU): launch list of the default bench, then full sections for the hot kernels
export CUDA_VISIBLE_DEVICES=0
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-f32-grade"
$CMD > gpurun_out/r02f_plain_c3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_c3_launches.csv $CMD > gpurun_out/r02f_ncu_list.log 2>&1
echo "ncu list rc=$?"
$CMD > gpurun_out/r02f_plain_c3b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:chain_kernel|gemm_bf16_kernel|update_w_kernel" -s 3 -c 3 -o gpurun_out/r02f_prof_c3 $CMD > gpurun_out/r02f_ncu_full.log 2>&1
echo "ncu full rc=$?"
CMD2="python bench.py --workload c3f32 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD2 > gpurun_out/r02f_plain_c3f32.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:chain_kernel|gemm_bf16_kernel" -s 2 -c 2 -o gpurun_out/r02f_prof_c3f32 $CMD2 > gpurun_out/r02f_ncu_full32.log 2>&1
echo "ncu full f32 rc=$?"
ls -la gpurun_out/*.ncu-rep
