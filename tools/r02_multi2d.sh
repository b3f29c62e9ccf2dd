#!/bin/bash
set -u
mkdir -p gpurun_out
N=$(python -c 'import torch; print(torch.cuda.device_count())')
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
for i in 1 2; do
timeout 600 $TR --master-port 2953$i tests/dp_check.py > gpurun_out/r02g_dp_check_n${N}_$i.log 2>&1
echo "dp_check rc=$?" >> gpurun_out/r02g_dp_check_n${N}_$i.log
grep -E "dp_check\] (units|PASS|FAIL)|rc=" gpurun_out/r02g_dp_check_n${N}_$i.log
done
show() {
  python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r02g_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print("%s: %.4f ms/step %.3f M samples/s [%s] launch_ms %s dw %s" % (sys.argv[1], d["ms_per_step"], d["value"] / 1e6,
          d["config"].get("exchange", "")[:30], r.get("launch_ms"), r.get("dw_launch_ms")))
    if r.get("step_breakdown_ms"): print("   breakdown", {k: v for k, v in r["step_breakdown_ms"].items() if k != "what"})
except Exception as ex:
    print(sys.argv[1], "no line", ex); print(open("gpurun_out/r02g_%s.err" % sys.argv[1]).read()[-1500:])
PY
}
multi() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 $TR --master-port 29541 bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline --no-e2e "$@" > gpurun_out/r02g_$name.json 2> gpurun_out/r02g_$name.err; show $name; }
multi c4_units_overlap X=1 -- --workload c4
multi c4_units_inorder KUCD_UNITS_OVERLAP=0 -- --workload c4
multi c4_units_overlap2 X=1 -- --workload c4
# ---- ncu (one GPU; plain launches of the chain kernels: ncu cannot replay the cooperative form)
export CUDA_VISIBLE_DEVICES=0
export KUCD_COOP=0
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-f32-grade"
$CMD > gpurun_out/r02g_plain_c3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02g_c3_launches.csv $CMD > gpurun_out/r02g_ncu_list.log 2>&1
echo "ncu list rc=$?"
$CMD > gpurun_out/r02g_plain_c3b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:chain_kernel|gemm_bf16_kernel|update_w_kernel" -s 3 -c 3 -o gpurun_out/r02g_prof_c3 $CMD > gpurun_out/r02g_ncu_full.log 2>&1
echo "ncu full rc=$?"
CMD2="python bench.py --workload c3f32 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD2 > gpurun_out/r02g_plain_c3f32.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:chain_kernel|gemm_bf16_kernel" -s 2 -c 2 -o gpurun_out/r02g_prof_c3f32 $CMD2 > gpurun_out/r02g_ncu_full32.log 2>&1
echo "ncu full f32 rc=$?"
CMD3="python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD3 > gpurun_out/r02g_plain_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:gemm_bf16_kernel|update_w_kernel" -s 2 -c 2 -o gpurun_out/r02g_prof_c4 $CMD3 > gpurun_out/r02g_ncu_full_c4.log 2>&1
echo "ncu full c4 rc=$?"
ls -la gpurun_out/r02g*.ncu-rep
