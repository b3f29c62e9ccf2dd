#!/bin/bash
# single-GPU call after the re-entry of round 2: the whole GPU suite on HEAD, smoke(), the driver's two bench arms,
# every other workload, then the ncu launch list and one --set full capture of the C3 step's kernels
set -u
mkdir -p gpurun_out
P=gpurun_out/r02i
timeout 900 python -m pytest tests -m gpu -q --durations=8 > ${P}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> ${P}_pytest_gpu.log
tail -25 ${P}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 ${P}_smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > ${P}_bench_reference.json 2> ${P}_bench_reference.err; echo "reference arm rc=$?"; tail -c 600 ${P}_bench_reference.json
for w in c3 c3f32 c1 c1f32 c2 c5 c4; do
  extra="--steps 60 --warmup 5 --no-cpu-baseline"; [ "$w" = "c3" ] && extra=""
  timeout 600 python bench.py --workload $w $extra > ${P}_bench_$w.json 2> ${P}_bench_$w.err
  python - "$w" <<'PY'
import json, sys
w = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/r02i_bench_%s.json" % w).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print("%s: %.4f ms/step  value %.4g %s  e2e %.4g  f32_grade %s  roof %s" % (w, d["ms_per_step"], d["value"], d["unit"], e.get("value", 0),
          (d.get("f32_grade") or {}).get("value"), {k: (d.get("roofline") or {}).get(k) for k in ("frac", "regime", "achieved", "launch_ms")}))
    if w == "c5":
        print("  " + "  ".join("n=%d: %.3f ms" % (r["rows"], r["transform_ms"]) for r in d["sweep"]))
    if w == "c3":
        print("  cpu", d.get("cpu_baseline")); print("  clocks", d.get("clocks"))
except Exception as ex:
    print(w, "no line", ex); print(open("gpurun_out/r02i_bench_%s.err" % w).read()[-1200:])
PY
done
# ---- ncu (plain launches of the chain kernels: ncu cannot replay the cooperative form)
export KUCD_COOP=0
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-f32-grade"
$CMD > ${P}_plain_c3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_c3_launches.csv $CMD > ${P}_ncu_list.log 2>&1
echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:chain_kernel|gemm_bf16_kernel|update_w_kernel" -s 3 -c 3 -o ${P}_prof_c3 $CMD > ${P}_ncu_full.log 2>&1
echo "ncu full rc=$?"
CMD3="python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD3 > ${P}_plain_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:gemm_bf16_kernel|update_w_kernel" -s 2 -c 2 -o ${P}_prof_c4 $CMD3 > ${P}_ncu_full_c4.log 2>&1
echo "ncu full c4 rc=$?"
ls -la gpurun_out/r02i*.ncu-rep
