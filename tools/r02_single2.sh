#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02d_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest_gpu.log
tail -40 gpurun_out/r02d_pytest_gpu.log
for w in c3 c3f32 c1 c1f32 c2 c5 c4; do
  extra=""; [ "$w" != "c3" ] && extra="--no-cpu-baseline"
  timeout 600 python bench.py --workload $w --steps 60 --warmup 5 $extra > gpurun_out/r02d_bench_$w.json 2> gpurun_out/r02d_bench_$w.err
  python - "$w" <<'PY'
import json, sys
w = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/r02d_bench_%s.json" % w).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print("%s: %.4f ms/step  value %.4g %s  e2e %.4g  f32_grade %s  roof %s" % (w, d["ms_per_step"], d["value"], d["unit"], e.get("value", 0),
          (d.get("f32_grade") or {}).get("value"), {k: (d.get("roofline") or {}).get(k) for k in ("frac", "regime", "achieved", "launch_ms")}))
    if w == "c5":
        print("  " + "  ".join("n=%d: %.3f ms" % (r["rows"], r["transform_ms"]) for r in d["sweep"]))
    if w == "c3":
        print("  cpu", d.get("cpu_baseline"))
except Exception as ex:
    print(w, "no line", ex); print(open("gpurun_out/r02d_bench_%s.err" % w).read()[-1200:])
PY
done
