#!/bin/bash
# single-GPU call: the GPU suite after the commuting column statistics + CH = 8, the fit that used to diverge 60 x 2 times,
# float32-grade and latency-bound bench lines
set -u
mkdir -p gpurun_out
P=gpurun_out/r02m
timeout 900 python -m pytest tests -m gpu -q --durations=6 > ${P}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> ${P}_pytest_gpu.log
tail -22 ${P}_pytest_gpu.log | cut -c1-300
timeout 300 python tools/diag/stream_vs_resident.py 60 > ${P}_diag_stream.log 2>&1; echo "diag rc=$?"; grep -c trial ${P}_diag_stream.log; grep -c "max|dW| 0.000e+00  max|dc| 0.000e+00  max|db| 0.000e+00" ${P}_diag_stream.log; grep "c\[" ${P}_diag_stream.log | head -3 | cut -c1-300
for w in c3 c3f32 c1 c1f32 c4; do
  timeout 600 python bench.py --workload $w --steps 60 --warmup 5 --no-cpu-baseline > ${P}_bench_$w.json 2> ${P}_bench_$w.err
  python - "$w" <<'PY'
import json, sys
w = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/r02m_bench_%s.json" % w).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print("%s: %.4f ms/step  value %.4g %s  e2e %.4g  f32_grade %s  roof %s" % (w, d["ms_per_step"], d["value"], d["unit"], e.get("value", 0),
          (d.get("f32_grade") or {}).get("value"), {k: (d.get("roofline") or {}).get(k) for k in ("frac", "regime", "achieved", "launch_ms")}))
except Exception as ex:
    print(w, "no line", ex); print(open("gpurun_out/r02m_bench_%s.err" % w).read()[-1200:])
PY
done
