#!/bin/bash
set -u
mkdir -p gpurun_out
P=gpurun_out/r02r
timeout 900 python -m pytest tests -x -q -m gpu > ${P}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> ${P}_pytest_gpu.log
tail -8 ${P}_pytest_gpu.log | cut -c1-400
for w in c1 c2; do
  timeout 300 python bench.py --workload $w --steps 469 --warmup 5 --no-cpu-baseline > ${P}_bench_$w.json 2> ${P}_bench_$w.err
  python - "$w" <<'PY'
import json, sys
w = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/r02r_bench_%s.json" % w).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print("%s: %.4f ms/step  value %.4g %s  e2e %.4g (%.4f ms/step)" % (w, d["ms_per_step"], d["value"], d["unit"], e.get("value", 0), e.get("ms_per_step", 0)))
except Exception as ex:
    print(w, "no line", ex); print(open("gpurun_out/r02r_bench_%s.err" % w).read()[-1200:])
PY
done
