#!/bin/bash
# GPU check of the multi-block reconstruction statistic; C1 end to end over a whole epoch
set -u
mkdir -p gpurun_out
P=gpurun_out/r02t
timeout 900 python -m pytest tests -x -q -m gpu > ${P}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> ${P}_pytest_gpu.log
tail -5 ${P}_pytest_gpu.log | cut -c1-300
for c in 8 32; do
  KUCD_STREAM_CHUNK=$c timeout 300 python bench.py --workload c1 --steps 469 --warmup 5 --no-cpu-baseline > ${P}_bench_c1_chunk$c.json 2> ${P}_bench_c1_chunk$c.err
  python - "$c" <<'PY'
import json, sys
c = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/r02t_bench_c1_chunk%s.json" % c).read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print("c1 chunk %s: %.4f ms/step resident, e2e %.4g samples/s (%.4f ms/step); uint8 %.4f ms; bits %.4f ms" % (c, d["ms_per_step"], e.get("value", 0), e.get("ms_per_step", 0),
          (e.get("uint8_input") or {}).get("ms_per_step", 0), (e.get("packed_bits_input") or {}).get("ms_per_step", 0)))
except Exception as ex:
    print(c, "no line", ex); print(open("gpurun_out/r02t_bench_c1_chunk%s.err" % c).read()[-1200:])
PY
done
