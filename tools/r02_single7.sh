#!/bin/bash
# single-GPU call: GPU suite on the final column-statistic grid; the per-rank share of C3 strong scaling at 8 GPUs (512
# and 1024 rows) with the mid chain variants on / off
set -u
mkdir -p gpurun_out
P=gpurun_out/r02n
timeout 900 python -m pytest tests -m gpu -q --durations=4 > ${P}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> ${P}_pytest_gpu.log
tail -14 ${P}_pytest_gpu.log | cut -c1-300
single() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-e2e --no-f32-grade "$@" > ${P}_$name.json 2> ${P}_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/r02n_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print("%s: %.4f ms/step %.3f M samples/s launch_ms %s dw %s kernel %s" % (sys.argv[1], d["ms_per_step"], d["value"] / 1e6, r.get("launch_ms"), r.get("dw_launch_ms"), (r.get("kernel") or "")[:50]))
except Exception as ex:
    print(sys.argv[1], "no line", ex); print(open("gpurun_out/r02n_%s.err" % sys.argv[1]).read()[-800:])
PY
}
single c3_b512_auto X=1 -- --workload c3 --batch 512
single c3_b512_bn256 KUCD_MID_BN=256 -- --workload c3 --batch 512
single c3_b512_nomid KUCD_MID_CHAIN=0 -- --workload c3 --batch 512
single c3_b1024_auto X=1 -- --workload c3 --batch 1024
single c3_b1024_nomid KUCD_MID_CHAIN=0 -- --workload c3 --batch 1024
single c3_b2048_auto X=1 -- --workload c3 --batch 2048
