#!/bin/bash
# GPU sanity of the split build (chain kernels launched through cudaLaunchKernelExC)
set -u
mkdir -p gpurun_out
P=gpurun_out/r02s
timeout 900 python -m pytest tests -x -q -m gpu > ${P}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> ${P}_pytest_gpu.log
tail -5 ${P}_pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1 | cut -c1-200
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > ${P}_bench_c3.json 2> ${P}_bench_c3.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02s_bench_c3.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "steps", "gpu_launches")}, "e2e", d["e2e"]["value"], "launch_ms", d["roofline"]["launch_ms"], "f32_grade", (d.get("f32_grade") or {}).get("value"))
PY
