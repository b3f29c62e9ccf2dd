#!/bin/bash
# last single-GPU call of the round: the final build exactly as the driver will run it
set -u
mkdir -p gpurun_out
P=gpurun_out/r02q
timeout 900 python -m pytest tests -x -q -m gpu > ${P}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> ${P}_pytest_gpu.log
tail -6 ${P}_pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${P}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > ${P}_bench_default.json 2> ${P}_bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02q_bench_default.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "steps", "gpu_launches", "dtype")})
print("e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "h2d_bytes_per_step")})
print("roofline", {k: d["roofline"].get(k) for k in ("achieved", "peak", "frac", "regime", "frac_of_sustained_peak", "launch_ms", "traffic")})
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "f32_grade", (d.get("f32_grade") or {}).get("value"))
print("clocks", d["clocks"])
PY
export KUCD_COOP=0
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-f32-grade"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${P}_c3_launches.csv $CMD > ${P}_ncu_list.log 2>&1
echo "ncu list rc=$?"
