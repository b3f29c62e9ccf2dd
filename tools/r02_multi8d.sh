#!/bin/bash
# 8 GPUs, one program: C3 strong scaling (512 rows per GPU) with the fused exchange forced (the default rule takes NCCL below 2048 rows)
set -u
mkdir -p gpurun_out
N=$(python -c 'import torch; print(torch.cuda.device_count())')
KUCD_FUSED_MIN_ROWS=1 timeout 200 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N --master-port 29541 \
  bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --workload c3 --scaling strong > gpurun_out/r02p_c3_strong_fused.json 2> gpurun_out/r02p_c3_strong_fused.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02p_c3_strong_fused.json").read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print("c3_strong_fused: %.4f ms/step %.3f M samples/s [%s]" % (d["ms_per_step"], d["value"] / 1e6, d["config"].get("exchange", "")[:40]))
    print("   breakdown", {k: v for k, v in (r.get("step_breakdown_ms") or {}).items() if k != "what"})
except Exception as e:
    print("no line", e); print(open("gpurun_out/r02p_c3_strong_fused.err").read()[-1500:])
PY
