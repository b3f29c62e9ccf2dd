#!/bin/bash
# last GPU sanity of the round on the final library
set -u
mkdir -p gpurun_out
P=gpurun_out/r02v
timeout 300 python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py tests/test_gpu_data_path.py -x -q -m gpu > ${P}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> ${P}_pytest_gpu.log
tail -4 ${P}_pytest_gpu.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1 | cut -c1-120
