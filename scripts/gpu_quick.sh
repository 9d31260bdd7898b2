#!/bin/bash
# gpurun --timeout 1800 -- 'bash scripts/gpu_quick.sh' : full GPU test suite + secondary timings (no ncu)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== other configs"; timeout 600 python scripts/bench_configs.py 2> gpurun_out/configs.err | tee gpurun_out/bench_configs.json | cut -c1-200
tail -5 gpurun_out/configs.err
