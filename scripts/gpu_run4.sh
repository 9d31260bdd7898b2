#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== sweep"; timeout 600 python scripts/sweep_fk.py 2> gpurun_out/sweep.err | tee gpurun_out/sweep_fk.json | cut -c1-200
echo "== configs"; timeout 600 python scripts/bench_configs.py 2> gpurun_out/configs.err | tee gpurun_out/bench_configs.json | cut -c1-3000
tail -5 gpurun_out/sweep.err gpurun_out/configs.err
