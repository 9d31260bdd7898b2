#!/bin/bash
# gpurun --timeout 1500 -- 'bash scripts/gpu_fd.sh' : forward-dynamics tests, examples, timings and ncu captures
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== pytest (forward dynamics + examples)"; timeout 900 python -m pytest tests/test_forward_dynamics_gpu.py tests/test_forward_dynamics_backward_gpu.py tests/test_examples_gpu.py -q 2>&1 | tail -15 | tee gpurun_out/pytest_fd.log
echo "== other configs"; timeout 600 python scripts/bench_configs.py 2> gpurun_out/configs.err | tee gpurun_out/bench_configs.json | cut -c1-200
echo "== ncu aba"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:aba -s 2 -c 2 -f -o gpurun_out/aba python scripts/profile_fd.py > gpurun_out/ncu_aba.log 2>&1; tail -3 gpurun_out/ncu_aba.log
