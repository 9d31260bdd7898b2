#!/bin/bash
# gpurun --timeout 1500 -- 'bash scripts/gpu_sanitize.sh' : compute-sanitizer memcheck / racecheck over the small-batch parity tests
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
echo "== memcheck"
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck.log \
    python -m pytest tests/test_forward_dynamics_backward_gpu.py tests/test_forward_dynamics_gpu.py tests/test_backward_gpu.py tests/test_kinematic_state_gpu.py tests/test_table_gpu.py \
    -q -x -k "oracle or golden or reference" 2>&1 | tail -4
echo "memcheck rc=$?"; tail -3 gpurun_out/memcheck.log
echo "== racecheck"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/racecheck.log \
    python -m pytest tests/test_forward_dynamics_backward_gpu.py tests/test_backward_gpu.py -q -x -k "fp64_oracle" 2>&1 | tail -4
echo "racecheck rc=$?"; tail -3 gpurun_out/racecheck.log
echo "== memcheck (forward kernels, host entry point, callers)"
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck2.log \
    python -m pytest tests/test_engine_gpu.py tests/test_callers_gpu.py -q -x -k "not full_size" 2>&1 | tail -4
echo "memcheck2 rc=$?"; grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/memcheck2.log; tail -2 gpurun_out/memcheck2.log
echo "== racecheck (forward kernels)"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/racecheck2.log \
    python -m pytest tests/test_engine_gpu.py tests/test_forward_dynamics_gpu.py -q -x -k "oracle and not full_size" 2>&1 | tail -4
echo "racecheck2 rc=$?"; tail -2 gpurun_out/racecheck2.log
