#!/bin/bash
# gpurun --timeout 1800 -- 'bash scripts/gpu_sanitize_r02.sh'
# (1) where does the single CUDA_ERROR_INVALID_HANDLE of the r01 memcheck logs come from?  (2) memcheck / racecheck over the
# kernels added in round 2 (multi-link tree walk, RNEA state dump, fused parametrisation, PDL launches, fused Adam).
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
cat > /tmp/first_launch.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
import differentiable_robot_model_b200 as drm
from differentiable_robot_model_b200 import engine
which = sys.argv[1]
m = drm.DifferentiableKUKAiiwa(device="cuda:0")
q = torch.zeros(8, 7, device="cuda:0")
if which == "fk_first":            # constant table built on the CPU path? no: force the FK kernel to be the library's first launch
    table = torch.zeros(9, 28, device="cuda:0")
    engine.fk_jacobian_raw(m._topology, 8, table, q)
else:
    m.compute_forward_kinematics(q, "iiwa_link_ee")
torch.cuda.synchronize()
print("done", which, engine.launch_count())
PY
for w in table_first fk_first; do
  for mode in LAZY EAGER; do
    CUDA_MODULE_LOADING=$mode timeout 300 compute-sanitizer --tool memcheck --log-file gpurun_out/first_${w}_${mode}.log python /tmp/first_launch.py $w > /dev/null 2>&1
    echo "first launch = $w, CUDA_MODULE_LOADING=$mode: $(grep -c 'Program hit' gpurun_out/first_${w}_${mode}.log) API errors; $(grep 'Host Frame: drm::' gpurun_out/first_${w}_${mode}.log | head -1 | sed 's/=========//'); $(tail -1 gpurun_out/first_${w}_${mode}.log)"
  done
done
echo "== memcheck (round-2 kernels)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck_r02.log \
    python -m pytest tests/test_fk_multi_gpu.py tests/test_kinematic_state_gpu.py tests/test_fused_params_gpu.py tests/test_sharded_gpu.py \
    -q -x -k "not 40000 and not sharded_equals" 2>&1 | tail -3
echo "memcheck rc=$?"; grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/memcheck_r02.log; tail -2 gpurun_out/memcheck_r02.log
echo "== racecheck (round-2 kernels)"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/racecheck_r02.log \
    python -m pytest tests/test_fk_multi_gpu.py tests/test_kinematic_state_gpu.py -q -x -k "bit_for_bit and (1000 or 31) or body_state" 2>&1 | tail -3
echo "racecheck rc=$?"; tail -2 gpurun_out/racecheck_r02.log
echo "== memcheck (PDL chain)"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck_pdl.log \
    python -m pytest tests/test_engine_gpu.py -q -x -k "pdl" 2>&1 | tail -3
tail -2 gpurun_out/memcheck_pdl.log
