#!/bin/bash
# gpurun --gpus N --timeout 1500 -- 'bash scripts/gpu_scale.sh "1 2 4 8" [steps] [warmup]'
# The driver's scaling protocol: bench.py at N = 1, 2, ... back to back on one box, K steps, W warm-up.
set -u
NS=${1:-"1 2"}
K=${2:-20}
W=${3:-5}
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi -L > gpurun_out/gpus.txt
for n in $NS; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps $K --warmup $W --no-cpu-baseline 2> gpurun_out/scale_n$n.err > gpurun_out/scale_n$n.json
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n \
        bench.py --gpus $n --steps $K --warmup $W --no-cpu-baseline 2> gpurun_out/scale_n$n.err > gpurun_out/scale_n$n.json
  fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_n$n.json").read().strip().splitlines()[-1])
    print("N=$n value %.4g  ms_per_step %.5f  e2e %.4g  regions %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["timing"]["region_ms_this_rank"]))
except Exception as e:
    print("N=$n failed:", e); print(open("gpurun_out/scale_n$n.err").read()[-1500:])
PY
done
