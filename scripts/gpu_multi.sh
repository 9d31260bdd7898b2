#!/bin/bash
# gpurun --gpus N --timeout 1500 -- 'bash scripts/gpu_multi.sh N'
set -u
N=${1:-2}
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi -L | tee gpurun_out/gpus_$N.txt
echo "== sharded correctness (N=$N)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    scripts/check_sharded.py 2>&1 | tail -5 | tee gpurun_out/check_sharded_$N.log
for n in 1 $N; do
  echo "== bench --gpus $n"
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --no-cpu-baseline 2> gpurun_out/bench_multi.err | tee gpurun_out/bench_gpus1.json | cut -c1-300
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 \
        bench.py --gpus $n --no-cpu-baseline 2>> gpurun_out/bench_multi.err | tee gpurun_out/bench_gpus$n.json | cut -c1-300
  fi
done
tail -5 gpurun_out/bench_multi.err
