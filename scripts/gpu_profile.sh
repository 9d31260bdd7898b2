#!/bin/bash
# gpurun --timeout 2400 -- 'bash scripts/gpu_profile.sh "fk_tree rnea_bwd ..."'
# ncu --set full captures of the named kernel groups (one GPU) + the steady-state FK capture at the contract batch.
set -u
SETS=${1:-"fk_tree fk_allegro rnea rnea_bwd fk_bwd aba fk_steady"}
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
for g in $SETS; do
  case $g in
    fk_tree)    K="regex:fk_tree_kernel"; S=1; C=2;;
    fk_allegro) K="regex:fk_tree_kernel|fk_jacobian_kernel"; S=2; C=2;;
    rnea)       K="regex:rnea_kernel"; S=1; C=1;;
    rnea_bwd)   K="regex:rnea_backward"; S=2; C=2;;
    fk_bwd)     K="regex:fk_jacobian_backward"; S=1; C=1;;
    aba)        K="regex:aba"; S=2; C=2;;
    fk_steady)  K="regex:fk_jacobian_kernel"; S=60; C=3;;
  esac
  if [ "$g" = "fk_steady" ]; then
    # steady state at the contract batch: launches 60.. of the rotation (buffers > L2), caches NOT flushed between replays
    timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k $K -s $S -c $C -f -o gpurun_out/ncu_$g \
        python bench.py --steps 96 --warmup 3 --reps 1 --no-cpu-baseline --no-e2e --no-large --no-modes > gpurun_out/ncu_$g.log 2>&1
  else
    WHICH=$g timeout 900 ncu --set full --clock-control none --import-source on -k $K -s $S -c $C -f -o gpurun_out/ncu_$g \
        python scripts/profile_kernels.py > gpurun_out/ncu_$g.log 2>&1
  fi
  tail -2 gpurun_out/ncu_$g.log
  # condense on the box (gpurun_out/ is capped at 64 MiB): metrics JSON + per-instruction source page of the first capture
  python scripts/ncu_summary.py gpurun_out/ncu_$g.summary.json $g=gpurun_out/ncu_$g.ncu-rep > gpurun_out/ncu_$g.summary.txt 2>&1
  ncu -i gpurun_out/ncu_$g.ncu-rep --page details --csv 2>/dev/null | grep -E "Stall|Issue Slots|Eligible|No Eligible|Active Warps|Theoretical|Achieved Occupancy|L2 Cache Throughput|DRAM Throughput|Registers|Bank" | cut -c1-300 | head -80 > gpurun_out/ncu_$g.details.txt
  ncu -i gpurun_out/ncu_$g.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rd=list(csv.reader(sys.stdin))
head=rd[0]
keep=[i for i,h in enumerate(head) if 'smsp__average_warp' in h or 'warps_issue_stalled' in h or h in ('Kernel Name','smsp__warps_eligible.avg.per_cycle_active','smsp__warps_active.avg.per_cycle_active')]
for r in rd[2:]:
    print({head[i]: r[i] for i in keep})
" > gpurun_out/ncu_$g.stalls.txt 2>&1
  rm -f gpurun_out/ncu_$g.ncu-rep
done
ls -la gpurun_out/
