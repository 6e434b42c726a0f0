#!/bin/bash
# Run on the GPU box (via gpurun): plain runs first, then the ncu launch list of the bench command and one full capture per kernel.
# Usage: scripts/gpu_profile_r02.sh <tag> [kernels: admm pcw wbc ...]
TAG=${1:-r02}
shift
KERNELS=${@:-admm pcw wbc}
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-closed-loop"
mkdir -p gpurun_out
$B > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/ncu_launch_$TAG.log 2>&1
for K in $KERNELS; do
  case $K in
    admm) CMD="$B"; RE="regex:mpc_admm_kernelILi64"; SKIP=3 ;;
    pcw) CMD="$B --workload config3"; RE="regex:mpc_pcw_kernel"; SKIP=2 ;;
    wbc) CMD="$B --workload config4"; RE="regex:wbc_qp_kernel"; SKIP=2 ;;
    seq) CMD="python scripts/sharded_e2e.py 1 65536"; RE="regex:sequencer_warp_kernel"; SKIP=1 ;;
  esac
  $CMD > gpurun_out/plain_${TAG}_$K.log 2>&1 &&
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k $RE -s $SKIP -c 1 \
      -o gpurun_out/prof_${TAG}_$K -f $CMD > gpurun_out/ncu_full_${TAG}_$K.log 2>&1
done
ls -la gpurun_out | tail -20
