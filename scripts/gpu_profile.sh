#!/bin/bash
# Run on the GPU box (via gpurun): plain bench, then the ncu launch list and one full capture of the dominant kernel.
# Usage: scripts/gpu_profile.sh <tag>
TAG=${1:-r01}
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:mpc_admm_kernelILi64 -s 3 -c 1 \
    -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out
