#!/bin/bash
# ncu launch list + one full capture of the transmission kernels (B200_PROFILING.md recipe). usage: gpu_profile.sh <tag>
TAG=${1:-r01}
CMD="python bench.py --steps 3 --warmup 1 --preroll 4 --cpu-days 1"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_transmit -s 20 -c 4 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
tail -2 gpurun_out/plain_$TAG.log
