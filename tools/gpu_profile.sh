#!/bin/bash
# ncu launch list + one full capture of one simulated day of our kernels (B200_PROFILING.md recipe).
# usage (under gpurun): tools/gpu_profile.sh <tag> [bench args...]
TAG=${1:-r01}
shift
CMD="python bench.py --steps 3 --warmup 3 --preroll 4 --cpu-days 1 $*"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "launch list rc=$?"
# one whole simulated day of the value pass (the record pass is 10 days of ~21 launches; skip well into the value pass)
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:^k_ -s 330 -c 24 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
tail -1 gpurun_out/plain_$TAG.log
