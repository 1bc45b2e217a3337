#!/bin/bash
# ncu launch list + one full capture of simulated days of our kernels (B200_PROFILING.md recipe).
# usage (under gpurun): tools/gpu_profile.sh <tag> <preroll days> [bench args...]
# The full capture skips the record pass and the pre-roll of the value pass (13 launches per day) and takes two days.
TAG=${1:-r02}
PRE=${2:-4}
shift; shift
CMD="python bench.py --steps 3 --warmup 3 --preroll $PRE --cpu-days 1 $*"
SKIP=$(( (PRE + 6) * 13 + PRE * 13 + 3 * 13 + 8 ))
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:^k_ -s $SKIP -c 26 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
tail -1 gpurun_out/plain_$TAG.log | cut -c1-400
