#!/bin/bash
# First GPU call of a round (one box, about 6 minutes):  gpurun --timeout 900 -- 'bash tools/first_gpu.sh r02a'
#   1. the whole -m gpu suite (log kept)          2. smoke()          3. the default bench line (plain run)
#   4. tools/gpu_profile.sh: ncu launch list + one --set full capture of a simulated day (only after 3. exited 0)
# Results land in gpurun_out/; summarise the ncu files here with tools/ncu_summary.py and copy what is to be judged
# into profiles/.
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_$TAG.log 2>&1
echo "gpu tests rc=$? : $(tail -1 gpurun_out/gpu_tests_$TAG.log)"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1
echo "smoke rc=$? : $(tail -1 gpurun_out/smoke_$TAG.log)"
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
rc=$?
echo "bench rc=$rc"
[ $rc -eq 0 ] && bash tools/gpu_profile.sh $TAG 18   # days 24-26: about 15 % prevalence
