#!/bin/bash
# One gpurun call: bench line, ncu launch list of an eager step, one `--set full` capture of the tcgen05 conv kernels.
# Outputs land in gpurun_out/ (copied into profiles/ by hand after reading them).
#   gpurun --timeout 560 -- 'bash tools/profile_round.sh r01b [pytest-seconds]'
TAG=${1:-r01}
PYT=${2:-0}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 600 gpurun_out/bench_$TAG.err
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    python tools/one_step.py 2 64 20 > gpurun_out/launches_$TAG.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k 'regex:conv_tc_kernel|wgrad_tc_kernel' \
    --launch-skip 22 --launch-count 14 -f -o gpurun_out/prof_tc_$TAG python tools/one_step.py 1 64 20 > gpurun_out/prof_tc_$TAG.log 2>&1
ncu -i gpurun_out/prof_tc_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_tc_$TAG.raw.csv 2>/dev/null
if [ "$PYT" != "0" ]; then
  timeout $PYT python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/pytest_$TAG.log 2>&1
  tail -n 25 gpurun_out/pytest_$TAG.log
fi
head -c 1500 gpurun_out/bench_$TAG.json
