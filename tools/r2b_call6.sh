#!/bin/bash
# call 6: 256-wide weight-gradient tiles for every channel count, BatchNorm statistics inside the first-layer kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_kernels.py -q -x > gpurun_out/b6_pytest_tc.log 2>&1
echo "pytest tc+kernels rc=$?"; tail -n 6 gpurun_out/b6_pytest_tc.log | cut -c1-300
FSDET_BENCH_NO_EXTRAS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b6_bench.json 2> gpurun_out/b6_bench.err
echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/b6_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], {k: round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()}, d['gpu_launches'])"
timeout 900 python -m pytest tests -q -x -m gpu --deselect tests/test_gpu_tc.py --deselect tests/test_gpu_kernels.py > gpurun_out/b6_pytest_all.log 2>&1
echo "pytest rest rc=$?"; tail -n 8 gpurun_out/b6_pytest_all.log | cut -c1-300
FSDET_BENCH_NO_EXTRAS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/b6_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b6_ncu_bench.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/b6_launches.csv
