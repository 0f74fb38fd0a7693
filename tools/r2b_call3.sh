#!/bin/bash
# call 3: lean epilogue (8 epilogue warps in the halo kernel), tests, flag sweep, sanitizer on the fused-MMA variant
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -q -x > gpurun_out/b3_pytest_tc.log 2>&1
echo "pytest tc rc=$?"; tail -n 6 gpurun_out/b3_pytest_tc.log | cut -c1-300
timeout 300 python tools/halo_bench.py 10 0,1,2,8,11 > gpurun_out/b3_halo_sweep.log 2>&1
echo "sweep rc=$?"; cat gpurun_out/b3_halo_sweep.log | cut -c1-250
FSDET_BENCH_NO_EXTRAS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b3_bench_halo.json 2> gpurun_out/b3_bench_halo.err
echo "bench(halo) rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/b3_bench_halo.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernels'])"
FSDET_TC_HALO=0 FSDET_BENCH_NO_EXTRAS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b3_bench_nohalo.json 2> gpurun_out/b3_bench_nohalo.err
echo "bench(no halo) rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/b3_bench_nohalo.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernels'])"
timeout 900 python -m pytest tests -q -x -m gpu --deselect tests/test_gpu_tc.py > gpurun_out/b3_pytest_all.log 2>&1
echo "pytest rest rc=$?"; tail -n 6 gpurun_out/b3_pytest_all.log | cut -c1-300
FSDET_HALO_FLAGS=4 timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/halo_one.py 2 > gpurun_out/b3_sanitizer.log 2>&1
echo "sanitizer rc=$?"; grep -v "^=========     Host Frame\|^=========         " gpurun_out/b3_sanitizer.log | head -n 40 | cut -c1-250
