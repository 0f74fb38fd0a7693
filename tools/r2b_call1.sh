#!/bin/bash
# Round 2 (second session), call 1: halo-tile kernel - GPU tests, A/B micro-benchmark, short bench with and without it.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -q -x -k "halo" > gpurun_out/b1_pytest_halo.log 2>&1
echo "pytest halo rc=$?"; tail -n 15 gpurun_out/b1_pytest_halo.log | cut -c1-300
timeout 300 python tools/halo_bench.py 10 > gpurun_out/b1_halo_bench.log 2>&1
echo "halo_bench rc=$?"; cat gpurun_out/b1_halo_bench.log | cut -c1-200
FSDET_BENCH_NO_EXTRAS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b1_bench_halo.json 2> gpurun_out/b1_bench_halo.err
echo "bench(halo) rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/b1_bench_halo.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernels'])"
FSDET_TC_HALO=0 FSDET_BENCH_NO_EXTRAS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b1_bench_nohalo.json 2> gpurun_out/b1_bench_nohalo.err
echo "bench(no halo) rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/b1_bench_nohalo.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernels'])"
timeout 900 python -m pytest tests -q -x -m gpu > gpurun_out/b1_pytest_all.log 2>&1
echo "pytest all rc=$?"; tail -n 6 gpurun_out/b1_pytest_all.log | cut -c1-300
