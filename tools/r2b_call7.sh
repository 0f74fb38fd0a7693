#!/bin/bash
# call 7: weight gradients on a second stream (overlap with the BatchNorm backward passes), wave-aware split-K
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -q -x -k "wgrad" > gpurun_out/b7_pytest_wg.log 2>&1
echo "pytest wgrad rc=$?"; tail -n 3 gpurun_out/b7_pytest_wg.log | cut -c1-300
for cfg in "" "FSDET_WGRAD_STREAM=0"; do
  env $cfg FSDET_BENCH_NO_EXTRAS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b7_bench.json 2> gpurun_out/b7_bench.err
  echo "bench [$cfg] rc=$?"; tail -n 3 gpurun_out/b7_bench.err | cut -c1-300; python -c "
import json; d=json.loads(open('gpurun_out/b7_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], {k: round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()}, d['gpu_launches'])"
done
timeout 1200 python -m pytest tests -q -x -m gpu --deselect tests/test_gpu_tc.py --deselect tests/test_gpu_kernels.py > gpurun_out/b7_pytest_all.log 2>&1
echo "pytest rest rc=$?"; tail -n 8 gpurun_out/b7_pytest_all.log | cut -c1-300
