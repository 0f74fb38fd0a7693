#!/bin/bash
# call E: first-layer forward kernel at 3 CTAs / SM (80 registers) against 2 (122 registers)
mkdir -p gpurun_out
FSDET_FIRST_MINB=3 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "first_layer" > gpurun_out/bE_pytest.log 2>&1
echo "pytest first (minb 3) rc=$?"; tail -n 2 gpurun_out/bE_pytest.log | cut -c1-200
for cfg in "FSDET_FIRST_MINB=3" "FSDET_FIRST_MINB=2"; do
  env $cfg FSDET_BENCH_NO_EXTRAS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bE_bench.json 2> gpurun_out/bE_bench.err
  echo "bench [$cfg] rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bE_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], {k: round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
done
