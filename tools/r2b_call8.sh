#!/bin/bash
# call 8: A_lo * B_hi accumulates away from the wide MMA's columns (dependent tcgen05.mma chains)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -q -x > gpurun_out/b8_pytest_tc.log 2>&1
echo "pytest tc rc=$?"; tail -n 4 gpurun_out/b8_pytest_tc.log | cut -c1-300
timeout 300 python tools/halo_bench.py 10 0,4 > gpurun_out/b8_halo_sweep.log 2>&1
echo "sweep rc=$?"; cat gpurun_out/b8_halo_sweep.log | cut -c1-200
for cfg in "" "FSDET_TC_FUSE=0"; do
  env $cfg FSDET_BENCH_NO_EXTRAS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b8_bench.json 2> gpurun_out/b8_bench.err
  echo "bench [$cfg] rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/b8_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], {k: round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
done
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_zz_configs.py -q -x > gpurun_out/b8_pytest_model.log 2>&1
echo "pytest model rc=$?"; tail -n 4 gpurun_out/b8_pytest_model.log | cut -c1-300
