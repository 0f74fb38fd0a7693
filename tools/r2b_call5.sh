#!/bin/bash
# call 5: fused pairs in the long-K flavour, persistent short-K flavour re-measured, mini-model gradient diagnosis
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -q -x > gpurun_out/b5_pytest_tc.log 2>&1
echo "pytest tc rc=$?"; tail -n 6 gpurun_out/b5_pytest_tc.log | cut -c1-300
timeout 300 python tools/diag_mini.py 6 > gpurun_out/b5_diag_mini.log 2>&1
echo "diag rc=$?"; grep "^seed" gpurun_out/b5_diag_mini.log | cut -c1-250
FSDET_TC_FUSE=0 timeout 300 python tools/diag_mini.py 3 > gpurun_out/b5_diag_mini_nofuse.log 2>&1
echo "diag(nofuse) rc=$?"; grep "^seed" gpurun_out/b5_diag_mini_nofuse.log | cut -c1-250
for cfg in "" "FSDET_TC_FUSE=0" "FSDET_TC_PERSIST=1"; do
  env $cfg FSDET_BENCH_NO_EXTRAS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b5_bench.json 2> gpurun_out/b5_bench.err
  echo "bench [$cfg] rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/b5_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], {k: round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
done
timeout 300 python tools/halo_bench.py 10 0,1,2,3 > gpurun_out/b5_halo_sweep.log 2>&1
echo "sweep rc=$?"; cat gpurun_out/b5_halo_sweep.log | cut -c1-200
