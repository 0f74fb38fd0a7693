#!/bin/bash
# call 4: elect_one issue pattern (no serialising loops around TMA / MMA), fused hi|lo MMA, lean epilogue
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -q -x > gpurun_out/b4_pytest_tc.log 2>&1
echo "pytest tc rc=$?"; tail -n 6 gpurun_out/b4_pytest_tc.log | cut -c1-300
for f in 0 4 8; do
  timeout 300 python tools/halo_bench.py 10 $f > gpurun_out/b4_halo_sweep_$f.log 2>&1
  echo "sweep flags=$f rc=$?"; cat gpurun_out/b4_halo_sweep_$f.log | cut -c1-200
done
FSDET_BENCH_NO_EXTRAS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b4_bench_halo.json 2> gpurun_out/b4_bench_halo.err
echo "bench(halo) rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/b4_bench_halo.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernels'])"
FSDET_TC_HALO=0 FSDET_BENCH_NO_EXTRAS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b4_bench_nohalo.json 2> gpurun_out/b4_bench_nohalo.err
echo "bench(no halo) rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/b4_bench_nohalo.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernels'])"
FSDET_TC_HALO=0 FSDET_TC_FUSE=0 FSDET_BENCH_NO_EXTRAS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b4_bench_nofuse.json 2> gpurun_out/b4_bench_nofuse.err
echo "bench(no halo, no fuse) rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/b4_bench_nofuse.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernels'])"
timeout 900 python -m pytest tests -q -x -m gpu --deselect tests/test_gpu_tc.py > gpurun_out/b4_pytest_all.log 2>&1
echo "pytest rest rc=$?"; tail -n 8 gpurun_out/b4_pytest_all.log | cut -c1-300
