#!/bin/bash
# call 9: evidence - ncu --set full of the tensor-core kernels and of the first-layer / BatchNorm kernels (one eager step)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k 'regex:conv_tc_kernel|conv_halo_kernel|wgrad_tc_kernel|conv_first|bn_act_pool_kernel|bn_act_bwd_pool_kernel' \
    --launch-skip 120 --launch-count 100 -f -o gpurun_out/b9_full python tools/one_step.py 2 64 20 > gpurun_out/b9_ncu.log 2>&1
echo "ncu rc=$?"; tail -n 3 gpurun_out/b9_ncu.log | cut -c1-200
ncu -i gpurun_out/b9_full.ncu-rep --page raw --csv > gpurun_out/b9_full.raw.csv 2>/dev/null
ls -la gpurun_out/b9_full*; rm -f gpurun_out/b9_full.ncu-rep
FSDET_BENCH_NO_EXTRAS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b9_bench.json 2> gpurun_out/b9_bench.err
echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/b9_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], {k: round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
