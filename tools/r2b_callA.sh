#!/bin/bash
# call A: whole GPU suite, smoke(), ncu --set full of the tensor-core / first-layer / BatchNorm kernels, short bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/bA_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 5 gpurun_out/bA_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/bA_smoke.log 2>&1
echo "smoke rc=$?"; tail -n 3 gpurun_out/bA_smoke.log | cut -c1-300
timeout 900 ncu --set full --clock-control none -k 'regex:conv_tc_kernel|conv_halo_kernel|wgrad_tc_kernel|conv_first|bn_act_pool_kernel|bn_act_bwd_pool_kernel' \
    --launch-skip 120 --launch-count 100 -f -o gpurun_out/bA_full python tools/one_step.py 2 64 20 > gpurun_out/bA_ncu.log 2>&1
echo "ncu rc=$?"; tail -n 3 gpurun_out/bA_ncu.log | cut -c1-200
ncu -i gpurun_out/bA_full.ncu-rep --page raw --csv > gpurun_out/bA_full.raw.csv 2>/dev/null
ls -la gpurun_out/bA_full*; rm -f gpurun_out/bA_full.ncu-rep
FSDET_BENCH_NO_EXTRAS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bA_bench.json 2> gpurun_out/bA_bench.err
echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bA_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], {k: round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()}, d['gpu_launches'])"
for cfg in "FSDET_TC_SMALLK_MAX=2303" "FSDET_TC_SMALLK_MAX=1151"; do
  env $cfg FSDET_BENCH_NO_EXTRAS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bA_bench2.json 2> gpurun_out/bA_bench2.err
  echo "bench [$cfg] rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bA_bench2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], {k: round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
done
