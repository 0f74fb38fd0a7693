#!/bin/bash
# Round 2, first GPU call: new unified conv_tc kernel (term modes, fused BN statistics, persistent loop), precision budget.
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_tc.py -q -m gpu -x --durations=5 > gpurun_out/r2c1_tc.log 2>&1
echo "tc tests rc=$?"; tail -n 6 gpurun_out/r2c1_tc.log
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_region.py -q -m gpu -x --durations=5 > gpurun_out/r2c1_model.log 2>&1
echo "model tests rc=$?"; tail -n 6 gpurun_out/r2c1_model.log
timeout 300 python tools/precision_budget.py 16 20 gpurun_out/precision_budget_r02.json > gpurun_out/r2c1_prec.log 2>&1
echo "precision rc=$?"; tail -n 22 gpurun_out/r2c1_prec.log | cut -c1-400
timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c1_bench_base.json 2> gpurun_out/r2c1_bench_base.err
FSDET_TC_PERSIST=1 timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c1_bench_persist.json 2> gpurun_out/r2c1_bench_persist.err
FSDET_TC_PERSIST=1 timeout 200 python -m pytest tests/test_gpu_model.py -q -m gpu -x -k "mini_all_tensors or full416_digest or cuda_graph" > gpurun_out/r2c1_model_persist.log 2>&1
echo "persist model tests rc=$?"; tail -n 4 gpurun_out/r2c1_model_persist.log
python - <<'PY'
import json
for t in ('base', 'persist'):
    try:
        d = json.load(open('gpurun_out/r2c1_bench_%s.json' % t))
        print(t, round(d['value'], 1), 'img/s', round(d['ms_per_step'], 2), 'ms e2e', round(d['e2e']['value'], 1),
              {k: round(v['ms_per_step'], 2) for k, v in d['roofline']['kernels'].items()})
    except Exception as e:
        print(t, 'bench failed', e)
PY
