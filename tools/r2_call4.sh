#!/bin/bash
# Round 2, fourth GPU call: recomputing first-layer kernels (first_tc), fixed tests, e2e train debugging.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_first_tc.py -q -m gpu -x > gpurun_out/r2c4_first.log 2>&1
echo "first_tc tests rc=$?"; tail -n 8 gpurun_out/r2c4_first.log | cut -c1-300
timeout 900 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/r2c4_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2c4_pytest.log | tail -n 20
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c4_smoke.log 2>&1
echo "smoke rc=$?"; tail -n 3 gpurun_out/r2c4_smoke.log
FSDET_DUMP_LAUNCHES=1 timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err
echo "bench rc=$?"; grep -i "failed\|error" gpurun_out/r2c4_bench.err | head -5
FSDET_TC_PARTS=fwd,dgrad,wgrad,head timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c4_bench_nofirst.json 2> /dev/null
timeout 400 python tools/e2e_train_synth.py 512 3 gpurun_out/r2c4_e2e_train.json > gpurun_out/r2c4_e2e_train.log 2>&1
rc=$?; echo "e2e train rc=$rc"; tail -n 25 gpurun_out/r2c4_e2e_train.log | cut -c1-300
if [ "$rc" != "0" ]; then
  FSDET_NO_GRAPH=1 CUDA_LAUNCH_BLOCKING=1 timeout 400 python tools/e2e_train_synth.py 512 2 gpurun_out/r2c4_e2e_train_eager.json > gpurun_out/r2c4_e2e_train_eager.log 2>&1
  echo "e2e eager+blocking rc=$?"; tail -n 25 gpurun_out/r2c4_e2e_train_eager.log | cut -c1-300
fi
python - <<'PY'
import json
for t in ('', '_nofirst'):
    try:
        d = json.load(open('gpurun_out/r2c4_bench%s.json' % t))
        print(t, round(d['value'], 1), 'img/s', round(d['ms_per_step'], 2), 'ms e2e', round(d['e2e']['value'], 1), 'launches', d['gpu_launches'],
              {k: round(v['ms_per_step'], 2) for k, v in d['roofline']['kernels'].items()})
        print('cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'], 3), {k: (round(v['value'], 1), round(v['ms_per_step'], 2)) if 'value' in v else v for k, v in d['extras'].items()})
    except Exception as e:
        print(t, 'bench failed', e)
PY
du -sh gpurun_out
