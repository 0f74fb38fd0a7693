#!/bin/bash
# Round 2, fifth GPU call: green check of everything (BN launch bounds, pool-kernel lanes, fixed tests, thread_local capture).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/r2c5_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2c5_pytest.log | tail -n 20
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c5_smoke.log 2>&1
echo "smoke rc=$?"; tail -n 3 gpurun_out/r2c5_smoke.log | cut -c1-300
FSDET_DUMP_LAUNCHES=1 timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c5_bench.json 2> gpurun_out/r2c5_bench.err
echo "bench rc=$?"; grep -i "failed\|error" gpurun_out/r2c5_bench.err | head -5
timeout 400 python tools/e2e_train_synth.py 512 4 gpurun_out/r2c5_e2e_train.json > gpurun_out/r2c5_e2e_train.log 2>&1
echo "e2e train rc=$?"; grep -v "^class_scale" gpurun_out/r2c5_e2e_train.log | grep -v "nGT" | tail -n 12 | cut -c1-600
grep "nGT" gpurun_out/r2c5_e2e_train.log | sed -n '1p;8p;16p;$p' | cut -c1-200
if grep -q "capture failed" gpurun_out/r2c5_e2e_train.log; then
  FSDET_NO_BG_PREP=1 timeout 400 python tools/e2e_train_synth.py 512 3 gpurun_out/r2c5_e2e_train_nobg.json > gpurun_out/r2c5_e2e_train_nobg.log 2>&1
  echo "e2e (no background prep) rc=$?"; grep -i "capture failed" -A3 gpurun_out/r2c5_e2e_train_nobg.log | head; tail -n 2 gpurun_out/r2c5_e2e_train_nobg.log | cut -c1-400
fi
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2c5_bench.json'))
    print(round(d['value'], 1), 'img/s', round(d['ms_per_step'], 2), 'ms e2e', round(d['e2e']['value'], 1), 'launches', d['gpu_launches'],
          {k: round(v['ms_per_step'], 2) for k, v in d['roofline']['kernels'].items()})
    print('cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'], 3), {k: (round(v['value'], 1), round(v['ms_per_step'], 2)) if 'value' in v else v for k, v in d['extras'].items()})
except Exception as e:
    print('bench failed', e)
PY
