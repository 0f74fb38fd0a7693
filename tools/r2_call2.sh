#!/bin/bash
# Round 2, second GPU call: whole GPU suite, smoke, bench (+extras, per-launch dump), launch list, ncu --set full, e2e train.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --durations=15 > gpurun_out/r2c2_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 25 gpurun_out/r2c2_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c2_smoke.log 2>&1
echo "smoke rc=$?"; tail -n 3 gpurun_out/r2c2_smoke.log
FSDET_DUMP_LAUNCHES=1 timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r2c2_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c2_launches.csv \
    python tools/one_step.py 2 64 20 > gpurun_out/r2c2_launches.log 2>&1
echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:conv_tc_kernel|wgrad_tc_kernel' \
    --launch-skip 85 --launch-count 60 -f -o gpurun_out/r2c2_prof_tc python tools/one_step.py 2 64 20 > gpurun_out/r2c2_prof_tc.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/r2c2_prof_tc.ncu-rep --page raw --csv > gpurun_out/r2c2_prof_tc.raw.csv 2>/dev/null
timeout 400 python tools/e2e_train_synth.py 512 3 gpurun_out/r2c2_e2e_train.json > gpurun_out/r2c2_e2e_train.log 2>&1
echo "e2e train rc=$?"; tail -n 4 gpurun_out/r2c2_e2e_train.log | cut -c1-600
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2c2_bench.json'))
    print(round(d['value'], 1), 'img/s', round(d['ms_per_step'], 2), 'ms e2e', round(d['e2e']['value'], 1), 'launches', d['gpu_launches'],
          {k: round(v['ms_per_step'], 2) for k, v in d['roofline']['kernels'].items()})
    print('cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'], 3), {k: (round(v['value'], 1), round(v['ms_per_step'], 2)) if 'value' in v else v for k, v in d['extras'].items()})
except Exception as e:
    print('bench failed', e)
PY
