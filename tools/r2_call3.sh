#!/bin/bash
# Round 2, third GPU call: rerun of call 2 (its outputs were lost: gpurun_out over 64 MiB) with fixes + diagnostics.
mkdir -p gpurun_out
timeout 200 python tools/diag_smoke.py 16 128 2 > gpurun_out/r2c3_diag.log 2>&1; tail -n 24 gpurun_out/r2c3_diag.log | cut -c1-200
timeout 900 python -m pytest tests -q -m gpu --durations=15 > gpurun_out/r2c3_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2c3_pytest.log | tail -n 20
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c3_smoke.log 2>&1
echo "smoke rc=$?"; tail -n 3 gpurun_out/r2c3_smoke.log
FSDET_DUMP_LAUNCHES=1 timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err
echo "bench rc=$?"; grep -i "failed\|error" gpurun_out/r2c3_bench.err | head -5
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c3_launches.csv \
    python tools/one_step.py 2 64 20 > gpurun_out/r2c3_launches.log 2>&1
echo "launch list rc=$?"
timeout 500 ncu --set full --clock-control none -k 'regex:conv_tc_kernel|wgrad_tc_kernel|bn_act' \
    --launch-skip 160 --launch-count 110 -f -o /tmp/r2c3_prof python tools/one_step.py 2 64 20 > gpurun_out/r2c3_prof.log 2>&1
echo "ncu full rc=$?"
ncu -i /tmp/r2c3_prof.ncu-rep --page raw --csv > gpurun_out/r2c3_prof.raw.csv 2>/dev/null
ls -la /tmp/r2c3_prof.ncu-rep gpurun_out/r2c3_prof.raw.csv
CUDA_LAUNCH_BLOCKING=1 timeout 500 python tools/e2e_train_synth.py 512 3 gpurun_out/r2c3_e2e_train.json > gpurun_out/r2c3_e2e_train.log 2>&1
echo "e2e train rc=$?"; tail -n 12 gpurun_out/r2c3_e2e_train.log | cut -c1-400
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2c3_bench.json'))
    print(round(d['value'], 1), 'img/s', round(d['ms_per_step'], 2), 'ms e2e', round(d['e2e']['value'], 1), 'launches', d['gpu_launches'],
          {k: round(v['ms_per_step'], 2) for k, v in d['roofline']['kernels'].items()})
    print('cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'], 3), {k: (round(v['value'], 1), round(v['ms_per_step'], 2)) if 'value' in v else v for k, v in d['extras'].items()})
except Exception as e:
    print('bench failed', e)
PY
du -sh gpurun_out
