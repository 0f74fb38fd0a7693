#!/bin/bash
# Round 2, eighth GPU call: end-to-end loop throughput after the async input path; bench after the BN grid change.
mkdir -p gpurun_out
timeout 600 python tools/e2e_train_synth.py 2048 5 gpurun_out/r2c8_e2e_train.json > gpurun_out/r2c8_e2e_train.log 2>&1
echo "e2e train rc=$?"; grep -i "capture failed" -A3 gpurun_out/r2c8_e2e_train.log | head -8; grep -v "^class_scale" gpurun_out/r2c8_e2e_train.log | grep -v nGT | tail -n 9 | cut -c1-500
FSDET_BENCH_NO_EXTRAS=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2c8_bench.json 2> gpurun_out/r2c8_bench.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2c8_bench.json'))
    print(round(d['value'], 1), 'img/s', round(d['ms_per_step'], 2), 'ms e2e', round(d['e2e']['value'], 1), {k: round(v['ms_per_step'], 2) for k, v in d['roofline']['kernels'].items()})
except Exception as e:
    print('bench failed', e)
PY
