#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_multiscale.py -q -m gpu --durations=5 > gpurun_out/pytest_multiscale.log 2>&1
tail -n 25 gpurun_out/pytest_multiscale.log
timeout 240 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01c.json 2> gpurun_out/bench_r01c.err
tail -c 400 gpurun_out/bench_r01c.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_r01c.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'e2e', 'detect_nms_ms', 'build_targets_ms', 'clocks', 'cpu_baseline')})
PY
