#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_multiscale.py -q -m gpu -k "reader or prefetcher" > gpurun_out/pytest_reader.log 2>&1
tail -n 5 gpurun_out/pytest_reader.log
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01e.json 2> gpurun_out/bench_r01e.err
tail -c 300 gpurun_out/bench_r01e.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_r01e.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'e2e', 'clocks')})
PY
timeout ${1:-200} python -m pytest tests -q -m gpu -x --durations=15 > gpurun_out/pytest_full.log 2>&1
echo "pytest rc=$?"
tail -n 30 gpurun_out/pytest_full.log
