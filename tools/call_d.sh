#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_multiscale.py -q -m gpu -k "reader or prefetcher" > gpurun_out/pytest_reader.log 2>&1
tail -n 8 gpurun_out/pytest_reader.log
timeout 240 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01d.json 2> gpurun_out/bench_r01d.err
tail -c 400 gpurun_out/bench_r01d.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_r01d.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'e2e', 'clocks')})
PY
