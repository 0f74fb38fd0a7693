#!/bin/bash
# call D: default bench line once more (extras now time every step: min / max)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bD_bench.json 2> gpurun_out/bD_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bD_bench.json').read().strip().splitlines()[-1])
print(round(d['value'], 1), 'img/s', round(d['ms_per_step'], 2), 'ms  e2e', round(d['e2e']['value'], 1), d['clocks'])
for k, v in d.get('extras', {}).items():
    print(k, {a: v.get(a) for a in ('value', 'ms_per_step', 'ms_per_step_min_max', 'error')})
PY
