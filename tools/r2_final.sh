#!/bin/bash
# Round 2, final 1-GPU call: whole GPU suite, smoke(), the default bench line, and the ncu launch list of the same bench command.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 4 gpurun_out/r2f_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1
echo "smoke rc=$?"; tail -n 3 gpurun_out/r2f_smoke.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
echo "bench rc=$?"; wc -l gpurun_out/r2f_bench.json
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
    print(round(d['value'], 1), 'img/s', round(d['ms_per_step'], 2), 'ms  e2e', round(d['e2e']['value'], 1), ' roofline', d['roofline'], ' cpu', d['cpu_baseline'])
    print({k: (round(v['value'], 1) if isinstance(v, dict) and 'value' in v else v) for k, v in d.get('extras', {}).items()})
except Exception as e:
    print('bench failed', e)
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err
echo "reference rc=$?"; cut -c1-400 gpurun_out/r2f_bench_ref.json
FSDET_BENCH_NO_EXTRAS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv \
    --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_ncu_bench.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r2f_launches.csv
