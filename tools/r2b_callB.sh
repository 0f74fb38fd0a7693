#!/bin/bash
# call B: the default bench line (extras, CPU arm), the reference arm, the ncu launch list of the bench command, short-K threshold A/B
mkdir -p gpurun_out
for cfg in "FSDET_TC_SMALLK_MAX=2303" "FSDET_TC_SMALLK_MAX=1151"; do
  env $cfg FSDET_BENCH_NO_EXTRAS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bB_bench2.json 2> gpurun_out/bB_bench2.err
  echo "bench [$cfg] rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bB_bench2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], {k: round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
done
timeout 900 python bench.py > gpurun_out/bB_bench.json 2> gpurun_out/bB_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bB_bench.json').read().strip().splitlines()[-1])
    print(round(d['value'], 1), 'img/s', round(d['ms_per_step'], 2), 'ms  e2e', round(d['e2e']['value'], 1), d['clocks'], ' roofline', {k: d['roofline'][k] for k in ('kernel', 'achieved', 'frac', 'traffic', 'share_of_step')}, ' cpu', d['cpu_baseline'])
    print({k: (round(v['value'], 1) if isinstance(v, dict) and 'value' in v else v) for k, v in d.get('extras', {}).items()})
except Exception as e:
    print('bench failed', e)
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bB_bench_ref.json 2> gpurun_out/bB_bench_ref.err
echo "reference rc=$?"; cut -c1-400 gpurun_out/bB_bench_ref.json
FSDET_BENCH_NO_EXTRAS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv \
    --log-file gpurun_out/bB_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bB_ncu_bench.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/bB_launches.csv
