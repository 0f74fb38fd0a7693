#!/bin/bash
# call G: CTA pairs sharing the weight tile (TMA multicast) re-measured on the new issue path
mkdir -p gpurun_out
for cfg in "FSDET_TC_CLUSTER=1" "FSDET_TC_CLUSTER=0"; do
  env $cfg FSDET_BENCH_NO_EXTRAS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bG_bench.json 2> gpurun_out/bG_bench.err
  echo "bench [$cfg] rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bG_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], {k: round(v['ms_per_step'],3) for k,v in d['roofline']['kernels'].items()})"
done
