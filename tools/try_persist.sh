#!/bin/bash
# Next round, early: does the experimental persistent conv kernel (FSDET_TC_PERSIST=1, csrc/conv_tc.cu) work, and
# is it faster?  Everything under `timeout`: a wrong barrier phase means a spinning kernel, not a crash.
#   gpurun --timeout 400 -- 'bash tools/try_persist.sh'
mkdir -p gpurun_out
export FSDET_TC_PERSIST=1
timeout 120 python -m pytest tests/test_gpu_tc.py -q -m gpu -x -k "conv_tc_fwd or padded" > gpurun_out/persist_tc.log 2>&1
echo "tc tests rc=$?"; tail -n 5 gpurun_out/persist_tc.log
timeout 150 python -m pytest tests/test_gpu_model.py -q -m gpu -x -k "mini_all_tensors or full416_digest" > gpurun_out/persist_model.log 2>&1
echo "model tests rc=$?"; tail -n 5 gpurun_out/persist_model.log
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_persist.json 2> gpurun_out/bench_persist.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_persist.json'))
    print('persist:', d['value'], 'img/s', d['ms_per_step'], 'ms', {k: v['ms_per_step'] for k, v in d['roofline']['kernels'].items()})
except Exception as e:
    print('bench failed', e)
PY
unset FSDET_TC_PERSIST
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nopersist.json 2> /dev/null
python -c "
import json; d = json.load(open('gpurun_out/bench_nopersist.json')); print('baseline:', d['value'], 'img/s', d['ms_per_step'], 'ms')"
