#!/bin/bash
# Round 2, 2-GPU call: all-reduce correctness + in-graph overlap, weak scaling 1 -> 2.
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/r2m_multi.log 2>&1
echo "multi test rc=$?"; tail -n 12 gpurun_out/r2m_multi.log | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_bench1.json 2> gpurun_out/r2m_bench1.err
FSDET_BENCH_NO_EXTRAS=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2m_bench2.json 2> gpurun_out/r2m_bench2.err
echo "bench2 rc=$?"; grep -i "in-graph\|falling back\|NVLS\|Connected all\|error" gpurun_out/r2m_bench2.err | head -12 | cut -c1-250
python - <<'PY'
import json
for t in ('1', '2'):
    try:
        d = json.loads([l for l in open('gpurun_out/r2m_bench%s.json' % t) if l.startswith('{')][-1])
        print(t, round(d['value'], 1), 'img/s', round(d['ms_per_step'], 2), 'ms e2e', round(d['e2e']['value'], 1))
    except Exception as e:
        print(t, 'bench failed', e)
PY
