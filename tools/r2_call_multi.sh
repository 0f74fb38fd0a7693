#!/bin/bash
# Round 2, 2-GPU call: all-reduce correctness + in-graph overlap, weak scaling 1 -> 2 (bench at N=1 comes from the 1-GPU calls).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > gpurun_out/r2m_multi.log 2>&1
echo "multi test rc=$?"; tail -n 6 gpurun_out/r2m_multi.log | cut -c1-300
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2m_bench2.json 2> gpurun_out/r2m_bench2.err
echo "bench2 rc=$?"; grep -i "in-graph\|falling back\|NVLS\|Connected all\|nChannels\|error" gpurun_out/r2m_bench2.err | head -8 | cut -c1-250
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/r2m_bench2.json') if l.startswith('{')][-1])
    print(2, round(d['value'], 1), 'img/s', round(d['ms_per_step'], 2), 'ms e2e', round(d['e2e']['value'], 1))
except Exception as e:
    print('bench failed', e)
PY
wc -l gpurun_out/r2m_bench2.json
