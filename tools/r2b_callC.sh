#!/bin/bash
# call C (2 GPUs): the multi-GPU parity test and the 2-GPU bench line (torchrun, NCCL)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x -m gpu > gpurun_out/bC_pytest_multi.log 2>&1
echo "pytest multi rc=$?"; tail -n 5 gpurun_out/bC_pytest_multi.log | cut -c1-300
FSDET_BENCH_NO_EXTRAS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bC_bench2.json 2> gpurun_out/bC_bench2.err
echo "bench 2 GPUs rc=$?"; tail -n 3 gpurun_out/bC_bench2.err | cut -c1-300; python -c "
import json; d=json.loads(open('gpurun_out/bC_bench2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['n_gpus'], d['e2e']['value'])"
