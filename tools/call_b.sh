#!/bin/bash
# gpurun call: the detection tests + smoke + (optionally) the whole GPU suite and the bench line
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_detect.py -q -m gpu --durations=5 > gpurun_out/pytest_detect.log 2>&1
tail -n 30 gpurun_out/pytest_detect.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -n 3 gpurun_out/smoke.log
