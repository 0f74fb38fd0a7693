#!/bin/bash
# call F: final state - whole GPU suite, smoke(), the command-line training driver on a synthetic VOC directory
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/bF_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 4 gpurun_out/bF_pytest.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/bF_smoke.log 2>&1
echo "smoke rc=$?"; tail -n 2 gpurun_out/bF_smoke.log | cut -c1-300
timeout 600 python tools/e2e_train_synth.py 2048 4 gpurun_out/bF_e2e_train.json > gpurun_out/bF_e2e_train.log 2>&1
echo "e2e train rc=$?"; grep -v "^class_scale" gpurun_out/bF_e2e_train.log | grep -v nGT | tail -n 6 | cut -c1-400
