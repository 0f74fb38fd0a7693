#!/bin/bash
# Round 2, seventh GPU call: green check + end-to-end training loop throughput.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/r2c7_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2c7_pytest.log | tail -n 12
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c7_smoke.log 2>&1
echo "smoke rc=$?"; tail -n 1 gpurun_out/r2c7_smoke.log | cut -c1-300
timeout 500 python tools/e2e_train_synth.py 1024 6 gpurun_out/r2c7_e2e_train.json > gpurun_out/r2c7_e2e_train.log 2>&1
echo "e2e train rc=$?"; grep -i "capture failed" -A3 gpurun_out/r2c7_e2e_train.log | head -8; grep -v "^class_scale" gpurun_out/r2c7_e2e_train.log | grep -v nGT | tail -n 8 | cut -c1-500
grep "nGT" gpurun_out/r2c7_e2e_train.log | sed -n '1p;16p;48p;$p' | cut -c1-220
