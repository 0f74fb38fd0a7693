#!/bin/bash
# Round 2, sixth GPU call: cluster-multicast flavour of conv_tc, e2e capture diagnosis.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tc.py -q -m gpu -x -k "conv_tc_fwd" > gpurun_out/r2c6_tc.log 2>&1
echo "tc tests rc=$?"; tail -n 8 gpurun_out/r2c6_tc.log | cut -c1-300
FSDET_TC_CLUSTER=1 timeout 300 python -m pytest tests/test_gpu_model.py tests/test_gpu_zz_configs.py -q -m gpu -x -k "mini_all_tensors or full416_digest or cuda_graph or configs1" > gpurun_out/r2c6_model_cluster.log 2>&1
echo "cluster model tests rc=$?"; tail -n 4 gpurun_out/r2c6_model_cluster.log | cut -c1-300
FSDET_BENCH_NO_EXTRAS=1 FSDET_DUMP_LAUNCHES=1 FSDET_TC_CLUSTER=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c6_bench_cluster.json 2> gpurun_out/r2c6_bench_cluster.err
FSDET_BENCH_NO_EXTRAS=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c6_bench_base.json 2> gpurun_out/r2c6_bench_base.err
TORCH_SHOW_CPP_STACKTRACES=1 FSDET_STRICT_CAPTURE=1 timeout 300 python tools/e2e_train_synth.py 512 2 > gpurun_out/r2c6_e2e_diag.log 2>&1
echo "e2e diag rc=$?"; grep -n "frame #" gpurun_out/r2c6_e2e_diag.log | head -40 | cut -c1-260
python - <<'PY'
import json
for t in ('cluster', 'base'):
    try:
        d = json.load(open('gpurun_out/r2c6_bench_%s.json' % t))
        print(t, round(d['value'], 1), 'img/s', round(d['ms_per_step'], 2), 'ms e2e', round(d['e2e']['value'], 1),
              {k: round(v['ms_per_step'], 2) for k, v in d['roofline']['kernels'].items()})
    except Exception as e:
        print(t, 'bench failed', e)
PY
