#!/bin/bash
mkdir -p gpurun_out
FSDET_TRAIN_PROFILE=1 timeout 600 python tools/e2e_train_synth.py 2048 4 gpurun_out/r2c9_e2e_train.json > gpurun_out/r2c9_e2e_train.log 2>&1
echo "e2e train rc=$?"; grep -v "^class_scale" gpurun_out/r2c9_e2e_train.log | grep -v nGT | tail -n 12 | cut -c1-400
python - <<'PY'
# host-side micro timing of the prepare half on this box
import sys, os, time
sys.path.insert(0, '.')
import importlib.util
from fewshot_detection_b200.cfg import cfg
from fewshot_detection_b200.dataset import DetectionBatcher
import torch
cfg.data = 'voc'; cfg.metayolo = True; cfg.multiscale = 1
root = '/tmp/fsdet_synth_voc_2048'
lines = [l for l in open(root + '/lists/train.txt')]
b = DetectionBatcher(lines, shape=(416, 416), shuffle=False, train=True, seen=0, batch_size=64, seen_step=1)
r = b.batch_ranges()
for i in range(3):
    t0 = time.time(); p = b.prepare(r[0]); t1 = time.time()
    q = b.finish(p); t2 = time.time(); torch.cuda.synchronize(); t3 = time.time()
    print('prepare %.1f ms  finish(host) %.1f ms  finish(device drain) %.1f ms' % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
PY
