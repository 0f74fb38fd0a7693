#!/bin/bash
# First gpurun call of the next round (~4 min of box time): what this round could not run on a GPU any more.
#   gpurun --timeout 420 -- 'bash tools/next_round.sh r02'
TAG=${1:-r02}
mkdir -p gpurun_out
# 1. the augmentation kernels through the C ABI (CPU-verified by host emulation only so far)
timeout 120 python -m pytest tests/test_gpu_zz_augment.py -q -m gpu --durations=5 > gpurun_out/pytest_augment_$TAG.log 2>&1
tail -n 15 gpurun_out/pytest_augment_$TAG.log
# 2. bench line (adds the `augment` key: ms/batch next to Pillow on one host core, and roofline.whole_step)
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 300 gpurun_out/bench_$TAG.err
# 3. ncu of the augmentation kernels (HBM roofline: source bytes once + 12 B per output pixel)
cat > /tmp/aug_probe.py <<'PY'
import sys, random, numpy as np, torch
sys.path.insert(0, '.')
from fewshot_detection_b200 import image as I
rs = np.random.RandomState(0); random.seed(0)
srcs = [torch.from_numpy(rs.randint(0, 256, (375, 500, 3)).astype(np.uint8)).cuda() for _ in range(64)]
ps = [I.draw_augmentation(500, 375, 0.2, 0.1, 1.5, 1.5) for _ in range(64)]
for _ in range(3):
    out = I.augment_batch(srcs, (416, 416), ps)
torch.cuda.synchronize(); print(out.mean().item())
PY
timeout 150 ncu --set full --clock-control none --import-source on -k 'regex:augment' --launch-skip 2 --launch-count 4 -f \
    -o gpurun_out/prof_augment_$TAG python /tmp/aug_probe.py > gpurun_out/prof_augment_$TAG.log 2>&1
ncu -i gpurun_out/prof_augment_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_augment_$TAG.raw.csv 2>/dev/null
head -c 1200 gpurun_out/bench_$TAG.json
