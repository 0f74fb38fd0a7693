"""Developer tool for profilers: runs N eager training steps of the config-2 workload and exits.
Usage: python tools/one_step.py [steps] [batch] [ncls]"""
import os, sys, contextlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch
from fewshot_detection_b200 import netcfg
from fewshot_detection_b200.darknet_meta import Darknet
from fewshot_detection_b200.optim import FusedSGD
from fewshot_detection_b200.distributed import GradAllReducer
from seeding import seeded_init
sys.path.insert(0, ROOT)
from bench import synth_batch
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ncls = int(sys.argv[3]) if len(sys.argv) > 3 else 20
with contextlib.redirect_stdout(sys.stderr):
    m = Darknet(netcfg.darknet_dynamic_blocks(), netcfg.reweighting_net_blocks())
seeded_init(m, 0); m = m.cuda().train()
L = m.loss; L.verbose = False; L.seen = 20000
opt = FusedSGD(m.parameters(), lr=1e-6, momentum=0.9, weight_decay=0.48)
red = GradAllReducer(m)
x, metax, mask, tgt = [t.cuda() for t in synth_batch(B, ncls, 416, 0)]
for i in range(steps):
    red.begin_step()
    loss = L(m(x, metax, mask), tgt)
    loss.backward(); red.finish(); opt.step()
torch.cuda.synchronize()
print('done', loss.item())
