"""Developer tool: the mini plain-Darknet training step of tests/test_gpu_model.py::test_tiny_mini_train_step_vs_oracle over
several input seeds - parameter-gradient error against the float64 oracle for this build and for the float32 oracle, and
how concentrated the difference is (a max-pool arg-max flip moves a few gradient paths completely; an arithmetic error
is spread).  Usage: python tools/diag_mini.py [seeds]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch

from fewshot_detection_b200 import netcfg
from fewshot_detection_b200.darknet import Darknet
from fewshot_detection_b200.cfg import cfg
from oracle import darknet as ODK, region_loss as ORL
from seeding import seeded_init, synth_targets

relt = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 6
blocks = netcfg.mini_tiny_blocks(128, 8)
for seed in range(6, 6 + nseeds):
    x = torch.rand(3, 3, 128, 128, generator=torch.Generator().manual_seed(seed))
    tgt = torch.from_numpy(synth_targets(3, 1, 7, max_gt=4)[:, 0, :])
    tgt[:, 0::5] = torch.floor(tgt[:, 0::5] * 0) + (torch.arange(50) % 20).double()

    def run_oracle(dtype):
        om = ODK.PlainDarknet([dict(b) for b in blocks])
        seeded_init(om, 5)
        om = om.to(dtype).train()
        oo = om(x.to(dtype))
        o32 = oo.detach().float().requires_grad_(True)
        lo = ORL.region_loss_plain(o32, tgt, om.anchors, 5, 20, seen=20000, metayolo=False)
        lo.backward()
        oo.backward(o32.grad.to(dtype))
        return oo.detach().double(), lo.item(), {n: p.grad.detach().double() for n, p in om.named_parameters()}

    o64, l64, g64 = run_oracle(torch.float64)
    o32, l32, g32 = run_oracle(torch.float32)
    m = Darknet([dict(b) for b in blocks])
    seeded_init(m, 5)
    m = m.cuda().train()
    cfg.metayolo = False
    try:
        out = m(x.cuda())
        L = m.models[len(m.models) - 1]
        L.seen = 20000
        loss = L(out, tgt)
        loss.backward()
    finally:
        cfg.metayolo = True
    worst = ('', 0.0, 0.0)
    for n, p in m.named_parameters():
        e = relt(p.grad.detach().cpu().contiguous(), g64[n])
        if e > worst[1]:
            worst = (n, e, relt(g32[n], g64[n]))
    n = 'models.0.conv1.weight'
    d = (dict(m.named_parameters())[n].grad.detach().cpu().double() - g64[n]).abs().flatten()
    print('seed %d: out %.2e (fp32 oracle %.2e)  loss %.2e  worst grad %s %.2e (fp32 oracle %.2e)  conv1.weight %.2e' % (
        seed, relt(out.detach().cpu(), o64), relt(o32, o64), abs(loss.item() - l64) / abs(l64), worst[0], worst[1], worst[2],
        relt(dict(m.named_parameters())[n].grad.detach().cpu(), g64[n])))
