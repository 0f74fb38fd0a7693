"""Developer diagnostic: per-tensor gradient error of the smoke() mini model under several term policies, against the
float64 oracle (and the float32 oracle's own distance).  python tools/diag_smoke.py [c] [side] [bs]"""
import contextlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch  # noqa: E402

from fewshot_detection_b200 import netcfg, engine  # noqa: E402
from fewshot_detection_b200.darknet_meta import Darknet  # noqa: E402
from oracle import darknet as ODK, region_loss as ORL  # noqa: E402
from seeding import seeded_init, synth_targets, synth_masks  # noqa: E402

c = int(sys.argv[1]) if len(sys.argv) > 1 else 16
side = int(sys.argv[2]) if len(sys.argv) > 2 else 128
bs = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cs = 3
det, ler = netcfg.mini_dynamic_blocks(side, c), netcfg.mini_reweighting_blocks(side // 2, c, 32 * c)
g = torch.Generator().manual_seed(2)
x = torch.rand(bs, 3, side, side, generator=g)
metax = torch.rand(cs, 3, side // 2, side // 2, generator=g)
mask = torch.from_numpy(synth_masks(cs, side // 2, 3))
tgt = torch.from_numpy(synth_targets(bs, cs, 4, max_gt=4))
rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def oracle(dtype):
    om = ODK.MetaDarknet([dict(b) for b in det], [dict(b) for b in ler])
    seeded_init(om, 1)
    om = om.to(dtype).train()
    oo = om(x.to(dtype), metax.to(dtype), mask.to(dtype))
    o32 = oo.detach().float().requires_grad_(True)
    lo = ORL.region_loss_v2(o32, tgt, om.anchors, 5, 1, seen=20000)
    lo.backward()
    oo.backward(o32.grad.to(dtype))
    return {n: p.grad.detach().double() for n, p in om.named_parameters()}, {n: tuple(p.shape) for n, p in om.named_parameters()}


g64, shapes = oracle(torch.float64)
g32, _ = oracle(torch.float32)
for pol in (dict(fwd=3, dgrad=3, wgrad=3, head=3), dict(fwd=3, dgrad=3, wgrad=0, head=3), dict(fwd=3, dgrad=3, wgrad=1, head=3)):
    engine.TC_TERMS.update(pol)
    with contextlib.redirect_stdout(sys.stderr):
        m = Darknet([dict(b) for b in det], [dict(b) for b in ler])
    seeded_init(m, 1)
    m = m.cuda().train()
    L = m.models[len(m.models) - 1]
    L.seen = 20000
    L.verbose = False
    loss = L(m(x.cuda(), metax.cuda(), mask.cuda()), tgt)
    loss.backward()
    errs = {n: rel(p.grad.detach().cpu().contiguous(), g64[n]) for n, p in m.named_parameters()}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    print(pol, 'loss', loss.item())
    for n, e in worst:
        print('   %-34s %-22s ours %.2e   fp32-oracle %.2e' % (n, shapes[n], e, rel(g32[n], g64[n])))
