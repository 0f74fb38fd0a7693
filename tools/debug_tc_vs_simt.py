"""Developer tool: tensor-core path vs exact-fp32 path on the same GPU, mini plain net, per tensor."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch
from fewshot_detection_b200 import netcfg, engine
from fewshot_detection_b200.darknet import Darknet
from fewshot_detection_b200.cfg import cfg
from seeding import seeded_init, synth_targets
cfg.metayolo = False
blocks = netcfg.mini_tiny_blocks(128, 8)
x = torch.rand(3, 3, 128, 128, generator=torch.Generator().manual_seed(6)).cuda()
tgt = torch.from_numpy(synth_targets(3, 1, 7, max_gt=4)[:, 0, :])
tgt[:, 0::5] = (torch.arange(50) % 20).double()
rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def run(use_tc, parts):
    engine.USE_TC = use_tc
    engine.TC_PARTS = set(parts)
    m = Darknet([dict(b) for b in blocks]); seeded_init(m, 5); m = m.cuda().train()
    out = m(x)
    L = m.models[len(m.models) - 1]; L.seen = 20000; L.verbose = False
    loss = L(out, tgt); loss.backward()
    return out.detach().clone(), {n: p.grad.detach().clone() for n, p in m.named_parameters()}


o0, g0 = run(False, [])
for parts in (['fwd'], ['dgrad'], ['wgrad'], ['fwd', 'dgrad', 'wgrad', 'head']):
    o1, g1 = run(True, parts)
    print('parts', parts, 'output rel', '%.2e' % rel(o1, o0))
    print('   ' + ' '.join('%s=%.1e' % (n.split('.')[1] + n.split('.')[-1][0], rel(g1[n], g0[n])) for n in g0))
