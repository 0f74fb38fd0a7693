"""Developer tool: per-tensor parity of the full 416 meta model against a float64
CPU oracle (ground truth), next to the float32 CPU oracle and torch-CUDA fp32.
Usage: python tools/debug_full.py [bs] [cs]"""
import copy
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from fewshot_detection_b200 import netcfg
from fewshot_detection_b200.darknet_meta import Darknet
from oracle import darknet as ODK, region_loss as ORL
from seeding import seeded_init, synth_targets, synth_masks

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
det, ler = netcfg.darknet_dynamic_blocks(), netcfg.reweighting_net_blocks()
g = torch.Generator().manual_seed(62)
x = torch.rand(bs, 3, 416, 416, generator=g); metax = torch.rand(cs, 3, 416, 416, generator=g)
mask = torch.from_numpy(synth_masks(cs, 416, 63)); tgt = torch.from_numpy(synth_targets(bs, cs, 64, max_gt=4))


def run_oracle(dtype, device):
    om = ODK.MetaDarknet(det, ler); seeded_init(om, 61); om = om.to(dtype).to(device).train()
    t0 = time.time()
    oo = om(x.to(dtype).to(device), metax.to(dtype).to(device), mask.to(dtype).to(device))
    # the loss itself is evaluated by the float32 CPU oracle on the float32-rounded output, then chained
    o32 = oo.detach().float().cpu().requires_grad_(True)
    lo = ORL.region_loss_v2(o32, tgt, om.anchors, 5, 1, seen=20000)
    lo.backward()
    oo.backward(o32.grad.to(dtype).to(device))
    print('oracle %s %s: %.1fs loss %.6f' % (dtype, device, time.time() - t0, lo.item()))
    return oo.detach().double().cpu(), {n: p.grad.detach().double().cpu() for n, p in om.named_parameters()}


o64, g64 = run_oracle(torch.float64, 'cpu')
o32, g32 = run_oracle(torch.float32, 'cpu')
o32g, g32g = run_oracle(torch.float32, 'cuda')
m = Darknet([dict(b) for b in det], [dict(b) for b in ler]); seeded_init(m, 61); m = m.cuda().train()
out = m(x.cuda(), metax.cuda(), mask.cuda())
L = m.models[len(m.models) - 1]; L.seen = 20000
loss = L(out, tgt); loss.backward()
rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()
print('output rel vs f64: cpu32 %.2e  cuda32 %.2e  ours %.2e' % (rel(o32, o64), rel(o32g, o64), rel(out.detach().cpu(), o64)))
print('%-40s %10s %10s %10s' % ('param grad rel err vs f64', 'cpu-f32', 'torch-cuda', 'ours'))
for n, p in m.named_parameters():
    print('%-40s %10.2e %10.2e %10.2e' % (n, rel(g32[n], g64[n]), rel(g32g[n], g64[n]),
                                          rel(p.grad.detach().cpu().contiguous(), g64[n])))

# are the residuals discrete (arg-max flips: a few output channels carry all the error) or diffuse?
for n in ('learnet_models.10.conv6.weight', 'learnet_models.8.conv5.weight', 'models.29.conv22.weight', 'models.23.conv19.weight'):
    p = dict(m.named_parameters())[n]
    a = p.grad.detach().cpu().contiguous().double(); b = g64[n]
    per = ((a - b).flatten(1).norm(dim=1) / b.flatten(1).norm(dim=1).clamp_min(1e-30))
    print(n, 'out-channels with rel err > 1e-3: %d of %d; median %.2e; max %.2e' % ((per > 1e-3).sum().item(), per.numel(), per.median().item(), per.max().item()))
    perin = ((a - b).transpose(0, 1).flatten(1).norm(dim=1) / b.transpose(0, 1).flatten(1).norm(dim=1).clamp_min(1e-30))
    print('   in-channels with rel err > 1e-3: %d of %d; median %.2e; max %.2e' % ((perin > 1e-3).sum().item(), perin.numel(), perin.median().item(), perin.max().item()))
