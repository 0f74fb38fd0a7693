"""Developer probe: where do the ~2.9 ms/step between the device-resident loop and the host-fed loop go?
Times variants of the config-2 loop (CUDA-graph step, B=64, n_cls=20) with CUDA events; prints ms/step per variant.
Usage (gpurun): python tools/e2e_probe.py [steps]"""
import contextlib
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch
from fewshot_detection_b200 import netcfg
from fewshot_detection_b200.cfg import cfg
from fewshot_detection_b200.darknet_meta import Darknet
from fewshot_detection_b200.optim import FusedSGD
from fewshot_detection_b200.distributed import GradAllReducer
from fewshot_detection_b200.graph import GraphedTrainStep
from fewshot_detection_b200.prefetch import DevicePrefetcher, AsyncLossReader
from seeding import seeded_init
from bench import synth_batch

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device('cuda', 0)
cfg.neg_ratio = 'full'
with contextlib.redirect_stdout(sys.stderr):
    m = Darknet(netcfg.darknet_dynamic_blocks(), netcfg.reweighting_net_blocks())
seeded_init(m, 0); m = m.to(dev).train()
L = m.loss; L.verbose = False; L.seen = 20000
opt = FusedSGD(m.parameters(), lr=1e-6, momentum=0.9, weight_decay=0.48)
g = GraphedTrainStep(m, L, opt, GradAllReducer(m))
host = [tuple(t.pin_memory() for t in synth_batch(64, 20, 416, 10 * i)) for i in range(2)]
res = [tuple(t.to(dev) for t in hb) for hb in host]
for i in range(4):
    g(*res[i % 2])
torch.cuda.synchronize()


def timed(fn, flush=None):
    for i in range(3):
        fn(i)
    if flush: flush()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    if flush: flush()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


out = {}
out['A resident, no readback'] = timed(lambda i: g(*res[i % 2]))
out['B resident + loss.item() every step'] = timed(lambda i: g(*res[i % 2]).item())
r = AsyncLossReader(2)
def c(i):
    r.push(g(*res[i % 2]))
    if r.count == 2: r.pop()
out['C resident + async readback'] = timed(c, r.drain)
def d(i):
    hb = host[i % 2]
    g(hb[0].to(dev, non_blocking=True), hb[1].to(dev, non_blocking=True), hb[2].to(dev, non_blocking=True), hb[3])
out['D serial H2D on the training stream, no readback'] = timed(d)
pf = DevicePrefetcher((host[i % 2] for i in range(10 ** 6)), dev, host_fields=(3,))
out['E prefetched H2D, no readback'] = timed(lambda i: g(*next(pf)))
pf2 = DevicePrefetcher(((res[i % 2][0], res[i % 2][1], res[i % 2][2], host[i % 2][3]) for i in range(10 ** 6)), dev, host_fields=(3,))
out['F prefetcher fed from device tensors (no PCIe), host target'] = timed(lambda i: g(*next(pf2)))
out['G resident inputs, host (pinned) target'] = timed(lambda i: g(res[i % 2][0], res[i % 2][1], res[i % 2][2], host[i % 2][3]))
def h(i):
    r.push(g(*next(pf)))
    if r.count == 2: r.pop()
out['H prefetched H2D + async readback (= bench e2e)'] = timed(h, r.drain)
# chunked H2D: 16 MB pieces on the copy stream
side = torch.cuda.Stream(dev)
stage = [[torch.empty_like(t, device=dev) for t in host[0][:3]] for _ in range(2)]
evs = [torch.cuda.Event() for _ in range(2)]
def chunked_fill(slot, hb):
    with torch.cuda.stream(side):
        for s, t in zip(stage[slot], hb[:3]):
            sv, tv = s.view(-1), t.view(-1)
            n = tv.numel(); c = 4 << 20
            for o in range(0, n, c):
                sv[o:o + c].copy_(tv[o:o + c], non_blocking=True)
        evs[slot].record(side)
chunked_fill(0, host[0])
state = {'k': 0}
def ch(i):
    k = state['k']
    torch.cuda.current_stream().wait_event(evs[k])
    free = torch.cuda.Event(); free.record()
    side.wait_event(free)
    chunked_fill(k ^ 1, host[(i + 1) % 2])
    state['k'] = k ^ 1
    g(stage[k][0], stage[k][1], stage[k][2], host[i % 2][3])
out['I prefetched H2D in 16 MB chunks, no readback'] = timed(ch)
for k, v in out.items():
    print('%-62s %7.3f ms/step' % (k, v))
