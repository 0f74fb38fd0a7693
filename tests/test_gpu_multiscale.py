"""BASELINE configs[4] (608x608, multi-scale) and the input staging of the e2e path.

The detector is fully convolutional: train_meta.py's multi-scale schedule (dataset.py:223-245) feeds query batches of
side 320..608 (multiples of 32) through the same weights while the support branch stays at 416.  Forward / loss
parity against the float32 CPU oracle at the two ends of that range, gradients as one concatenated vector (entry-wise
gradient parity on tiny batches is the business of test_gpu_model.py, see DESIGN.md 'Parity')."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-3


def relt(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize('side,bs,cs', [(608, 1, 3), (320, 2, 4)])
def test_meta_full_arch_other_input_sizes_vs_oracle(side, bs, cs):
    from fewshot_detection_b200 import netcfg
    from fewshot_detection_b200.darknet_meta import Darknet
    from oracle import darknet as ODK, region_loss as ORL
    from seeding import seeded_init, synth_targets, synth_masks
    det, ler = netcfg.darknet_dynamic_blocks(side, side), netcfg.reweighting_net_blocks()
    g = torch.Generator().manual_seed(71)
    x = torch.rand(bs, 3, side, side, generator=g)
    metax = torch.rand(cs, 3, 416, 416, generator=g)
    mask = torch.from_numpy(synth_masks(cs, 416, 72))
    tgt = torch.from_numpy(synth_targets(bs, cs, 73, max_gt=4))
    om = ODK.MetaDarknet([dict(b) for b in det], [dict(b) for b in ler])
    seeded_init(om, 70)
    om.train()
    oo = om(x, metax, mask)
    lo = ORL.region_loss_v2(oo, tgt, om.anchors, 5, 1, seen=20000)
    lo.backward()
    m = Darknet([dict(b) for b in det], [dict(b) for b in ler])
    seeded_init(m, 70)
    m = m.cuda().train()
    out = m(x.cuda(), metax.cuda(), mask.cuda())
    G = side // 32
    assert tuple(out.shape) == (bs * cs, 30, G, G)
    assert relt(out.detach().cpu(), oo.detach()) < TOL
    L = m.models[len(m.models) - 1]
    L.seen = 20000
    L.verbose = False
    loss = L(out, tgt)
    loss.backward()
    assert abs(loss.item() - lo.item()) < TOL * abs(lo.item())
    ours = torch.cat([p.grad.detach().cpu().contiguous().reshape(-1).double() for p in m.parameters()])
    ref = torch.cat([p.grad.detach().reshape(-1).double() for p in om.parameters()])
    assert torch.isfinite(ours).all()
    assert relt(ours, ref) < 5e-2


def test_device_prefetcher_order_values_and_host_fields():
    from fewshot_detection_b200.prefetch import DevicePrefetcher
    g = torch.Generator().manual_seed(3)
    host = [(torch.rand(4, 3, 32, 32, generator=g).pin_memory(), torch.rand(2, 1, 8, 8, generator=g).pin_memory(),
             torch.rand(4, 2, 250, generator=g, dtype=torch.float64), i) for i in range(5)]
    pf = DevicePrefetcher(host, 'cuda', host_fields=(2,))
    seen = 0
    ptrs = set()
    for i, (a, b, t, k) in enumerate(pf):
        assert a.is_cuda and b.is_cuda and not t.is_cuda and k == i
        ptrs.add(a.data_ptr())
        y = (a * 2).sum() + b.sum()                                    # consume on the current stream
        want = (host[i][0].double() * 2).sum() + host[i][1].double().sum()
        assert abs(y.item() - want.item()) < 1e-2
        assert t is host[i][2]
        seen += 1
    assert seen == 5
    assert len(ptrs) == 2                                              # two staging slots, reused
    assert pf.h2d_bytes == sum(h[0].numel() * 4 + h[1].numel() * 4 for h in host)
    with pytest.raises(StopIteration):
        next(pf)
    with pytest.raises(TypeError):
        DevicePrefetcher(host, 'cpu')


def test_async_loss_reader_values_in_order():
    from fewshot_detection_b200.prefetch import AsyncLossReader
    r = AsyncLossReader(depth=2)
    static = torch.zeros((), device='cuda')          # a graph's static loss tensor is overwritten every step
    got = []
    for i in range(7):
        static.fill_(float(i) + 0.5)
        r.push(static)
        if r.count == 2:
            got.append(r.pop())
    got += r.drain()
    assert got == [i + 0.5 for i in range(7)]
    with pytest.raises(IndexError):
        r.pop()
    r.push(static)
    r.push(static)
    with pytest.raises(RuntimeError):
        r.push(static)
