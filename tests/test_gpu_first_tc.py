"""The recomputing first-layer kernels (csrc/conv_first_tc.cuh) through the C ABI against torch float64: statistics,
BN + LeakyReLU + max-pool output (fp32 and fp16 planes), and the backward pair (reduce -> bn_bwd_finalize -> weight
gradient formed without a dz tensor) against autograd of conv -> batch_norm(train) -> leaky_relu -> max_pool2d."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def st():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize('B,C0,C1,H,W,Cout', [(2, 3, 1, 64, 64, 32), (3, 3, 0, 32, 48, 8), (1, 3, 0, 128, 128, 32), (2, 3, 1, 416, 416, 32)])
def test_first_block_passes_vs_torch(B, C0, C1, H, W, Cout):
    from fewshot_detection_b200 import _lib as L
    P = lambda t: t.data_ptr() if t is not None else None
    g = torch.Generator(device='cuda').manual_seed(B + H + Cout)
    x0 = torch.rand(B, C0, H, W, device='cuda', generator=g)
    x1 = (torch.rand(B, C1, H, W, device='cuda', generator=g) > 0.5).float() if C1 else None
    x = torch.cat([x0, x1], 1) if C1 else x0
    C = C0 + C1
    w = (torch.randn(Cout, C, 3, 3, device='cuda', generator=g) * 0.3)
    gamma = torch.rand(Cout, device='cuda', generator=g) + 0.5
    beta = torch.randn(Cout, device='cuda', generator=g) * 0.1
    w4 = torch.zeros(Cout, 9, 4, device='cuda')
    w4[:, :, :C] = w.permute(0, 2, 3, 1).reshape(Cout, 9, C)
    amax_x = x.abs().max().reshape(1).float()
    assert L.lib.fsdet_conv_first_tc_supported(H, W, Cout)
    # ---- float64 reference chain
    wd = w.double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    z = F.conv2d(x.double(), wd, None, 1, 1)
    y = F.leaky_relu(F.batch_norm(z, None, None, gd, bd, True, 0.1, 1e-5), 0.1)
    yp = F.max_pool2d(y, 2, 2)
    dyp = torch.randn(yp.shape, device='cuda', generator=g, dtype=torch.float32)
    yp.backward(dyp.double())
    zz = z.detach().permute(0, 2, 3, 1).reshape(-1, Cout)
    npix = B * H * W
    # ---- pass 0 + finalize
    rows = L.lib.fsdet_conv_first_tc_rows(B, H, W)
    stat = torch.empty(rows + L.lib.fsdet_bn_stat_scratch_rows(), 4 * Cout, device='cuda')
    L.call('fsdet_conv_first_tc_stats', P(x0), C0, P(x1), C1, P(w4), P(amax_x), P(stat), B, H, W, Cout, st())
    s = stat[:rows].double().sum(0)
    assert rel(s[:Cout], zz.sum(0)) < 1e-5 or (s[:Cout] - zz.sum(0)).abs().max() < 1e-2
    assert rel(s[Cout:2 * Cout], (zz * zz).sum(0)) < 1e-5
    assert rel(stat[:rows, 2 * Cout:3 * Cout].min(0)[0], zz.min(0)[0]) < 1e-5
    assert rel(stat[:rows, 3 * Cout:].max(0)[0], zz.max(0)[0]) < 1e-5
    vec = torch.empty(5, Cout, device='cuda')
    amax_y = torch.empty(1, device='cuda')
    L.call('fsdet_bn_finalize', P(stat), rows, float(npix), P(gamma), P(beta), None, None, 0.1, 1e-5, P(vec[0]), P(vec[1]), P(vec[2]),
           P(vec[3]), 0.1, P(amax_y), P(vec[4]), Cout, 1, st())
    assert rel(vec[0], zz.mean(0)) < 1e-5
    assert amax_y.item() >= y.abs().max().item() * (1 - 1e-5)
    # ---- pass 1
    Hp, Wp = H // 2, W // 2
    yp32 = torch.full((B * Hp * Wp, Cout + 4), 7.0, device='cuda')
    ph = torch.full((B * Hp * Wp, 64), 99.0, dtype=torch.float16, device='cuda')
    pl = torch.full((B * Hp * Wp, 64), 99.0, dtype=torch.float16, device='cuda')
    L.call('fsdet_conv_first_tc_apply', P(x0), C0, P(x1), C1, P(w4), P(amax_x), P(vec[2]), P(vec[3]), 0.1, P(yp32), Cout + 4, P(ph), P(pl),
           64, P(amax_y), B, H, W, Cout, st())
    ref_p = yp.detach().permute(0, 2, 3, 1).reshape(-1, Cout)
    assert rel(yp32[:, :Cout], ref_p) < 1e-5
    assert (yp32[:, Cout:] == 7.0).all()
    import math
    sc = 2.0 ** (10 - math.frexp(amax_y.item())[1])
    assert rel((ph.double() + pl.double())[:, :Cout] / sc, ref_p) < 1e-5
    assert (ph[:, 32:] == 0).all() and (pl[:, 32:] == 0).all()
    if Cout < 32:
        assert (ph[:, Cout:32] == 0).all()
    # ---- pass 2 + finalize + pass 3
    dyn = dyp.permute(0, 2, 3, 1).reshape(-1, Cout).contiguous()
    part = torch.empty(rows + 1, 3 * Cout, dtype=torch.float64, device='cuda')
    L.call('fsdet_conv_first_tc_bwd_reduce', P(x0), C0, P(x1), C1, P(w4), P(amax_x), P(vec[2]), P(vec[3]), P(vec[0]), P(vec[1]), 0.1,
           P(dyn), Cout, P(part), B, H, W, Cout, st())
    dgamma = torch.empty(Cout, device='cuda')
    dbeta = torch.empty(Cout, device='cuda')
    coef = torch.empty(2, Cout, dtype=torch.float64, device='cuda')
    amax_dz = torch.empty(1, device='cuda')
    L.call('fsdet_bn_bwd_finalize', P(part), rows, float(npix), P(gamma), P(vec[1]), P(vec[4]), P(dgamma), P(dbeta), P(coef), P(amax_dz),
           Cout, 1, st())
    assert rel(dbeta, bd.grad) < 1e-4
    assert rel(dgamma, gd.grad) < 1e-4
    nws = L.lib.fsdet_conv_first_tc_wgrad_workspace_floats(B, H, W)
    ws = torch.empty(nws, device='cuda')
    dw4 = torch.full((Cout, 9, 4), 5.0, device='cuda')
    L.call('fsdet_conv_first_tc_bwd_wgrad', P(x0), C0, P(x1), C1, P(w4), P(amax_x), P(vec[2]), P(vec[3]), P(vec[0]), P(vec[1]), P(coef), 0.1,
           P(dyn), Cout, P(amax_dz), P(dw4), P(ws), nws, B, H, W, Cout, st())
    torch.cuda.synchronize()
    ref_w = wd.grad.permute(0, 2, 3, 1).reshape(Cout, 9, C)
    e = rel(dw4[:, :, :C], ref_w)
    assert e < 1e-3, e                 # x exact, dz rounded to fp16 (the 2-term weight-gradient mode)
    if C < 4:
        assert dw4[:, :, C:].abs().max().item() < 1e-3 * ref_w.abs().max().item()
    print('first block: dW rel err %.2e' % e)


def test_model_with_recomputing_first_block_matches_default_path():
    """The opt-in first-block path ('first' in engine.TC_PARTS) through the public API: same output, loss and
    parameter gradients as the default path (which stores the first layer's pre-BN tensor)."""
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
    from fewshot_detection_b200 import netcfg, engine
    from fewshot_detection_b200.darknet_meta import Darknet
    from seeding import seeded_init, synth_targets, synth_masks
    det, ler = netcfg.mini_dynamic_blocks(128, 8), netcfg.mini_reweighting_blocks(64, 8, 256)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(3, 3, 128, 128, generator=g).cuda()
    metax = torch.rand(2, 3, 64, 64, generator=g).cuda()
    mask = torch.from_numpy(synth_masks(2, 64, 6)).cuda()
    tgt = torch.from_numpy(synth_targets(3, 2, 7, max_gt=3))
    runs = []
    had = 'first' in engine.TC_PARTS
    try:
        for first in (False, True):
            (engine.TC_PARTS.add if first else engine.TC_PARTS.discard)('first')
            m = Darknet([dict(b) for b in det], [dict(b) for b in ler])
            seeded_init(m, 4)
            m = m.cuda().train()
            L = m.models[len(m.models) - 1]
            L.verbose = False
            L.seen = 20000
            out = m(x, metax, mask)
            loss = L(out, tgt)
            loss.backward()
            runs.append((out.detach().clone(), loss.item(), {n: p.grad.detach().clone() for n, p in m.named_parameters()},
                         {n: b.detach().clone() for n, b in m.named_buffers() if 'running' in n}))
    finally:
        (engine.TC_PARTS.add if had else engine.TC_PARTS.discard)('first')
    (o0, l0, g0, b0), (o1, l1, g1, b1) = runs
    assert rel(o1, o0) < 1e-5 and abs(l1 - l0) < 1e-5 * abs(l0)
    for n in g0:
        assert rel(g1[n], g0[n]) < (1e-3 if n.endswith('conv1.weight') else 2e-4), n     # conv1: dz rounded to fp16 in the new path
    for n in b0:
        assert rel(b1[n], b0[n]) < 1e-5, n
