"""tcgen05 / TMA-im2col convolution path (csrc/conv_tc.cu) against torch fp32."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).norm() / max(b.norm().item(), 1e-30)).item()


@pytest.fixture(scope='module')
def L():
    from fewshot_detection_b200 import _lib
    assert torch.cuda.is_available()
    return _lib


def st():
    return torch.cuda.current_stream().cuda_stream


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().view(-1, x.shape[1])


def split(L, t2d, cpad=None):
    rows, C = t2d.shape
    cpad = cpad or C
    hi = torch.empty(rows, cpad, dtype=torch.bfloat16, device='cuda')
    lo = torch.empty(rows, cpad, dtype=torch.bfloat16, device='cuda')
    L.call('fsdet_split_bf16', t2d.data_ptr(), C, C, cpad, rows, hi.data_ptr(), lo.data_ptr(), st())
    return hi, lo


def test_split_bf16(L):
    x = torch.randn(1000, 64, device='cuda') * 3
    hi, lo = split(L, x)
    assert torch.equal(hi, x.to(torch.bfloat16))
    assert rel(hi.float() + lo.float(), x) < 2e-5
    assert torch.equal(lo, (x - hi.float()).to(torch.bfloat16))


@pytest.mark.parametrize('B,H,W,C,ks,m0,c0,tap', [
    (2, 13, 13, 64, 3, 0, 0, 0), (2, 13, 13, 128, 3, 128, 64, 4), (2, 13, 13, 64, 3, 256, 0, 8),
    (3, 26, 26, 64, 1, 640, 0, 0), (1, 52, 52, 64, 3, 2560, 0, 2), (5, 6, 6, 64, 3, 128, 0, 6), (1, 8, 8, 64, 3, 0, 0, 5)])
def test_tma_im2col_tile(L, B, H, W, C, ks, m0, c0, tap):
    g = torch.Generator(device='cuda').manual_seed(m0 + tap)
    x = torch.randn(B, H, W, C, device='cuda', generator=g).to(torch.bfloat16)
    out = torch.zeros(128, 64, dtype=torch.bfloat16, device='cuda')
    L.call('fsdet_debug_im2col_tile', x.data_ptr(), B, H, W, C, ks, m0, c0, tap, out.data_ptr(), st())
    torch.cuda.synchronize()
    pad = (ks - 1) // 2
    r, s = tap // ks, tap % ks
    exp = torch.zeros(128, 64, dtype=torch.bfloat16, device='cuda')
    for i in range(128):
        m = m0 + i
        n, rem = divmod(m, H * W)
        p, q = divmod(rem, W)
        hh, ww = p + r - pad, q + s - pad
        if n < B and 0 <= hh < H and 0 <= ww < W:
            exp[i] = x[n, hh, ww, c0:c0 + 64]
    assert torch.equal(out, exp), (out.float() - exp.float()).abs().max().item()


TC_CASES = [
    # B, H, W, Cin, Cout, k
    (2, 13, 13, 64, 128, 3), (4, 26, 26, 128, 64, 1), (1, 52, 52, 64, 128, 3), (2, 13, 13, 1024, 480, 1),
    (2, 13, 13, 1280, 1024, 3), (3, 6, 6, 1024, 1024, 3), (2, 19, 19, 256, 512, 3), (1, 104, 104, 128, 256, 3),
    (3, 4, 4, 128, 256, 3), (3, 8, 8, 64, 128, 3), (3, 4, 4, 256, 256, 3), (1, 2, 2, 64, 64, 3), (3, 16, 16, 64, 64, 1),
]


@pytest.mark.parametrize('B,H,W,Cin,Cout,k', TC_CASES)
def test_conv_tc_fwd(L, B, H, W, Cin, Cout, k):
    g = torch.Generator(device='cuda').manual_seed(B + H + Cin + Cout)
    x = torch.randn(B, Cin, H, W, device='cuda', generator=g)
    w = torch.randn(Cout, Cin, k, k, device='cuda', generator=g) * 0.05
    ref = F.conv2d(x, w, None, 1, (k - 1) // 2)
    xh, xl = split(L, nhwc(x))
    wh, wl = split(L, w.permute(0, 2, 3, 1).contiguous().view(Cout, -1))
    assert L.lib.fsdet_conv_tc_supported(Cin, Cout, k)
    ld = Cout + 4
    z = torch.zeros(B * H * W, ld, device='cuda')
    L.call('fsdet_conv_tc_fwd', xh.data_ptr(), xl.data_ptr(), wh.data_ptr(), wl.data_ptr(), z.data_ptr(), ld, B, H, W, Cin, Cout,
           k, 0, st())
    torch.cuda.synchronize()
    got = z[:, :Cout].contiguous().view(B, H, W, Cout).permute(0, 3, 1, 2)
    assert rel(got, ref) < 5e-5
    assert (z[:, Cout:] == 0).all()
    L.call('fsdet_conv_tc_fwd', xh.data_ptr(), xl.data_ptr(), wh.data_ptr(), wl.data_ptr(), z.data_ptr(), ld, B, H, W, Cin, Cout,
           k, 1, st())
    got2 = z[:, :Cout].contiguous().view(B, H, W, Cout).permute(0, 3, 1, 2)
    assert rel(got2, 2 * ref) < 5e-5


def test_colstats(L):
    z = torch.randn(5000, 96, device='cuda') * 2 + 1
    buf = torch.zeros(5000, 100, device='cuda')
    buf[:, :96] = z
    rows = L.lib.fsdet_colstats_rows(5000)
    part = torch.zeros(rows + 2, 192, device='cuda')
    L.call('fsdet_colstats', buf.data_ptr(), 100, 5000, 96, part.data_ptr(), st())
    s = part[:rows].double().sum(0)
    assert rel(s[:96], z.double().sum(0)) < 1e-5
    assert rel(s[96:], (z.double() ** 2).sum(0)) < 1e-5


WG_CASES = [
    # B, H, W, Cin, Cout, k
    (2, 13, 13, 64, 128, 3), (4, 26, 26, 128, 64, 1), (1, 52, 52, 64, 128, 3), (2, 13, 13, 1024, 640, 1),
    (2, 13, 13, 1280, 1024, 3), (3, 6, 6, 1024, 1024, 3), (2, 19, 19, 256, 512, 3), (8, 104, 104, 64, 128, 3),
    (1, 26, 26, 512, 64, 1), (3, 4, 4, 128, 256, 3), (3, 8, 8, 64, 128, 3), (3, 4, 4, 256, 256, 3), (1, 2, 2, 64, 64, 3),
]


@pytest.mark.parametrize('B,H,W,Cin,Cout,k', WG_CASES)
def test_conv_tc_wgrad(L, B, H, W, Cin, Cout, k):
    g = torch.Generator(device='cuda').manual_seed(B + H + Cin + Cout + 1)
    x = torch.randn(B, Cin, H, W, device='cuda', generator=g)
    w = (torch.randn(Cout, Cin, k, k, device='cuda', generator=g) * 0.05).requires_grad_(True)
    dz = torch.randn(B, Cout, H, W, device='cuda', generator=g)
    F.conv2d(x, w, None, 1, (k - 1) // 2).backward(dz)
    xh, xl = split(L, nhwc(x))
    dh, dl = split(L, nhwc(dz))
    assert L.lib.fsdet_conv_tc_wgrad_supported(Cin, Cout, k)
    nws = L.lib.fsdet_conv_tc_wgrad_workspace_floats(B, H, W, Cin, Cout, k)
    ws = torch.empty(max(nws, 4), device='cuda')
    dw = torch.full((Cout, k * k, Cin), 7.0, device='cuda')
    L.call('fsdet_conv_tc_wgrad', xh.data_ptr(), xl.data_ptr(), dh.data_ptr(), dl.data_ptr(), dw.data_ptr(), ws.data_ptr(), nws,
           B, H, W, Cin, Cout, k, st())
    torch.cuda.synchronize()
    assert rel(dw.view(Cout, k, k, Cin).permute(0, 3, 1, 2), w.grad) < 5e-5


def test_conv_tc_padded_channels_and_small_cout(L):
    """32-channel layers: planes zero-padded to 64 channels; dgrad with 32 output channels (weight rows < tile)."""
    B, H, W, Cin, Cout, k = 2, 26, 26, 32, 64, 3
    g = torch.Generator(device='cuda').manual_seed(3)
    x = torch.randn(B, Cin, H, W, device='cuda', generator=g, requires_grad=True)
    w = (torch.randn(Cout, Cin, k, k, device='cuda', generator=g) * 0.05).requires_grad_(True)
    dz = torch.randn(B, Cout, H, W, device='cuda', generator=g)
    ref = F.conv2d(x, w, None, 1, 1)
    ref.backward(dz)
    xh, xl = split(L, nhwc(x.detach()), 64)
    wrows = w.detach().permute(0, 2, 3, 1).contiguous().view(Cout * k * k, Cin)
    wh, wl = split(L, wrows, 64)
    z = torch.zeros(B * H * W, Cout, device='cuda')
    L.call('fsdet_conv_tc_fwd', xh.data_ptr(), xl.data_ptr(), wh.data_ptr(), wl.data_ptr(), z.data_ptr(), Cout, B, H, W, 64, Cout, k, 0, st())
    assert rel(z.view(B, H, W, Cout).permute(0, 3, 1, 2), ref) < 5e-5
    # dgrad: GEMM Cin = 64 (dz channels), Cout = 32
    wt = torch.empty(Cin, k * k, Cout, device='cuda')
    L.call('fsdet_weight_flip_transpose', w.detach().permute(0, 2, 3, 1).contiguous().data_ptr(), wt.data_ptr(), Cout, k * k, Cin, st())
    th, tl = split(L, wt.view(Cin, k * k * Cout))
    dh, dl = split(L, nhwc(dz))
    dx = torch.zeros(B * H * W, Cin, device='cuda')
    L.call('fsdet_conv_tc_fwd', dh.data_ptr(), dl.data_ptr(), th.data_ptr(), tl.data_ptr(), dx.data_ptr(), Cin, B, H, W, Cout, Cin, k, 0, st())
    assert rel(dx.view(B, H, W, Cin).permute(0, 3, 1, 2), x.grad) < 5e-5
    # wgrad with padded input channels: result [Cout][9][64], first 32 channels valid, rest zero
    nws = L.lib.fsdet_conv_tc_wgrad_workspace_floats(B, H, W, 64, Cout, k)
    ws = torch.empty(max(nws, 4), device='cuda')
    dw = torch.full((Cout, k * k, 64), 7.0, device='cuda')
    L.call('fsdet_conv_tc_wgrad', xh.data_ptr(), xl.data_ptr(), dh.data_ptr(), dl.data_ptr(), dw.data_ptr(), ws.data_ptr(), nws, B, H, W, 64, Cout, k, st())
    assert rel(dw[:, :, :32].reshape(Cout, k, k, Cin).permute(0, 3, 1, 2), w.grad) < 5e-5
    assert (dw[:, :, 32:] == 0).all()
