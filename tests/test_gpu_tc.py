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


def split(L, t2d):
    rows, C = t2d.shape
    hi = torch.empty(rows, C, dtype=torch.bfloat16, device='cuda')
    lo = torch.empty(rows, C, dtype=torch.bfloat16, device='cuda')
    L.call('fsdet_split_bf16', t2d.data_ptr(), C, C, rows, hi.data_ptr(), lo.data_ptr(), st())
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
