"""Control flow of the EXPERIMENTAL persistent tcgen05 convolution (csrc/conv_tc_persist.cuh, FSDET_TC_PERSIST=1, not
yet run on a GPU) on the CPU: the kernel source is compiled against functional models of its PTX wrappers
(tools/host_emul/conv_persist_emul.cpp: mbarrier phases and transaction counts, im2col / tiled TMA loads, tcgen05.mma
into a TMEM array, commit, tcgen05.ld, swizzled TMA store) and must reproduce the convolution for persistent grids
smaller than, equal to and larger than the tile count.  A wrong barrier phase deadlocks (reported as -100 after a
timeout) or corrupts the result.  Descriptors, swizzle modes and the instruction descriptor are NOT what is tested
here - the kernel shares those, unchanged, with the GPU-verified conv_tc_kernel."""
import ctypes

import numpy as np
import pytest

from emul_util import build_emul


@pytest.fixture(scope='module')
def emul():
    return build_emul('conv_persist', 'conv_tc_persist.cuh')


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def scale_from_amax(a):
    """conv_tc.cu: the power of two that maps the absolute maximum into [512, 1024)."""
    if not (a > 0) or not np.isfinite(a):
        return 1.0
    m, ex = np.frexp(np.float32(a))          # a = m * 2^ex, m in [0.5, 1)
    return float(2.0 ** (10 - int(ex)))


def split_planes(x):
    """fp32 tensor -> (hi, lo) fp16 planes of s*x as uint16 bit patterns, amax, and the exactly representable values."""
    amax = np.float32(np.abs(x).max())
    s = np.float32(scale_from_amax(amax))
    f = (x.astype(np.float32) * s).astype(np.float32)
    hi = f.astype(np.float16)
    lo = (f - hi.astype(np.float32)).astype(np.float16)
    exact = (hi.astype(np.float64) + lo.astype(np.float64)) / float(s)
    return hi.view(np.uint16), lo.view(np.uint16), np.array([amax], dtype=np.float32), exact


def conv_ref(x, w, k):
    """x [B,H,W,Cin], w [Cout,k*k,Cin] (float64) -> [B*H*W, Cout], stride 1, same padding."""
    B, H, W, Cin = x.shape
    pad = (k - 1) // 2
    xp = np.zeros((B, H + 2 * pad, W + 2 * pad, Cin))
    xp[:, pad:pad + H, pad:pad + W] = x
    out = np.zeros((B, H, W, w.shape[0]))
    for r in range(k):
        for s in range(k):
            out += np.einsum('bhwc,oc->bhwo', xp[:, r:r + H, s:s + W], w[:, r * k + s])
    return out.reshape(B * H * W, -1)


CASES = [
    # B, H, W, Cin, Cout, k, bn, stages, ctas, accumulate
    (2, 16, 16, 32, 64, 3, 64, 6, 3, 0),      # conv2-like: 4 tiles over 3 CTAs (uneven), 9 k-blocks per tile
    (2, 16, 16, 32, 64, 3, 64, 6, 1, 0),      # one CTA walks all tiles: accumulator sets alternate 4 times
    (2, 16, 16, 32, 64, 3, 64, 6, 7, 0),      # more CTAs than tiles: some CTAs have nothing to do
    (1, 13, 13, 64, 32, 3, 64, 6, 2, 0),      # dgrad-like: Cout < BN, M = 169 not a multiple of 128 (clipped rows)
    (2, 12, 12, 64, 200, 1, 128, 4, 2, 0),    # 1x1, BN = 128, two N tiles (the second one partial), 2 k-blocks per tile
    (1, 20, 20, 32, 64, 3, 64, 2, 2, 1),      # accumulate into z (TMA reduce-add), only 2 stages: every stage wraps many times
    (3, 8, 8, 96, 64, 3, 64, 6, 2, 0),        # 27 k-blocks per tile, 3 channel chunks per tap
]


@pytest.mark.parametrize('B,H,W,Cin,Cout,k,bn,stages,ctas,acc', CASES)
def test_persistent_kernel_control_flow(emul, B, H, W, Cin, Cout, k, bn, stages, ctas, acc):
    rs = np.random.RandomState(B * 100 + H + Cin + Cout + ctas)
    x = rs.randn(B, H, W, Cin).astype(np.float32)
    w = (rs.randn(Cout, k * k, Cin) * 0.1).astype(np.float32)
    xh, xl, ax, xe = split_planes(x)
    wh, wl, aw, we = split_planes(w)
    M = B * H * W
    ld = Cout + 4
    z0 = rs.randn(M, ld).astype(np.float32) if acc else np.full((M, ld), 7.0, dtype=np.float32)
    z = z0.copy()
    rc = emul.emul_conv_tc_persist(P(xh), P(xl), P(wh), P(wl), P(ax), P(aw), P(z), ld, B, H, W, Cin, Cin, Cout, k, acc, bn,
                                   stages, ctas)
    assert rc == 0, 'barrier deadlock in the persistent kernel' if rc == -100 else rc
    ref = conv_ref(xe, we, k)                     # exact product of the (hi + lo) operands; the lo*lo term is dropped
    got = z[:, :Cout].astype(np.float64) - (z0[:, :Cout] if acc else 0)
    err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    assert err < 2e-6, err
    assert np.array_equal(z[:, Cout:], z0[:, Cout:])      # columns beyond Cout are never written (clipped stores)


def test_slow_epilogue_does_not_lose_accumulators(emul):
    """With a slow epilogue the MMA issuer runs ahead: it must wait until the epilogue has handed an accumulator set
    back (acc_empty) before overwriting it - one CTA, four tiles, two sets."""
    emul.emul_set_ld_delay_us(30000)
    try:
        test_persistent_kernel_control_flow(emul, *CASES[1])
    finally:
        emul.emul_set_ld_delay_us(0)
