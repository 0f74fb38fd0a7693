"""csrc/sgd.cu (the fused multi-tensor SGD of the training step) on the CPU through tools/host_emul, against
torch.optim.SGD as the reference driver configures it (train_meta.py:143-147: momentum 0.9, dampening 0, weight decay
on every parameter), incl. the chunk table, unaligned tensors, the first step and device-resident hyper-parameters."""
import ctypes

import numpy as np
import pytest
import torch

from emul_util import build_emul


@pytest.fixture(scope='module')
def emul():
    return build_emul('sgd', 'sgd.cu')


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def chunk_table(sizes, chunk):
    tens, offs = [], []
    for t, n in enumerate(sizes):
        for o in range(0, n, chunk):
            tens.append(t)
            offs.append(o)
    return np.array(tens, dtype=np.int32), np.array(offs, dtype=np.int64)


@pytest.mark.parametrize('use_hyper', [False, True])
def test_fused_sgd_equals_torch_sgd(emul, use_hyper):
    rs = np.random.RandomState(3)
    sizes = [1000, 7, 4096, 33, 12 * 1024 + 5]
    lr, mom, damp, wd = 1e-2, 0.9, 0.0, 0.48
    # one flat backing store with deliberately unaligned tensor starts for some tensors
    pad = [0, 1, 0, 3, 0]
    params = [rs.randn(n + p).astype(np.float32)[p:] for n, p in zip(sizes, pad)]
    moms = [np.full(n + p, 123.0, dtype=np.float32)[p:] for n, p in zip(sizes, pad)]     # garbage: the first step must ignore it
    tp = [torch.from_numpy(p.copy()).requires_grad_(True) for p in params]
    opt = torch.optim.SGD(tp, lr=lr, momentum=mom, dampening=damp, weight_decay=wd)
    chunk = 1024
    ct, co = chunk_table(sizes, chunk)
    sz = np.array(sizes, dtype=np.int64)
    hyper = np.array([lr, mom, damp, wd], dtype=np.float32)
    for step in range(3):
        grads = [rs.randn(n).astype(np.float32) for n in sizes]
        for t, g in zip(tp, grads):
            t.grad = torch.from_numpy(g.copy())
        opt.step()
        pt = (ctypes.c_void_p * len(sizes))(*[p.ctypes.data for p in params])
        gt = (ctypes.c_void_p * len(sizes))(*[g.ctypes.data for g in grads])
        mt = (ctypes.c_void_p * len(sizes))(*[m.ctypes.data for m in moms])
        args = (lr, mom, damp, wd) if not use_hyper else (0.0, 0.0, 0.0, 0.0)   # overridden by the table when given
        emul.emul_sgd_step(pt, gt, mt, P(sz), P(ct), P(co), len(ct), chunk, *[ctypes.c_float(a) for a in args], int(step == 0),
                           P(hyper) if use_hyper else None)
        for mine, ref in zip(params, tp):
            np.testing.assert_allclose(mine, ref.detach().numpy(), rtol=2e-6, atol=1e-7)
    assert all(np.isfinite(m).all() and not (m == 123.0).any() for m in moms)
