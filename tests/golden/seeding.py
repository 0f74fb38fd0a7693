"""Deterministic, constructor-independent parameter / input synthesis.

Shared by tests/golden/make_golden.py (which applies it to the *reference*
model imported from /root/reference), by the oracle tests and by the GPU parity
tests, so that all three see bit-identical weights and inputs without relying
on the RNG consumption order of any constructor.
"""
import math

import numpy as np
import torch


def seeded_init(model, seed):
    """Overwrite every parameter and BN buffer of `model` from one generator.

    Order = `named_parameters()` order, then BN buffers in `named_buffers()`
    order; both are determined by the module tree, which our Darknet mirrors
    from the reference (darknet_meta.py:208-353).
    """
    g = torch.Generator().manual_seed(int(seed))
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() == 4:  # conv weight [O, I, kh, kw]
                fan_in = p.shape[1] * p.shape[2] * p.shape[3]
                v = torch.randn(p.shape, generator=g) * math.sqrt(2.0 / fan_in)
            elif name.endswith('weight'):  # BN gamma
                v = torch.rand(p.shape, generator=g) + 0.5
            else:  # BN beta / conv bias
                v = torch.randn(p.shape, generator=g) * 0.1
            p.copy_(v.to(p.dtype).view_as(p))
        for name, b in model.named_buffers():
            if name.endswith('running_mean'):
                b.copy_(torch.randn(b.shape, generator=g) * 0.1)
            elif name.endswith('running_var'):
                b.copy_(torch.rand(b.shape, generator=g) + 0.5)
    return model


def synth_targets(bs, cs, seed, max_gt=5, empty_prob=0.0, slots=50):
    """Synthetic label tensor f64 [bs, cs, slots*5] in the layout written by the
    reference's image.fill_truth_detection_meta (image.py:144-192): per class
    row, consecutive 5-slots [class_idx, cx, cy, w, h], zero padded, list ends
    at the first slot whose cx == 0."""
    rs = np.random.RandomState(seed)
    t = np.zeros((bs, cs, slots * 5), dtype=np.float64)
    for b in range(bs):
        if rs.rand() < empty_prob:
            continue
        k = rs.randint(1, max_gt + 1)
        nfill = np.zeros(cs, dtype=np.int64)
        for _ in range(k):
            c = rs.randint(0, cs)
            w = rs.uniform(0.05, 0.8)
            h = rs.uniform(0.05, 0.8)
            cx = rs.uniform(w / 2, 0.999 - w / 2)
            cy = rs.uniform(h / 2, 0.999 - h / 2)
            s = nfill[c]
            if s >= slots:
                continue
            t[b, c, s * 5:(s + 1) * 5] = [c, cx, cy, w, h]
            nfill[c] += 1
    return t


def synth_masks(cs, side, seed):
    """One axis-aligned rectangle of ones per class (dataset.py:378-398)."""
    rs = np.random.RandomState(seed)
    m = np.zeros((cs, 1, side, side), dtype=np.float32)
    for c in range(cs):
        w = int(rs.uniform(0.1, 0.9) * side)
        h = int(rs.uniform(0.1, 0.9) * side)
        x0 = rs.randint(0, side - w + 1)
        y0 = rs.randint(0, side - h + 1)
        m[c, 0, y0:y0 + h, x0:x0 + w] = 1.0
    return m

