#!/usr/bin/env python
"""Mint tests/golden/weights.npz with the REFERENCE's own darknet_meta.Darknet.save_weights / load_weights
(imported through make_golden.load_ref: mechanical Py3 patches only) - build container only:

    python tests/golden/make_golden_weights.py

Stored: the byte streams the reference WROTE for a seeded mini meta-model (complete, and with cutoff = 12), and what
the reference's loader LEFT in a differently seeded model after reading (a) the complete stream and (b) a stream that
ends after the detector's third convolution (the darknet19_448.conv.23 situation: loading stops silently and every
later tensor keeps its initialisation) - as per-parameter float64 sums plus the first values.
"""
import io
import os
import sys
import tempfile
from contextlib import redirect_stdout

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG                       # noqa: E402  (sets sys.path for the reference and the shims)
from seeding import seeded_init                # noqa: E402
from fewshot_detection_b200 import netcfg      # noqa: E402


def digest(m):
    out = {}
    for name, t in list(m.named_parameters()) + [(n, b) for n, b in m.named_buffers() if 'running' in n]:
        v = t.detach().double().reshape(-1)
        out['sum/' + name] = np.float64(v.sum().item())
        out['head/' + name] = v[:8].numpy().copy()
    return out


class legacy_copy(object):
    """torch 0.3.1's Tensor.copy_ accepted a source of equal numel and different shape (cfg.py:411-470 copies flat
    slices of the weight stream into 4-D weights); torch >= 0.4 wants broadcastable shapes.  Re-create the old rule
    around the reference's loader."""

    def __enter__(self):
        self.orig = torch.Tensor.copy_
        orig = self.orig

        def copy_(dst, src, *a, **k):
            if src.shape != dst.shape and src.numel() == dst.numel():
                src = src.reshape(dst.shape)
            return orig(dst, src, *a, **k)
        torch.Tensor.copy_ = copy_

    def __exit__(self, *exc):
        torch.Tensor.copy_ = self.orig


def main():
    with redirect_stdout(io.StringIO()):
        MG.importlib.import_module('utils')
        MG.importlib.import_module('cfg')
    DM = MG.load_ref('darknet_meta')
    det, ler = netcfg.mini_dynamic_blocks(128, 4), netcfg.mini_reweighting_blocks(64, 4, 128)

    def model(seed):
        with redirect_stdout(io.StringIO()):
            m = DM.Darknet([dict(b) for b in det], [dict(b) for b in ler])
        return seeded_init(m, seed)

    src = model(81)
    src.seen = 4242
    out = {'seed_written': 81, 'seed_loaded_into': 82, 'seen': 4242}
    with tempfile.TemporaryDirectory() as td:
        f_all, f_cut = os.path.join(td, 'all.weights'), os.path.join(td, 'cut.weights')
        src.save_weights(f_all)
        src.save_weights(f_cut, cutoff=12)
        raw_all = np.fromfile(f_all, dtype=np.uint8)
        raw_cut = np.fromfile(f_cut, dtype=np.uint8)
        out['stream_all'], out['stream_cutoff12'] = raw_all, raw_cut
        # (a) complete stream into a differently initialised model
        dst = model(82)
        with legacy_copy():
            dst.load_weights(f_all)
        out['loaded_seen'] = int(dst.seen)
        out.update({'all/' + k: v for k, v in digest(dst).items()})
        # (b) truncated stream: header + the first three detector convolutions (conv+BN each)
        n = 0
        convs = [mod for mod in src.models if isinstance(mod, torch.nn.Sequential)][:3]
        for seq in convs:
            n += seq[0].weight.numel() + 4 * seq[1].weight.numel()
        f_tr = os.path.join(td, 'trunc.weights')
        raw_all[:16 + 4 * n].tofile(f_tr)
        out['truncated_floats'] = n
        dst = model(82)
        with legacy_copy():
            dst.load_weights(f_tr)
        out.update({'trunc/' + k: v for k, v in digest(dst).items()})
    np.savez_compressed(os.path.join(HERE, 'weights.npz'), **out)
    print('wrote weights.npz: %d + %d stream bytes, %d truncated floats' % (raw_all.size, raw_cut.size, n))


if __name__ == '__main__':
    main()
