"""fewshot_detection_b200.trainer (the loop of train_meta.py:86-255) on the CPU: schedule arithmetic against values
worked out from the reference's formulas and cfg/darknet_dynamic.cfg, and the whole epoch loop with a stub model
(the data side is the real DetectionBatcher / MetaBatcher, C-ABI calls routed to the host-emulated kernels)."""
import os
import random

import numpy as np
import pytest
import torch
import torch.nn as nn

from emul_util import build_emul, route_image_calls_to_emulation
from fewshot_detection_b200 import trainer as T

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
STEPS, SCALES = [-1, 500, 40000, 60000], [0.1, 10, .1, .1]      # cfg/darknet_dynamic.cfg:24-25


def test_lr_factor_and_sgd_hyper_parameters():
    assert [T.lr_factor(n, 15) for n in ('full', 1, 0, 5, 3)] == [15., 3.0, 1.5, 8.0, 15]
    hp = T.sgd_hyper_parameters(0.001, 0.9, 0.0005, 64, 15.)
    assert hp['lr'] == 0.001 / 15. / 64 and hp['weight_decay'] == 0.0005 * 64 * 15. and hp['dampening'] == 0


def test_learning_rate_schedule():
    lr0 = 0.001 / 15.
    f = lambda b: T.learning_rate_at(b, lr0, STEPS, SCALES)
    # the reference's loop: burn-in at 0.1x until batch 500, back to 1x, /10 at 40000, /100 at 60000
    assert f(0) == lr0 * 0.1 and f(499) == lr0 * 0.1
    assert f(500) == lr0 * 0.1 * 10 and f(501) == lr0 * 0.1 * 10
    assert f(40000) == lr0 * 0.1 * 10 * .1 and f(59999) == lr0 * 0.1 * 10 * .1
    assert f(60000) == lr0 * 0.1 * 10 * .1 * .1 and f(10 ** 6) == lr0 * 0.1 * 10 * .1 * .1
    assert T.learning_rate_at(7, 1.0, [5, 10], [0.5]) == 0.5          # fewer scales than steps: missing scale = 1
    assert T.learning_rate_at(12, 1.0, [5, 10], [0.5]) == 0.5


def test_epoch_plan():
    assert T.epoch_plan(0, 1000, 64, 80200) == (0, 0, 80200 * 64 // 1000 + 1)
    assert T.epoch_plan(64000, 1000, 64, 80200) == (1000, 64, 5133)
    assert T.epoch_plan(64000, 1000, 64, 80200, tuning=True, max_epoch=500, repeat=200) == (0, 0, 3)


class StubModel(nn.Module):
    """The Darknet surface MetaTrainer touches: forward(x, metax, mask) -> [B*n_cls, 30, G, G], .loss, .seen,
    save_weights."""

    def __init__(self, n_cls):
        super().__init__()
        self.w = nn.Parameter(torch.ones(1))
        self.n_cls, self.seen, self.saved = n_cls, 0, []
        self.loss = StubLoss()

    def forward(self, x, metax, mask):
        assert x.dim() == 4 and x.size(1) == 3 and tuple(metax.shape[:2]) == (self.n_cls, 3) and mask.size(1) == 1
        g = x.size(-1) // 32
        return (x.mean() * self.w).expand(x.size(0) * self.n_cls, 30, g, g)

    def save_weights(self, path):
        self.saved.append((path, self.seen))


class StubLoss(nn.Module):
    seen = 0

    def forward(self, output, target):
        assert target.dtype == torch.float64 and target.size(0) * target.size(1) == output.size(0)
        return output.sum() * 1e-3


def test_epoch_loop_bookkeeping(monkeypatch, tmp_path):
    emul = build_emul('augment', 'augment.cu')
    route_image_calls_to_emulation(monkeypatch, emul)
    from fewshot_detection_b200.cfg import cfg
    from fewshot_detection_b200.dataset import DetectionBatcher, MetaBatcher
    gold = np.load(os.path.join(G, 'dataset.npz'), allow_pickle=False)
    saved = {k: cfg.get(k) for k in ('base_classes', 'base_ids', 'metain_type', 'meta_width', 'meta_height', 'mask_width',
                                      'mask_height', 'multiscale', 'metayolo')}
    ncls = 3
    cfg.base_classes, cfg.base_ids, cfg.metain_type, cfg.multiscale, cfg.metayolo = cfg.voc_classes[:ncls], list(range(ncls)), 2, 0, True
    cfg.meta_width = cfg.meta_height = cfg.mask_width = cfg.mask_height = 48
    try:
        lines = [(gold['src%d' % i], gold['lab%d' % i]) for i in range(8)]
        pool = gold['meta/pool']
        metalines = [[(gold['src%d' % i], gold['meta_lab/%d/%d' % (c, i)]) for i in pool[c] if i >= 0] for c in range(ncls)]
        inds = [tuple(int(v) for v in r) for r in gold['meta/inds']]
        model = StubModel(ncls)
        model.seen = 128
        opt = torch.optim.SGD(model.parameters(), **T.sgd_hyper_parameters(0.001, 0.9, 0.0005, 4, 15.))
        logs = []
        tr = T.MetaTrainer(model, opt, 0.001 / 15., 4, [-1, 1, 3], [0.1, 10, 0.1],
                           lambda seen: DetectionBatcher(lines, shape=(64, 64), shuffle=False, train=True, seen=seen, batch_size=4,
                                                         num_workers=1),
                           lambda: MetaBatcher(metalines, inds, train=True), backupdir=str(tmp_path), save_interval=2,
                           processed_batches=0, log=logs.append)
        assert model.loss.seen == 128                         # region_loss.seen = model.seen
        random.seed(3)
        lrs = []
        orig = tr.train_step
        tr.train_step = lambda *a: (lrs.append(opt.param_groups[0]['lr']), orig(*a))[1]
        tr.fit(0, 2)
        assert tr.processed_batches == 4 and len(tr.losses) == 4
        base = 0.001 / 15.
        want = [base * 0.1 / 4, base * 0.1 * 10 / 4, base * 0.1 * 10 / 4, base * 0.1 * 10 * 0.1 / 4]   # batches 0, 1, 2, 3
        assert lrs == want
        assert model.loss.seen == 128 + 4 * 4
        assert model.saved == [('%s/%06d.weights' % (tmp_path, 2), 2 * 8)]     # saved after the 2nd epoch, seen = 2 * len(dataset)
        assert model.w.grad is not None and float(model.w) != 1.0
        assert any(l.startswith('epoch 0/2') for l in logs) and any('samples/s' in l for l in logs)
    finally:
        for k, v in saved.items():
            if v is None:
                cfg.pop(k, None)
            else:
                cfg[k] = v


def test_cli_helpers(tmp_path):
    """The command-line front end's own helpers (list / index construction lives in lists.py, tests/test_lists.py)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('train_meta_b200', os.path.join(os.path.dirname(G), '..', 'tools', 'train_meta_b200.py'))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    (tmp_path / 'bird.txt').write_text(''.join('/d/JPEGImages/b%d.jpg\n' % i for i in range(3)) + '\n')
    assert cli.read_list(str(tmp_path / 'bird.txt')) == ['/d/JPEGImages/b0.jpg', '/d/JPEGImages/b1.jpg', '/d/JPEGImages/b2.jpg']
    assert cli.main.__doc__ is None and callable(cli.broadcast_parameters)
    import sys as _sys
    old = _sys.argv
    _sys.argv = ['train_meta_b200.py']
    try:
        assert cli.main() == 1          # usage message, like the reference script without its four arguments
    finally:
        _sys.argv = old