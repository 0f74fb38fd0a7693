"""build_targets / RegionLoss / RegionLossV2 on the GPU against the golden fixtures
(reference outputs) and against the CPU oracle on larger seeded inputs."""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
NAMES = ['coord_mask', 'conf_mask', 'cls_mask', 'tx', 'ty', 'tw', 'th', 'tconf', 'tcls']


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def _bt(pred, target, anchors, A, H, W, seen):
    from fewshot_detection_b200.region_loss import build_targets
    r = build_targets(torch.from_numpy(pred).cuda(), torch.from_numpy(target), anchors, A, 1, H, W, 1.0, 5.0, 0.6, seen)
    return r[0], r[1], [t.cpu().numpy() for t in r[2:]]


@pytest.mark.parametrize('tag', ['g13_seen0', 'g13_seen20000', 'g19_seen20000', 'g10_seen12800'])
def test_build_targets_golden_bit_exact(tag):
    d = np.load(os.path.join(G, 'build_targets_%s.npz' % tag))
    nGT, nCorrect, outs = _bt(d['pred_boxes'], d['target'], [float(a) for a in d['anchors']], 5, int(d['nH']), int(d['nW']),
                              int(d['seen']))
    assert nGT == int(d['nGT']) and nCorrect == int(d['nCorrect'])
    for k, v in zip(NAMES, outs):
        if k in ('tw', 'th'):
            # double log() may differ from glibc by 1 ulp before the float32 rounding
            assert np.allclose(v, d[k], rtol=0, atol=1e-6), k
        else:
            assert np.array_equal(v.view(np.uint32), d[k].view(np.uint32)), k


@pytest.mark.parametrize('bs,cs,Gs,seen', [(64, 15, 13, 20000), (8, 20, 19, 0), (5, 80, 19, 20000)])
def test_build_targets_vs_oracle_large(bs, cs, Gs, seen):
    from oracle import region_loss as ORL
    from fewshot_detection_b200 import netcfg
    from seeding import synth_targets
    anchors = [float(a) for a in netcfg.VOC_ANCHORS.split(',')]
    tgt = synth_targets(bs, cs, 5, max_gt=5).reshape(bs * cs, 250)
    nB = bs * cs
    g = torch.Generator().manual_seed(3)
    n = nB * 5 * Gs * Gs
    o = torch.randn(n, 4, generator=g)
    cell = torch.arange(n) % (Gs * Gs)
    aw = torch.tensor(anchors[0::2]).repeat_interleave(Gs * Gs).repeat(nB)
    ah = torch.tensor(anchors[1::2]).repeat_interleave(Gs * Gs).repeat(nB)
    pred = torch.stack([torch.sigmoid(o[:, 0]) + (cell % Gs).float(), torch.sigmoid(o[:, 1]) + (cell // Gs).float(),
                        torch.exp(o[:, 2] * 0.5) * aw, torch.exp(o[:, 3] * 0.5) * ah], 1).contiguous().numpy()
    ref = ORL.build_targets(pred, tgt, anchors, 5, Gs, Gs, 1.0, 5.0, 0.6, seen)
    nGT, nCorrect, outs = _bt(pred, tgt, anchors, 5, Gs, Gs, seen)
    assert (nGT, nCorrect) == (ref[0], ref[1]) and nGT > 0
    for k, v, r in zip(NAMES, outs, ref[2:]):
        if k in ('tw', 'th'):
            assert np.allclose(v, r, rtol=0, atol=1e-6), k
        else:
            assert np.array_equal(v.view(np.uint32), r.view(np.uint32)), k


def test_build_targets_edge_cases():
    from fewshot_detection_b200 import netcfg
    anchors = [float(a) for a in netcfg.VOC_ANCHORS.split(',')]
    # all-empty targets: nothing assigned, conf_mask = noobject everywhere
    pred = np.abs(np.random.RandomState(0).randn(2 * 5 * 169, 4)).astype(np.float32) + 0.1
    tgt = np.zeros((2, 250))
    nGT, nC, outs = _bt(pred, tgt, anchors, 5, 13, 13, 20000)
    assert nGT == 0 and nC == 0
    assert (outs[1] == 1).all() and all((o == 0).all() for i, o in enumerate(outs) if i != 1)
    # degenerate box (h == 0): the reference raises (math.log(0)); so do we
    tgt[0, :5] = [0, 0.5, 0.5, 0.2, 0.0]
    with pytest.raises(ValueError):
        _bt(pred, tgt, anchors, 5, 13, 13, 20000)


def _loss_mod(cls, anchors, nC, seen):
    L = cls()
    L.anchors, L.num_anchors, L.anchor_step, L.num_classes = anchors, 5, 2, nC
    L.object_scale, L.noobject_scale, L.class_scale, L.coord_scale = 5.0, 1.0, 1.0, 1.0
    L.seen = seen
    return L


@pytest.mark.parametrize('name', ['region_loss_v2_full', 'region_loss_v2_full_warm', 'region_loss_v2_neg1', 'region_loss_v2_neg0'])
def test_region_loss_v2_golden(name, capsys):
    from fewshot_detection_b200.region_loss import RegionLossV2
    from fewshot_detection_b200.cfg import cfg
    d = np.load(os.path.join(G, name + '.npz'))
    nr = str(d['neg_ratio'])
    cfg.neg_ratio = nr if nr == 'full' else int(nr)
    try:
        L = _loss_mod(RegionLossV2, [float(a) for a in d['anchors']], 1, int(d['seen']))
        o = torch.from_numpy(d['output']).cuda().requires_grad_(True)
        random.seed(int(d['pyseed']))
        loss = L(o, torch.from_numpy(d['target']))
        loss.backward()
    finally:
        cfg.neg_ratio = 'full'
    assert abs(loss.item() - float(d['loss'])) <= 1e-5 * abs(float(d['loss']))
    assert rel(o.grad.cpu().numpy(), d['grad']) < 1e-5
    # integer fields of the reference's own log line
    ref_line = str(d['log_line'])
    assert 'nGT %d, recall %d, proposals %d,' % (L.last['nGT'], L.last['nCorrect'], L.last['nProposals']) in ref_line
    out = capsys.readouterr().out.strip().splitlines()[-1]
    assert out.split(', loss')[0] == ref_line.split(', loss')[0]


def test_region_loss_plain_golden():
    from fewshot_detection_b200.region_loss import RegionLoss
    from fewshot_detection_b200.cfg import cfg
    d = np.load(os.path.join(G, 'region_loss_plain.npz'))
    for my in (True, False):
        k = 'metayolo1' if my else 'metayolo0'
        cfg.metayolo = my
        try:
            L = _loss_mod(RegionLoss, [float(a) for a in d['anchors']], 20, int(d['seen']))
            o = torch.from_numpy(d['output']).cuda().requires_grad_(True)
            loss = L(o, torch.from_numpy(d['target']))
            loss.backward()
        finally:
            cfg.metayolo = True
        assert abs(loss.item() - float(d['loss_' + k])) <= 1e-5 * abs(float(d['loss_' + k]))
        assert rel(o.grad.cpu().numpy(), d['grad_' + k]) < 1e-5


def test_region_loss_v2_vs_oracle_config2_shape():
    """B=64, n_cls=15 head output (the BASELINE config #2 shape), neg=full."""
    from fewshot_detection_b200.region_loss import RegionLossV2
    from fewshot_detection_b200 import netcfg
    from oracle import region_loss as ORL
    from seeding import synth_targets
    anchors = [float(a) for a in netcfg.VOC_ANCHORS.split(',')]
    bs, cs = 64, 15
    tgt = torch.from_numpy(synth_targets(bs, cs, 9, max_gt=5))
    o_cpu = (torch.randn(bs * cs, 30, 13, 13, generator=torch.Generator().manual_seed(9)) * 0.7)
    oc = o_cpu.clone().requires_grad_(True)
    ref, parts = ORL.region_loss_v2(oc, tgt, anchors, 5, 1, seen=20000, return_parts=True)
    ref.backward()
    L = _loss_mod(RegionLossV2, anchors, 1, 20000)
    og = o_cpu.cuda().requires_grad_(True)
    loss = L(og, tgt)
    loss.backward()
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item())
    assert rel(og.grad.cpu().numpy(), oc.grad.numpy()) < 1e-5
    assert (L.last['nGT'], L.last['nCorrect'], L.last['nProposals']) == (parts['nGT'], parts['nCorrect'], parts['nProposals'])
