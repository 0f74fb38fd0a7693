"""Pin the CPU oracle against fixtures produced by the reference itself
(tests/golden/make_golden.py).  CPU only."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import region_loss as ORL
from oracle import darknet as ODK
from fewshot_detection_b200 import netcfg
from seeding import seeded_init, synth_masks

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def test_iou_bit_exact():
    d = load('iou.npz')
    got = ORL.bbox_ious(d['b1'], d['b2'])
    assert np.array_equal(got.view(np.uint32), d['ious_f32'].view(np.uint32))
    sc = np.array([ORL.bbox_iou([float(v) for v in d['b1'][:, i]], [float(v) for v in d['b2'][:, i]])
                   for i in range(d['b1'].shape[1])])
    assert np.array_equal(sc, d['ious_f64'])


@pytest.mark.parametrize('tag', ['g13_seen0', 'g13_seen20000', 'g19_seen20000', 'g10_seen12800'])
def test_build_targets_bit_exact(tag):
    d = load('build_targets_%s.npz' % tag)
    r = ORL.build_targets(d['pred_boxes'], d['target'], [float(a) for a in d['anchors']], 5, int(d['nH']), int(d['nW']),
                          1.0, 5.0, 0.6, int(d['seen']))
    names = ['nGT', 'nCorrect', 'coord_mask', 'conf_mask', 'cls_mask', 'tx', 'ty', 'tw', 'th', 'tconf', 'tcls']
    for k, v in zip(names, r):
        if isinstance(v, int):
            assert v == int(d[k]), k
        else:
            assert np.array_equal(v.view(np.uint32), d[k].view(np.uint32)), k
    assert int(d['nGT']) > 0


@pytest.mark.parametrize('name', ['region_loss_v2_full', 'region_loss_v2_full_warm', 'region_loss_v2_neg1', 'region_loss_v2_neg0'])
def test_region_loss_v2(name):
    d = load(name + '.npz')
    nr = str(d['neg_ratio'])
    nr = nr if nr == 'full' else int(nr)
    o = torch.from_numpy(d['output']).requires_grad_(True)
    random.seed(int(d['pyseed']))
    loss, parts = ORL.region_loss_v2(o, torch.from_numpy(d['target']), [float(a) for a in d['anchors']], 5, 1,
                                     seen=int(d['seen']), neg_ratio=nr, return_parts=True)
    loss.backward()
    assert list(parts['inds']) == list(d['inds'])
    assert abs(loss.item() - float(d['loss'])) <= 1e-6 * abs(float(d['loss']))
    assert rel(o.grad.numpy(), d['grad']) < 1e-6
    # the reference's own log line carries nGT / recall / proposals
    line = str(d['log_line'])
    assert 'nGT %d, recall %d, proposals %d,' % (parts['nGT'], parts['nCorrect'], parts['nProposals']) in line


def test_region_loss_plain():
    d = load('region_loss_plain.npz')
    for my in (True, False):
        k = 'metayolo1' if my else 'metayolo0'
        o = torch.from_numpy(d['output']).requires_grad_(True)
        loss = ORL.region_loss_plain(o, torch.from_numpy(d['target']), [float(a) for a in d['anchors']], 5, 20,
                                     seen=int(d['seen']), metayolo=my)
        loss.backward()
        assert abs(loss.item() - float(d['loss_' + k])) <= 1e-6 * abs(float(d['loss_' + k]))
        assert rel(o.grad.numpy(), d['grad_' + k]) < 1e-6


def test_layers():
    d = load('layers.npz')
    x = torch.from_numpy(d['x'])
    assert np.array_equal(ODK.Reorg(2)(x).numpy(), d['reorg'])
    assert np.array_equal(ODK.MaxPoolStride1()(x).numpy(), d['maxpool_stride1'])
    assert np.array_equal(ODK.GlobalMaxPool2d()(x).numpy(), d['globalmax'])
    y = ODK.DynamicConv2d()((x, torch.from_numpy(d['dyn_w'])))
    assert np.array_equal(y.numpy(), d['dyn_out'])


def _run_meta(d, det, ler, regen_inputs):
    seed = int(d['seed'])
    m = ODK.MetaDarknet(det, ler)
    seeded_init(m, seed)
    m.train()
    bs, cs, side, ms = int(d['bs']), int(d['cs']), int(d['side']), int(d['meta_side'])
    if regen_inputs:
        g = torch.Generator().manual_seed(seed + 1)
        x = torch.rand(bs, 3, side, side, generator=g)
        metax = torch.rand(cs, 3, ms, ms, generator=g)
        mask = torch.from_numpy(synth_masks(cs, ms, seed + 2))
    else:
        x, metax, mask = (torch.from_numpy(d[k]) for k in ('x', 'metax', 'mask'))
    out = m(x, metax, mask)
    loss = ORL.region_loss_v2(out, torch.from_numpy(d['target']), m.anchors, m.num_anchors, m.num_classes,
                              seen=int(d['seen']))
    loss.backward()
    return m, out, loss


def test_meta_mini_full_tensors():
    d = load('meta_mini.npz')
    m, out, loss = _run_meta(d, netcfg.mini_dynamic_blocks(128, 4), netcfg.mini_reweighting_blocks(64, 4, 128), False)
    assert rel(out.detach().numpy(), d['output']) < 1e-6
    assert abs(loss.item() - float(d['loss'])) < 1e-5 * abs(float(d['loss']))
    n = 0
    for name, p in m.named_parameters():
        assert rel(p.grad.numpy(), d['grad/' + name]) < 1e-5, name
        n += 1
    assert n == len([k for k in d.files if k.startswith('grad/')])
    # the generator ran the support branch a second time (train mode, no_grad)
    with torch.no_grad():
        dw = m.meta_forward(torch.from_numpy(d['metax']), torch.from_numpy(d['mask']))
    assert rel(dw[0].numpy(), d['dynamic_weights_2nd_pass']) < 1e-6
    for name, b in m.named_buffers():
        if 'running' in name:
            assert rel(b.numpy(), d['buf/' + name]) < 1e-6, name


def test_meta_full416_digest():
    d = load('meta_full416.npz')
    m, out, loss = _run_meta(d, netcfg.darknet_dynamic_blocks(), netcfg.reweighting_net_blocks(), True)
    assert rel(out.detach().numpy(), d['output']) < 1e-5
    assert abs(loss.item() - float(d['loss'])) < 1e-5 * abs(float(d['loss']))
    for name, p in m.named_parameters():
        gn = float(d['gradnorm/' + name])
        assert abs(p.grad.double().norm().item() - gn) < 1e-4 * gn + 1e-12, name
        assert rel(p.grad.reshape(-1)[:64].numpy(), d['gradhead/' + name]) < 1e-3, name


def test_tiny_yolo_416_config1():
    d = load('tiny_yolo_416.npz')
    m = ODK.PlainDarknet(netcfg.tiny_yolo_voc_blocks())
    seeded_init(m, int(d['w_seed']))
    x = torch.rand(1, 3, 416, 416, generator=torch.Generator().manual_seed(int(d['x_seed'])))
    m.eval()
    with torch.no_grad():
        y = m(x).numpy()
    assert y.shape == (1, 125, 13, 13)
    assert rel(y, d['y_eval']) < 1e-6
    m.train()
    assert rel(m(x).detach().numpy(), d['y_train']) < 1e-5
