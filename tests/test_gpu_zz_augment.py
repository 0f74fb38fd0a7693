"""Training-input augmentation on the GPU (csrc/augment.cu) through the C ABI and the image.py mirror, against
tests/golden/augment.npz = outputs of the reference's own image.data_augmentation (Pillow 12.2).  Bar: bit-exact
(uint8 pipeline; the float output is uint8 / 255 in float32).  The same kernel source is checked against Pillow on
the CPU by tests/test_augment_host_emul.py.  (File name sorts last on purpose: newest GPU tests run last.)"""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
IMG_CASES = ['a', 'b', 'c', 'd', 'e', 'f', 'g']


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(G, 'augment.npz'), allow_pickle=False)


def to_tensor(u8):
    return np.ascontiguousarray((u8.astype(np.float32) / np.float32(255)).transpose(2, 0, 1))


@pytest.mark.parametrize('tag', IMG_CASES)
def test_data_augmentation_bit_exact_vs_reference(gold, tag):
    from fewshot_detection_b200 import image as I
    W, H, seed, flag = [int(v) for v in gold[tag + '/args']]
    src = gold[tag + '/src']
    random.seed(seed)
    img, flip, dx, dy, sx, sy = I.data_augmentation(src, (W, H), 0.2, 0.1, 1.5, 1.5, flag=bool(flag))
    assert img.is_cuda and tuple(img.shape) == (3, H, W)
    assert (flip, dx, dy, sx, sy) == tuple([int(gold[tag + '/params'][0])] + list(gold[tag + '/params'][1:]))
    assert np.array_equal(img.cpu().numpy().view(np.uint32), to_tensor(gold[tag + '/img']).view(np.uint32))
    # the uint8 image PIL would hold, and the NEAREST filter of older Pillow
    random.seed(seed)
    oh, ow = src.shape[:2]
    p = I.draw_augmentation(ow, oh, 0.2, 0.1, 1.5, 1.5) if flag else I.identity_augmentation(ow, oh)
    out, u8 = I.augment_batch([torch.from_numpy(src)], (W, H), [p], return_uint8=True)
    assert np.array_equal(u8[0].cpu().numpy(), gold[tag + '/img'])
    out0, u80 = I.augment_batch([src], (W, H), [p], filter=I.NEAREST, return_uint8=True)
    assert np.array_equal(u80[0].cpu().numpy(), gold['nearest/' + tag])
    assert np.array_equal(out0[0].cpu().numpy(), to_tensor(gold['nearest/' + tag]))


def test_augment_batch_mixed_source_sizes_one_launch(gold):
    from fewshot_detection_b200 import image as I
    tags = ['a', 'c', 'f', 'b']
    srcs, ps = [], []
    for t in tags:
        W, H, seed, flag = [int(v) for v in gold[t + '/args']]
        random.seed(seed)
        src = gold[t + '/src']
        srcs.append(torch.from_numpy(src).cuda() if t == 'c' else src)       # device-resident sources work too
        ps.append(I.draw_augmentation(src.shape[1], src.shape[0], 0.2, 0.1, 1.5, 1.5))
    out = torch.full((4, 3, 64, 64), -1.0, device='cuda')
    res = I.augment_batch(srcs, (64, 64), ps, out=out)
    assert res.data_ptr() == out.data_ptr()
    for i, t in enumerate(tags):
        if tuple(int(v) for v in gold[t + '/args'][:2]) == (64, 64):
            assert np.array_equal(out[i].cpu().numpy(), to_tensor(gold[t + '/img'])), t
    single = I.augment_batch([srcs[3]], (64, 64), [ps[3]])
    assert torch.equal(single[0], out[3])


def test_full_batch_properties():
    """BASELINE-sized batch (64 VOC-sized images -> 416x416): float output == uint8 output / 255, flips mirror."""
    from fewshot_detection_b200 import image as I
    rs = np.random.RandomState(3)
    random.seed(12)
    srcs = [rs.randint(0, 256, (int(rs.randint(300, 400)), int(rs.randint(400, 500)), 3)).astype(np.uint8) for _ in range(64)]
    ps = [I.draw_augmentation(s.shape[1], s.shape[0], 0.2, 0.1, 1.5, 1.5) for s in srcs]
    out, u8 = I.augment_batch(srcs, (416, 416), ps, return_uint8=True)
    torch.cuda.synchronize()
    assert tuple(out.shape) == (64, 3, 416, 416)
    # ToTensor divides on the CPU (IEEE division); torch's CUDA tensor/scalar division multiplies by a reciprocal
    assert torch.equal(out.cpu(), u8.cpu().permute(0, 3, 1, 2).float() / 255)
    assert 0.0 <= out.min().item() and out.max().item() <= 1.0
    k = next(i for i, p in enumerate(ps) if p['flip'])
    unflipped = dict(ps[k], flip=0)
    o2 = I.augment_batch([srcs[k]], (416, 416), [unflipped])
    assert torch.equal(o2[0].flip(-1), out[k])
    with pytest.raises(TypeError):
        I.augment_batch([srcs[0].astype(np.float32)], (416, 416), [ps[0]])


def test_box_masks_and_load_data_detection(gold):
    from fewshot_detection_b200 import image as I
    from fewshot_detection_b200.cfg import cfg
    boxes = [[0.5, 0.5, 0.25, 0.5], [0.1, 0.9, 0.4, 0.4], [0.3, 0.3, 0.0, 0.2]]
    m = I.box_masks(boxes, 64, 48)
    assert tuple(m.shape) == (3, 1, 48, 64)
    for i, b in enumerate(boxes):
        x1, y1, x2, y2 = I.mask_rect(b, 64, 48)
        want = np.zeros((48, 64), dtype=np.float32)
        want[y1:y2, x1:x2] = 1
        assert np.array_equal(m[i, 0].cpu().numpy(), want)
    old = (cfg.base_classes, cfg.base_ids, cfg.yolo_joint, cfg.metayolo)
    cfg.base_classes, cfg.base_ids, cfg.yolo_joint, cfg.metayolo = cfg.voc_classes[:15], list(range(15)), False, True
    try:
        random.seed(1)
        img, label = I.load_data_detection(gold['a/src'], gold['l1/boxes'].copy(), (64, 64), 0.2, 0.1, 1.5, 1.5)
        assert np.array_equal(img.cpu().numpy(), to_tensor(gold['a/img']))
        assert label.shape == (15, 250)
        flip, dx, dy, sx, sy = gold['a/params']
        want = I.fill_truth_detection_meta(gold['l1/boxes'].copy(), 64, 64, int(flip), dx, dy, 1. / sx, 1. / sy)
        assert np.array_equal(label, want)
    finally:
        cfg.base_classes, cfg.base_ids, cfg.yolo_joint, cfg.metayolo = old
