"""fewshot_detection_b200.dataset (batch-at-a-time listDataset / MetaDataset) against tests/golden/dataset.npz =
what the reference's own dataset.py returned sample by sample (tests/golden/make_golden_dataset.py).  Runs on the
CPU: the two C-ABI calls go to the host-emulated kernels (tests/emul_util.py).  Bit-exact."""
import os
import random

import numpy as np
import pytest

from emul_util import build_emul, route_image_calls_to_emulation

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.fixture(scope='module')
def emul():
    return build_emul('augment', 'augment.cu')


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(G, 'dataset.npz'), allow_pickle=False)


@pytest.fixture()
def cfg5():
    from fewshot_detection_b200.cfg import cfg
    keys = ('data', 'multiscale', 'metayolo', 'yolo_joint', 'classes', 'base_classes', 'base_ids', 'novel_ids', 'metain_type',
            'meta_width', 'meta_height', 'mask_width', 'mask_height')
    old = {k: cfg.get(k) for k in keys}
    cfg.data, cfg.multiscale, cfg.metayolo, cfg.yolo_joint = 'voc', 0, True, False
    cfg.classes = cfg.voc_classes
    cfg.base_classes, cfg.base_ids, cfg.novel_ids = cfg.voc_classes[:5], list(range(5)), [5, 6]
    yield cfg
    for k, v in old.items():
        if v is None:
            cfg.pop(k, None)
        else:
            cfg[k] = v


@pytest.mark.parametrize('mode,train', [('train', True), ('test', False)])
def test_detection_batcher_equals_reference_listDataset(emul, gold, cfg5, monkeypatch, mode, train):
    route_image_calls_to_emulation(monkeypatch, emul)
    from fewshot_detection_b200.dataset import DetectionBatcher
    lines = [(gold['src%d' % i], gold['lab%d' % i]) for i in range(8)]
    random.seed(31)
    ds = DetectionBatcher(lines, shape=(64, 64), shuffle=False, train=train, seen=0, batch_size=4, num_workers=1)
    imgs, labels = [], []
    for data, target in ds:                       # two batches of four, one launch each
        assert tuple(data.shape) == (4, 3, 64, 64) and tuple(target.shape) == (4, 5, 250)
        imgs.append(data.numpy())
        labels.append(target.numpy())
    assert np.array_equal(np.concatenate(imgs), gold['list_%s/img' % mode])
    assert np.array_equal(np.concatenate(labels), gold['list_%s/label' % mode])
    assert ds.seen == int(gold['list_%s/seen_after' % mode])


def test_multiscale_schedule_equals_reference(gold):
    """The fixture called the reference's __getitem__(64 * k) with `seen` forced into each regime; between two size
    draws the reference also consumes the augmentation draws of that sample, replayed here."""
    from fewshot_detection_b200.dataset import multiscale_width
    from fewshot_detection_b200 import image as I
    random.seed(32)
    got = []
    for k, s in enumerate(gold['multiscale/seens']):
        got.append(multiscale_width(int(s)))
        src = gold['src%d' % ((64 * k) % 8)]
        I.draw_augmentation(src.shape[1], src.shape[0], 0.2, 0.1, 1.5, 1.5)
    got.append(multiscale_width(0, first_batch=True))
    assert got == gold['multiscale/widths'].tolist()
    assert got[0] == 416 and got[-1] == 608 and len(set(got)) > 3


def test_batcher_applies_the_schedule_at_batch_starts(emul, gold, cfg5, monkeypatch):
    route_image_calls_to_emulation(monkeypatch, emul)
    from fewshot_detection_b200.dataset import DetectionBatcher
    cfg5.multiscale = 1
    lines = [(gold['src%d' % (i % 8)], gold['lab%d' % (i % 8)]) for i in range(66)]
    ds = DetectionBatcher(lines, shape=(64, 64), shuffle=False, train=True, seen=4000 * 64, batch_size=2, num_workers=1)
    random.seed(5)
    data, _ = ds.batch([0, 1])
    assert data.shape[-1] in (416, 448, 480, 512) and data.shape[-2] == data.shape[-1]
    with pytest.raises(ValueError):
        ds.batch([63, 64])                        # index 64 redraws the size: a batch must not straddle it


def test_meta_batcher_equals_reference_MetaDataset(emul, gold, cfg5, monkeypatch):
    route_image_calls_to_emulation(monkeypatch, emul)
    from fewshot_detection_b200.dataset import MetaBatcher
    ncls = 3
    cfg5.base_classes, cfg5.base_ids, cfg5.metain_type = cfg5.voc_classes[:ncls], list(range(ncls)), 2
    cfg5.meta_width = cfg5.meta_height = cfg5.mask_width = cfg5.mask_height = 48
    pool = gold['meta/pool']
    metalines = [[(gold['src%d' % i], gold['meta_lab/%d/%d' % (c, i)]) for i in pool[c] if i >= 0] for c in range(ncls)]
    inds = [tuple(int(v) for v in r) for r in gold['meta/inds']]
    mb = MetaBatcher(metalines, inds, train=True, with_ids=True)
    random.seed(43)
    imgs, masks, ids = [], [], []
    for b in range(len(inds) // ncls):            # one batch = one support image per class, as MetaDataset.batch_size
        metax, mask, clsids = mb.batch(range(b * ncls, (b + 1) * ncls))
        imgs.append(metax.numpy())
        masks.append(mask.numpy())
        ids += clsids
    assert ids == [c for c, _ in inds]
    assert np.array_equal(np.concatenate(masks), gold['meta/mask'])
    assert np.array_equal(np.concatenate(imgs), gold['meta/img'])
    assert gold['meta/mask'].sum() > 0


def test_path_entries_equal_in_memory_entries(emul, gold, cfg5, monkeypatch, tmp_path):
    """Image / label FILES (PNG decoded on the host with PIL, label path derived like listDataset.get_labpath)."""
    pytest.importorskip('PIL')
    from PIL import Image
    route_image_calls_to_emulation(monkeypatch, emul)
    from fewshot_detection_b200.dataset import DetectionBatcher, get_labpath, get_meta_labpath
    (tmp_path / 'JPEGImages').mkdir()
    (tmp_path / 'labels').mkdir()
    paths = []
    for i in range(4):
        p = tmp_path / 'JPEGImages' / ('%06d.png' % i)
        Image.fromarray(gold['src%d' % i], 'RGB').save(str(p))
        (tmp_path / 'labels' / ('%06d.txt' % i)).write_text(
            ''.join('%d %.6f %.6f %.6f %.6f\n' % tuple(r) for r in gold['lab%d' % i]))
        paths.append(str(p) + '\n')
    assert get_labpath(paths[0].rstrip()) == str(tmp_path / 'labels' / '000000.txt')
    assert get_meta_labpath('/d/JPEGImages/1.jpg', 'cat') == '/d/labels_1c/cat/1.txt'
    random.seed(31)
    a = DetectionBatcher(paths, shape=(64, 64), shuffle=False, train=True, batch_size=4, num_workers=1).batch(range(4))
    assert np.array_equal(a[0].numpy(), gold['list_train/img'][:4])
    assert np.array_equal(a[1].numpy(), gold['list_train/label'][:4])
