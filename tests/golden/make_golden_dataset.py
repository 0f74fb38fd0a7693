#!/usr/bin/env python
"""Mint tests/golden/dataset.npz from the REFERENCE's own dataset.py (listDataset / MetaDataset), run UNMODIFIED on
a throw-away VOC-shaped directory of synthetic PNG images + label files - build container only:

    python tests/golden/make_golden_dataset.py

Stored: the source pixels and label arrays (so that the tests can feed the same data in memory), the Python `random`
/ numpy seeds, and what the reference returned: per-sample tensors of listDataset.__getitem__ (ToTensor'd image,
float64 label), the multi-scale shapes it chose for several `seen` regimes, and MetaDataset.__getitem__'s
(image, mask) pairs together with its `inds`.
"""
import io
import os
import random
import sys
import tempfile
from contextlib import redirect_stdout

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, '_shims'))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, HERE)
with redirect_stdout(io.StringIO()):
    import dataset as RD        # the reference's dataset.py
from cfg import cfg as RC
from PIL import Image
from torchvision import transforms
from make_golden_augment import synth_image


def main():
    out = {}
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, 'JPEGImages'))
    os.makedirs(os.path.join(tmp, 'labels'))
    classes = RC.voc_classes
    RC.data, RC.multiscale, RC.metayolo, RC.yolo_joint = 'voc', 0, True, False
    RC.classes = classes
    RC.base_classes = classes[:5]
    RC.base_ids = list(range(5))
    RC.novel_ids = [5, 6]
    rs = np.random.RandomState(21)
    paths = []
    for i in range(8):
        h, w = int(rs.randint(50, 90)), int(rs.randint(60, 110))
        a = synth_image(h, w, 300 + i, smooth=(i % 2 == 0))
        p = os.path.join(tmp, 'JPEGImages', '%06d.png' % i)
        Image.fromarray(a, 'RGB').save(p)
        rows = []
        for _ in range(int(rs.randint(1, 6))):
            bw, bh = rs.uniform(0.1, 0.6, 2)
            rows.append([rs.randint(0, 7), rs.uniform(bw / 2, 1 - bw / 2), rs.uniform(bh / 2, 1 - bh / 2), bw, bh])
        with open(os.path.join(tmp, 'labels', '%06d.txt' % i), 'w') as f:
            for r in rows:
                f.write('%d %.6f %.6f %.6f %.6f\n' % tuple(r))
        out['src%d' % i] = a
        out['lab%d' % i] = np.loadtxt(os.path.join(tmp, 'labels', '%06d.txt' % i)).reshape(-1, 5)
        paths.append(p + '\n')

    # ---- listDataset, training mode, fixed 64x64
    for mode, train in (('train', True), ('test', False)):
        random.seed(31)
        with redirect_stdout(io.StringIO()):
            ds = RD.listDataset(list(paths), shape=(64, 64), shuffle=False, transform=transforms.ToTensor(), train=train,
                                seen=0, batch_size=4, num_workers=1)
        imgs, labs = [], []
        for i in range(8):
            img, lab = ds[i]
            imgs.append(img.numpy())
            labs.append(lab.numpy())
        out['list_%s/img' % mode] = np.stack(imgs)
        out['list_%s/label' % mode] = np.stack(labs)
        out['list_%s/seen_after' % mode] = np.int64(ds.seen)
    # ---- multi-scale schedule: which shape does __getitem__(index = 0, 64, ...) pick in each `seen` regime
    RC.multiscale = 1
    shapes = []
    seens = [0, 4000 * 64, 2 * 4000 * 64 + 5, 3 * 4000 * 64, 4 * 4000 * 64, 9 * 4000 * 64]
    random.seed(32)
    with redirect_stdout(io.StringIO()):
        ds = RD.listDataset(list(paths) * 80, shape=(416, 416), shuffle=False, transform=transforms.ToTensor(), train=True,
                            seen=0, batch_size=64, num_workers=1)
    for k, s in enumerate(seens):
        ds.seen = s
        ds[64 * k]
        shapes.append(ds.shape[0])
    ds.first_batch = True
    ds[64]
    shapes.append(ds.shape[0])
    out['multiscale/seens'] = np.array(seens, dtype=np.int64)
    out['multiscale/widths'] = np.array(shapes, dtype=np.int64)
    RC.multiscale = 0

    # ---- MetaDataset (metain_type 2: image + mask), training mode
    RC.num_gpus, RC.batch_size, RC.randmeta, RC.metain_type = 1, 64, False, 2
    RC.meta_width = RC.meta_height = RC.mask_width = RC.mask_height = 48
    ncls = 3
    RC.base_classes = classes[:ncls]
    RC.base_ids = list(range(ncls))
    metadict = os.path.join(tmp, 'metadict.txt')
    with open(metadict, 'w') as f:
        for c in range(ncls):
            lst = os.path.join(tmp, 'meta_%s.txt' % classes[c])
            os.makedirs(os.path.join(tmp, 'labels_1c', classes[c]), exist_ok=True)
            with open(lst, 'w') as g:
                for i in range(8):
                    if (i + c) % 3 == 0:
                        continue
                    g.write(os.path.join(tmp, 'JPEGImages', '%06d.png' % i) + '\n')
                    rows = []
                    for _ in range(int(rs.randint(0, 3))):
                        bw, bh = rs.uniform(0.004, 0.5, 2)      # some boxes are too small for a 48-pixel mask
                        rows.append([c, rs.uniform(bw / 2, 1 - bw / 2), rs.uniform(bh / 2, 1 - bh / 2), bw, bh])
                    with open(os.path.join(tmp, 'labels_1c', classes[c], '%06d.txt' % i), 'w') as lf:
                        for r in rows:
                            lf.write('%d %.6f %.6f %.6f %.6f\n' % tuple(r))
                    out['meta_lab/%d/%d' % (c, i)] = np.array(rows, dtype=np.float64).reshape(-1, 5)
            f.write('%s %s\n' % (classes[c], lst))
    np.random.seed(41)
    random.seed(42)
    with redirect_stdout(io.StringIO()):
        ms = RD.MetaDataset(metadict, train=True, num_workers=0)
    inds = list(ms.inds)[:4 * ncls]
    out['meta/inds'] = np.array(inds, dtype=np.int64)
    out['meta/pool'] = np.array([[int(os.path.basename(l.strip())[:6]) for l in ms.metalines[c]] + [-1] * (8 - len(ms.metalines[c]))
                                 for c in range(ncls)], dtype=np.int64)
    imgs, masks = [], []
    random.seed(43)
    for k in range(len(inds)):
        img, mask = ms[k]
        imgs.append(img.numpy())
        masks.append(mask.numpy())
    out['meta/img'] = np.stack(imgs)
    out['meta/mask'] = np.stack(masks)
    np.savez_compressed(os.path.join(HERE, 'dataset.npz'), **out)
    print('wrote dataset.npz', os.path.getsize(os.path.join(HERE, 'dataset.npz')), 'multiscale widths', shapes,
          'meta inds', inds[:6])


if __name__ == '__main__':
    main()
