#!/usr/bin/env python
"""Mint tests/golden/augment.npz from the REFERENCE's own image.py (run UNMODIFIED: it imports cleanly under
Python 3.12 / Pillow 12.2) - build container only (needs /root/reference and Pillow; never at test time):

    python tests/golden/make_golden_augment.py

Per case: a synthetic uint8 RGB source image, the Python `random` seed, the arguments, and what
`image.data_augmentation` returned: the augmented uint8 image (np.asarray of the PIL result) and
(flip, dx, dy, sx, sy).  Label cases: a label file written to a temp dir, run through
`image.fill_truth_detection`, `fill_truth_detection_meta` and `load_label` with the transform parameters of the
matching image case.  Pillow's default resize filter is what the reference gets (`cropped.resize(shape)`,
image.py:77): BICUBIC under the container's Pillow 12.2 - recorded in `pillow_version`.
Also stored: PIL NEAREST resizes of the same crops (the default filter of the Pillow of the reference's time), made
with explicit PIL calls in this script (not reference code) and labelled `nearest/...`.
"""
import io
import os
import random
import sys
import tempfile
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, '_shims'))
sys.path.insert(0, REF)

with redirect_stdout(io.StringIO()):
    import image as RI          # the reference's image.py
import PIL
from PIL import Image


def synth_image(h, w, seed, smooth=True):
    rs = np.random.RandomState(seed)
    if smooth:   # smooth gradients + blobs + a little noise: compresses well, still exercises every filter tap
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
        chans = []
        for c in range(3):
            f = rs.uniform(0.01, 0.08, 4)
            v = 127 + 60 * np.sin(f[0] * xx + f[1] * yy + c) + 50 * np.cos(f[2] * xx - f[3] * yy) + rs.randn(h, w) * 6
            chans.append(v)
        a = np.stack(chans, -1)
        for _ in range(6):     # saturated rectangles (hard edges -> bicubic overshoot -> clip8)
            y0, x0 = rs.randint(0, h - 4), rs.randint(0, w - 4)
            a[y0:y0 + rs.randint(3, h // 3), x0:x0 + rs.randint(3, w // 3)] = rs.choice([0, 255], 3)
        return np.clip(a, 0, 255).astype(np.uint8)
    return rs.randint(0, 256, (h, w, 3)).astype(np.uint8)


def main():
    out = {'pillow_version': np.array(PIL.__version__)}
    cases = [  # tag, h, w, shape (W, H), seed, flag, smooth
        ('a', 75, 100, (64, 64), 1, True, False),
        ('b', 90, 60, (96, 96), 2, True, False),
        ('c', 60, 80, (64, 32), 3, True, True),
        ('d', 50, 50, (50, 64), 4, False, False),       # flag=False: plain resize, one pass skipped (width unchanged)
        ('e', 333, 500, (416, 416), 5, True, True),     # VOC-sized
        ('f', 120, 90, (64, 64), 6, True, True),
        ('g', 40, 40, (160, 160), 7, True, False),      # up-scaling x4
    ]
    for tag, h, w, shape, seed, flag, smooth in cases:
        a = synth_image(h, w, 100 + seed, smooth)
        random.seed(seed)
        img, flip, dx, dy, sx, sy = RI.data_augmentation(Image.fromarray(a, 'RGB'), shape, 0.2, 0.1, 1.5, 1.5, flag=flag)
        out[tag + '/src'] = a
        out[tag + '/args'] = np.array([shape[0], shape[1], seed, int(flag)], dtype=np.int64)
        out[tag + '/img'] = np.asarray(img)
        out[tag + '/params'] = np.array([flip, dx, dy, sx, sy], dtype=np.float64)
        # explicit-PIL NEAREST variant of the same augmentation (same draws)
        random.seed(seed)
        if flag:
            dw, dh = int(w * 0.2), int(h * 0.2)
            pleft, pright = random.randint(-dw, dw), random.randint(-dw, dw)
            ptop, pbot = random.randint(-dh, dh), random.randint(-dh, dh)
            fl = random.randint(1, 10000) % 2
            sw, sh = w - pleft - pright, h - ptop - pbot
            im = Image.fromarray(a, 'RGB').crop((pleft, ptop, pleft + sw - 1, ptop + sh - 1)).resize(shape, Image.NEAREST)
            if fl:
                im = im.transpose(Image.FLIP_LEFT_RIGHT)
            im = RI.random_distort_image(im, 0.1, 1.5, 1.5)
        else:
            im = Image.fromarray(a, 'RGB').resize(shape, Image.NEAREST)
        out['nearest/' + tag] = np.asarray(im)
        print(tag, a.shape, '->', np.asarray(img).shape, 'flip', flip, 'sx', sx, 'sy', sy)

    # ---- label transforms
    RI.cfg.base_classes = RI.cfg.voc_classes[:15]
    RI.cfg.base_ids = list(range(15))
    RI.cfg.yolo_joint = False
    rs = np.random.RandomState(9)
    tmp = tempfile.mkdtemp()
    for tag, nbox in (('l1', 7), ('l2', 60), ('l3', 0), ('l4', 1)):
        rows = []
        for _ in range(nbox):
            wv, hv = rs.uniform(0.02, 0.6, 2)
            rows.append([rs.randint(0, 20), rs.uniform(wv / 2, 1 - wv / 2), rs.uniform(hv / 2, 1 - hv / 2), wv, hv])
        lab = os.path.join(tmp, '0000%s.txt' % tag)
        with open(lab, 'w') as f:
            for r in rows:
                f.write('%d %.6f %.6f %.6f %.6f\n' % tuple(r))
        boxes = np.loadtxt(lab).reshape(-1, 5) if nbox else np.zeros((0, 5))
        flip, dx, dy, sx, sy = (1, -0.07, 0.031, 0.88, 1.12) if tag != 'l4' else (0, 0.11, -0.05, 1.2, 0.9)
        out[tag + '/boxes'] = boxes
        out[tag + '/transform'] = np.array([flip, dx, dy, 1. / sx, 1. / sy])
        out[tag + '/fill'] = RI.fill_truth_detection(lab, 416, 416, flip, dx, dy, 1. / sx, 1. / sy)
        if tag != 'l2':   # 60 boxes of random classes can exceed... the reference calls pdb.set_trace() on overflow
            out[tag + '/fill_meta'] = RI.fill_truth_detection_meta(lab, 416, 416, flip, dx, dy, 1. / sx, 1. / sy)
        ll = RI.load_label(lab, 416, 416, flip, dx, dy, 1. / sx, 1. / sy) if nbox else []
        out[tag + '/load_label'] = np.array(ll, dtype=np.float64).reshape(-1, 4)
        print(tag, nbox, 'kept', int((out[tag + '/fill'].reshape(-1, 5)[:, 3] > 0).sum()))
    np.savez_compressed(os.path.join(HERE, 'augment.npz'), **out)
    print('wrote', os.path.join(HERE, 'augment.npz'), os.path.getsize(os.path.join(HERE, 'augment.npz')))


if __name__ == '__main__':
    main()
