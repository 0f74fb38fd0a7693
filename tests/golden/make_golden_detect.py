#!/usr/bin/env python
"""Mint tests/golden/detect_*.npz (detection decode + NMS + reweight ensembling) from the REFERENCE ITSELF.

Run in the build container only (needs /root/reference; never at test time):

    python tests/golden/make_golden_detect.py

The reference's utils.py (get_region_boxes :112-193, get_region_boxes_v2 :195-290, nms :85-104) is exec'd from where
it lies with the mechanical Py3 / torch-2 substitutions listed in PATCHES (integer division, `.cuda()` removal).
torch-0.3.1 semantics are re-created where they matter: `convert2cpu*` (utils.py:106-110) return a FloatTensor /
LongTensor that the triple loop indexes element by element, which in 0.3.1 yields Python floats / ints — here they
return nested Python lists with exactly those values.  The running-mean ensembling loop of
valid_ensemble.py:86-100 sits inside the `valid()` driver (dataset + file IO), so it is replayed with the same
expression on tensors (`_ensemble`), nothing else.

Ragged box lists are stored flat: `<case>/len` (int64, entries per box), `<case>/rows` (boxes per row), `<case>/vals`.
"""
import io
import os
import re
import sys
import types
from contextlib import redirect_stdout

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, '_shims'))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

from fewshot_detection_b200 import netcfg  # noqa: E402

PATCHES = [
    (r'\.cuda\(\)', ''),
    (r'len\(anchors\)/num_anchors', 'len(anchors)//num_anchors'),
    (r'\bxrange\b', 'range'),
]


def load_ref_utils():
    if 'imghdr' not in sys.modules:
        try:
            import imghdr  # noqa: F401
        except Exception:
            sys.modules['imghdr'] = types.ModuleType('imghdr')
    src = open(os.path.join(REF, 'utils.py')).read()
    for pat, rep in PATCHES:
        src = re.sub(pat, rep, src)
    mod = types.ModuleType('ref_utils')
    mod.__file__ = os.path.join(REF, 'utils.py')
    with redirect_stdout(io.StringIO()):
        exec(compile(src, mod.__file__, 'exec'), mod.__dict__)
    # torch 0.3.1: FloatTensor[int] -> Python float, LongTensor[int] -> Python int
    mod.convert2cpu = lambda t: t.float().tolist()
    mod.convert2cpu_long = lambda t: t.long().tolist()
    return mod


def flatten(all_boxes):
    rows = np.array([len(b) for b in all_boxes], dtype=np.int64)
    lens = np.array([len(box) for boxes in all_boxes for box in boxes], dtype=np.int64)
    vals = np.array([float(v) for boxes in all_boxes for box in boxes for v in box], dtype=np.float64)
    return rows, lens, vals


def plant(o, nA, nC, G, rs, n_per_row=3):
    """Plant clusters of strong, overlapping detections so NMS has work that matters."""
    N = o.shape[0]
    ov = o.view(N, nA, 5 + nC, G, G)
    for n in range(N):
        for _ in range(n_per_row):
            cy, cx = rs.randint(1, G - 1, 2)
            for dy, dx in ((0, 0), (0, 1), (1, 0)):
                for a in rs.choice(nA, 2, replace=False):
                    ov[n, a, 4, cy + dy, cx + dx] = float(rs.uniform(1.5, 4.0))
                    ov[n, a, 2:4, cy + dy, cx + dx] = torch.from_numpy(rs.uniform(-0.3, 0.6, 2).astype(np.float32))
                    ov[n, a, 5:, cy + dy, cx + dx] += float(rs.uniform(1.0, 3.0))
    return o


def _ensemble(batches, n_cls):
    enews = [0.0] * n_cls
    cnt = [0.0] * n_cls
    for dw, clsids in batches:
        for ci, c in enumerate(clsids):
            enews[c] = enews[c] * cnt[c] / (cnt[c] + 1) + dw[ci] / (cnt[c] + 1)   # valid_ensemble.py:97
            cnt[c] += 1
    return torch.stack(enews)


def main():
    U = load_ref_utils()
    voc = [float(a) for a in netcfg.VOC_ANCHORS.split(',')]
    tiny = [float(a) for a in netcfg.TINY_VOC_ANCHORS.split(',')]
    out = {}

    def run(tag, fn, o, args, nms_thresh):
        with redirect_stdout(io.StringIO()):
            boxes = fn(o, *args)
        rows, lens, vals = flatten(boxes)
        out[tag + '/output'] = o.numpy()
        out[tag + '/rows'], out[tag + '/len'], out[tag + '/vals'] = rows, lens, vals
        kept = [U.nms([list(b) for b in row], nms_thresh) for row in boxes]
        krows, klens, kvals = flatten(kept)
        out[tag + '/nms_rows'], out[tag + '/nms_len'], out[tag + '/nms_vals'] = krows, klens, kvals
        out[tag + '/nms_thresh'] = np.float64(nms_thresh)
        print(tag, 'candidates', rows.tolist(), 'kept', krows.tolist())

    # (1) the evaluation call of valid_ensemble.py:140-145: n_cls rows per image, nC = 1, thresh 0.005, nms 0.45
    rs = np.random.RandomState(61)
    g = torch.Generator().manual_seed(61)
    bs, cs, G = 2, 3, 13
    o = plant(torch.randn(bs * cs, 30, G, G, generator=g) * 1.2 - 0.5, 5, 1, G, rs)
    o.view(bs * cs, 5, 6, G, G)[:, :, 4] -= 3.0     # most cells below the objectness that matters
    run('v2_g13', U.get_region_boxes_v2, o, (cs, 0.005, 1, voc, 5, 0, 1), 0.45)
    out['v2_g13/params'] = np.array([cs, 0.005, 1, 5, 0, 1], dtype=np.float64)
    # (2) 608 input (G = 19), 4 classes, higher threshold
    bs, cs, G = 1, 4, 19
    o = plant(torch.randn(bs * cs, 30, G, G, generator=g) * 1.0, 5, 1, G, rs, n_per_row=5)
    o.view(bs * cs, 5, 6, G, G)[:, :, 4] -= 2.0
    run('v2_g19', U.get_region_boxes_v2, o, (cs, 0.1, 1, voc, 5, 0, 1), 0.45)
    out['v2_g19/params'] = np.array([cs, 0.1, 1, 5, 0, 1], dtype=np.float64)
    # (3) objectness-only flavour of the same function (only_objectness=1, validation=False)
    bs, cs, G = 2, 2, 10
    o = plant(torch.randn(bs * cs, 30, G, G, generator=g) * 1.0, 5, 1, G, rs)
    run('v2_g10_obj', U.get_region_boxes_v2, o, (cs, 0.6, 1, voc, 5, 1, False), 0.4)
    out['v2_g10_obj/params'] = np.array([cs, 0.6, 1, 5, 1, 0], dtype=np.float64)
    # (4) plain detector (tiny-yolo-voc: nC = 20): valid.py-style call with the extra (conf, id) pairs, and the
    #     do_detect-style call (utils.py:410-458: conf_thresh 0.5, only_objectness default)
    G = 13
    o = plant(torch.randn(2, 125, G, G, generator=g) * 1.5, 5, 20, G, rs)
    o.view(2, 5, 25, G, G)[:, :, 4] -= 2.5
    run('v1_valid', U.get_region_boxes, o, (0.005, 20, tiny, 5, 0, 1), 0.45)
    out['v1_valid/params'] = np.array([1, 0.005, 20, 5, 0, 1], dtype=np.float64)
    run('v1_detect', U.get_region_boxes, o, (0.5, 20, tiny, 5), 0.4)
    out['v1_detect/params'] = np.array([1, 0.5, 20, 5, 1, 0], dtype=np.float64)
    # (5) empty result (nothing above threshold) and a single-image / 3-D input
    o = torch.full((2, 30, 5, 5), -9.0)
    run('v2_empty', U.get_region_boxes_v2, o, (2, 0.5, 1, voc, 5, 0, 1), 0.45)
    out['v2_empty/params'] = np.array([2, 0.5, 1, 5, 0, 1], dtype=np.float64)
    out['anchors_voc'] = np.array(voc)
    out['anchors_tiny'] = np.array(tiny)

    # (6) reweighting-vector ensembling (valid_ensemble.py:86-100)
    g = torch.Generator().manual_seed(62)
    n_cls, C = 5, 64
    ids = [[0, 1, 2, 3, 4, 0, 0, 1], [2, 2, 4, 4, 4, 3, 1, 0], [0, 3, 3]]
    dws = [torch.randn(len(i), C, generator=g) for i in ids]
    ens = _ensemble(list(zip(dws, ids)), n_cls)
    out['ens/n_cls'] = np.int64(n_cls)
    for k, (d, i) in enumerate(zip(dws, ids)):
        out['ens/dw%d' % k] = d.numpy()
        out['ens/ids%d' % k] = np.array(i, dtype=np.int64)
    out['ens/result'] = ens.numpy()

    np.savez_compressed(os.path.join(HERE, 'detect.npz'), **out)
    print('wrote', os.path.join(HERE, 'detect.npz'))


if __name__ == '__main__':
    main()
