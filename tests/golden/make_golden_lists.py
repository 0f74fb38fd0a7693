#!/usr/bin/env python
"""Mint tests/golden/lists.json from the REFERENCE's own dataset.py list builders (loadlines / build_dataset /
load_metadict / build_fewset / MetaDataset.__init__), run UNMODIFIED (the stray pdb.set_trace() of load_metadict is
neutralised) on a throw-away VOC-shaped directory of label and list files - build container only:

    python tests/golden/make_golden_lists.py

Stored: the label / list / dict file contents (paths relative to the directory, spelled <ROOT>), the seeds, and
what the reference returned.  tests/test_lists.py recreates the directory and compares fewshot_detection_b200.lists.
"""
import io
import json
import os
import random
import sys
import tempfile
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, '_shims'))
sys.path.insert(0, '/root/reference')
if not hasattr(np, 'int'):
    np.int = int                      # numpy >= 1.24 removed the alias the reference uses (dataset.py:48,107,139,280)
with redirect_stdout(io.StringIO()):
    import dataset as RD              # the reference's dataset.py
from cfg import cfg as RC

RD.pdb.set_trace = lambda *a, **k: None


def main():
    root = tempfile.mkdtemp()
    os.makedirs(os.path.join(root, 'JPEGImages'))
    os.makedirs(os.path.join(root, 'labels'))
    os.makedirs(os.path.join(root, 'lists'))
    classes = RC.voc_classes
    novel = ['bird', 'bus', 'cow', 'motorbike', 'sofa']
    rs = np.random.RandomState(7)
    files = {}
    imgs = []
    for i in range(400):
        p = os.path.join(root, 'JPEGImages', '%06d.jpg' % i)
        imgs.append(p)
        k = int(rs.choice([0, 1, 1, 1, 1, 2, 2, 3, 4, 6]))
        rows = []
        for _ in range(k):
            c = int(rs.randint(0, 20))
            w, h = rs.uniform(0.1, 0.5, 2)
            rows.append('%d %.4f %.4f %.4f %.4f' % (c, rs.uniform(w / 2, 1 - w / 2), rs.uniform(h / 2, 1 - h / 2), w, h))
        files['labels/%06d.txt' % i] = '\n'.join(rows) + ('\n' if rows else '')
    files['lists/train.txt'] = ''.join(p + '\n' for p in imgs)
    # per-class lists: every image that shows the class; k-shot lists: the first 2 of them
    full, shot = [], []
    for ci, name in enumerate(classes):
        has = [p for i, p in enumerate(imgs) if any(r.split()[0] == str(ci) for r in files['labels/%06d.txt' % i].splitlines())]
        if not has:
            has = [imgs[ci]]
        files['lists/full_%s.txt' % name] = ''.join(p + '\n' for p in has)
        files['lists/2shot_%s.txt' % name] = ''.join(p + '\n' for p in has[:2])
        full.append('%s %s' % (name, os.path.join(root, 'lists', 'full_%s.txt' % name)))
        shot.append('%s %s' % (name, os.path.join(root, 'lists', '2shot_%s.txt' % name)))
    files['lists/dict_full.txt'] = '\n'.join(full) + '\n'
    files['lists/dict_2shot.txt'] = '\n'.join(shot) + '\n'
    for rel, text in files.items():
        with open(os.path.join(root, rel), 'w') as f:
            f.write(text)

    P = lambda rel: os.path.join(root, rel)
    out = {'classes': classes, 'novel': novel, 'files': {k: v.replace(root, '<ROOT>') for k, v in files.items()}, 'cases': {}}

    def configure(tuning, repeat=1, shot=2):
        RC.data, RC.classes, RC.tuning, RC.repeat, RC.shot = 'voc', classes, tuning, repeat, shot
        RC.novel_classes = novel
        RC.base_classes = list(classes) if tuning else [c for c in classes if c not in novel]
        RC.base_ids = [classes.index(c) for c in RC.base_classes]
        RC.novel_ids = [classes.index(c) for c in novel]
        RC.num_gpus, RC.batch_size, RC.randmeta = 1, 64, False
        RC.meta_width = RC.meta_height = RC.mask_width = RC.mask_height = 64

    rel = lambda lines: [l.replace(root, '<ROOT>') for l in lines]
    with redirect_stdout(io.StringIO()):
        configure(False)
        out['cases']['base_plain'] = rel(RD.loadlines(P('lists/train.txt')))
        out['cases']['base_dict'] = rel(RD.loadlines(P('lists/dict_full.txt')))
        out['cases']['base_dict_nocheck'] = rel(RD.loadlines(P('lists/dict_full.txt'), checkvalid=False))
        out['cases']['build_base'] = rel(RD.build_dataset({'train': P('lists/train.txt'), 'meta': P('lists/dict_full.txt')}))
        configure(True, repeat=1)
        out['cases']['tune_repeat1'] = rel(RD.build_dataset({'train': P('lists/train.txt'), 'meta': P('lists/dict_2shot.txt')}))
        configure(True, repeat=3)
        out['cases']['tune_repeat3'] = rel(RD.build_dataset({'train': P('lists/train.txt'), 'meta': P('lists/dict_2shot.txt'),
                                                             'dynamic': '0'}))
        configure(True, repeat=2)
        ml, mc = RD.load_metadict(P('lists/dict_2shot.txt'), 2)
        out['cases']['metadict_list_sorted'] = sorted(rel(ml))
        out['cases']['metadict_counts'] = mc
        random.seed(11)
        out['cases']['tune_dynamic_seed11'] = rel(RD.build_dataset({'train': P('lists/train.txt'), 'meta': P('lists/dict_2shot.txt'),
                                                                    'dynamic': '1'}))
        configure(False)
        np.random.seed(3)
        md = RD.MetaDataset(P('lists/dict_full.txt'), train=True)
        out['cases']['support_train_seed3'] = {'inds': [list(map(int, t)) for t in md.inds[:600]], 'n': len(md.inds),
                                                'meta_cnts': md.meta_cnts, 'batch_size': md.batch_size}
    with open(os.path.join(HERE, 'lists.json'), 'w') as f:
        json.dump(out, f)
    print('wrote lists.json:', {k: (len(v) if hasattr(v, '__len__') else v) for k, v in out['cases'].items()})


if __name__ == '__main__':
    main()
