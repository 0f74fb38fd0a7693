#!/usr/bin/env python
"""Mint tests/golden/voc_eval.npz from the REFERENCE's scripts/voc_eval.py (build container only).

The script is Python 2 (three `print` statements, cPickle, text-mode pickle files, np.bool, `from termcolor import
colored`, a module-global `args` that only its __main__ block defines): it is exec'd from where it lies with the
mechanical substitutions in PATCHES; no arithmetic is touched.  Inputs are synthetic VOC XML annotations and
detection files written to a temp dir; they are stored in the fixture so that the test can re-create them.
"""
import io
import os
import re
import sys
import tempfile
import types
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/scripts/voc_eval.py'
PATCHES = [
    (r'import cPickle', 'import pickle as cPickle'),
    (r"print 'Reading annotation for \{:d\}/\{:d\}'\.format\(\s*i \+ 1, len\(imagenames\)\)", 'pass'),
    (r"print 'Saving cached annotations to \{:s\}'\.format\(cachefile\)", 'pass'),
    (r"print 'VOC07 metric\? ' \+ \('Yes' if use_07_metric else 'No'\)", 'pass'),
    (r"open\(cachefile, 'w'\)", "open(cachefile, 'wb')"),
    (r"open\(cachefile, 'r'\)", "open(cachefile, 'rb')"),
    (r'np\.bool\b', 'bool'),
]
XML = '<annotation><filename>{name}.jpg</filename>{objs}</annotation>'
OBJ = ('<object><name>{cls}</name><pose>Unspecified</pose><truncated>{tr}</truncated><difficult>{df}</difficult>'
       '<bndbox><xmin>{x1}</xmin><ymin>{y1}</ymin><xmax>{x2}</xmax><ymax>{y2}</ymax></bndbox></object>')


def load_ref():
    sys.modules.setdefault('termcolor', types.SimpleNamespace(colored=lambda s, *a, **k: s))
    src = open(REF).read()
    for pat, rep in PATCHES:
        src, n = re.subn(pat, rep, src)
        assert n >= 1, pat
    mod = types.ModuleType('ref_voc_eval')
    mod.__file__ = REF
    mod.args = types.SimpleNamespace(single=False)
    exec(compile(src, REF, 'exec'), mod.__dict__)
    return mod


def main():
    R = load_ref()
    rs = np.random.RandomState(77)
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, 'Annotations'))
    classes = ['bird', 'bus', 'cow']
    names = ['%06d' % i for i in range(30)]
    gts = []
    for name in names:
        objs = ''
        for _ in range(int(rs.randint(0, 5))):
            x1, y1 = int(rs.randint(1, 300)), int(rs.randint(1, 200))
            w, h = int(rs.randint(20, 180)), int(rs.randint(20, 150))
            c = classes[int(rs.randint(0, 3))]
            df = int(rs.rand() < 0.2)
            objs += OBJ.format(cls=c, tr=int(rs.rand() < 0.3), df=df, x1=x1, y1=y1, x2=x1 + w, y2=y1 + h)
            gts.append((name, c, df, x1, y1, x1 + w, y1 + h))
        with open(os.path.join(tmp, 'Annotations', name + '.xml'), 'w') as f:
            f.write(XML.format(name=name, objs=objs))
    with open(os.path.join(tmp, 'test.txt'), 'w') as f:
        f.write('\n'.join(names) + '\n')
    out = {'names': np.array(names), 'classes': np.array(classes),
           'gt': np.array([(n, c, str(d), str(a), str(b), str(e), str(g)) for n, c, d, a, b, e, g in gts])}
    for c in classes:
        lines = []
        for n, gc, df, x1, y1, x2, y2 in gts:            # jittered copies of the ground truth (some duplicated) ...
            if gc != c:
                continue
            for _ in range(int(rs.randint(0, 3))):
                j = rs.uniform(-25, 25, 4)
                lines.append('%s %f %f %f %f %f' % (n, rs.uniform(0.05, 1.0), x1 + j[0], y1 + j[1], x2 + j[2], y2 + j[3]))
        for _ in range(25):                              # ... plus random false positives
            n = names[int(rs.randint(0, len(names)))]
            x1, y1 = rs.uniform(0, 300), rs.uniform(0, 200)
            lines.append('%s %f %f %f %f %f' % (n, rs.uniform(0.005, 0.6), x1, y1, x1 + rs.uniform(10, 150), y1 + rs.uniform(10, 150)))
        rs.shuffle(lines)
        with open(os.path.join(tmp, 'det_%s.txt' % c), 'w') as f:
            f.write('\n'.join(lines) + '\n')
        out['det/' + c] = np.array(lines)
        for m07 in (True, False):
            with redirect_stdout(io.StringIO()):
                rec, prec, ap = R.voc_eval(os.path.join(tmp, 'det_{}.txt'), os.path.join(tmp, 'Annotations', '{}.xml'),
                                           os.path.join(tmp, 'test.txt'), c, os.path.join(tmp, 'cache'), 0.5, m07)
            k = '%s/%d' % (c, int(m07))
            out['rec/' + k], out['prec/' + k], out['ap/' + k] = rec, prec, np.float64(ap)
            print(c, 'voc07' if m07 else 'area', 'ap = %.4f' % ap, 'dets', len(lines))
    np.savez_compressed(os.path.join(HERE, 'voc_eval.npz'), **out)
    print('wrote voc_eval.npz')


if __name__ == '__main__':
    main()
