"""Pin the detection decode / NMS / ensembling oracle (oracle/utils.py) against the fixture minted from the
reference's own utils.py (tests/golden/make_golden_detect.py).  CPU only, bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import utils as OU

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = ['v2_g13', 'v2_g19', 'v2_g10_obj', 'v1_valid', 'v1_detect', 'v2_empty']


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(G, 'detect.npz'), allow_pickle=False)


def flatten(all_boxes):
    rows = np.array([len(b) for b in all_boxes], dtype=np.int64)
    lens = np.array([len(box) for boxes in all_boxes for box in boxes], dtype=np.int64)
    vals = np.array([float(v) for boxes in all_boxes for box in boxes for v in box], dtype=np.float64)
    return rows, lens, vals


def oracle_boxes(gold, tag):
    cs, thr, nC, nA, only_obj, val = gold[tag + '/params']
    o = torch.from_numpy(gold[tag + '/output'])
    if tag.startswith('v2'):
        return OU.get_region_boxes_v2(o, int(cs), float(thr), int(nC), gold['anchors_voc'].tolist(), int(nA), int(only_obj),
                                      bool(val))
    return OU.get_region_boxes(o, float(thr), int(nC), gold['anchors_tiny'].tolist(), int(nA), int(only_obj), bool(val))


@pytest.mark.parametrize('tag', CASES)
def test_region_boxes_and_nms_bit_exact(gold, tag):
    boxes = oracle_boxes(gold, tag)
    rows, lens, vals = flatten(boxes)
    assert np.array_equal(rows, gold[tag + '/rows'])
    assert np.array_equal(lens, gold[tag + '/len'])
    assert np.array_equal(vals, gold[tag + '/vals'])          # float64 values, exact
    kept = [OU.nms([list(b) for b in row], float(gold[tag + '/nms_thresh'])) for row in boxes]
    krows, klens, kvals = flatten(kept)
    assert np.array_equal(krows, gold[tag + '/nms_rows'])
    assert np.array_equal(klens, gold[tag + '/nms_len'])
    assert np.array_equal(kvals, gold[tag + '/nms_vals'])
    if tag != 'v2_empty':
        assert rows.sum() > krows.sum() > 0                    # NMS really suppressed something


def test_nms_edge_cases():
    assert OU.nms([], 0.45) == []
    one = [[0.5, 0.5, 0.2, 0.2, 0.9, 1.0, 0]]
    assert OU.nms([list(one[0])], 0.45) == one
    # identical boxes: the first in list order survives (stable sort on equal keys)
    a = [0.5, 0.5, 0.2, 0.2, 0.7, 1.0, 0]
    b = [0.5, 0.5, 0.2, 0.2, 0.7, 0.5, 1]
    kept = OU.nms([list(a), list(b)], 0.45)
    assert kept == [a]
    # disjoint boxes both survive, highest confidence first
    c = [0.1, 0.1, 0.1, 0.1, 0.3, 1.0, 0]
    d = [0.8, 0.8, 0.1, 0.1, 0.6, 1.0, 0]
    assert OU.nms([list(c), list(d)], 0.45) == [d, c]


def test_ensemble_reweights_bit_exact(gold):
    n_cls = int(gold['ens/n_cls'])
    batches = [(gold['ens/dw%d' % k], gold['ens/ids%d' % k].tolist()) for k in range(3)]
    got = OU.ensemble_reweights(batches, n_cls).numpy()
    assert np.array_equal(got.view(np.uint32), gold['ens/result'].view(np.uint32))


def test_detection_lines_format():
    box = [0.5, 0.5, 0.2, 0.4, 0.8, 0.5, 0]
    lines = OU.detection_lines([box], '000012', 500, 375)
    assert lines == ['000012 0.400000 200.000000 112.500000 300.000000 262.500000\n']
