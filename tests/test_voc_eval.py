"""fewshot_detection_b200.voc_eval against tests/golden/voc_eval.npz = the reference's scripts/voc_eval.py on the same
synthetic annotations / detection files (tests/golden/make_golden_voc_eval.py).  CPU only; exact (float64)."""
import os

import numpy as np
import pytest

from fewshot_detection_b200 import voc_eval as V

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
XML = '<annotation><filename>{name}.jpg</filename>{objs}</annotation>'
OBJ = ('<object><name>{cls}</name><pose>Unspecified</pose><truncated>0</truncated><difficult>{df}</difficult>'
       '<bndbox><xmin>{x1}</xmin><ymin>{y1}</ymin><xmax>{x2}</xmax><ymax>{y2}</ymax></bndbox></object>')


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(G, 'voc_eval.npz'), allow_pickle=False)


@pytest.fixture()
def devkit(gold, tmp_path):
    (tmp_path / 'Annotations').mkdir()
    names = [str(n) for n in gold['names']]
    objs = dict((n, '') for n in names)
    for n, c, df, x1, y1, x2, y2 in gold['gt']:
        objs[str(n)] += OBJ.format(cls=c, df=df, x1=x1, y1=y1, x2=x2, y2=y2)
    for n in names:
        (tmp_path / 'Annotations' / (n + '.xml')).write_text(XML.format(name=n, objs=objs[n]))
    (tmp_path / 'test.txt').write_text('\n'.join(names) + '\n')
    for c in gold['classes']:
        (tmp_path / ('det_%s.txt' % c)).write_text('\n'.join(str(l) for l in gold['det/' + str(c)]) + '\n')
    return tmp_path


@pytest.mark.parametrize('m07', [True, False])
def test_voc_eval_equals_reference(gold, devkit, m07):
    for c in gold['classes']:
        c = str(c)
        rec, prec, ap = V.voc_eval(str(devkit / 'det_{}.txt'), str(devkit / 'Annotations' / '{}.xml'), str(devkit / 'test.txt'),
                                   c, str(devkit / 'cache'), 0.5, m07)
        k = '%s/%d' % (c, int(m07))
        assert np.array_equal(rec, gold['rec/' + k])
        assert np.array_equal(prec, gold['prec/' + k])
        assert ap == float(gold['ap/' + k])
    assert (devkit / 'cache' / 'annots.pkl').exists()        # annotation cache written like the reference


def test_voc_ap_known_answers():
    rec = np.array([0.25, 0.25, 0.5, 0.75, 1.0])
    prec = np.array([1.0, 0.5, 2 / 3., 0.75, 0.8])
    assert abs(V.voc_ap(rec, prec) - 0.25 * (1.0 + 0.8 + 0.8 + 0.8)) < 1e-12
    assert abs(V.voc_ap(rec, prec, True) - (3 * 1.0 + 8 * 0.8) / 11.) < 1e-12
    assert V.voc_ap(np.array([]), np.array([]), True) == 0.0
    assert V.voc_ap(np.array([]), np.array([])) == 0.0


def test_matching_rules():
    gt = {'a': (np.array([[10, 10, 50, 50], [100, 100, 150, 150]]), np.array([False, True])), 'b': (np.zeros((0, 4)), np.zeros(0, bool))}
    ids = ['a', 'a', 'a', 'b', 'a']
    conf = [0.9, 0.8, 0.7, 0.6, 0.5]
    boxes = [[10, 10, 50, 50],        # TP
             [12, 12, 50, 50],        # duplicate of a claimed box -> FP
             [100, 100, 150, 150],    # matches a difficult box -> ignored
             [0, 0, 10, 10],          # image without ground truth -> FP
             [300, 300, 320, 320]]    # no overlap -> FP
    tp, fp = V.match_detections(ids, conf, boxes, gt)
    assert tp.tolist() == [1, 0, 0, 0, 0] and fp.tolist() == [0, 1, 0, 1, 1]


def test_mean_ap_groups(devkit, gold):
    r = V.mean_ap(str(devkit / 'det_{}.txt'), str(devkit / 'Annotations' / '{}.xml'), str(devkit / 'test.txt'),
                  [str(c) for c in gold['classes']], str(devkit / 'cache'), True, novel_classes=('cow',))
    assert abs(r['mean_novel'] - float(gold['ap/cow/1'])) < 1e-15
    assert abs(r['mean_base'] - (float(gold['ap/bird/1']) + float(gold['ap/bus/1'])) / 2) < 1e-15


def test_precision_recall_counts_on_reference_decode_fixture():
    """evaluate.count_matches (train_meta.test()'s loop body) on boxes decoded by the oracle from the reference's
    own fixture, against an independent set-based recount."""
    import torch
    from fewshot_detection_b200 import evaluate as E
    from oracle import utils as OU
    d = np.load(os.path.join(G, 'detect.npz'), allow_pickle=False)
    out = torch.from_numpy(d['v1_detect/output'])
    boxes = OU.get_region_boxes(out, 0.25, 20, d['anchors_tiny'].tolist(), 5)
    target = np.zeros((2, 250))
    kept0 = OU.nms([list(b) for b in boxes[0]], 0.4)
    # ground truths: two surviving detections of image 0 (one with the right class, one with a wrong class), one box
    # far from everything; image 1 has no ground truth
    target[0, 0:5] = [kept0[0][6], kept0[0][0], kept0[0][1], kept0[0][2], kept0[0][3]]
    target[0, 5:10] = [(kept0[1][6] + 1) % 20, kept0[1][0], kept0[1][1], kept0[1][2], kept0[1][3]]
    target[0, 10:15] = [3, 0.02, 0.02, 0.01, 0.01]
    total, proposals, correct = E.count_matches([[list(b) for b in r] for r in boxes], target, 0.25, 0.4, 0.5, nms_fn=OU.nms)
    assert total == 3
    assert proposals == sum(1 for r in boxes for b in OU.nms([list(x) for x in r], 0.4) if b[4] > 0.25)
    assert correct == 1
    p, r, f = E.precision_recall(total, proposals, correct)
    assert abs(r - 1 / (3 + 1e-5)) < 1e-12 and 0 < p < 1 and 0 < f < 1
    assert E.truths_length(target[1].reshape(-1, 5).tolist()) == 0
