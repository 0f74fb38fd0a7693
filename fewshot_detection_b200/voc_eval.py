"""PASCAL VOC detection AP from per-class result files.

Host-side companion of `valid.write_detections`: consumes the `imgid prob x1 y1 x2 y2` files and the VOC XML
annotations and returns (recall, precision, AP) per class - the numbers the reference reports through
scripts/voc_eval.py (voc_ap :63-94, voc_eval :96-243; itself the py-faster-rcnn evaluator).  Same function names,
arguments and return values; written from the PASCAL VOC devkit definition:

  * detections of a class are ranked by confidence (descending);
  * a detection is a true positive if its best-overlapping ground-truth box of that class in its image has
    IoU > ovthresh (pixel-inclusive areas: width = xmax - xmin + 1), is not `difficult`, and has not been claimed by a
    higher-ranked detection; a match to a `difficult` box is ignored; anything else is a false positive;
  * AP = area under the monotone precision envelope (VOC10+) or the 11-point average (VOC07).

This is small host work (text parsing + a few thousand IoUs per class): numpy on the CPU, like the reference.
"""
import os
import pickle
import xml.etree.ElementTree as ET

import numpy as np


def parse_rec(filename):
    """One VOC annotation file -> list of {'name', 'pose', 'truncated', 'difficult', 'bbox': [xmin, ymin, xmax, ymax]}."""
    objects = []
    for node in ET.parse(filename).getroot().iter('object'):
        box = node.find('bndbox')
        objects.append({
            'name': node.findtext('name'),
            'pose': node.findtext('pose'),
            'truncated': int(node.findtext('truncated')),
            'difficult': int(node.findtext('difficult')),
            'bbox': [int(box.findtext(k)) for k in ('xmin', 'ymin', 'xmax', 'ymax')],
        })
    return objects


def voc_ap(rec, prec, use_07_metric=False):
    """AP of a precision/recall curve (arrays ordered by decreasing confidence)."""
    rec = np.asarray(rec, dtype=np.float64)
    prec = np.asarray(prec, dtype=np.float64)
    if use_07_metric:
        ap = 0.
        for t in np.arange(0., 1.1, 0.1):
            above = rec >= t
            ap = ap + (np.max(prec[above]) if above.any() else 0) / 11.
        return ap
    r = np.concatenate(([0.], rec, [1.]))
    p = np.concatenate(([0.], prec, [0.]))
    p = np.maximum.accumulate(p[::-1])[::-1]          # precision envelope: best precision at any higher recall
    step = np.nonzero(r[1:] != r[:-1])[0]
    return np.sum((r[step + 1] - r[step]) * p[step + 1])


def load_annotations(annopath, imagenames, cachedir=None):
    """{imagename: parse_rec(annopath.format(imagename))}, cached as <cachedir>/annots.pkl like the reference."""
    cachefile = os.path.join(cachedir, 'annots.pkl') if cachedir else None
    if cachefile and os.path.isfile(cachefile):
        with open(cachefile, 'rb') as f:
            return pickle.load(f)
    recs = dict((name, parse_rec(annopath.format(name))) for name in imagenames)
    if cachefile:
        if not os.path.isdir(cachedir):
            os.mkdir(cachedir)
        with open(cachefile, 'wb') as f:
            pickle.dump(recs, f)
    return recs


def match_detections(image_ids, confidence, boxes, gt, ovthresh=0.5):
    """Rank the detections and mark true / false positives.

    gt: {image id: (bbox int/float [k, 4], difficult bool [k])}.  Returns (tp, fp) float arrays in rank order."""
    order = np.argsort(-np.asarray(confidence, dtype=np.float64))
    boxes = np.asarray(boxes, dtype=np.float64).reshape(-1, 4)
    claimed = dict((k, np.zeros(len(v[0]), dtype=bool)) for k, v in gt.items())
    tp = np.zeros(len(order))
    fp = np.zeros(len(order))
    for rank, d in enumerate(order):
        img = image_ids[d]
        gboxes, difficult = gt[img]
        best, j = -np.inf, -1
        if len(gboxes):
            g = np.asarray(gboxes, dtype=np.float64)
            b = boxes[d]
            iw = np.minimum(g[:, 2], b[2]) - np.maximum(g[:, 0], b[0]) + 1.
            ih = np.minimum(g[:, 3], b[3]) - np.maximum(g[:, 1], b[1]) + 1.
            inter = np.maximum(iw, 0.) * np.maximum(ih, 0.)
            union = (b[2] - b[0] + 1.) * (b[3] - b[1] + 1.) + (g[:, 2] - g[:, 0] + 1.) * (g[:, 3] - g[:, 1] + 1.) - inter
            iou = inter / union
            j = int(np.argmax(iou))
            best = iou[j]
        if best > ovthresh:
            if difficult[j]:
                continue                      # neither TP nor FP
            if claimed[img][j]:
                fp[rank] = 1.
            else:
                tp[rank] = 1.
                claimed[img][j] = True
        else:
            fp[rank] = 1.
    return tp, fp


def voc_eval(detpath, annopath, imagesetfile, classname, cachedir, ovthresh=0.5, use_07_metric=False):
    """rec, prec, ap = voc_eval(...) as scripts/voc_eval.py:96-243: `detpath.format(classname)` is the result file,
    `annopath.format(imagename)` the XML annotation, `imagesetfile` the list of image names."""
    with open(imagesetfile, 'r') as f:
        imagenames = [x.strip() for x in f.readlines()]
    recs = load_annotations(annopath, imagenames, cachedir)
    gt, npos = {}, 0
    for name in imagenames:
        objs = [o for o in recs[name] if o['name'] == classname]
        difficult = np.array([o['difficult'] for o in objs]).astype(bool)
        gt[name] = (np.array([o['bbox'] for o in objs]), difficult)
        npos += int(np.sum(~difficult))
    with open(detpath.format(classname), 'r') as f:
        rows = [x.strip().split(' ') for x in f.readlines()]
    image_ids = [r[0] for r in rows]
    confidence = np.array([float(r[1]) for r in rows])
    boxes = np.array([[float(z) for z in r[2:]] for r in rows])
    tp, fp = match_detections(image_ids, confidence, boxes, gt, ovthresh)
    tp, fp = np.cumsum(tp), np.cumsum(fp)
    rec = tp / float(npos)
    prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
    return rec, prec, voc_ap(rec, prec, use_07_metric)


def mean_ap(detpath, annopath, imagesetfile, classes, cachedir, use_07_metric=True, novel_classes=()):
    """Per-class AP plus the base / novel means the reference prints (scripts/voc_eval.py:_do_python_eval)."""
    aps = dict((c, voc_eval(detpath, annopath, imagesetfile, c, cachedir, 0.5, use_07_metric)[2]) for c in classes)
    base = [aps[c] for c in classes if c not in novel_classes]
    novel = [aps[c] for c in classes if c in novel_classes]
    return {'ap': aps, 'mean': float(np.mean(list(aps.values()))), 'mean_base': float(np.mean(base)) if base else None,
            'mean_novel': float(np.mean(novel)) if novel else None}
