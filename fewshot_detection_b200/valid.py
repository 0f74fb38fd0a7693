"""Evaluation path of the meta detector (valid_ensemble.py:13-181) on device-resident tensors.

The reference's `valid()` interleaves four things: dataset / file IO (dataset.listDataset, dataset.MetaDataset: the
input pipeline, SURVEY.md 8f row 3, not part of this build), the ensembling of the support net's reweighting vectors
(:86-100), the query forward + decode + NMS (:140-162) and the result-file format (:163-178).  This module provides the
last three with the loaders replaced by plain iterables of tensors, so a caller that owns a data pipeline gets the
same files:

    dw   = ensemble_dynamic_weights(m, meta_batches, n_cls)             # [ [n_cls, C, 1, 1] ]
    dets = detect(m, data, dw, n_cls)                                    # Detections, NMS done, still on the device
    write_detections(fps, dets, imgids, sizes, n_cls)                    # 'imgid prob x1 y1 x2 y2' per class file

CUDA only (libfsdet.so); no host fallback.
"""
import os

import torch

from ._lib import call, ptr
from .utils import region_detections

CONF_THRESH = 0.005   # valid_ensemble.py:137
NMS_THRESH = 0.45     # valid_ensemble.py:138


def _st():
    return torch.cuda.current_stream().cuda_stream


class ReweightEnsembler(object):
    """Running mean of the reweighting vectors per class (valid_ensemble.py:86-100):
    enews[c] = enews[c]*cnt[c]/(cnt[c]+1) + dw[ci]/(cnt[c]+1); cnt[c] += 1, in sample order, float32."""

    def __init__(self, n_cls, channels, device):
        self.n_cls, self.C = n_cls, channels
        self.enews = torch.zeros(n_cls, channels, dtype=torch.float32, device=device)
        self._cnt = [torch.zeros(n_cls, dtype=torch.int32, device=device) for _ in range(2)]
        self._cur = 0

    def update(self, dw, clsids):
        """dw: CUDA float32 [n, C(,1,1)]; clsids: n class indices (list / tensor)."""
        n = int(dw.size(0))
        dw = dw.detach().reshape(n, -1).float().contiguous()
        assert dw.size(1) == self.C
        ids = torch.as_tensor([int(c) for c in clsids], dtype=torch.int32).to(dw.device)
        assert ids.numel() == n
        if n and (int(ids.min()) < 0 or int(ids.max()) >= self.n_cls):
            raise IndexError('class id out of range')        # the reference's `enews[c]` raises IndexError too
        cin, cout = self._cnt[self._cur], self._cnt[1 - self._cur]
        call('fsdet_rw_running_mean', ptr(self.enews), ptr(cin), ptr(cout), ptr(dw), ptr(ids), n, self.n_cls, self.C, _st())
        self._cur = 1 - self._cur

    @property
    def counts(self):
        return self._cnt[self._cur]

    def result(self):
        """`[torch.stack(enews)]`: [ [n_cls, C, 1, 1] ]."""
        return [self.enews.view(self.n_cls, self.C, 1, 1)]


def ensemble_dynamic_weights(m, meta_batches, n_cls):
    """valid_ensemble.py:86-100.  `meta_batches` yields (metax [n,3,S,S], mask [n,1,S,S], clsids [n]) like the
    reference's MetaDataset(ensemble=True, with_ids=True) loader."""
    ens = None
    with torch.no_grad():
        for metax, mask, clsids in meta_batches:
            dev = next(m.parameters()).device
            dw = m.meta_forward(metax.to(dev), mask.to(dev))[0]
            if ens is None:
                ens = ReweightEnsembler(n_cls, dw[0].numel(), dw.device)
            ens.update(dw, clsids)
    if ens is None:
        raise ValueError('no support batches')
    return ens.result()


def detect(m, data, dynamic_weights, n_cls, conf_thresh=CONF_THRESH, nms_thresh=NMS_THRESH):
    """valid_ensemble.py:140-162 for one batch: detect_forward -> get_region_boxes_v2(only_objectness=0,
    validation=1) -> nms for every (image, class) row.  Returns utils.Detections (device resident)."""
    with torch.no_grad():
        output = m.detect_forward(data, dynamic_weights)
    dets = region_detections(output, conf_thresh, m.num_classes, m.anchors, m.num_anchors, 0, 1, n_models=n_cls)
    return dets.nms(nms_thresh)


def detection_lines(dets, imgids, sizes, n_cls, nms_thresh=NMS_THRESH):
    """valid_ensemble.py:153-178: {class index: [lines]} with `imgid prob x1 y1 x2 y2` per surviving box.
    imgids[b], sizes[b] = (width, height) of image b of the batch."""
    kept = dets.kept_boxes(nms_thresh)
    bs = dets.N // n_cls
    assert len(imgids) == bs and len(sizes) == bs
    out = dict((i, []) for i in range(n_cls))
    for b in range(bs):
        width, height = sizes[b]
        for i in range(n_cls):
            for box in kept[b * n_cls + i]:
                x1 = (box[0] - box[2] / 2.0) * width
                y1 = (box[1] - box[3] / 2.0) * height
                x2 = (box[0] + box[2] / 2.0) * width
                y2 = (box[1] + box[3] / 2.0) * height
                det_conf = box[4]
                for j in range((len(box) - 5) // 2):
                    prob = det_conf * box[5 + 2 * j]
                    out[i].append('%s %f %f %f %f %f\n' % (imgids[b], prob, x1, y1, x2, y2))
    return out


def write_detections(fps, dets, imgids, sizes, n_cls, nms_thresh=NMS_THRESH):
    """Append the batch's lines to the per-class files `fps[i]` (valid_ensemble.py:128-131, :178)."""
    lines = detection_lines(dets, imgids, sizes, n_cls, nms_thresh)
    for i in range(n_cls):
        fps[i].writelines(lines[i])


def valid_batches(m, meta_batches, image_batches, class_names, prefix, outfile):
    """The body of valid_ensemble.valid() (:86-181) over iterables: `meta_batches` as in ensemble_dynamic_weights,
    `image_batches` yields (data [b,3,H,W], imgids, sizes).  Writes `<prefix>/<outfile><class>.txt`."""
    n_cls = len(class_names)
    m.eval()
    dynamic_weights = ensemble_dynamic_weights(m, meta_batches, n_cls)
    if not os.path.exists(prefix):
        os.makedirs(prefix)
    fps = [open('%s/%s%s.txt' % (prefix, outfile, name), 'w') for name in class_names]
    try:
        dev = next(m.parameters()).device
        for data, imgids, sizes in image_batches:
            dets = detect(m, data.to(dev), dynamic_weights, n_cls)
            write_detections(fps, dets, imgids, sizes, n_cls)
    finally:
        for fp in fps:
            fp.close()
    return dynamic_weights
