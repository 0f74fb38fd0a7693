"""RegionLoss / RegionLossV2 with the whole loss on the device.

Same Python surface as the reference's region_loss.py: modules constructed with
no arguments and attribute-assigned by the network builder
(darknet_meta.py:322-334), `loss = region_loss(output, target)` with `target`
a CPU float64 tensor ([B, n_cls, 250] for V2, [B, 250] or [B, n, 250] for the
plain loss), `.seen` mutated by the training driver, `cfg.neg_ratio /
cfg.metayolo / cfg.max_boxes` read at call time.

What changed underneath: the reference decodes boxes with ~25 small kernels,
copies pred_boxes to the host, runs `build_targets` as a Python double loop
(region_loss.py:37-132), copies nine target tensors back and evaluates six
losses with autograd.  Here three kernels do it all (csrc/region.cu):
decode -> build_targets (bit-exact masks / indices) -> loss + d(loss)/d(output).
`neg_filter` keeps its host RNG semantics (region_loss.py:15-34).
"""
from numbers import Number
from random import random

import numpy as np
import torch
import torch.nn as nn

from .cfg import cfg
from ._lib import call, ptr


def _st():
    return torch.cuda.current_stream().cuda_stream


def kept_rows(target):
    """The row selection of region_loss.py:15-34 on the [rows, 250] label matrix: every row with a label is kept,
    every empty row with probability neg_ratio * #labelled / #empty, one Python `random()` draw per empty row in row
    order (the reference's RNG consumption).  Returns the kept row indices (ascending list)."""
    n = target.size(0)
    if cfg.neg_ratio == 'full':
        return list(range(n))
    if isinstance(cfg.neg_ratio, Number):
        flags = (torch.sum(target, 1) != 0).cpu().tolist()
        ratio = cfg.neg_ratio * sum(flags) * 1. / (len(flags) - sum(flags))
        if ratio >= 1:
            return list(range(n))
        flags = [0 if f == 0 and random() > ratio else 1 for f in flags]
        return [i for i, f in enumerate(flags) if f]
    raise NotImplementedError('neg_ratio not recognized')


def neg_filter(pred_boxes, target, withids=False):
    """region_loss.py:15-34 with the reference's signature. `target` is the float64 [rows, 250] label
    matrix. Returns the kept row indices (python list, ascending); the rows
    themselves are gathered on the device by the caller."""
    assert pred_boxes.size(0) == target.size(0)
    inds = kept_rows(target)
    if withids:
        return pred_boxes, target, inds
    return pred_boxes, target


def _to_device_f64(target_rows, device):
    if target_rows.is_cuda:
        return target_rows.to(torch.float64).contiguous()
    t = target_rows.to(torch.float64).contiguous()
    if not t.is_pinned():
        try:
            t = t.pin_memory()
        except RuntimeError:
            pass
    return t.to(device, non_blocking=True)


def build_targets(pred_boxes, target, anchors, num_anchors, num_classes, nH, nW, noobject_scale, object_scale,
                  sil_thresh, seen, sync=True):
    """Device version of region_loss.py:37-132 with the reference's signature.

    pred_boxes: float32 CUDA [nB*nA*nH*nW, 4]; target: float64 [nB, 250] (CPU or
    CUDA). Returns (nGT, nCorrect, coord_mask, conf_mask, cls_mask, tx, ty, tw,
    th, tconf, tcls) with the nine tensors float32 CUDA [nB, nA, nH, nW].  With
    sync=False nGT/nCorrect are returned as a CUDA int32 tensor [4] instead of
    Python ints (no host synchronisation)."""
    dev = pred_boxes.device
    if not pred_boxes.is_cuda:
        raise TypeError('build_targets runs on the GPU only; pred_boxes must be a CUDA tensor')
    nB = target.size(0)
    nA = num_anchors
    assert len(anchors) // num_anchors == 2, 'anchor_step must be 2'
    tgt = _to_device_f64(target, dev)
    out = torch.empty(9, nB, nA, nH, nW, device=dev)
    counters = torch.empty(4, dtype=torch.int32, device=dev)
    anc = torch.tensor([float(a) for a in anchors], dtype=torch.float64).to(dev)
    pb = pred_boxes.contiguous()
    call('fsdet_build_targets', ptr(pb), ptr(tgt), ptr(anc), nB, nA, nH, nW, int(cfg.max_boxes),
         float(noobject_scale), float(object_scale), float(sil_thresh), int(seen),
         *[ptr(out[i]) for i in range(9)], ptr(counters), None, None, _st())
    if not sync:
        return (counters,) + tuple(out[i] for i in range(9))
    c = counters.tolist()
    if c[2]:
        raise ValueError('math domain error: %d ground-truth boxes with zero width/height or outside the grid' % c[2])
    return (c[0], c[1]) + tuple(out[i] for i in range(9))


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, loss, grad):
        ctx.save_for_backward(grad)
        return loss.clone()

    @staticmethod
    def backward(ctx, gout):
        (grad,) = ctx.saved_tensors
        return grad * gout, None, None


class _RegionBase(nn.Module):
    def __init__(self, num_classes=0, anchors=[], num_anchors=1):
        super(_RegionBase, self).__init__()
        self.num_classes = num_classes
        self.anchors = anchors
        self.num_anchors = num_anchors
        self.anchor_step = len(anchors) // num_anchors
        self.coord_scale = 1
        self.noobject_scale = 1
        self.object_scale = 5
        self.class_scale = 1
        self.thresh = 0.6
        self.seen = 0
        self.verbose = True   # print the reference's per-step log line (forces a host sync)
        self.last = None      # dict with the logged scalars of the last call (device tensors when not verbose)
        self.static = None    # fixed-capacity device buffers of a CUDA-graph-captured step (graph.GraphedTrainStep)
        self._pending = None  # (pinned counters, event) of the last non-verbose call, checked one call late

    # -- degenerate ground truth (the reference raises `math domain error` from math.log, region_loss.py:114) -----
    def _note_counters(self, counters):
        """Non-verbose calls do not synchronise: the kernel's degenerate-box counter is copied to pinned memory behind
        the step and examined at the NEXT call (or by poll()), so a zero-size / out-of-grid label is still reported,
        one step late, instead of silently training on garbage."""
        if torch.cuda.is_current_stream_capturing():
            return
        if self._pending is None:
            self._pending = [torch.zeros(4, dtype=torch.int32).pin_memory(), torch.cuda.Event(), False]
        host, ev, armed = self._pending
        if armed:
            ev.synchronize()
            self._raise_if_degenerate(host)
        host.copy_(counters, non_blocking=True)
        ev.record()
        self._pending[2] = True

    @staticmethod
    def _raise_if_degenerate(host):
        if int(host[2]):
            raise ValueError('math domain error: %d ground-truth boxes with zero width/height or outside the grid '
                             '(reported one step late: verbose=False)' % int(host[2]))

    def poll(self):
        """Check the degenerate-label counter of the last non-verbose call (blocks until that step has finished)."""
        if self._pending is not None and self._pending[2]:
            self._pending[1].synchronize()
            self._pending[2] = False
            self._raise_if_degenerate(self._pending[0])

    # -- CUDA-graph form -------------------------------------------------------------------------------------------
    def warm_caches(self, device, bs=0, cs=0):
        """Upload the small constants a call needs (anchors, the per-image row prefix of the unsampled loss) NOW: a
        host -> device copy of pageable memory is not allowed inside a CUDA-graph capture."""
        key = (str(device), tuple(float(a) for a in self.anchors))
        if getattr(self, '_anchor_cache', (None,))[0] != key:
            self._anchor_cache = (key, torch.tensor(key[1], dtype=torch.float32).to(device),
                                  torch.tensor(key[1], dtype=torch.float64).to(device))
        if bs:
            ikey = (str(device), tuple(range(0, bs * cs + 1, cs)))
            if getattr(self, '_img_cache', (None,))[0] != ikey:
                self._img_cache = (ikey, torch.tensor(ikey[1], dtype=torch.int32).to(device))

    def make_static(self, rows_total, bs, device):
        """Allocate the fixed-capacity buffers a captured step reads: [0] = number of live rows, [1 : bs+2] = per-image
        prefix of the kept rows (V2), [bs+2 :] = kept row indices.  The kernels are launched for `rows_total` slots and
        skip the dead ones, so ONE graph serves every outcome of neg_filter."""
        n = 1 + (bs + 1) + rows_total
        return {'rows_total': rows_total, 'bs': bs, 'dev': torch.zeros(n, dtype=torch.int32, device=device),
                'ring': [[torch.zeros(n, dtype=torch.int32).pin_memory(), torch.cuda.Event(), False] for _ in range(3)],
                'turn': 0}

    @staticmethod
    def stage_filter(st, target):
        """Host half of a captured step: neg_filter on the CPU label tensor (the reference's `random()` draws) and one
        asynchronous upload of (live rows, img_start, inds) into the static buffers `st` (from make_static); call
        before every replay.  While `self.static = st` is set, forward() takes the fixed-capacity form."""
        t2 = target.view(-1, target.size(-1))
        if t2.is_cuda:
            raise TypeError('stage_filter needs the CPU label tensor (the reference keeps `target` on the host)')
        if t2.size(0) != st['rows_total']:
            raise ValueError('label rows %d != captured capacity %d' % (t2.size(0), st['rows_total']))
        inds = kept_rows(t2)
        host, ev, armed = st['ring'][st['turn']]
        if armed:
            ev.synchronize()
        host[0] = len(inds)
        bs = st['bs']
        if bs:
            cs = st['rows_total'] // bs
            counts, _ = np.histogram(inds, bins=bs, range=(0, bs * cs))
            host[1:bs + 2] = torch.from_numpy(np.concatenate([[0], np.cumsum(counts)]).astype(np.int32))
        if inds:
            host[bs + 2:bs + 2 + len(inds)] = torch.tensor(inds, dtype=torch.int32)
        st['dev'].copy_(host, non_blocking=True)
        ev.record()
        st['ring'][st['turn']][2] = True
        st['turn'] = (st['turn'] + 1) % len(st['ring'])
        return inds

    def _run(self, output, target2d, inds, mode, bs, cs, img_start):
        if not output.is_cuda:
            raise TypeError('the region loss runs on the GPU only (no CPU fallback)')
        dev = output.device
        st = _st()
        out_c = output.detach().contiguous()
        rows_total = out_c.size(0)
        nA, nC = int(self.num_anchors), int(self.num_classes)
        nH, nW = out_c.size(2), out_c.size(3)
        assert out_c.size(1) == nA * (5 + nC)
        static = self.static if inds is None else None
        nb_dev = bt_inds = None
        if static is not None:
            # captured step: launches at full capacity, live-row count / indices / image prefix read from device memory
            if static['rows_total'] != rows_total or (mode == 0 and static['bs'] != bs):
                raise ValueError('static loss buffers were made for another shape')
            nB = rows_total
            sb = static['dev']
            nb_dev = sb[0:1]
            inds_t = bt_inds = sb[static['bs'] + 2:]
            tgt = _to_device_f64(target2d, dev)          # the FULL label matrix, rows gathered by the kernel
        else:
            nB = len(inds)
            full = nB == rows_total
            inds_t = None
            if not full:
                inds_t = torch.tensor(inds, dtype=torch.int32).to(dev, non_blocking=True)
            if full:
                tgt_rows = target2d
            elif target2d.is_cuda:
                tgt_rows = target2d[inds_t.long()]
            else:
                tgt_rows = target2d[torch.as_tensor(inds, dtype=torch.long)]
            tgt = _to_device_f64(tgt_rows, dev)
        self.warm_caches(dev)                 # small constants are uploaded once (CUDA-graph friendly)
        anc32, anc64 = self._anchor_cache[1], self._anchor_cache[2]
        pred = torch.empty(max(nB, 1) * nA * nH * nW, 4, device=dev)
        call('fsdet_region_decode', ptr(out_c), ptr(inds_t), nB, ptr(nb_dev), nA, nC, nH, nW, ptr(anc32), ptr(pred), st)
        tg = torch.empty(9, max(nB, 1), nA, nH, nW, device=dev)
        counters = torch.empty(4, dtype=torch.int32, device=dev)
        call('fsdet_build_targets', ptr(pred), ptr(tgt), ptr(anc64), nB, nA, nH, nW, int(cfg.max_boxes),
             float(self.noobject_scale), float(self.object_scale), float(self.thresh), int(self.seen),
             *[ptr(tg[i]) for i in range(9)], ptr(counters), ptr(bt_inds), ptr(nb_dev), st)
        grad = torch.empty_like(out_c)
        losses = torch.empty(8, dtype=torch.float64, device=dev)
        imgs_t = None
        if mode == 0 and static is not None:
            imgs_t = static['dev'][1:static['bs'] + 2]
        elif mode == 0:
            ikey = (str(dev), tuple(img_start))
            if getattr(self, '_img_cache', (None,))[0] != ikey:
                self._img_cache = (ikey, torch.tensor(img_start, dtype=torch.int32).to(dev))
            imgs_t = self._img_cache[1]
        call('fsdet_region_loss_grad', ptr(out_c), ptr(grad), ptr(inds_t), ptr(nb_dev), ptr(imgs_t), rows_total, nB, bs, cs, nA,
             nC, nH, nW, *[ptr(tg[i]) for i in range(9)], float(self.coord_scale), float(self.class_scale), mode,
             1 if cfg.metayolo else 0, ptr(losses), st)
        loss = losses[6].to(torch.float32)
        self.last = {'losses': losses, 'counters': counters}
        if self.verbose:
            host = losses.tolist()
            c = counters.tolist()
            if c[2]:
                raise ValueError('math domain error: %d ground-truth boxes with zero width/height or outside the grid'
                                 % c[2])
            print('%d: nGT %d, recall %d, proposals %d, loss: x %f, y %f, w %f, h %f, conf %f, cls %f, total %f' % (
                self.seen, c[0], c[1], int(host[7]), host[0], host[1], host[2], host[3], host[4], host[5], host[6]))
            self.last.update(nGT=c[0], nCorrect=c[1], nProposals=int(host[7]), loss_x=host[0], loss_y=host[1],
                             loss_w=host[2], loss_h=host[3], loss_conf=host[4], loss_cls=host[5], loss=host[6])
        else:
            self._note_counters(counters)
        if output.requires_grad and torch.is_grad_enabled():
            return _LossFn.apply(output, loss, grad)
        return loss


class RegionLoss(_RegionBase):
    """region_loss.py:134-232 (plain YOLOv2 region loss)."""

    def forward(self, output, target):
        if target.dim() == 3:
            target = target.view(-1, target.size(-1))
        if self.static is not None:       # captured step: the row selection was staged by stage_filter()
            return self._run(output, target, None, 1, 0, 0, None)
        _, _, inds = neg_filter(output, target, withids=True)
        return self._run(output, target, inds, 1, 0, 0, None)


class RegionLossV2(_RegionBase):
    """region_loss.py:235-366: region loss + softmax classification across the
    n_cls class branches of every image."""

    def __init__(self, num_classes=0, anchors=[], num_anchors=1):
        super(RegionLossV2, self).__init__(num_classes, anchors, num_anchors)
        print('class_scale', self.class_scale)

    def forward(self, output, target):
        bs = target.size(0)
        cs = target.size(1)
        target2d = target.view(-1, target.size(-1))
        if self.static is not None:       # captured step: the row selection was staged by stage_filter()
            return self._run(output, target2d, None, 0, bs, cs, None)
        _, _, inds = neg_filter(output, target2d, withids=True)
        counts, _ = np.histogram(inds, bins=bs, range=(0, bs * cs))
        img_start = [0] + [int(v) for v in np.cumsum(counts)]
        return self._run(output, target2d, inds, 0, bs, cs, img_start)
