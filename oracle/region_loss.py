"""Oracle restatement of the reference's region loss (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/region_loss.py and utils.py:
  bbox_iou      utils.py:21-52        (scalar, Python float64)
  bbox_ious     utils.py:54-83        (vector, float32, same op order)
  neg_filter    region_loss.py:15-34
  build_targets region_loss.py:37-132
  RegionLoss    region_loss.py:134-232
  RegionLossV2  region_loss.py:235-366
"""
import math
from numbers import Number
from random import random

import numpy as np
import torch
import torch.nn.functional as F

MAX_BOXES = 50  # cfg.max_boxes, cfg.py:29


def bbox_iou(box1, box2):
    """utils.py:21-52 with x1y1x2y2=False; Python floats (f64)."""
    mx = min(box1[0] - box1[2] / 2.0, box2[0] - box2[2] / 2.0)
    Mx = max(box1[0] + box1[2] / 2.0, box2[0] + box2[2] / 2.0)
    my = min(box1[1] - box1[3] / 2.0, box2[1] - box2[3] / 2.0)
    My = max(box1[1] + box1[3] / 2.0, box2[1] + box2[3] / 2.0)
    w1, h1, w2, h2 = box1[2], box1[3], box2[2], box2[3]
    uw = Mx - mx
    uh = My - my
    cw = w1 + w2 - uw
    ch = h1 + h2 - uh
    if cw <= 0 or ch <= 0:
        return 0.0
    area1 = w1 * h1
    area2 = w2 * h2
    carea = cw * ch
    uarea = area1 + area2 - carea
    return carea / uarea


def bbox_ious(b1, b2):
    """utils.py:54-83 with x1y1x2y2=False on float32 arrays of shape [4, n]."""
    assert b1.dtype == np.float32 and b2.dtype == np.float32
    mx = np.minimum(b1[0] - b1[2] / 2.0, b2[0] - b2[2] / 2.0)
    Mx = np.maximum(b1[0] + b1[2] / 2.0, b2[0] + b2[2] / 2.0)
    my = np.minimum(b1[1] - b1[3] / 2.0, b2[1] - b2[3] / 2.0)
    My = np.maximum(b1[1] + b1[3] / 2.0, b2[1] + b2[3] / 2.0)
    w1, h1, w2, h2 = b1[2], b1[3], b2[2], b2[3]
    uw = Mx - mx
    uh = My - my
    cw = w1 + w2 - uw
    ch = h1 + h2 - uh
    mask = (cw <= 0) | (ch <= 0)
    area1 = w1 * h1
    area2 = w2 * h2
    carea = cw * ch
    carea = np.where(mask, np.float32(0), carea)
    uarea = area1 + area2 - carea
    with np.errstate(divide='ignore', invalid='ignore'):
        return (carea / uarea).astype(np.float32)


def neg_filter(target2d, neg_ratio, rng=random):
    """region_loss.py:15-34. `target2d` float64 [rows, 250]. Returns kept row
    indices (ascending). One `rng()` draw per empty row, in row order."""
    n = target2d.shape[0]
    if neg_ratio == 'full':
        return list(range(n))
    if isinstance(neg_ratio, Number):
        flags = [bool(f) for f in (np.asarray(target2d).sum(1) != 0)]
        npos = sum(flags)
        ratio = neg_ratio * npos * 1. / (len(flags) - npos)  # ZeroDivisionError if every row is positive, as the reference
        if ratio >= 1:
            return list(range(n))
        keep = [0 if (not f) and rng() > ratio else 1 for f in flags]
        return [i for i, k in enumerate(keep) if k]
    raise NotImplementedError('neg_ratio not recognized')


def build_targets(pred_boxes, target, anchors, nA, nH, nW, noobject_scale, object_scale, sil_thresh, seen,
                  max_boxes=MAX_BOXES):
    """region_loss.py:37-132.

    pred_boxes float32 [nB*nA*nH*nW, 4] (cx, cy, w, h in grid units), target
    float64 [nB, 250]; anchors list of Python floats (2 per anchor).
    """
    pred_boxes = np.ascontiguousarray(pred_boxes, dtype=np.float32)
    target = np.ascontiguousarray(target, dtype=np.float64)
    nB = target.shape[0]
    anchor_step = len(anchors) // nA
    assert anchor_step == 2
    shape = (nB, nA, nH, nW)
    conf_mask = np.full(shape, noobject_scale, dtype=np.float32)
    coord_mask = np.zeros(shape, np.float32)
    cls_mask = np.zeros(shape, np.float32)
    tx = np.zeros(shape, np.float32)
    ty = np.zeros(shape, np.float32)
    tw = np.zeros(shape, np.float32)
    th = np.zeros(shape, np.float32)
    tconf = np.zeros(shape, np.float32)
    tcls = np.zeros(shape, np.float32)
    nAnchors = nA * nH * nW
    nPixels = nH * nW
    thr = np.float32(sil_thresh)  # float tensor > python scalar compares in float32
    tl = target.tolist()          # torch-0.3.1: element indexing yields Python floats
    for b in range(nB):           # phase 1, :55-69
        cur_pred = pred_boxes[b * nAnchors:(b + 1) * nAnchors].T
        cur_ious = np.zeros(nAnchors, np.float32)
        row = tl[b]
        for t in range(max_boxes):
            if row[t * 5 + 1] == 0:
                break
            gx = row[t * 5 + 1] * nW
            gy = row[t * 5 + 2] * nH
            gw = row[t * 5 + 3] * nW
            gh = row[t * 5 + 4] * nH
            gt = np.array([gx, gy, gw, gh], dtype=np.float32)[:, None].repeat(nAnchors, 1)
            iou = bbox_ious(cur_pred, gt)
            # torch.max(a, b) propagates NaN
            cur_ious = np.where(np.isnan(iou) | np.isnan(cur_ious), np.float32('nan'), np.maximum(cur_ious, iou))
        conf_mask[b].reshape(-1)[cur_ious > thr] = 0
    if seen < 12800:              # :70-79
        tx.fill(0.5)
        ty.fill(0.5)
        tw.fill(0)
        th.fill(0)
        coord_mask.fill(1)
    nGT = 0
    nCorrect = 0
    pl = None
    for b in range(nB):           # phase 2, :83-130
        row = tl[b]
        for t in range(50):
            if row[t * 5 + 1] == 0:
                break
            nGT += 1
            best_iou = 0.0
            best_n = -1
            gx = row[t * 5 + 1] * nW
            gy = row[t * 5 + 2] * nH
            gi = int(gx)
            gj = int(gy)
            gw = row[t * 5 + 3] * nW
            gh = row[t * 5 + 4] * nH
            gt_box = [0, 0, gw, gh]
            for n in range(nA):
                aw = anchors[anchor_step * n]
                ah = anchors[anchor_step * n + 1]
                iou = bbox_iou([0, 0, aw, ah], gt_box)
                if iou > best_iou:
                    best_iou = iou
                    best_n = n
            gt_box = [gx, gy, gw, gh]
            pred_box = [float(v) for v in pred_boxes[b * nAnchors + best_n * nPixels + gj * nW + gi]]
            coord_mask[b, best_n, gj, gi] = 1
            cls_mask[b, best_n, gj, gi] = 1
            conf_mask[b, best_n, gj, gi] = object_scale
            tx[b, best_n, gj, gi] = row[t * 5 + 1] * nW - gi
            ty[b, best_n, gj, gi] = row[t * 5 + 2] * nH - gj
            tw[b, best_n, gj, gi] = math.log(gw / anchors[anchor_step * best_n])
            th[b, best_n, gj, gi] = math.log(gh / anchors[anchor_step * best_n + 1])
            iou = bbox_iou(gt_box, pred_box)
            tconf[b, best_n, gj, gi] = iou
            tcls[b, best_n, gj, gi] = row[t * 5]
            if iou > 0.5:
                nCorrect += 1
    return nGT, nCorrect, coord_mask, conf_mask, cls_mask, tx, ty, tw, th, tconf, tcls


def _decode(output, nA, nC, anchors):
    """region_loss.py:276-298: activations and pred_boxes (detached, float32)."""
    nB, _, nH, nW = output.shape
    o = output.view(nB, nA, 5 + nC, nH, nW)
    x = torch.sigmoid(o[:, :, 0])
    y = torch.sigmoid(o[:, :, 1])
    w = o[:, :, 2]
    h = o[:, :, 3]
    conf = torch.sigmoid(o[:, :, 4])
    n = nB * nA * nH * nW
    grid_x = torch.linspace(0, nW - 1, nW).repeat(nH, 1).repeat(nB * nA, 1, 1).view(n)
    grid_y = torch.linspace(0, nH - 1, nH).repeat(nW, 1).t().repeat(nB * nA, 1, 1).view(n)
    aw = torch.Tensor(anchors).view(nA, 2)[:, 0:1].repeat(nB, 1).repeat(1, 1, nH * nW).view(n)
    ah = torch.Tensor(anchors).view(nA, 2)[:, 1:2].repeat(nB, 1).repeat(1, 1, nH * nW).view(n)
    pb = torch.empty(4, n)
    pb[0] = x.detach().reshape(-1) + grid_x
    pb[1] = y.detach().reshape(-1) + grid_y
    pb[2] = torch.exp(w.detach()).reshape(-1) * aw
    pb[3] = torch.exp(h.detach()).reshape(-1) * ah
    return x, y, w, h, conf, pb.t().contiguous()


def region_loss_v2(output, target, anchors, num_anchors, num_classes=1, seen=0, coord_scale=1.0,
                   noobject_scale=1.0, object_scale=5.0, class_scale=1.0, thresh=0.6, neg_ratio='full',
                   rng=random, return_parts=False):
    """region_loss.py:252-366 (RegionLossV2.forward). `output` float32
    [bs*cs, nA*(5+nC), nH, nW] (autograd leaf or graph tensor), `target`
    float64 [bs, cs, 250]. Returns the scalar loss (autograd-connected)."""
    bs, cs = target.shape[0], target.shape[1]
    nA, nC = num_anchors, num_classes
    nH, nW = output.shape[2], output.shape[3]
    cls = output.view(output.size(0), nA, 5 + nC, nH, nW)[:, :, 5:5 + nC].squeeze(2)
    cls = cls.reshape(bs, cs, nA * nC * nH * nW).transpose(1, 2).contiguous().view(bs * nA * nC * nH * nW, cs)
    target2d = target.reshape(-1, target.shape[-1])
    inds = neg_filter(target2d.numpy(), neg_ratio, rng)
    counts, _ = np.histogram(inds, bins=bs, range=(0, bs * cs))
    ind_t = torch.as_tensor(inds, dtype=torch.long)
    out_f = output[ind_t]
    tgt_f = target2d[ind_t]
    nB = out_f.shape[0]
    x, y, w, h, conf, pred_boxes = _decode(out_f, nA, nC, anchors)
    nGT, nCorrect, coord_mask, conf_mask, cls_mask, tx, ty, tw, th, tconf, tcls = build_targets(
        pred_boxes.numpy(), tgt_f.numpy(), anchors, nA, nH, nW, noobject_scale, object_scale, thresh, seen)
    # per-image merge of the class mask, :303-319
    idx = 0
    cm_list, tc_list = [], []
    for i in range(len(counts)):
        if counts[i] == 0:
            cm_list.append(np.zeros((nA, nH, nW), np.float32))
            tc_list.append(np.zeros((nA, nH, nW), np.float32))
        else:
            cm_list.append(cls_mask[idx:idx + counts[i]].sum(0))
            tc_list.append(tcls[idx:idx + counts[i]].sum(0))
        idx += counts[i]
    cls_mask_img = torch.from_numpy(np.stack(cm_list)) == 1
    tcls_img = torch.from_numpy(np.stack(tc_list))
    nProposals = int((conf > 0.25).float().sum().item())
    T = lambda a: torch.from_numpy(a)
    coord_m = T(coord_mask)
    conf_m = T(conf_mask).sqrt()
    cls_sel = cls[cls_mask_img.view(-1, 1).repeat(1, cs)].view(-1, cs)
    tcls_sel = tcls_img.view(-1)[cls_mask_img.view(-1)].long()
    mse = lambda a, b: F.mse_loss(a, b, reduction='sum')
    loss_x = coord_scale * mse(x * coord_m, T(tx) * coord_m) / 2.0
    loss_y = coord_scale * mse(y * coord_m, T(ty) * coord_m) / 2.0
    loss_w = coord_scale * mse(w * coord_m, T(tw) * coord_m) / 2.0
    loss_h = coord_scale * mse(h * coord_m, T(th) * coord_m) / 2.0
    loss_conf = mse(conf * conf_m, T(tconf) * conf_m) / 2.0
    loss_cls = class_scale * F.cross_entropy(cls_sel, tcls_sel, reduction='sum')
    loss = loss_x + loss_y + loss_w + loss_h + loss_conf + loss_cls
    if return_parts:
        parts = dict(nGT=nGT, nCorrect=nCorrect, nProposals=nProposals, loss_x=loss_x.item(), loss_y=loss_y.item(),
                     loss_w=loss_w.item(), loss_h=loss_h.item(), loss_conf=loss_conf.item(),
                     loss_cls=loss_cls.item(), loss=loss.item(), inds=inds)
        return loss, parts
    return loss


def region_loss_plain(output, target, anchors, num_anchors, num_classes, seen=0, coord_scale=1.0,
                      noobject_scale=1.0, object_scale=5.0, class_scale=1.0, thresh=0.6, neg_ratio='full',
                      metayolo=True, rng=random, return_parts=False):
    """region_loss.py:148-232 (RegionLoss.forward). `target` float64 [B, 250]
    (or [B, n, 250], flattened as the reference does)."""
    if target.dim() == 3:
        target = target.reshape(-1, target.shape[-1])
    inds = neg_filter(target.numpy(), neg_ratio, rng)
    ind_t = torch.as_tensor(inds, dtype=torch.long)
    output = output[ind_t]
    target = target[ind_t]
    nB = output.shape[0]
    nA, nC = num_anchors, num_classes
    nH, nW = output.shape[2], output.shape[3]
    x, y, w, h, conf, pred_boxes = _decode(output, nA, nC, anchors)
    cls = output.view(nB, nA, 5 + nC, nH, nW)[:, :, 5:5 + nC]
    cls = cls.reshape(nB * nA, nC, nH * nW).transpose(1, 2).contiguous().view(nB * nA * nH * nW, nC)
    nGT, nCorrect, coord_mask, conf_mask, cls_mask, tx, ty, tw, th, tconf, tcls = build_targets(
        pred_boxes.numpy(), target.numpy(), anchors, nA, nH, nW, noobject_scale, object_scale, thresh, seen)
    T = lambda a: torch.from_numpy(a)
    cls_mask_b = T(cls_mask) == 1
    if metayolo:
        tcls = np.zeros_like(tcls)
    nProposals = int((conf > 0.25).float().sum().item())
    tcls_sel = T(tcls).view(-1)[cls_mask_b.view(-1)].long()
    coord_m = T(coord_mask)
    conf_m = T(conf_mask).sqrt()
    cls_sel = cls[cls_mask_b.view(-1, 1).repeat(1, nC)].view(-1, nC)
    mse = lambda a, b: F.mse_loss(a, b, reduction='sum')
    loss_x = coord_scale * mse(x * coord_m, T(tx) * coord_m) / 2.0
    loss_y = coord_scale * mse(y * coord_m, T(ty) * coord_m) / 2.0
    loss_w = coord_scale * mse(w * coord_m, T(tw) * coord_m) / 2.0
    loss_h = coord_scale * mse(h * coord_m, T(th) * coord_m) / 2.0
    loss_conf = mse(conf * conf_m, T(tconf) * conf_m) / 2.0
    loss_cls = class_scale * F.cross_entropy(cls_sel, tcls_sel, reduction='sum')
    loss = loss_x + loss_y + loss_w + loss_h + loss_conf + loss_cls
    if return_parts:
        parts = dict(nGT=nGT, nCorrect=nCorrect, nProposals=nProposals, loss_x=loss_x.item(), loss_y=loss_y.item(),
                     loss_w=loss_w.item(), loss_h=loss_h.item(), loss_conf=loss_conf.item(),
                     loss_cls=loss_cls.item(), loss=loss.item(), inds=inds)
        return loss, parts
    return loss
