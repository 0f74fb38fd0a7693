"""The in-training precision / recall check of the reference's driver, `train_meta.test()` (train_meta.py:258-315),
as a function over tensors: decode (`utils.get_region_boxes`) -> `utils.nms` -> greedy best-IoU match of every
ground-truth box -> precision, recall, F-score.

Counting rules kept from the reference: `proposals` counts NMS survivors with det_conf > conf_thresh; a ground truth
is `correct` when its best-IoU survivor has IoU > iou_thresh and the same class id; `total` is the number of target
rows before the first all-zero row (truths_length).  eps = 1e-5 and the default thresholds as in the driver (train_meta.py:63-70).
"""
from .utils import bbox_iou, get_region_boxes, nms as device_nms


def truths_length(truths):
    """train_meta.py:259-262 (returns 50 where the reference falls through with None on a full row set)."""
    for i in range(50):
        if truths[i][1] == 0:
            return i
    return 50


def count_matches(all_boxes, target, conf_thresh, nms_thresh, iou_thresh, nms_fn=None):
    """(total, proposals, correct) over a batch.  all_boxes: get_region_boxes' list per output row; target:
    [rows, 250] (or [rows, 50, 5]) labels, one row of up to 50 (cls, x, y, w, h) boxes per output row."""
    nms_fn = device_nms if nms_fn is None else nms_fn
    total = proposals = correct = 0.0
    for i in range(len(all_boxes)):
        boxes = nms_fn(all_boxes[i], nms_thresh)
        truths = target[i].reshape(-1, 5).tolist()
        num_gts = truths_length(truths)
        total = total + num_gts
        for b in boxes:
            if b[4] > conf_thresh:
                proposals = proposals + 1
        for t in range(num_gts):
            box_gt = [truths[t][1], truths[t][2], truths[t][3], truths[t][4], 1.0, 1.0, truths[t][0]]
            best_iou, best_j = 0, -1
            for j in range(len(boxes)):
                iou = bbox_iou(box_gt, boxes[j], x1y1x2y2=False)
                if iou > best_iou:
                    best_j, best_iou = j, iou
            if best_iou > iou_thresh and boxes[best_j][6] == box_gt[6]:
                correct = correct + 1
    return total, proposals, correct


def precision_recall(total, proposals, correct, eps=1e-5):
    precision = 1.0 * correct / (proposals + eps)
    recall = 1.0 * correct / (total + eps)
    fscore = 2.0 * precision * recall / (precision + recall + eps)
    return precision, recall, fscore


def evaluate_batches(model, batches, conf_thresh=0.25, nms_thresh=0.4, iou_thresh=0.5):
    """train_meta.test(): `batches` yields (data, metax, mask, target) like zip(test_loader, test_metaloader);
    thresholds default to the driver's."""
    import torch
    model.eval()
    tot = prop = corr = 0.0
    with torch.no_grad():
        for data, metax, mask, target in batches:
            output = model(data, metax, mask)
            all_boxes = get_region_boxes(output, conf_thresh, model.num_classes, model.anchors, model.num_anchors)
            t, p, c = count_matches(all_boxes, target.reshape(output.size(0), -1), conf_thresh, nms_thresh, iou_thresh)
            tot, prop, corr = tot + t, prop + p, corr + c
    return precision_recall(tot, prop, corr)
