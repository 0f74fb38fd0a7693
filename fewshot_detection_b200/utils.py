"""Small host helpers of the reference's utils.py that the training driver uses."""
import time


def read_data_cfg(datacfg):
    """`.data` file -> dict (utils.py:460-475), same defaults."""
    options = dict()
    options['gpus'] = '0,1,2,3'
    options['num_workers'] = '10'
    with open(datacfg, 'r') as fp:
        for line in fp.readlines():
            line = line.strip()
            if line == '':
                continue
            key, value = line.split('=')
            options[key.strip()] = value.strip()
    return options


def logging(message):
    print('%s %s' % (time.strftime("%Y-%m-%d %H:%M:%S", time.localtime()), message))


# --------------------------------------------------------------------------------------------------------------------
# Detection decode + NMS (evaluation).  Same names, arguments and return values as the reference's
# utils.py; the tensor prologue, the confidence filter and the O(n^2) suppression loops run in libfsdet.so
# (csrc/detect.cu) for ALL rows of the batch at once.  CUDA only: there is no host fallback.
def bbox_iou(box1, box2, x1y1x2y2=True):
    """utils.py:21-52 on Python floats (host helper, e.g. for train_meta.test()'s recall count)."""
    if x1y1x2y2:
        mx, Mx = min(box1[0], box2[0]), max(box1[2], box2[2])
        my, My = min(box1[1], box2[1]), max(box1[3], box2[3])
        w1, h1, w2, h2 = box1[2] - box1[0], box1[3] - box1[1], box2[2] - box2[0], box2[3] - box2[1]
    else:
        mx = min(box1[0] - box1[2] / 2.0, box2[0] - box2[2] / 2.0)
        Mx = max(box1[0] + box1[2] / 2.0, box2[0] + box2[2] / 2.0)
        my = min(box1[1] - box1[3] / 2.0, box2[1] - box2[3] / 2.0)
        My = max(box1[1] + box1[3] / 2.0, box2[1] + box2[3] / 2.0)
        w1, h1, w2, h2 = box1[2], box1[3], box2[2], box2[3]
    cw = w1 + w2 - (Mx - mx)
    ch = h1 + h2 - (My - my)
    if cw <= 0 or ch <= 0:
        return 0.0
    carea = cw * ch
    return carea / (w1 * h1 + w2 * h2 - carea)


def _stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


class Detections(object):
    """Device-resident result of the decode (+ NMS) of one head output.

    cand   float32 [N, A*H*W, 8]: per row the candidates above the confidence threshold in the reference's loop order,
           {xs, ys, ws, hs (grid units), det_conf, cls_max_conf, int32 cls_max_id, int32 a*H*W+cell}
    count  int32 [N]
    keep / keep_count (after `.nms(thresh)`): candidate slots of the NMS survivors per row, best first.
    Nothing is copied to the host until `.boxes()` / `.kept_boxes()` / `.lines()` is called."""

    def __init__(self, cand, count, cls_dense, N, A, nC, H, W, only_objectness, validation, conf_thresh):
        self.cand, self.count, self.cls_dense = cand, count, cls_dense
        self.N, self.A, self.nC, self.H, self.W = N, A, nC, H, W
        self.only_objectness, self.validation, self.conf_thresh = only_objectness, validation, conf_thresh
        self.keep = self.keep_count = None
        self._nms_thresh = None
        self._host = None
        self._rows = None

    # ---- device side
    def nms(self, nms_thresh):
        """utils.nms (utils.py:85-104) for every row in one launch.  Returns self."""
        import torch
        from ._lib import call, ptr
        if self._nms_thresh != nms_thresh:
            cap = self.A * self.H * self.W
            self.keep = torch.empty(self.N, cap, dtype=torch.int32, device=self.cand.device)
            self.keep_count = torch.zeros(self.N, dtype=torch.int32, device=self.cand.device)
            call('fsdet_nms', ptr(self.cand), ptr(self.count), self.N, cap, self.H, self.W, float(nms_thresh),
                 ptr(self.keep), ptr(self.keep_count), _stream())
            self._nms_thresh = nms_thresh
            self._kept_host = None
        return self

    # ---- host side (one D2H copy each)
    def _fetch(self):
        if self._host is None:
            count = self.count.cpu().numpy()
            mx = int(count.max()) if self.N else 0
            cand = self.cand[:, :mx].cpu().numpy()
            dense = self.cls_dense.cpu().numpy() if (self.cls_dense is not None and self.validation
                                                     and not self.only_objectness and self.nC > 1) else None
            self._host = (count, cand, dense)
        return self._host

    def _box(self, n, slot, cand, dense):
        """One box in the reference's list form (utils.py:175-181 / :270-276), Python floats."""
        import numpy as np
        v = cand[n, slot]
        ints = v[6:8].view(np.int32)
        det, cid = float(v[4]), int(ints[0])
        box = [float(v[0]) / self.W, float(v[1]) / self.H, float(v[2]) / self.W, float(v[3]) / self.H, det, float(v[5]), cid]
        if dense is not None:
            row = dense[n * self.A * self.H * self.W + int(ints[1])]
            for c in range(self.nC):
                tmp = float(row[c])
                if c != cid and det * tmp > self.conf_thresh:
                    box.append(tmp)
                    box.append(c)
        return box

    def boxes(self):
        """all_boxes of get_region_boxes(_v2): list (rows) of lists (boxes) of Python numbers."""
        if self._rows is None:
            count, cand, dense = self._fetch()
            self._rows = [_Row([self._box(n, s, cand, dense) for s in range(int(count[n]))], self, n)
                          for n in range(self.N)]
        return self._rows

    def kept_boxes(self, nms_thresh):
        """[nms(row, nms_thresh) for row in boxes()] without building the un-kept boxes' lists."""
        self.nms(nms_thresh)
        count, cand, dense = self._fetch()
        kc = self.keep_count.cpu().numpy()
        mx = int(kc.max()) if self.N else 0
        keep = self.keep[:, :mx].cpu().numpy()
        return [[self._box(n, int(keep[n, i]), cand, dense) for i in range(int(kc[n]))] for n in range(self.N)]

    def _nms_row(self, index, row, nms_thresh):
        """Reference semantics for one row of boxes(): survivors returned best first (the same list objects),
        suppressed boxes get box[4] = 0 in place."""
        self.nms(nms_thresh)
        if getattr(self, '_kept_host', None) is None:
            self._kept_host = (self.keep_count.cpu().numpy(), self.keep.cpu().numpy())
        kc, keep = self._kept_host
        slots = [int(s) for s in keep[index, :int(kc[index])]]
        alive = set(slots)
        for s, box in enumerate(row):
            if s not in alive:
                box[4] = 0
        return [row[s] for s in slots]


class _Row(list):
    """A row of get_region_boxes(_v2)'s result that remembers where it came from, so that `nms(row, t)` can use
    the batched device NMS instead of re-uploading the boxes."""

    def __init__(self, boxes, parent, index):
        super(_Row, self).__init__(boxes)
        self._parent, self._index, self._n0 = parent, index, len(boxes)
        self._sig = [b[4] for b in boxes]

    def _pristine(self):
        return len(self) == self._n0 and all(b[4] == s for b, s in zip(self, self._sig))


def _detect(output, n_models, v2, conf_thresh, num_classes, anchors, num_anchors, only_objectness, validation):
    import torch
    from ._lib import call, ptr
    if not torch.is_tensor(output) or not output.is_cuda:
        raise TypeError('get_region_boxes runs on the GPU only; `output` must be a CUDA tensor (no CPU fallback)')
    if output.dim() == 3:
        output = output.unsqueeze(0)
    output = output.detach().float().contiguous()
    N, ch, H, W = output.shape
    A, nC = int(num_anchors), int(num_classes)
    assert ch == (5 + nC) * A
    assert len(anchors) // A == 2, 'anchor_step must be 2'
    if v2:
        assert N % n_models == 0
    dev = output.device
    cap = A * H * W
    cand = torch.empty(N, cap, 8, dtype=torch.float32, device=dev)
    count = torch.zeros(N, dtype=torch.int32, device=dev)
    want_dense = bool(validation) and not only_objectness and nC > 1
    dense = torch.empty(N * cap, nC, dtype=torch.float32, device=dev) if want_dense else None
    anc = torch.tensor([float(a) for a in anchors], dtype=torch.float32).to(dev)
    call('fsdet_region_detect', ptr(output), ptr(anc), N, A, nC, H, W, int(n_models), int(v2), int(bool(only_objectness)),
         float(conf_thresh), ptr(cand), ptr(count), ptr(dense), _stream())
    return Detections(cand, count, dense, N, A, nC, H, W, bool(only_objectness), bool(validation), float(conf_thresh))


def region_detections(output, conf_thresh, num_classes, anchors, num_anchors, only_objectness=1, validation=False,
                      n_models=None):
    """Device-resident form of get_region_boxes (n_models=None) / get_region_boxes_v2: returns `Detections`."""
    if n_models is None:
        return _detect(output, 1, 0, conf_thresh, num_classes, anchors, num_anchors, only_objectness, validation)
    return _detect(output, n_models, 1, conf_thresh, num_classes, anchors, num_anchors, only_objectness, validation)


def get_region_boxes(output, conf_thresh, num_classes, anchors, num_anchors, only_objectness=1, validation=False):
    """utils.py:112-193: list (images) of lists of [x, y, w, h, det_conf, cls_max_conf, cls_max_id(, conf, id)*]."""
    return region_detections(output, conf_thresh, num_classes, anchors, num_anchors, only_objectness, validation).boxes()


def get_region_boxes_v2(output, n_models, conf_thresh, num_classes, anchors, num_anchors, only_objectness=1,
                        validation=False):
    """utils.py:195-290: rows are (image, class) pairs, image-major (`oi = b * n_cls + i`, valid_ensemble.py:158)."""
    return region_detections(output, conf_thresh, num_classes, anchors, num_anchors, only_objectness, validation,
                             n_models=n_models).boxes()


def nms(boxes, nms_thresh):
    """utils.py:85-104.  A row of get_region_boxes(_v2) uses the batched device NMS of its batch (computed once for
    all rows); any other list of boxes is uploaded as float64 and suppressed by the same kernel."""
    if len(boxes) == 0:
        return boxes
    if isinstance(boxes, _Row) and boxes._pristine():
        return boxes._parent._nms_row(boxes._index, boxes, nms_thresh)
    import numpy as np
    import torch
    from ._lib import call, ptr
    if not torch.cuda.is_available():
        raise RuntimeError('nms runs on the GPU only (no CPU fallback)')
    n = len(boxes)
    if n > 4096:
        raise ValueError('nms: %d boxes in one row (max 4096)' % n)
    host = np.array([[float(b[0]), float(b[1]), float(b[2]), float(b[3]), float(b[4])] for b in boxes], dtype=np.float64)
    dev = torch.device('cuda', torch.cuda.current_device())
    b64 = torch.from_numpy(host).to(dev)
    count = torch.tensor([n], dtype=torch.int32, device=dev)
    keep = torch.empty(1, n, dtype=torch.int32, device=dev)
    kc = torch.zeros(1, dtype=torch.int32, device=dev)
    call('fsdet_nms_boxes64', ptr(b64), ptr(count), 1, n, float(nms_thresh), ptr(keep), ptr(kc), _stream())
    slots = keep[0, :int(kc.item())].tolist()
    alive = set(slots)
    for s, box in enumerate(boxes):
        if s not in alive:
            box[4] = 0
    return [boxes[s] for s in slots]
