#!/usr/bin/env python
"""Mint the golden fixtures in tests/golden/*.npz from the REFERENCE ITSELF.

Run in the build container only (needs /root/reference; never at test time):

    python tests/golden/make_golden.py

The reference is Python-2.7 / torch-0.3.1 code (README.md:9,
requirements.txt:3).  It is imported from where it lies, read-only, and made
runnable under Python 3.12 / torch 2.11 by *mechanical* textual substitutions
applied in memory (the table PATCHES below: integer division, xrange, the
removed `torch.cuda.*Tensor` constructors, `.data[0]`, `size_average`, mask
indexing by shape).  No arithmetic and no operation order is touched.

One semantic difference between torch 0.3.1 and torch >= 0.4 matters for
`build_targets` (region_loss.py:37-132): in 0.3.1 indexing a tensor down to one
element returns a *Python float* (so all of phase 2 runs in float64 on values
promoted exactly from the stored dtype), whereas torch >= 0.4 returns 0-dim
tensors and would do part of that arithmetic in float32.  `_Legacy2D` below
re-creates the 0.3.1 indexing behaviour around the inputs so that the patched
reference computes what the original computed.

Nothing from /root/reference is copied into this repository: only the numeric
inputs/outputs land in the .npz files.
"""
import importlib
import io
import os
import random
import re
import sys
import types
from contextlib import redirect_stdout

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, '_shims'))
sys.path.insert(0, REF)
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

from seeding import seeded_init, synth_targets, synth_masks  # noqa: E402
from fewshot_detection_b200 import netcfg  # noqa: E402

COMMON = [
    (r'\bxrange\b', 'range'),
    (r'torch\.cuda\.LongTensor', 'torch.LongTensor'),
    (r'torch\.cuda\.FloatTensor', 'torch.FloatTensor'),
    (r'\.cuda\(\)', ''),
    (r'\.data\[0\]', '.item()'),
    (r'size_average=False', "reduction='sum'"),
    (r'size_average=True', "reduction='mean'"),
]
PATCHES = {
    'region_loss': COMMON + [
        (r'len\(anchors\)/num_anchors', 'len(anchors)//num_anchors'),
        (r'conf_mask\[b\]\[cur_ious>sil_thresh\] = 0', 'conf_mask[b].view(-1)[cur_ious>sil_thresh] = 0'),
        (r'tcls\.view\(-1\)\[cls_mask\]', 'tcls.view(-1)[cls_mask.view(-1)]'),
        # torch 0.3.1 added tensors of equal numel but different shape element-wise
        (r'x\.data \+ grid_x', 'x.data.view(-1) + grid_x'),
        (r'y\.data \+ grid_y', 'y.data.view(-1) + grid_y'),
        (r'torch\.exp\(w\.data\) \* anchor_w', 'torch.exp(w.data).view(-1) * anchor_w'),
        (r'torch\.exp\(h\.data\) \* anchor_h', 'torch.exp(h.data).view(-1) * anchor_h'),
    ],
    'darknet_meta': COMMON + [
        (r'\(kernel_size-1\)/2', '(kernel_size-1)//2'),
        (r'H/hs\*W/ws', '(H//hs)*(W//ws)'),
        (r'H/hs', 'H//hs'), (r'W/ws', 'W//ws'),
        (r'len\(loss\.anchors\)/loss\.num_anchors', 'len(loss.anchors)//loss.num_anchors'),
    ],
    'darknet': COMMON + [
        (r'\(kernel_size-1\)/2', '(kernel_size-1)//2'),
        (r'H/hs\*W/ws', '(H//hs)*(W//ws)'),
        (r'H/hs', 'H//hs'), (r'W/ws', 'W//ws'),
        (r'len\(loss\.anchors\)/loss\.num_anchors', 'len(loss.anchors)//loss.num_anchors'),
    ],
}


def load_ref(name):
    """exec the reference module `name`.py with PATCHES[name] applied in memory."""
    if 'imghdr' not in sys.modules:
        try:
            import imghdr  # noqa: F401  (utils.py imports it; removed in py3.13)
        except Exception:
            sys.modules['imghdr'] = types.ModuleType('imghdr')
    src = open(os.path.join(REF, name + '.py')).read()
    for pat, rep in PATCHES[name]:
        src = re.sub(pat, rep, src)
    mod = types.ModuleType(name)
    mod.__file__ = os.path.join(REF, name + '.py')
    sys.modules[name] = mod
    with redirect_stdout(io.StringIO()):
        exec(compile(src, mod.__file__, 'exec'), mod.__dict__)
    return mod


class _Legacy2D(object):
    """torch-0.3.1 indexing semantics for a 2-D tensor: t[i] with an int gives a
    row whose elements are Python floats; slices give tensors."""

    def __init__(self, t):
        self.t = t
        self.rows = t.tolist()  # exact promotion to Python float (f64)

    def size(self, d):
        return self.t.size(d)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return self.t[i]
        return self.rows[i]


def main():
    out = HERE
    with redirect_stdout(io.StringIO()):
        ref_utils = importlib.import_module('utils')
        ref_cfg = importlib.import_module('cfg')
    RL = load_ref('region_loss')
    DM = load_ref('darknet_meta')
    DK = load_ref('darknet')
    cfg = ref_cfg.cfg

    # ------------------------------------------------------------ G1: IoU
    rs = np.random.RandomState(1)
    n = 256
    b1 = np.stack([rs.uniform(0, 13, n), rs.uniform(0, 13, n), rs.uniform(0.01, 9, n), rs.uniform(0.01, 9, n)]).astype(np.float32)
    b2 = np.stack([rs.uniform(0, 13, n), rs.uniform(0, 13, n), rs.uniform(0.01, 9, n), rs.uniform(0.01, 9, n)]).astype(np.float32)
    b2[:, :16] = b1[:, :16]                       # identical boxes
    b2[0, 16:32] = b1[0, 16:32] + 40.0            # disjoint boxes
    ious_vec = ref_utils.bbox_ious(torch.from_numpy(b1), torch.from_numpy(b2), x1y1x2y2=False).numpy()
    ious_sc = np.array([ref_utils.bbox_iou([float(v) for v in b1[:, i]], [float(v) for v in b2[:, i]], x1y1x2y2=False)
                        for i in range(n)], dtype=np.float64)
    np.savez_compressed(os.path.join(out, 'iou.npz'), b1=b1, b2=b2, ious_f32=ious_vec, ious_f64=ious_sc)

    # ---------------------------------------------------- G2: build_targets
    anchors = [float(a) for a in netcfg.VOC_ANCHORS.split(',')]

    def run_bt(pred, target, nA, nH, nW, seen):
        r = RL.build_targets(_Legacy2D(pred), _Legacy2D(target), anchors, nA, 1, nH, nW, 1.0, 5.0, 0.6, seen)
        names = ['nGT', 'nCorrect', 'coord_mask', 'conf_mask', 'cls_mask', 'tx', 'ty', 'tw', 'th', 'tconf', 'tcls']
        return {k: (np.int64(v) if isinstance(v, int) else v.numpy()) for k, v in zip(names, r)}

    def synth_pred(nB, nA, nH, nW, seed, target=None):
        g = torch.Generator().manual_seed(seed)
        n = nB * nA * nH * nW
        o = torch.randn(n, 4, generator=g)
        gx = torch.arange(nW).repeat(nH, 1).repeat(nB * nA, 1, 1).view(n).float()
        gy = torch.arange(nH).repeat(nW, 1).t().repeat(nB * nA, 1, 1).view(n).float()
        aw = torch.tensor(anchors[0::2]).view(1, nA, 1).repeat(nB, 1, nH * nW).view(n)
        ah = torch.tensor(anchors[1::2]).view(1, nA, 1).repeat(nB, 1, nH * nW).view(n)
        pb = torch.stack([torch.sigmoid(o[:, 0]) + gx, torch.sigmoid(o[:, 1]) + gy,
                          torch.exp(o[:, 2] * 0.4) * aw, torch.exp(o[:, 3] * 0.4) * ah], 1).contiguous()
        if target is not None:
            # plant near-perfect predictions on ~half of the GT cells so that the
            # IoU>0.6 silencing and the IoU>0.5 recall branches are both exercised
            rs = np.random.RandomState(seed)
            for b in range(nB):
                for t in range(50):
                    if target[b, t * 5 + 1] == 0:
                        break
                    if rs.rand() < 0.5:
                        x, y, w, h = [float(v) for v in target[b, t * 5 + 1:t * 5 + 5]]
                        gi, gj = int(x * nW), int(y * nH)
                        for a in range(nA):
                            i = b * nA * nH * nW + a * nH * nW + gj * nW + gi
                            pb[i] = torch.tensor([x * nW, y * nH, w * nW * rs.uniform(0.8, 1.2), h * nH * rs.uniform(0.8, 1.2)])
        return pb

    cases = {}
    for tag, (bs, cs, G, seen, seed) in {
            'g13_seen0': (3, 4, 13, 0, 11), 'g13_seen20000': (3, 4, 13, 20000, 12),
            'g19_seen20000': (2, 3, 19, 20000, 13), 'g10_seen12800': (2, 3, 10, 12800, 14)}.items():
        tgt = synth_targets(bs, cs, seed, max_gt=6).reshape(bs * cs, 250)
        if tag == 'g13_seen20000':
            # collision: two GTs of one row in the same cell with the same best anchor
            tgt[1, 0:5] = [1, 0.52, 0.52, 0.30, 0.40]
            tgt[1, 5:10] = [1, 0.53, 0.51, 0.31, 0.41]
            tgt[1, 10:] = 0
            # a full row of 50 boxes
            rs2 = np.random.RandomState(99)
            for t in range(50):
                w, h = rs2.uniform(0.05, 0.5, 2)
                tgt[2, t * 5:(t + 1) * 5] = [2, rs2.uniform(w / 2, 0.999 - w / 2), rs2.uniform(h / 2, 0.999 - h / 2), w, h]
        pred = synth_pred(bs * cs, 5, G, G, seed, tgt)
        r = run_bt(pred, torch.from_numpy(tgt), 5, G, G, seen)
        r.update(pred_boxes=pred.numpy(), target=tgt, anchors=np.array(anchors), nH=G, nW=G, seen=seen)
        cases[tag] = r
        np.savez_compressed(os.path.join(out, 'build_targets_%s.npz' % tag), **r)

    # ------------------------------------------------- G3: RegionLossV2 / RegionLoss
    def run_loss_v2(bs, cs, G, seen, seed, neg_ratio, pyseed=0, empty_prob=0.3):
        cfg.neg_ratio = neg_ratio
        tgt = synth_targets(bs, cs, seed, max_gt=5, empty_prob=empty_prob)
        g = torch.Generator().manual_seed(seed)
        o = (torch.randn(bs * cs, 30, G, G, generator=g) * 0.7).requires_grad_(True)
        L = RL.RegionLossV2()
        L.anchors, L.num_anchors, L.anchor_step, L.num_classes = anchors, 5, 2, 1
        L.object_scale, L.noobject_scale, L.class_scale, L.coord_scale = 5.0, 1.0, 1.0, 1.0
        L.seen = seen
        orig_bt = RL.build_targets
        RL.build_targets = lambda pb, tg, *a: orig_bt(_Legacy2D(pb), _Legacy2D(tg), *a)
        random.seed(pyseed)
        buf = io.StringIO()
        try:
            with redirect_stdout(buf):
                loss = L(o, torch.from_numpy(tgt))
        finally:
            RL.build_targets = orig_bt
        loss.backward()
        # replay neg_filter's draws to record which rows were kept
        random.seed(pyseed)
        _, _, inds = RL.neg_filter(o.detach(), torch.from_numpy(tgt).view(-1, 250), withids=True)
        line = buf.getvalue().strip().splitlines()[-1]
        return dict(output=o.detach().numpy(), target=tgt, loss=np.float64(loss.item()), grad=o.grad.numpy(),
                    inds=np.atleast_1d(np.asarray(inds)).astype(np.int64), seen=seen, pyseed=pyseed,
                    neg_ratio=str(neg_ratio), log_line=line, anchors=np.array(anchors))

    np.savez_compressed(os.path.join(out, 'region_loss_v2_full.npz'), **run_loss_v2(3, 4, 13, 20000, 21, 'full'))
    np.savez_compressed(os.path.join(out, 'region_loss_v2_full_warm.npz'), **run_loss_v2(2, 3, 13, 64, 22, 'full'))
    np.savez_compressed(os.path.join(out, 'region_loss_v2_neg1.npz'), **run_loss_v2(4, 5, 13, 20000, 23, 1, pyseed=7))
    np.savez_compressed(os.path.join(out, 'region_loss_v2_neg0.npz'), **run_loss_v2(4, 5, 10, 20000, 24, 0, pyseed=8))
    cfg.neg_ratio = 'full'

    def run_loss_plain(bs, G, seen, seed):
        nC = 20
        tanch = [float(a) for a in netcfg.TINY_VOC_ANCHORS.split(',')]
        tgt = synth_targets(bs, 1, seed, max_gt=5)[:, 0, :]
        rs = np.random.RandomState(seed)
        for b in range(bs):
            for t in range(50):
                if tgt[b, t * 5 + 1] == 0:
                    break
                tgt[b, t * 5] = rs.randint(0, nC)
        g = torch.Generator().manual_seed(seed)
        o = (torch.randn(bs, 5 * (5 + nC), G, G, generator=g) * 0.7).requires_grad_(True)
        res = {}
        for my in (True, False):
            cfg.metayolo = my
            L = RL.RegionLoss()
            L.anchors, L.num_anchors, L.anchor_step, L.num_classes = tanch, 5, 2, nC
            L.object_scale, L.noobject_scale, L.class_scale, L.coord_scale = 5.0, 1.0, 1.0, 1.0
            L.seen = seen
            orig_bt = RL.build_targets
            RL.build_targets = lambda pb, tg, *a: orig_bt(_Legacy2D(pb), _Legacy2D(tg), *a)
            buf = io.StringIO()
            try:
                with redirect_stdout(buf):
                    loss = L(o, torch.from_numpy(tgt))
            finally:
                RL.build_targets = orig_bt
            o.grad = None
            loss.backward()
            k = 'metayolo1' if my else 'metayolo0'
            res['loss_' + k] = np.float64(loss.item())
            res['grad_' + k] = o.grad.numpy().copy()
            res['log_' + k] = buf.getvalue().strip().splitlines()[-1]
        cfg.metayolo = True
        res.update(output=o.detach().numpy(), target=tgt, seen=seen, anchors=np.array(tanch))
        return res

    np.savez_compressed(os.path.join(out, 'region_loss_plain.npz'), **run_loss_plain(2, 13, 20000, 31))

    # ------------------------------------------------------------ G4: models
    # (a) layer primitives
    g = torch.Generator().manual_seed(41)
    x = torch.randn(2, 6, 8, 8, generator=g)
    reorg = DM.Reorg(2)(x).numpy()
    mp1 = DM.MaxPoolStride1()(x).numpy()
    from pooling import GlobalMaxPool2d
    gmax = GlobalMaxPool2d()(x).numpy()
    dyn = DM.dynamic_conv2d(True, None)(6, 6, 1, 1, 0, bias=False)
    dw = torch.randn(3, 6, 1, 1, generator=g)
    dyn_out = dyn((x, dw)).detach().numpy()
    np.savez_compressed(os.path.join(out, 'layers.npz'), x=x.numpy(), reorg=reorg, maxpool_stride1=mp1,
                        globalmax=gmax, dyn_w=dw.numpy(), dyn_out=dyn_out)

    # (b) mini meta model: full forward + RegionLossV2 + backward, every tensor
    def run_meta(det_blocks, ler_blocks, bs, cs, side, meta_side, seed, seen, full_grads):
        cfg.neg_ratio = 'full'
        with redirect_stdout(io.StringIO()):
            m = DM.Darknet([dict(b) for b in det_blocks], [dict(b) for b in ler_blocks])
        seeded_init(m, seed)
        m.train()
        g = torch.Generator().manual_seed(seed + 1)
        x = torch.rand(bs, 3, side, side, generator=g)
        metax = torch.rand(cs, 3, meta_side, meta_side, generator=g)
        mask = torch.from_numpy(synth_masks(cs, meta_side, seed + 2))
        tgt = synth_targets(bs, cs, seed + 3, max_gt=4)
        out_t = m(x, metax, mask)
        L = m.models[len(m.models) - 1]
        L.seen = seen
        orig_bt = RL.build_targets
        RL.build_targets = lambda pb, tg, *a: orig_bt(_Legacy2D(pb), _Legacy2D(tg), *a)
        buf = io.StringIO()
        try:
            with redirect_stdout(buf):
                loss = L(out_t, torch.from_numpy(tgt))
        finally:
            RL.build_targets = orig_bt
        loss.backward()
        r = dict(target=tgt, output=out_t.detach().numpy(), bs=bs, cs=cs, side=side, meta_side=meta_side,
                 loss=np.float64(loss.item()), seed=seed, seen=seen, log_line=buf.getvalue().strip().splitlines()[-1])
        if full_grads:  # inputs are regenerable from `seed` (see tests/golden/seeding.py users); keep a copy for the small case
            r.update(x=x.numpy(), metax=metax.numpy(), mask=mask.numpy())
        with torch.no_grad():
            dws = m.meta_forward(metax, mask)
        r['dynamic_weights_2nd_pass'] = dws[0].numpy()
        for name, p in m.named_parameters():
            if full_grads:
                r['grad/' + name] = p.grad.numpy()
            else:
                r['gradnorm/' + name] = np.float64(p.grad.double().norm().item())
                r['gradhead/' + name] = p.grad.reshape(-1)[:64].numpy().copy()
        for name, b in m.named_buffers():
            if 'running' in name and (full_grads or b.numel() <= 64):
                r['buf/' + name] = b.numpy().copy()
        return r

    np.savez_compressed(os.path.join(out, 'meta_mini.npz'),
                        **run_meta(netcfg.mini_dynamic_blocks(128, 4), netcfg.mini_reweighting_blocks(64, 4, 128),
                                   bs=2, cs=3, side=128, meta_side=64, seed=51, seen=20000, full_grads=True))
    # (c) the real architectures at 416: output + gradient digests only
    np.savez_compressed(os.path.join(out, 'meta_full416.npz'),
                        **run_meta(netcfg.darknet_dynamic_blocks(), netcfg.reweighting_net_blocks(),
                                   bs=1, cs=2, side=416, meta_side=416, seed=61, seen=20000, full_grads=False))

    # (d) BASELINE config #1: tiny-yolo-voc forward, one 416x416 image (needs a cfg *file*)
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, 'tiny.cfg')
        netcfg.write_cfg(netcfg.tiny_yolo_voc_blocks(), p)
        # round-trip check of the serialiser against the reference's own cfg file
        assert ref_cfg.parse_cfg(p)[1:] == ref_cfg.parse_cfg(os.path.join(REF, 'cfg/tiny-yolo-voc.cfg'))[1:]
        with redirect_stdout(io.StringIO()):
            tm = DK.Darknet(p)
        p2 = os.path.join(td, 'dyn.cfg')
        netcfg.write_cfg(netcfg.darknet_dynamic_blocks(), p2)
        assert ref_cfg.parse_cfg(p2)[1:] == ref_cfg.parse_cfg(os.path.join(REF, 'cfg/darknet_dynamic.cfg'))[1:]
        p3 = os.path.join(td, 'rw.cfg')
        netcfg.write_cfg(netcfg.reweighting_net_blocks(), p3)
        assert ref_cfg.parse_cfg(p3)[1:] == ref_cfg.parse_cfg(os.path.join(REF, 'cfg/reweighting_net.cfg'))[1:]
    seeded_init(tm, 71)
    g = torch.Generator().manual_seed(72)
    x = torch.rand(1, 3, 416, 416, generator=g)
    tm.eval()
    with torch.no_grad():
        y_eval = tm(x).numpy()
    tm.train()
    y_train = tm(x).detach().numpy()
    np.savez_compressed(os.path.join(out, 'tiny_yolo_416.npz'), x_seed=72, w_seed=71, y_eval=y_eval, y_train=y_train)
    print('golden fixtures written to', out)
    for f in sorted(os.listdir(out)):
        if f.endswith('.npz'):
            print('  %-40s %8.1f KB' % (f, os.path.getsize(os.path.join(out, f)) / 1024))


if __name__ == '__main__':
    main()
