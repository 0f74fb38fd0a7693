"""Detection decode + NMS + reweight ensembling on the GPU (csrc/detect.cu) through the C ABI and the utils / valid
mirrors, against the CPU oracle (oracle/utils.py) and the fixture minted from the reference (tests/golden/detect.npz).

Bars: integer work (the set / order of NMS survivors given identical candidate bytes, the ensembling in float32 with
IEEE division) bit-exact; float32 decode values <= 1e-5 relative (device expf vs torch-CPU expf differ by an ulp), and
the candidate SET equal except for anchor-cells whose confidence is within 1e-5 relative of the threshold."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = ['v2_g13', 'v2_g19', 'v2_g10_obj', 'v1_valid', 'v1_detect', 'v2_empty']


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(G, 'detect.npz'), allow_pickle=False)


def params(gold, tag):
    cs, thr, nC, nA, only_obj, val = gold[tag + '/params']
    v2 = tag.startswith('v2')
    anchors = gold['anchors_voc' if v2 else 'anchors_tiny'].tolist()
    return dict(cs=int(cs), thr=float(thr), nC=int(nC), nA=int(nA), only_obj=int(only_obj), val=bool(val), v2=v2,
                anchors=anchors, nms=float(gold[tag + '/nms_thresh']))


def unflatten(rows, lens, vals):
    out, k, p = [], 0, 0
    for r in rows:
        boxes = []
        for _ in range(int(r)):
            n = int(lens[k])
            boxes.append([float(v) for v in vals[p:p + n]])
            p += n
            k += 1
        out.append(boxes)
    return out


def detections(gold, tag):
    from fewshot_detection_b200 import utils as U
    p = params(gold, tag)
    o = torch.from_numpy(gold[tag + '/output']).cuda()
    return U.region_detections(o, p['thr'], p['nC'], p['anchors'], p['nA'], p['only_obj'], p['val'],
                               n_models=p['cs'] if p['v2'] else None), p


def oracle_arrays(gold, tag, p):
    from oracle import utils as OU
    o = torch.from_numpy(gold[tag + '/output'])
    xs, ys, ws, hs, det, cmax, cid, cls = OU.region_arrays(o, p['nC'], p['anchors'], p['nA'], p['cs'] if p['v2'] else None)
    return [t.numpy() for t in (xs, ys, ws, hs, det, cmax)] + [cid.numpy(), cls.numpy()]


@pytest.mark.parametrize('tag', CASES)
def test_region_detect_candidates_vs_oracle(gold, tag):
    d, p = detections(gold, tag)
    torch.cuda.synchronize()
    N, _, H, W = gold[tag + '/output'].shape
    K = p['nA'] * H * W
    xs, ys, ws, hs, det, cmax, cid, cls = oracle_arrays(gold, tag, p)
    conf = det.astype(np.float64) if p['only_obj'] else det.astype(np.float64) * cmax.astype(np.float64)
    count = d.count.cpu().numpy()
    cand = d.cand.cpu().numpy()
    for n in range(N):
        c = cand[n, :count[n]]
        ints = c[:, 6:8].copy().view(np.int32)
        ind = ints[:, 1].astype(np.int64)                       # a*HW + cell
        # the reference's loop order (cy, cx, a): cell-major, anchor-minor
        a, cell = ind // (H * W), ind % (H * W)
        order_key = cell * p['nA'] + a
        assert np.all(np.diff(order_key) > 0), 'candidates not in (cy, cx, anchor) order'
        got = set((n * K + ind).tolist())
        want = set((n * K + np.nonzero(conf[n * K:(n + 1) * K] > p['thr'])[0]).tolist())
        for i in got ^ want:                                    # only borderline cells may differ
            assert abs(conf[i] - p['thr']) <= 1e-5 * p['thr'], (n, i, conf[i])
        g = n * K + ind
        for k, ref in enumerate((xs, ys, ws, hs, det, cmax)):
            np.testing.assert_allclose(c[:, k], ref[g], rtol=1e-5, atol=1e-7)
        assert np.array_equal(ints[:, 0], cid[g].astype(np.int32))
    # counts equal the reference's own (fixture) unless a borderline cell flipped
    assert np.abs(count.astype(np.int64) - gold[tag + '/rows']).max(initial=0) <= 1
    if d.cls_dense is not None:
        np.testing.assert_allclose(d.cls_dense.cpu().numpy(), cls.reshape(-1, p['nC']), rtol=1e-5, atol=1e-8)


def _cand_from_oracle(gold, tag, p):
    """Candidate arrays with the ORACLE's float32 bytes, in the reference's loop order (what fsdet_region_detect
    would write if its transcendental functions were bit-identical to torch-CPU's)."""
    N, _, H, W = gold[tag + '/output'].shape
    K = p['nA'] * H * W
    xs, ys, ws, hs, det, cmax, cid, _ = oracle_arrays(gold, tag, p)
    conf = det.astype(np.float64) if p['only_obj'] else det.astype(np.float64) * cmax.astype(np.float64)
    cand = np.zeros((N, K, 8), dtype=np.float32)
    count = np.zeros(N, dtype=np.int32)
    boxes = []
    for n in range(N):
        row = []
        for cell in range(H * W):
            for a in range(p['nA']):
                i = n * K + a * H * W + cell
                if conf[i] > p['thr']:
                    s = count[n]
                    cand[n, s, :6] = [xs[i], ys[i], ws[i], hs[i], det[i], cmax[i]]
                    cand[n, s, 6:8] = np.array([cid[i], a * H * W + cell], dtype=np.int32).view(np.float32)
                    count[n] += 1
                    row.append([float(xs[i]) / W, float(ys[i]) / H, float(ws[i]) / W, float(hs[i]) / H, float(det[i]),
                                float(cmax[i]), int(cid[i])])
        boxes.append(row)
    return cand, count, boxes


@pytest.mark.parametrize('tag', CASES)
def test_nms_bit_exact_given_identical_candidates(gold, tag):
    from fewshot_detection_b200._lib import call, ptr
    from oracle import utils as OU
    p = params(gold, tag)
    N, _, H, W = gold[tag + '/output'].shape
    cand, count, boxes = _cand_from_oracle(gold, tag, p)
    K = cand.shape[1]
    dc, dn = torch.from_numpy(cand).cuda(), torch.from_numpy(count).cuda()
    keep = torch.full((N, K), -1, dtype=torch.int32, device='cuda')
    kc = torch.full((N,), -1, dtype=torch.int32, device='cuda')
    call('fsdet_nms', ptr(dc), ptr(dn), N, K, H, W, p['nms'], ptr(keep), ptr(kc), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    keep, kc = keep.cpu().numpy(), kc.cpu().numpy()
    for n in range(N):
        row = [list(b) + [s] for s, b in enumerate(boxes[n])]   # tag every box with its slot
        want = [b[-1] for b in OU.nms(row, p['nms'])]
        assert kc[n] == len(want), (n, kc[n], len(want))
        assert keep[n, :kc[n]].tolist() == want                  # same survivors, same order
    assert np.array_equal(kc, gold[tag + '/nms_rows'])           # and as many as the reference itself kept


@pytest.mark.parametrize('tag', ['v2_g13', 'v1_valid', 'v1_detect'])
def test_list_form_nms_bit_exact_vs_reference_fixture(gold, tag):
    """utils.nms on plain Python lists (the reference's own box values) -> float64 upload -> same kernel."""
    from fewshot_detection_b200 import utils as U
    rows = unflatten(gold[tag + '/rows'], gold[tag + '/len'], gold[tag + '/vals'])
    want = unflatten(gold[tag + '/nms_rows'], gold[tag + '/nms_len'], gold[tag + '/nms_vals'])
    for row, w in zip(rows[:3], want[:3]):
        got = U.nms(row, float(gold[tag + '/nms_thresh']))
        assert got == w
        assert sum(1 for b in row if b[4] == 0) == len(row) - len(w)   # suppressed in place, like the reference


def _match(got, want, tol=1e-5):
    """Number of `want` boxes without a counterpart in `got` (first 7 entries within tol)."""
    if not want:
        return len(got)
    g = np.array([b[:7] for b in got], dtype=np.float64).reshape(-1, 7)
    miss = 0
    for b in want:
        w = np.array(b[:7])
        if not len(g) or np.abs(g - w).max(axis=1).min() > tol * max(1.0, np.abs(w).max()):
            miss += 1
    return miss


@pytest.mark.parametrize('tag', CASES)
def test_drop_in_calls_vs_reference_fixture(gold, tag):
    """The reference's call sequence (valid_ensemble.py:145-162): boxes = get_region_boxes_v2(...);
    nms(boxes[oi], nms_thresh) per row - against the reference's own outputs."""
    from fewshot_detection_b200 import utils as U
    p = params(gold, tag)
    o = torch.from_numpy(gold[tag + '/output']).cuda()
    if p['v2']:
        boxes = U.get_region_boxes_v2(o, p['cs'], p['thr'], p['nC'], p['anchors'], p['nA'], p['only_obj'], p['val'])
    else:
        boxes = U.get_region_boxes(o, p['thr'], p['nC'], p['anchors'], p['nA'], p['only_obj'], p['val'])
    want_all = unflatten(gold[tag + '/rows'], gold[tag + '/len'], gold[tag + '/vals'])
    want_kept = unflatten(gold[tag + '/nms_rows'], gold[tag + '/nms_len'], gold[tag + '/nms_vals'])
    assert len(boxes) == len(want_all)
    for n, (row, wa, wk) in enumerate(zip(boxes, want_all, want_kept)):
        assert abs(len(row) - len(wa)) <= 1
        if len(row) == len(wa):
            assert [len(b) for b in row] == [len(b) for b in wa]      # incl. the validation=(conf, id) extras
            for b, w in zip(row, wa):
                np.testing.assert_allclose(b, w, rtol=1e-5, atol=1e-7)
        kept = U.nms(row, p['nms'])
        assert all(b[4] > 0 for b in kept)
        assert all(kept[i][4] >= kept[i + 1][4] - 1e-6 for i in range(len(kept) - 1))
        # device expf vs CPU expf can reorder near-ties / flip an IoU at the threshold: allow 1 % of the row
        slack = max(1, len(wk) // 100)
        assert abs(len(kept) - len(wk)) <= slack, (n, len(kept), len(wk))
        assert _match(kept, wk) <= slack and _match(wk, kept) <= slack


def test_detections_kept_boxes_equals_rowwise_nms(gold):
    d, p = detections(gold, 'v2_g13')
    kept = d.kept_boxes(p['nms'])
    from fewshot_detection_b200 import utils as U
    rows = d.boxes()
    for n in range(d.N):
        assert kept[n] == U.nms(rows[n], p['nms'])


def test_rw_running_mean_bit_exact(gold):
    from fewshot_detection_b200.valid import ReweightEnsembler
    n_cls = int(gold['ens/n_cls'])
    ens = ReweightEnsembler(n_cls, gold['ens/dw0'].shape[1], torch.device('cuda'))
    for k in range(3):
        ens.update(torch.from_numpy(gold['ens/dw%d' % k]).cuda(), gold['ens/ids%d' % k].tolist())
    got = ens.result()[0]
    assert tuple(got.shape) == (n_cls, gold['ens/dw0'].shape[1], 1, 1)
    assert np.array_equal(got.view(n_cls, -1).cpu().numpy().view(np.uint32), gold['ens/result'].view(np.uint32))
    assert ens.counts.tolist() == [sum(int(c == i) for k in range(3) for c in gold['ens/ids%d' % k]) for i in range(n_cls)]


def test_full_size_properties():
    """BASELINE-size batch (64 images x 20 classes, 13x13): properties that need no oracle - survivors are unique
    candidate slots, sorted by confidence, pairwise IoU <= thresh; NMS is idempotent."""
    from fewshot_detection_b200 import utils as U, netcfg
    from oracle.region_loss import bbox_iou
    anchors = [float(a) for a in netcfg.VOC_ANCHORS.split(',')]
    g = torch.Generator().manual_seed(5)
    bs, cs = 64, 20
    o = torch.randn(bs * cs, 30, 13, 13, generator=g)
    o.view(bs * cs, 5, 6, 13, 13)[:, :, 4] -= 2.0
    d = U.region_detections(o.cuda(), 0.005, 1, anchors, 5, 0, 1, n_models=cs).nms(0.45)
    torch.cuda.synchronize()
    count, kc = d.count.cpu().numpy(), d.keep_count.cpu().numpy()
    assert count.min() > 0 and np.all(kc <= count) and kc.min() > 0
    keep, cand = d.keep.cpu().numpy(), d.cand.cpu().numpy()
    for n in (0, 777, bs * cs - 1):
        slots = keep[n, :kc[n]]
        assert len(set(slots.tolist())) == kc[n] and slots.max() < count[n]
        det = cand[n, slots, 4]
        assert np.all(np.diff((1 - det.astype(np.float64)).astype(np.float32)) >= 0)
        bx = (cand[n, slots, :4].astype(np.float64) / 13.0).tolist()
        for i in range(0, len(bx), 7):
            for j in range(i + 1, len(bx), 5):
                assert bbox_iou(bx[i], bx[j]) <= 0.45
    # idempotence: NMS of the survivors keeps all of them
    kept = d.kept_boxes(0.45)
    again = U.nms([list(b) for b in kept[3]], 0.45)
    assert again == kept[3]


def test_valid_detect_end_to_end_mini_model(tmp_path):
    """valid.valid_batches on the mini meta model: ensembling -> detect_forward (eval) -> decode -> NMS -> files,
    against the oracle fed with the SAME head output (model-level parity is test_gpu_model's job)."""
    from fewshot_detection_b200 import netcfg, valid
    from fewshot_detection_b200.darknet_meta import Darknet
    from oracle import utils as OU
    from seeding import seeded_init, synth_masks
    det, ler = netcfg.mini_dynamic_blocks(128, 4), netcfg.mini_reweighting_blocks(64, 4, 128)
    m = Darknet([dict(b) for b in det], [dict(b) for b in ler])
    seeded_init(m, 3)
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(9)
    n_cls = 3
    ids = [[0, 1, 2, 0], [2, 1]]
    meta = [(torch.rand(len(i), 3, 64, 64, generator=g), torch.from_numpy(synth_masks(len(i), 64, 7 + k)), i)
            for k, i in enumerate(ids)]
    data = torch.rand(2, 3, 128, 128, generator=g)
    names = ['aeroplane', 'bird', 'cat']
    dw = valid.valid_batches(m, meta, [(data, ['000001', '000002'], [(500, 375), (320, 240)])], names,
                             str(tmp_path / 'res'), 'comp4_det_test_')
    # ensembling vs the oracle's running mean of the same vectors
    with torch.no_grad():
        vecs = [(m.meta_forward(mx.cuda(), mk.cuda())[0].view(len(i), -1).cpu(), i) for mx, mk, i in meta]
    want_dw = OU.ensemble_reweights(vecs, n_cls)
    np.testing.assert_allclose(dw[0].view(n_cls, -1).cpu().numpy(), want_dw.numpy(), rtol=1e-6, atol=1e-9)
    with torch.no_grad():
        out = m.detect_forward(data.cuda(), dw).cpu()
    boxes = OU.get_region_boxes_v2(out, n_cls, 0.005, m.num_classes, m.anchors, m.num_anchors, 0, 1)
    for i, name in enumerate(names):
        got = open(str(tmp_path / 'res' / ('comp4_det_test_%s.txt' % name))).read().splitlines()
        want = []
        for b, (imgid, (w, h)) in enumerate(zip(['000001', '000002'], [(500, 375), (320, 240)])):
            want += OU.detection_lines(OU.nms(boxes[b * n_cls + i], 0.45), imgid, w, h)
        want = [l.rstrip('\n') for l in want]
        assert abs(len(got) - len(want)) <= max(1, len(want) // 100)
        if len(got) == len(want):
            for a, b in zip(got, want):
                fa, fb = a.split(), b.split()
                assert fa[0] == fb[0]
                np.testing.assert_allclose([float(v) for v in fa[1:]], [float(v) for v in fb[1:]], rtol=1e-4, atol=2e-3)
