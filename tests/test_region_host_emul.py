"""csrc/region.cu's decode and build_targets kernels without a GPU (tools/host_emul: the kernel source compiled by
g++, CUDA threads = OS threads) against the fixtures minted from the reference's own region_loss.build_targets:
masks, indices, counters and targets BIT-EXACT - the north star's 'anchor/index paths exactly' checked in the CPU
tier as well (the -m gpu twin through the C ABI is tests/test_gpu_region.py)."""
import ctypes
import os

import numpy as np
import pytest

from emul_util import build_emul

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
NAMES = ['coord_mask', 'conf_mask', 'cls_mask', 'tx', 'ty', 'tw', 'th', 'tconf', 'tcls']


@pytest.fixture(scope='module')
def emul():
    return build_emul('region', 'region.cu')


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def run_build_targets(emul, pred, target, anchors, A, H, W, seen, max_boxes=50, inds=None, live=None, cap=None):
    """inds / live / cap: the CUDA-graph form - `cap` slots are launched, `live` of them (device scalar) are real, slot
    b reads row inds[b] of the FULL label matrix."""
    nB = target.shape[0] if cap is None else cap
    outs = [np.full((nB, A, H, W), 7.0, dtype=np.float32) for _ in NAMES]
    counters = np.full(4, -1, dtype=np.int32)
    pred = np.ascontiguousarray(pred, dtype=np.float32)
    target = np.ascontiguousarray(target, dtype=np.float64)
    anchors = np.ascontiguousarray(anchors, dtype=np.float64)
    emul.emul_build_targets(P(pred), P(target), P(anchors), nB, A, H, W, max_boxes, ctypes.c_float(1.0), ctypes.c_float(5.0),
                            ctypes.c_float(0.6), ctypes.c_longlong(seen), *[P(o) for o in outs], P(counters),
                            P(inds) if inds is not None else None, P(live) if live is not None else None)
    return outs, counters


@pytest.mark.parametrize('tag', ['g13_seen0', 'g13_seen20000', 'g19_seen20000', 'g10_seen12800'])
def test_build_targets_kernel_bit_exact_vs_reference(emul, tag):
    d = np.load(os.path.join(G, 'build_targets_%s.npz' % tag))
    H, W = int(d['nH']), int(d['nW'])
    outs, counters = run_build_targets(emul, d['pred_boxes'], d['target'], d['anchors'], 5, H, W, int(d['seen']))
    assert counters[0] == int(d['nGT']) and counters[1] == int(d['nCorrect']) and counters[2] == 0
    for name, got in zip(NAMES, outs):
        want = d[name]
        if name in ('tw', 'th'):      # float32(log(double)): libm vs the reference's math.log agree to the last bit here
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)) or np.allclose(got, want, rtol=1e-6, atol=0), name
        else:
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), name
    assert int(d['nGT']) > 0


def test_build_targets_edge_rows(emul):
    """Empty rows, a row of max_boxes boxes, a degenerate (zero-size) box."""
    anchors = np.array([1.3221, 1.73145, 3.19275, 4.00944, 5.05587, 8.09892, 9.47112, 4.84053, 11.2364, 10.0071])
    rs = np.random.RandomState(0)
    target = np.zeros((3, 250))
    for t in range(50):
        w, h = rs.uniform(0.05, 0.4, 2)
        target[1, t * 5:(t + 1) * 5] = [1, rs.uniform(w / 2, 0.999 - w / 2), rs.uniform(h / 2, 0.999 - h / 2), w, h]
    target[2, 0:5] = [0, 0.5, 0.5, 0.0, 0.3]          # zero width: the reference raises (log(0)); counted in counters[2]
    pred = np.abs(rs.randn(3 * 5 * 13 * 13, 4)).astype(np.float32) + 0.1
    outs, counters = run_build_targets(emul, pred, target, anchors, 5, 13, 13, 20000)
    assert counters.tolist()[:3] == [51, counters[1], 1]
    assert not outs[0][0].any() and np.all(outs[1][0] == 1.0)           # empty row: no coord mask, conf_mask = noobject
    assert outs[2][1].sum() <= 50 and outs[2][1].sum() > 30             # collisions overwrite, never add


def test_region_decode_kernel_vs_numpy(emul):
    rs = np.random.RandomState(1)
    nB, A, nC, H, W = 3, 5, 1, 13, 13
    out = rs.randn(nB, A * (5 + nC), H, W).astype(np.float32)
    anchors = np.array([1.3221, 1.73145, 3.19275, 4.00944, 5.05587, 8.09892, 9.47112, 4.84053, 11.2364, 10.0071], dtype=np.float32)
    pb = np.zeros((nB * A * H * W, 4), dtype=np.float32)
    emul.emul_region_decode(P(out), None, nB, None, A, nC, H, W, P(anchors), P(pb))
    o = out.reshape(nB, A, 5 + nC, H * W)
    sig = lambda v: 1.0 / (1.0 + np.exp(-v.astype(np.float64)))
    col = np.tile(np.arange(W), H)
    row = np.repeat(np.arange(H), W)
    want = np.stack([sig(o[:, :, 0]) + col, sig(o[:, :, 1]) + row, np.exp(o[:, :, 2].astype(np.float64)) * anchors[0::2][None, :, None],
                     np.exp(o[:, :, 3].astype(np.float64)) * anchors[1::2][None, :, None]], -1).reshape(-1, 4)
    np.testing.assert_allclose(pb, want, rtol=2e-6, atol=1e-6)


def test_fixed_capacity_form_equals_compacted_rows(emul):
    """GraphedTrainStep with neg_filter: the launches are frozen at the capacity of all rows; only `live` slots are
    processed and slot b gathers row inds[b] of the full label matrix.  Result == the eager form on the gathered rows;
    slots beyond `live` are not touched."""
    d = np.load(os.path.join(G, 'build_targets_g13_seen20000.npz'))
    H, W = int(d['nH']), int(d['nW'])
    target = d['target']
    nB = target.shape[0]
    keep = np.array([i for i in range(nB) if i % 3 != 1], dtype=np.int32)
    per = 5 * H * W
    pred_all = d['pred_boxes'].reshape(nB, per, 4)
    pred_kept = np.ascontiguousarray(pred_all[keep])
    want, cw = run_build_targets(emul, pred_kept.reshape(-1, 4), target[keep], d['anchors'], 5, H, W, int(d['seen']))
    # capacity form: pred_boxes compacted in slot order (what region_decode writes), padded to the capacity
    pred_cap = np.zeros((nB, per, 4), dtype=np.float32)
    pred_cap[:len(keep)] = pred_kept
    inds = np.zeros(nB, dtype=np.int32)
    inds[:len(keep)] = keep
    live = np.array([len(keep)], dtype=np.int32)
    got, cg = run_build_targets(emul, pred_cap.reshape(-1, 4), target, d['anchors'], 5, H, W, int(d['seen']), inds=inds, live=live,
                                cap=nB)
    assert cg.tolist() == cw.tolist()
    for name, g, w_ in zip(NAMES, got, want):
        assert np.array_equal(g[:len(keep)].view(np.uint32), w_.view(np.uint32)), name
        assert np.all(g[len(keep):] == 7.0), name
    # decode with a live count: slots beyond it are not written
    rs = np.random.RandomState(3)
    out = rs.randn(nB, 5 * 6, H, W).astype(np.float32)
    anchors32 = d['anchors'].astype(np.float32)
    pb_full = np.full((nB * per, 4), -1.0, dtype=np.float32)
    emul.emul_region_decode(P(out), P(inds), nB, P(live), 5, 1, H, W, P(anchors32), P(pb_full))
    pb_ref = np.full((len(keep) * per, 4), -1.0, dtype=np.float32)
    emul.emul_region_decode(P(out), P(keep), len(keep), None, 5, 1, H, W, P(anchors32), P(pb_ref))
    assert np.array_equal(pb_full[:len(keep) * per], pb_ref)
    assert np.all(pb_full[len(keep) * per:] == -1.0)
