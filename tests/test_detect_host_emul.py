"""Block-level LOGIC of csrc/detect.cu without a GPU: the kernel source itself is compiled by g++ against
tools/host_emul/cuda_host_emul.h (CUDA threads = OS threads, __syncthreads = pthread barrier, warp ballots = warp
barrier) and run on the fixtures.  Covers what a value-level oracle cannot see from the host side: the ordered
block compaction, the bitonic sort, the barrier placement of the greedy suppression loop, indexing.
Device arithmetic (expf) is libm's here, so float bars are the GPU tests' bars; the parity tests proper are
tests/test_gpu_detect.py (-m gpu, through the C ABI)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from test_gpu_detect import CASES, params, oracle_arrays, _cand_from_oracle, unflatten

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, 'tests', 'golden')
SRC = os.path.join(ROOT, 'tools', 'host_emul', 'detect_emul.cpp')
HDR = os.path.join(ROOT, 'tools', 'host_emul', 'cuda_host_emul.h')
KSRC = os.path.join(ROOT, 'fewshot_detection_b200', 'csrc', 'detect.cu')
LIB = os.path.join(ROOT, 'build', 'libdetect_emul.so')


def _cuda_include():
    for d in ('/usr/local/cuda/include', os.path.join(os.environ.get('CUDA_HOME', ''), 'include')):
        if os.path.exists(os.path.join(d, 'cuda_runtime.h')):
            return d
    return None


@pytest.fixture(scope='module')
def emul():
    inc = _cuda_include()
    if inc is None:
        pytest.skip('CUDA headers not found')
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(p) for p in (SRC, HDR, KSRC)):
        cmd = ['g++', '-O1', '-std=c++17', '-ffp-contract=off', '-fPIC', '-shared', '-pthread', '-w',
               '-DFSDET_HOST_EMULATION', '-I' + inc, '-include', HDR, '-x', 'c++', SRC, '-o', LIB]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout
    return ctypes.CDLL(LIB)


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(G, 'detect.npz'), allow_pickle=False)


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def run_detect(emul, out, p):
    N, _, H, W = out.shape
    K = p['nA'] * H * W
    cand = np.zeros((N, K, 8), dtype=np.float32)
    count = np.full(N, -1, dtype=np.int32)
    dense = np.zeros((N * K, p['nC']), dtype=np.float32) if (p['val'] and not p['only_obj'] and p['nC'] > 1) else None
    anc = np.array(p['anchors'], dtype=np.float32)
    out = np.ascontiguousarray(out, dtype=np.float32)
    emul.emul_region_detect(P(out), P(anc), N, p['nA'], p['nC'], H, W, p['cs'] if p['v2'] else 1, int(p['v2']),
                            p['only_obj'], ctypes.c_double(p['thr']), P(cand), P(count), P(dense))
    return cand, count, dense


@pytest.mark.parametrize('tag', CASES)
def test_emulated_region_detect(emul, gold, tag):
    p = params(gold, tag)
    out = gold[tag + '/output']
    N, _, H, W = out.shape
    K = p['nA'] * H * W
    cand, count, dense = run_detect(emul, out, p)
    xs, ys, ws, hs, det, cmax, cid, cls = oracle_arrays(gold, tag, p)
    conf = det.astype(np.float64) if p['only_obj'] else det.astype(np.float64) * cmax.astype(np.float64)
    for n in range(N):
        c = cand[n, :count[n]]
        ints = c[:, 6:8].copy().view(np.int32)
        ind = ints[:, 1].astype(np.int64)
        a, cell = ind // (H * W), ind % (H * W)
        assert np.all(np.diff(cell * p['nA'] + a) > 0)
        got = set((n * K + ind).tolist())
        want = set((n * K + np.nonzero(conf[n * K:(n + 1) * K] > p['thr'])[0]).tolist())
        for i in got ^ want:
            assert abs(conf[i] - p['thr']) <= 1e-5 * p['thr']
        g = n * K + ind
        for k, ref in enumerate((xs, ys, ws, hs, det, cmax)):
            np.testing.assert_allclose(c[:, k], ref[g], rtol=1e-5, atol=1e-7)
        assert np.array_equal(ints[:, 0], cid[g].astype(np.int32))
    assert np.abs(count.astype(np.int64) - gold[tag + '/rows']).max(initial=0) <= 1
    if dense is not None:
        np.testing.assert_allclose(dense, cls.reshape(-1, p['nC']), rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize('tag', CASES)
def test_emulated_nms_bit_exact(emul, gold, tag):
    from oracle import utils as OU
    p = params(gold, tag)
    N, _, H, W = gold[tag + '/output'].shape
    cand, count, boxes = _cand_from_oracle(gold, tag, p)
    K = cand.shape[1]
    keep = np.full((N, K), -1, dtype=np.int32)
    kc = np.full(N, -1, dtype=np.int32)
    emul.emul_nms(P(cand), None, P(count), N, K, H, W, ctypes.c_double(p['nms']), P(keep), P(kc))
    for n in range(N):
        row = [list(b) + [s] for s, b in enumerate(boxes[n])]
        want = [b[-1] for b in OU.nms(row, p['nms'])]
        assert kc[n] == len(want)
        assert keep[n, :kc[n]].tolist() == want
    assert np.array_equal(kc, gold[tag + '/nms_rows'])


def test_emulated_nms_boxes64_bit_exact(emul, gold):
    tag = 'v1_detect'
    rows = unflatten(gold[tag + '/rows'], gold[tag + '/len'], gold[tag + '/vals'])
    want = unflatten(gold[tag + '/nms_rows'], gold[tag + '/nms_len'], gold[tag + '/nms_vals'])
    for row, w in zip(rows, want):
        n = len(row)
        b64 = np.array([b[:5] for b in row], dtype=np.float64)
        count = np.array([n], dtype=np.int32)
        keep = np.full((1, n), -1, dtype=np.int32)
        kc = np.full(1, -1, dtype=np.int32)
        emul.emul_nms(None, P(b64), P(count), 1, n, 1, 1, ctypes.c_double(float(gold[tag + '/nms_thresh'])), P(keep), P(kc))
        assert [row[s] for s in keep[0, :kc[0]]] == w


def test_emulated_rw_running_mean_bit_exact(emul, gold):
    n_cls = int(gold['ens/n_cls'])
    C = gold['ens/dw0'].shape[1]
    e = np.zeros((n_cls, C), dtype=np.float32)
    cnt = [np.zeros(n_cls, dtype=np.int32), np.zeros(n_cls, dtype=np.int32)]
    cur = 0
    for k in range(3):
        dw = np.ascontiguousarray(gold['ens/dw%d' % k])
        ids = gold['ens/ids%d' % k].astype(np.int32)
        emul.emul_rw_running_mean(P(e), P(cnt[cur]), P(cnt[1 - cur]), P(dw), P(ids), len(ids), n_cls, C)
        cur = 1 - cur
    assert np.array_equal(e.view(np.uint32), gold['ens/result'].view(np.uint32))
    assert cnt[cur].sum() == sum(len(gold['ens/ids%d' % k]) for k in range(3))


@pytest.mark.parametrize('tag', ['v2_g13', 'v1_valid'])
def test_host_glue_of_detections_on_emulated_buffers(emul, gold, tag):
    """utils.Detections' host side (list building incl. the validation extras, row-wise nms bookkeeping) fed with
    the emulated kernels' buffers as CPU tensors - no C-ABI call is made (the device NMS result is pre-seeded)."""
    from fewshot_detection_b200 import utils as U
    p = params(gold, tag)
    out = gold[tag + '/output']
    N, _, H, W = out.shape
    K = p['nA'] * H * W
    cand, count, dense = run_detect(emul, out, p)
    keep = np.full((N, K), -1, dtype=np.int32)
    kc = np.full(N, -1, dtype=np.int32)
    emul.emul_nms(P(cand), None, P(count), N, K, H, W, ctypes.c_double(p['nms']), P(keep), P(kc))
    d = U.Detections(torch.from_numpy(cand), torch.from_numpy(count), torch.from_numpy(dense) if dense is not None else None,
                     N, p['nA'], p['nC'], H, W, bool(p['only_obj']), p['val'], p['thr'])
    d.keep, d.keep_count, d._nms_thresh, d._kept_host = torch.from_numpy(keep), torch.from_numpy(kc), p['nms'], None
    want_all = unflatten(gold[tag + '/rows'], gold[tag + '/len'], gold[tag + '/vals'])
    want_kept = unflatten(gold[tag + '/nms_rows'], gold[tag + '/nms_len'], gold[tag + '/nms_vals'])
    kept_fast = d.kept_boxes(p['nms'])
    rows = d.boxes()
    for n in range(N):
        assert [len(b) for b in rows[n]] == [len(b) for b in want_all[n]]
        for b, w in zip(rows[n], want_all[n]):
            np.testing.assert_allclose(b, w, rtol=1e-5, atol=1e-7)
        kept = U.nms(rows[n], p['nms'])
        assert kept == kept_fast[n]
        assert len(kept) == len(want_kept[n])
        for b, w in zip(kept, want_kept[n]):
            np.testing.assert_allclose(b, w, rtol=1e-5, atol=1e-7)
        assert sum(1 for b in rows[n] if b[4] == 0) == len(rows[n]) - len(kept)


def _random_head(rs, N, A, nC, H, W, conf_shift):
    o = (rs.randn(N, A * (5 + nC), H, W) * 1.3).astype(np.float32)
    o.reshape(N, A, 5 + nC, H, W)[:, :, 4] += conf_shift
    return o


@pytest.mark.parametrize('N,cs,A,nC,H,W,thr,shift', [
    (2, 2, 5, 1, 19, 19, 1e-9, 2.0),      # every one of the 1805 anchor-cells is a candidate: P = 2048, 8 compaction chunks
    (3, 3, 3, 1, 1, 1, 0.2, 0.0),         # a 1x1 grid
    (2, 1, 5, 4, 7, 9, 0.3, 0.0),         # plain detector, non-square grid, K = 315 (not a multiple of 256)
    (4, 2, 5, 1, 16, 16, 0.999999, 0.0),  # nothing passes
])
def test_emulated_detect_and_nms_edge_shapes(emul, N, cs, A, nC, H, W, thr, shift):
    from oracle import utils as OU
    rs = np.random.RandomState(N * 100 + H)
    out = _random_head(rs, N, A, nC, H, W, shift)
    anchors = (rs.uniform(0.5, 6.0, 2 * A)).round(3).tolist()
    v2 = nC == 1
    p = dict(nA=A, nC=nC, cs=cs, v2=v2, only_obj=0, val=False, thr=thr, anchors=anchors)
    cand, count, _ = run_detect(emul, out, p)
    import torch
    o = torch.from_numpy(out)
    boxes = OU.get_region_boxes_v2(o, cs, thr, nC, anchors, A, 0, False) if v2 else OU.get_region_boxes(o, thr, nC, anchors, A, 0, False)
    assert count.tolist() == [len(r) for r in boxes]
    K = A * H * W
    keep = np.full((N, K), -1, dtype=np.int32)
    kc = np.full(N, -1, dtype=np.int32)
    emul.emul_nms(P(cand), None, P(count), N, K, H, W, ctypes.c_double(0.45), P(keep), P(kc))
    for n in range(N):
        c = cand[n, :count[n]]
        for k in range(4):
            np.testing.assert_allclose(c[:, k] / (W if k % 2 == 0 else H), [b[k] for b in boxes[n]], rtol=1e-5, atol=1e-7)
        row = [list(b) + [s] for s, b in enumerate(boxes[n])]
        want = [b[-1] for b in OU.nms(row, 0.45)]
        # candidates carry libm floats here and torch floats in the oracle: identical survivors unless an IoU or a
        # sort key sits within an ulp of a tie, which these seeds do not produce
        assert keep[n, :kc[n]].tolist() == want


def test_emulated_nms_ties_duplicates_and_zero_confidence(emul):
    """Equal keys keep candidate order; identical boxes suppress each other; det_conf == 0 never survives."""
    H = W = 13
    K = 8
    cand = np.zeros((1, K, 8), dtype=np.float32)
    rows = [(3.5, 3.5, 2.0, 2.0, 0.7), (3.5, 3.5, 2.0, 2.0, 0.7), (9.5, 9.5, 1.0, 1.0, 0.7), (9.6, 9.5, 1.0, 1.0, 0.9),
            (1.0, 12.0, 1.0, 1.0, 0.0), (6.0, 6.0, 3.0, 3.0, 0.70000005), (6.0, 6.0, 3.0, 3.0, 0.70000001)]
    for s, r in enumerate(rows):
        cand[0, s, :5] = r
    count = np.array([len(rows)], dtype=np.int32)
    keep = np.full((1, K), -1, dtype=np.int32)
    kc = np.full(1, -1, dtype=np.int32)
    emul.emul_nms(P(cand), None, P(count), 1, K, H, W, ctypes.c_double(0.45), P(keep), P(kc))
    from oracle import utils as OU
    boxes = [[float(np.float32(r[0])) / W, float(np.float32(r[1])) / H, float(np.float32(r[2])) / W, float(np.float32(r[3])) / H,
              float(np.float32(r[4])), 1.0, 0, s] for s, r in enumerate(rows)]
    want = [b[-1] for b in OU.nms(boxes, 0.45)]
    assert keep[0, :kc[0]].tolist() == want
    assert 4 not in want and want[0] == 3 and (0 in want) != (1 in want)
