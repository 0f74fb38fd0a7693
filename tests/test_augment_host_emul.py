"""csrc/augment.cu without a GPU: the kernel source compiled by g++ (tools/host_emul) against
 (a) tests/golden/augment.npz - outputs of the reference's own image.data_augmentation / fill_truth_detection(_meta) /
     load_label (tests/golden/make_golden_augment.py), bit-exact;
 (b) Pillow itself where it is installed: both colour conversions over all 2^24 byte triples, random crop/resize
     geometries with both filters, the `point` tables.
The host side (fewshot_detection_b200.image: random draws in the reference's order, label transforms) is covered
here too.  The -m gpu twin through the C ABI is tests/test_gpu_zz_augment.py."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, 'tests', 'golden')
SRC = os.path.join(ROOT, 'tools', 'host_emul', 'augment_emul.cpp')
HDR = os.path.join(ROOT, 'tools', 'host_emul', 'cuda_host_emul.h')
KSRC = os.path.join(ROOT, 'fewshot_detection_b200', 'csrc', 'augment.cu')
LIB = os.path.join(ROOT, 'build', 'libaugment_emul.so')
IMG_CASES = ['a', 'b', 'c', 'd', 'e', 'f', 'g']

try:
    from PIL import Image
    HAVE_PIL = True
except Exception:  # pragma: no cover
    HAVE_PIL = False


@pytest.fixture(scope='module')
def emul():
    inc = '/usr/local/cuda/include'
    if not os.path.exists(os.path.join(inc, 'cuda_runtime.h')):
        pytest.skip('CUDA headers not found')
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(p) for p in (SRC, HDR, KSRC)):
        cmd = ['g++', '-O2', '-std=c++17', '-ffp-contract=off', '-fPIC', '-shared', '-pthread', '-w',
               '-DFSDET_HOST_EMULATION', '-I' + inc, '-include', HDR, '-x', 'c++', SRC, '-o', LIB]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout
    return ctypes.CDLL(LIB)


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(G, 'augment.npz'), allow_pickle=False)


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def run_augment(emul, images, shape, params, filt):
    """The emulated fsdet_augment_batch on host arrays; returns (float32 [n,3,H,W], uint8 [n,H,W,3])."""
    from fewshot_detection_b200.image import marshal_params
    W, H = shape
    n = len(images)
    images = [np.ascontiguousarray(im) for im in images]
    geom, color, kmax = marshal_params([im.shape[:2] for im in images], params, W, H)
    L = max(W, H)
    tables = np.zeros((n, 2, L, 2 + kmax), dtype=np.int32)
    luts = np.zeros((n, 3, 256), dtype=np.uint8)
    out = np.zeros((n, 3, H, W), dtype=np.float32)
    u8 = np.zeros((n, H, W, 3), dtype=np.uint8)
    status = np.zeros(1, dtype=np.int32)
    ptrs = (ctypes.c_void_p * n)(*[im.ctypes.data for im in images])
    emul.emul_augment_batch(ptrs, P(geom), P(color), n, W, H, kmax, filt, P(tables), P(luts), P(out), P(u8), P(status))
    assert status[0] == 0
    return out, u8


def params_for(gold, tag):
    from fewshot_detection_b200 import image as I
    W, H, seed, flag = [int(v) for v in gold[tag + '/args']]
    src = gold[tag + '/src']
    random.seed(seed)
    oh, ow = src.shape[:2]
    p = I.draw_augmentation(ow, oh, 0.2, 0.1, 1.5, 1.5) if flag else I.identity_augmentation(ow, oh)
    return src, (W, H), p


@pytest.mark.parametrize('tag', IMG_CASES)
def test_augment_bit_exact_vs_reference(emul, gold, tag):
    src, shape, p = params_for(gold, tag)
    flip, dx, dy, sx, sy = gold[tag + '/params']
    assert (p['flip'], p['dx'], p['dy'], p['sx'], p['sy']) == (int(flip), dx, dy, sx, sy)   # same draws, same float64 math
    out, u8 = run_augment(emul, [src], shape, [p], 3)
    assert np.array_equal(u8[0], gold[tag + '/img'])
    want = (gold[tag + '/img'].astype(np.float32) / np.float32(255)).transpose(2, 0, 1)      # ToTensor
    assert np.array_equal(out[0].view(np.uint32), np.ascontiguousarray(want).view(np.uint32))
    out0, u80 = run_augment(emul, [src], shape, [p], 0)
    assert np.array_equal(u80[0], gold['nearest/' + tag])


def test_augment_batch_of_mixed_sizes_equals_single_calls(emul, gold):
    tags = ['a', 'c', 'f']
    srcs, ps = [], []
    for t in tags:
        src, shape, p = params_for(gold, t)
        srcs.append(src)
        ps.append(p)
    out, u8 = run_augment(emul, srcs, (64, 64), ps, 3)
    for i, t in enumerate(tags):
        if tuple(gold[t + '/args'][:2]) == (64, 64):
            assert np.array_equal(u8[i], gold[t + '/img'])
        one, one8 = run_augment(emul, [srcs[i]], (64, 64), [ps[i]], 3)
        assert np.array_equal(one8[0], u8[i])


@pytest.mark.skipif(not HAVE_PIL, reason='Pillow not installed')
def test_colour_conversions_exhaustive_vs_pillow(emul):
    v = np.arange(256, dtype=np.uint8)
    allc = np.ascontiguousarray(np.stack(np.meshgrid(v, v, v, indexing='ij'), -1).reshape(4096, 4096, 3))
    buf = np.zeros((4096, 4096, 3), dtype=np.uint8)
    emul.emul_rgb2hsv_all(P(buf))
    assert np.array_equal(buf, np.asarray(Image.fromarray(allc, 'RGB').convert('HSV')))
    emul.emul_hsv2rgb_all(P(buf))
    assert np.array_equal(buf, np.asarray(Image.fromarray(allc, 'HSV').convert('RGB')))


@pytest.mark.skipif(not HAVE_PIL, reason='Pillow not installed')
@pytest.mark.parametrize('filt', [3, 0])
def test_random_geometries_vs_pillow(emul, filt):
    """crop (incl. boxes reaching outside the image) + resize + flip + distort against explicit Pillow calls."""
    rs = np.random.RandomState(11 + filt)
    pil_filter = Image.BICUBIC if filt == 3 else Image.NEAREST
    for trial in range(14):
        h, w = [int(v) for v in rs.randint(20, 140, 2)]
        W, H = [int(v) for v in rs.randint(16, 150, 2)]
        if trial == 0:
            W = w - 9          # horizontal pass skipped after the crop below
        a = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        pleft, ptop = int(rs.randint(-12, 13)), int(rs.randint(-12, 13))
        cw, ch = int(rs.randint(max(8, w // 2), w + 14)), int(rs.randint(max(8, h // 2), h + 14))
        if trial == 0:
            pleft, cw = 4, w - 9
        flip = int(rs.randint(0, 2))
        dhue, dsat, dexp = float(rs.uniform(-0.1, 0.1)), float(rs.uniform(0.66, 1.5)), float(rs.uniform(0.66, 1.5))
        p = dict(pleft=pleft, ptop=ptop, cw=cw, ch=ch, flip=flip, distort=1, dhue=dhue, dsat=dsat, dexp=dexp)
        im = Image.fromarray(a, 'RGB').crop((pleft, ptop, pleft + cw, ptop + ch)).resize((W, H), pil_filter)
        if flip:
            im = im.transpose(Image.FLIP_LEFT_RIGHT)
        hsv = list(im.convert('HSV').split())

        def change_hue(x, hue=dhue):
            x += hue * 255
            if x > 255:
                x -= 255
            if x < 0:
                x += 255
            return x
        hsv[0] = hsv[0].point(change_hue)
        hsv[1] = hsv[1].point(lambda i: i * dsat)
        hsv[2] = hsv[2].point(lambda i: i * dexp)
        want = np.asarray(Image.merge('HSV', tuple(hsv)).convert('RGB'))
        out, u8 = run_augment(emul, [a], (W, H), [p], filt)
        assert np.array_equal(u8[0], want), (trial, h, w, W, H, p)


def test_label_transforms_bit_exact_vs_reference(gold):
    from fewshot_detection_b200 import image as I
    from fewshot_detection_b200.cfg import cfg
    old = (cfg.base_classes, cfg.base_ids, cfg.yolo_joint)
    cfg.base_classes, cfg.base_ids, cfg.yolo_joint = cfg.voc_classes[:15], list(range(15)), False
    try:
        for tag in ('l1', 'l2', 'l3', 'l4'):
            boxes = gold[tag + '/boxes']
            flip, dx, dy, sx, sy = gold[tag + '/transform']
            got = I.fill_truth_detection(boxes.copy(), 416, 416, int(flip), dx, dy, sx, sy)
            assert np.array_equal(got, gold[tag + '/fill']), tag
            if tag + '/fill_meta' in gold:
                gm = I.fill_truth_detection_meta(boxes.copy(), 416, 416, int(flip), dx, dy, sx, sy)
                assert np.array_equal(gm, gold[tag + '/fill_meta']), tag
            ll = np.array(I.load_label(boxes.copy(), 416, 416, int(flip), dx, dy, sx, sy), dtype=np.float64).reshape(-1, 4)
            assert np.array_equal(ll, gold[tag + '/load_label']), tag
        assert gold['l1/fill'].reshape(-1, 5)[:, 3].max() > 0
    finally:
        cfg.base_classes, cfg.base_ids, cfg.yolo_joint = old


def test_label_file_path_form(tmp_path, gold):
    from fewshot_detection_b200 import image as I
    boxes = gold['l1/boxes']
    lab = tmp_path / '000007.txt'
    lab.write_text(''.join('%d %.6f %.6f %.6f %.6f\n' % tuple(r) for r in boxes))
    flip, dx, dy, sx, sy = gold['l1/transform']
    a = I.fill_truth_detection(str(lab), 416, 416, int(flip), dx, dy, sx, sy)
    b = I.fill_truth_detection(np.loadtxt(str(lab)), 416, 416, int(flip), dx, dy, sx, sy)
    assert np.array_equal(a, b)
    assert not I.fill_truth_detection(str(tmp_path / 'missing.txt'), 416, 416, 0, 0, 0, 1, 1).any()


def test_box_masks_and_rects(emul):
    from fewshot_detection_b200.image import mask_rect
    boxes = [[0.5, 0.5, 0.25, 0.5], [0.1, 0.9, 0.4, 0.4], [0.3, 0.3, 0.0, 0.2], [0.505, 0.495, 0.01, 0.01]]
    w, h = 64, 48
    rects = np.array([mask_rect(b, w, h) for b in boxes], dtype=np.int32)
    assert rects[2][0] == rects[2][2]                      # empty rectangle (the reference's mask=None case)
    out = np.full((len(boxes), h, w), -1, dtype=np.float32)
    emul.emul_box_masks(P(rects), len(boxes), h, w, P(out))
    for i, (x1, y1, x2, y2) in enumerate(rects):
        want = np.zeros((h, w), dtype=np.float32)
        want[y1:y2, x1:x2] = 1
        assert np.array_equal(out[i], want)


def test_python_glue_of_augment_batch_through_the_emulated_entry_point(emul, gold, monkeypatch):
    """image.augment_batch / data_augmentation / load_data_detection end to end on CPU tensors: the C-ABI call is
    redirected to the host-emulated kernels with the SAME argument list (pointer table, geom / color tables, workspace
    split), so argument order, dtypes and shapes of the real call are what is being tested."""
    import torch
    from fewshot_detection_b200 import image as I

    def fake_call(name, *a):
        assert name == 'fsdet_augment_batch'
        src, geom, color, n, W, H, kmax, filt, ws, ws_bytes, out, out_u8, status, stream = a
        L = max(W, H)
        tbytes = n * 2 * L * (2 + kmax) * 4
        assert ws_bytes >= tbytes + n * 768
        emul.emul_augment_batch(ctypes.c_void_p(src), ctypes.c_void_p(geom), ctypes.c_void_p(color), n, W, H, kmax, filt,
                                ctypes.c_void_p(ws), ctypes.c_void_p(ws + tbytes), ctypes.c_void_p(out),
                                ctypes.c_void_p(out_u8) if out_u8 else None, ctypes.c_void_p(status))
        return 0
    monkeypatch.setattr(I, 'call', fake_call)
    monkeypatch.setattr(I, '_st', lambda: None)
    monkeypatch.setattr(torch.cuda, 'is_available', lambda: True)
    for tag in ('a', 'd', 'f'):
        W, H, seed, flag = [int(v) for v in gold[tag + '/args']]
        random.seed(seed)
        src = gold[tag + '/src']
        p = I.draw_augmentation(src.shape[1], src.shape[0], 0.2, 0.1, 1.5, 1.5) if flag else \
            I.identity_augmentation(src.shape[1], src.shape[0])
        out, u8 = I.augment_batch([src], (W, H), [p], device='cpu', return_uint8=True)
        assert np.array_equal(u8[0].numpy(), gold[tag + '/img'])
        want = np.ascontiguousarray((gold[tag + '/img'].astype(np.float32) / np.float32(255)).transpose(2, 0, 1))
        assert np.array_equal(out[0].numpy(), want)
    # a batch of three differently sized sources into a caller-provided output
    srcs, ps = [], []
    for tag in ('a', 'f', 'c'):
        random.seed(int(gold[tag + '/args'][2]))
        s = gold[tag + '/src']
        srcs.append(torch.from_numpy(s))
        ps.append(I.draw_augmentation(s.shape[1], s.shape[0], 0.2, 0.1, 1.5, 1.5))
    dst = torch.full((3, 3, 64, 64), -1.0)
    I.augment_batch(srcs, (64, 64), ps, device='cpu', out=dst)
    for i, tag in enumerate(('a', 'f')):
        want = np.ascontiguousarray((gold[tag + '/img'].astype(np.float32) / np.float32(255)).transpose(2, 0, 1))
        assert np.array_equal(dst[i].numpy(), want)
    assert dst[2].min() >= 0


@pytest.mark.skipif(not HAVE_PIL, reason='Pillow not installed')
@pytest.mark.parametrize('h,w,W,H,pleft,ptop,cw,ch', [
    (300, 400, 32, 24, 0, 0, 400, 300),        # 12.5x down-scaling: 51 taps per axis
    (40, 50, 200, 160, 5, 5, 8, 6),            # 25x up-scaling of a tiny crop
    (30, 30, 20, 20, 40, 40, 12, 12),          # crop entirely outside the image: all zero -> black
    (25, 31, 17, 13, -3, 28, 40, 5),           # crop hanging over two edges
    (64, 64, 64, 64, 0, 0, 64, 64),            # identity geometry: both passes skipped
    (50, 70, 7, 1, 0, 0, 70, 50),              # single output row
])
def test_extreme_geometries_vs_pillow(emul, h, w, W, H, pleft, ptop, cw, ch):
    rs = np.random.RandomState(h * 7 + W)
    a = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
    for filt, pil_filter in ((3, Image.BICUBIC), (0, Image.NEAREST)):
        p = dict(pleft=pleft, ptop=ptop, cw=cw, ch=ch, flip=0, distort=0, dhue=0.0, dsat=1.0, dexp=1.0)
        want = np.asarray(Image.fromarray(a, 'RGB').crop((pleft, ptop, pleft + cw, ptop + ch)).resize((W, H), pil_filter))
        out, u8 = run_augment(emul, [a], (W, H), [p], filt)
        assert np.array_equal(u8[0], want), filt


def test_oracle_image_equals_emulated_kernel_on_random_parameters(emul):
    """Differential test oracle/image.py (numpy) vs the kernel source, no Pillow needed: 10 seeded augmentations."""
    from oracle import image as OI
    from fewshot_detection_b200 import image as I
    rs = np.random.RandomState(5)
    for seed in range(10):
        h, w = int(rs.randint(24, 90)), int(rs.randint(24, 90))
        W, H = int(rs.randint(16, 80)), int(rs.randint(16, 80))
        a = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        random.seed(seed)
        want, flip, dx, dy, sx, sy = OI.data_augmentation(a, (W, H), 0.2, 0.1, 1.5, 1.5)
        random.seed(seed)
        p = I.draw_augmentation(w, h, 0.2, 0.1, 1.5, 1.5)
        assert (p['flip'], p['dx'], p['dy'], p['sx'], p['sy']) == (flip, dx, dy, sx, sy)
        out, u8 = run_augment(emul, [a], (W, H), [p], 3)
        assert np.array_equal(u8[0], want), seed
        assert np.array_equal(out[0], OI.to_tensor(want))
