"""Oracle restatement of the reference's training-input augmentation (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/image.py:
  rand_scale / random_distort_image / distort_image   image.py:19-50
  data_augmentation                                   image.py:52-87

Where the arithmetic really lives: the reference delegates every pixel operation to the un-vendored third-party
dependency Pillow (unpinned in requirements.txt; `from PIL import Image`, image.py:7): `Image.crop`, `Image.resize`
with the default filter, `Image.transpose`, `Image.convert('HSV'|'RGB')`, `Image.point`.  This module restates
Pillow's published uint8 algorithms in numpy:
  resize BICUBIC   libImaging/Resample.c: precompute_coeffs (double), normalize_coeffs_8bpc (22-bit fixed point),
                   horizontal pass then vertical pass, clip8 after each; a pass is skipped when the size is unchanged
  resize NEAREST   libImaging/Geometry.c ImagingScaleAffine: source index = (int)(a*0.5 + a + a + ...) accumulated
  RGB<->HSV        libImaging/Convert.c rgb2hsv_row / hsv2rgb (float/double mix as in the C source)
  point(f)         PIL/Image.py: table [round(f(i)) for i in range(256)] (round-half-even), clipped to 0..255
and is pinned bit-exactly against tests/golden/augment.npz (outputs of the reference's own image.data_augmentation
under Pillow 12.2, tests/golden/make_golden_augment.py) by tests/test_oracle_augment.py.
"""
import random

import numpy as np

PB = 32 - 8 - 2


def _bicubic(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def _coeffs(insize, outsize):
    scale = filterscale = float(insize) / outsize
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    bounds, kk = [], np.zeros((outsize, ksize), dtype=np.int64)
    for xx in range(outsize):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), insize) - xmin
        k = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        for x in range(xmax):
            v = k[x] / ww if ww != 0.0 else k[x]
            kk[xx, x] = int(-0.5 + v * (1 << PB)) if v < 0 else int(0.5 + v * (1 << PB))
        bounds.append((xmin, xmax))
    return bounds, kk


def _clip8(v):
    return np.clip(v >> PB, 0, 255)


def resize_bicubic(a, W, H):
    h, w, _ = a.shape
    cur = a.astype(np.int64)
    if W != w:
        b, kk = _coeffs(w, W)
        out = np.zeros((h, W, 3), dtype=np.int64)
        for xx in range(W):
            xmin, n = b[xx]
            out[:, xx] = _clip8((1 << (PB - 1)) + np.tensordot(cur[:, xmin:xmin + n], kk[xx, :n], axes=([1], [0])))
        cur = out
    if H != h:
        b, kk = _coeffs(h, H)
        out = np.zeros((H, cur.shape[1], 3), dtype=np.int64)
        for yy in range(H):
            ymin, n = b[yy]
            out[yy] = _clip8((1 << (PB - 1)) + np.tensordot(kk[yy, :n], cur[ymin:ymin + n], axes=([0], [0])))
        cur = out
    return cur.astype(np.uint8)


def _nearest_index(insize, outsize):
    a = float(insize) / outsize
    xo = a * 0.5
    idx = np.zeros(outsize, dtype=np.int64)
    for x in range(outsize):
        xin = int(xo)
        idx[x] = xin if 0 <= xin < insize else -1
        xo += a
    return idx


def resize_nearest(a, W, H):
    h, w, _ = a.shape
    sx, sy = _nearest_index(w, W), _nearest_index(h, H)
    out = a[np.maximum(sy, 0)][:, np.maximum(sx, 0)].copy()
    out[sy < 0] = 0
    out[:, sx < 0] = 0
    return out


def crop(a, left, top, right, bottom):
    """PIL Image.crop: size (right-left, bottom-top), zero outside the source."""
    h, w, _ = a.shape
    out = np.zeros((bottom - top, right - left, 3), dtype=np.uint8)
    x0, x1, y0, y1 = max(left, 0), min(right, w), max(top, 0), min(bottom, h)
    if x1 > x0 and y1 > y0:
        out[y0 - top:y1 - top, x0 - left:x1 - left] = a[y0:y1, x0:x1]
    return out


def rgb2hsv(a):
    r, g, b = [a[..., i].astype(np.int32) for i in range(3)]
    maxc = np.maximum(r, np.maximum(g, b))
    minc = np.minimum(r, np.minimum(g, b))
    with np.errstate(divide='ignore', invalid='ignore'):
        cr = (maxc - minc).astype(np.float32)
        s = cr / maxc.astype(np.float32)
        rc = (maxc - r).astype(np.float32) / cr
        gc = (maxc - g).astype(np.float32) / cr
        bc = (maxc - b).astype(np.float32) / cr
        h = np.where(r == maxc, bc - gc,
                     np.where(g == maxc, (2.0 + rc.astype(np.float64) - bc).astype(np.float32),
                              (4.0 + gc.astype(np.float64) - rc).astype(np.float32))).astype(np.float32)
        h = np.fmod(h.astype(np.float64) / 6.0 + 1.0, 1.0).astype(np.float32)
        uh = np.clip(np.nan_to_num(h.astype(np.float64) * 255.0).astype(np.int64), 0, 255)
        us = np.clip(np.nan_to_num(s.astype(np.float64) * 255.0).astype(np.int64), 0, 255)
    gray = minc == maxc
    return np.stack([np.where(gray, 0, uh), np.where(gray, 0, us), maxc], -1).astype(np.uint8)


def _cround(x):
    return np.where(x >= 0, np.floor(x + 0.5), np.ceil(x - 0.5))


def hsv2rgb(a):
    h, s, v = [a[..., i].astype(np.float32) for i in range(3)]
    h6 = h.astype(np.float64) * 6.0 / 255.0
    i = np.floor(h6).astype(np.int64)
    f = (h6 - i.astype(np.float32).astype(np.float64)).astype(np.float32)
    sd = s.astype(np.float64) / 255.0
    fs = (sd * f.astype(np.float64)).astype(np.float32).astype(np.float64)
    vd = v.astype(np.float64)
    p = np.clip(_cround(vd * (1.0 - sd)), 0, 255)
    q = np.clip(_cround(vd * (1.0 - fs)), 0, 255)
    t = np.clip(_cround(vd * (1.0 - sd + fs)), 0, 255)
    m = i % 6
    r = np.choose(m, [vd, q, p, p, t, vd])
    g = np.choose(m, [t, vd, vd, q, p, p])
    b = np.choose(m, [p, p, t, vd, vd, q])
    gray = a[..., 1] == 0
    return np.stack([np.where(gray, vd, r), np.where(gray, vd, g), np.where(gray, vd, b)], -1).astype(np.uint8)


def _point(chan, f):
    lut = np.array([min(255, max(0, round(f(i)))) for i in range(256)], dtype=np.uint8)
    return lut[chan]


def distort_image(a, hue, sat, val):
    """image.py:19-37."""
    hsv = rgb2hsv(a)

    def change_hue(x):
        x += hue * 255
        if x > 255:
            x -= 255
        if x < 0:
            x += 255
        return x
    out = np.stack([_point(hsv[..., 0], change_hue), _point(hsv[..., 1], lambda i: i * sat),
                    _point(hsv[..., 2], lambda i: i * val)], -1)
    return hsv2rgb(out)


def rand_scale(s):
    """image.py:39-43."""
    scale = random.uniform(1, s)
    if random.randint(1, 10000) % 2:
        return scale
    return 1. / scale


def data_augmentation(a, shape, jitter, hue, saturation, exposure, flag=True, bicubic=True):
    """image.py:52-87 on a uint8 [h, w, 3] array; returns (uint8 [H, W, 3], flip, dx, dy, sx, sy)."""
    resize = resize_bicubic if bicubic else resize_nearest
    oh, ow = a.shape[:2]
    dw = int(ow * jitter)
    dh = int(oh * jitter)
    if flag:
        pleft = random.randint(-dw, dw)
        pright = random.randint(-dw, dw)
        ptop = random.randint(-dh, dh)
        pbot = random.randint(-dh, dh)
        flip = random.randint(1, 10000) % 2
        swidth = ow - pleft - pright
        sheight = oh - ptop - pbot
        sx = float(swidth) / ow
        sy = float(sheight) / oh
        cropped = crop(a, pleft, ptop, pleft + swidth - 1, ptop + sheight - 1)
        dx = (float(pleft) / ow) / sx
        dy = (float(ptop) / oh) / sy
        sized = resize(cropped, shape[0], shape[1])
        if flip:
            sized = sized[:, ::-1]
        dhue = random.uniform(-hue, hue)
        dsat = rand_scale(saturation)
        dexp = rand_scale(exposure)
        img = distort_image(sized, dhue, dsat, dexp)
    else:
        flip, dx, dy, sx, sy = 0, 0, 0, 1, 1
        img = resize(a, shape[0], shape[1])
    return img, flip, dx, dy, sx, sy


def to_tensor(img):
    """transforms.ToTensor: uint8 HWC -> float32 CHW / 255."""
    return np.ascontiguousarray((img.astype(np.float32) / np.float32(255)).transpose(2, 0, 1))
