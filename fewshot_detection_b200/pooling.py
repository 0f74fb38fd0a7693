"""Parameter-free layer modules that appear in the reference's ModuleLists.

They keep the reference's class names (pooling.py:8-60, darknet_meta.py:37-83)
so that printed networks and isinstance checks in user code keep working; the
compute is done by the engine (fewshot_detection_b200/engine.py) on libfsdet.so.
Called on their own they run the same CUDA kernels on a single tensor.
"""
import torch
import torch.nn as nn

from . import _lib
from ._lib import call, ptr


def _st():
    return torch.cuda.current_stream().cuda_stream


def _to_nhwc(x):
    B, C, H, W = x.shape
    cp = (C + 3) // 4 * 4
    buf = torch.empty(B * H * W, cp, device=x.device)
    call('fsdet_nchw_to_nhwc', ptr(x.contiguous()), C, None, 0, ptr(buf), cp, cp, B, H * W, _st())
    return buf, cp


def _to_nchw(buf, B, C, H, W):
    out = torch.empty(B, C, H, W, device=buf.device)
    call('fsdet_nhwc_to_nchw', ptr(buf), buf.shape[1], None, ptr(out), B, C, H * W, _st())
    return out


def _check(x):
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4):
        raise TypeError('expects a float32 CUDA NCHW tensor (there is no CPU fallback)')


class GlobalMaxPool2d(nn.Module):
    """pooling.py:8-27 (forward only when used stand-alone)."""

    def forward(self, x):
        _check(x)
        B, C, H, W = x.shape
        buf, cp = _to_nhwc(x)
        y = torch.empty(B, cp, device=x.device)
        arg = torch.empty(B, cp, dtype=torch.int32, device=x.device)
        call('fsdet_globalmax_fwd', ptr(buf), cp, ptr(y), ptr(arg), B, H * W, cp, _st())
        return y[:, :C].reshape(B, C, 1, 1)


class MaxPoolStride1(nn.Module):
    """darknet_meta.py:47-53."""

    def forward(self, x):
        _check(x)
        B, C, H, W = x.shape
        buf, cp = _to_nhwc(x)
        y = torch.empty(B * H * W, cp, device=x.device)
        call('fsdet_maxpool_fwd', ptr(buf), cp, ptr(y), cp, B, H, W, cp, 1, _st())
        return _to_nchw(y, B, C, H, W)


class MaxPool2x2(nn.MaxPool2d):
    """nn.MaxPool2d(2, 2) placeholder (darknet_meta.py:260-266)."""

    def forward(self, x):
        _check(x)
        B, C, H, W = x.shape
        buf, cp = _to_nhwc(x)
        y = torch.empty(B * (H // 2) * (W // 2), cp, device=x.device)
        call('fsdet_maxpool_fwd', ptr(buf), cp, ptr(y), cp, B, H, W, cp, 2, _st())
        return _to_nchw(y, B, C, H // 2, W // 2)


class Reorg(nn.Module):
    """darknet_meta.py:55-74: out[b,(i*2+j)*C+c,h,w] = x[b,c,2h+i,2w+j]."""

    def __init__(self, stride=2):
        super(Reorg, self).__init__()
        self.stride = stride

    def forward(self, x):
        _check(x)
        assert self.stride == 2
        B, C, H, W = x.shape
        assert H % 2 == 0 and W % 2 == 0
        if C % 4:
            raise NotImplementedError('stand-alone Reorg needs C % 4 == 0')
        buf, cp = _to_nhwc(x)
        y = torch.empty(B * (H // 2) * (W // 2), 4 * C, device=x.device)
        call('fsdet_reorg_fwd', ptr(buf), cp, ptr(y), 4 * C, B, H, W, C, _st())
        return _to_nchw(y, B, 4 * C, H // 2, W // 2)


class EmptyModule(nn.Module):
    """route / shortcut / region placeholder (darknet_meta.py:77-83)."""

    def forward(self, x):
        return x
