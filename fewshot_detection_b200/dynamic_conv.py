"""`dynamic_conv2d(is_first, partial)` class factory (dynamic_conv.py:110-168).

The reference's DynamicConv2d tiles the input n_cls times and runs a grouped 1x1
convolution with groups = n_cls * C, materialising a [B*n_cls, C, H, W] tensor.
Here the module is a parameter-less marker: the engine fuses the per-class channel
reweighting into the following 1x1 detection convolution (engine.NetRunner._head_fwd),
so that tensor is never built.  Only the configuration the shipped cfgs use is
supported (is_first=True, partial=None: no shared weight, no bias).
"""
import torch.nn as nn


def dynamic_conv2d(is_first, partial=None):
    if partial is not None:
        raise NotImplementedError('dynamic conv with partial= (shared weights) is not used by the shipped cfgs')

    class DynamicConv2d(nn.Module):
        is_dynamic_conv = True

        def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                     bias=False):
            super(DynamicConv2d, self).__init__()
            self.in_channels = in_channels
            self.out_channels = out_channels
            self.kernel_size = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
            self.register_parameter('weight', None)  # dynamic_conv.py:47-48
            self.register_parameter('bias', None)

        def forward(self, inputs):
            raise RuntimeError('DynamicConv2d is fused into the detection head; call Darknet.detect_forward')

        def extra_repr(self):
            return '{}, {}, kernel_size={}, bias=False [fused per-class reweighting]'.format(
                self.in_channels, self.out_channels, self.kernel_size)

    DynamicConv2d.is_first = is_first
    DynamicConv2d.partial = partial
    return DynamicConv2d
