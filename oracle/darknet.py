"""Oracle restatement of the reference's cfg-driven networks on torch-CPU ops
(TEST INFRASTRUCTURE ONLY).

Follows /root/reference:
  darknet_meta.py:86-201, 208-353   Darknet(darknet, learnet): meta_forward / detect_forward
  darknet.py:61-129, 134-245        Darknet(cfg): plain YOLOv2 interpreter
  darknet_meta.py:47-74             MaxPoolStride1, Reorg
  dynamic_conv.py:125-164           DynamicConv2d.forward (is_first=True, partial=None)
  pooling.py:8-27                   GlobalMaxPool2d

The module tree (ModuleList of Sequential(conv{i}[, bn{i}][, leaky{i}])) and
hence `named_parameters()` order mirror the reference so that
tests/golden/seeding.seeded_init yields the same weights in both.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class MaxPoolStride1(nn.Module):
    def forward(self, x):  # darknet_meta.py:47-53
        return F.max_pool2d(F.pad(x, (0, 1, 0, 1), mode='replicate'), 2, stride=1)


class Reorg(nn.Module):
    def __init__(self, stride=2):
        super().__init__()
        self.stride = stride

    def forward(self, x):  # darknet_meta.py:55-74
        s = self.stride
        B, C, H, W = x.shape
        assert H % s == 0 and W % s == 0
        x = x.view(B, C, H // s, s, W // s, s).transpose(3, 4).contiguous()
        x = x.view(B, C, (H // s) * (W // s), s * s).transpose(2, 3).contiguous()
        x = x.view(B, C, s * s, H // s, W // s).transpose(1, 2).contiguous()
        return x.view(B, s * s * C, H // s, W // s)


class GlobalMaxPool2d(nn.Module):
    def forward(self, x):  # pooling.py:23-27
        return F.max_pool2d(x, x.size(-1), 1)


class DynamicConv2d(nn.Module):
    """dynamic_conv.py:110-168 with is_first=True, partial=None: tile the input
    n_cls times along channels and apply a grouped 1x1 conv whose weights are
    the dynamic (per-class) vectors."""

    def __init__(self):
        super().__init__()
        self.register_parameter('weight', None)
        self.register_parameter('bias', None)

    def forward(self, inputs):
        x, dw = inputs
        n_cls = dw.size(0)
        n_channels = x.size(1)
        x = x.repeat(1, n_cls, 1, 1)
        group_size = dw.size(1) // n_channels
        groups = n_cls * n_channels // group_size
        dw = dw.reshape(-1, group_size, dw.size(2), dw.size(3))
        y = F.conv2d(x, dw, None, 1, 0, 1, groups)
        return y.view(-1, n_channels, y.size(-2), y.size(-1))


class EmptyModule(nn.Module):
    def forward(self, x):
        return x


def is_dynamic(block):
    return 'dynamic' in block and int(block['dynamic']) == 1


def create_network(blocks):
    """darknet_meta.py:208-353 (region blocks become EmptyModule: the loss is
    applied by the caller through oracle.region_loss)."""
    models = nn.ModuleList()
    prev_filters = 3
    out_filters = []
    conv_id = 0
    for block in blocks:
        t = block['type']
        if t in ('net', 'learnet'):
            prev_filters = int(block['channels'])
            continue
        elif t == 'convolutional':
            conv_id += 1
            bn = int(block['batch_normalize'])
            filters = int(block['filters'])
            k = int(block['size'])
            stride = int(block['stride'])
            pad = (k - 1) // 2 if int(block['pad']) else 0
            act = block['activation']
            model = nn.Sequential()
            if is_dynamic(block):
                model.add_module('conv{0}'.format(conv_id), DynamicConv2d())
            elif bn:
                model.add_module('conv{0}'.format(conv_id), nn.Conv2d(prev_filters, filters, k, stride, pad, bias=False))
                model.add_module('bn{0}'.format(conv_id), nn.BatchNorm2d(filters))
            else:
                model.add_module('conv{0}'.format(conv_id), nn.Conv2d(prev_filters, filters, k, stride, pad))
            if act == 'leaky':
                model.add_module('leaky{0}'.format(conv_id), nn.LeakyReLU(0.1, inplace=True))
            elif act == 'relu':
                model.add_module('relu{0}'.format(conv_id), nn.ReLU(inplace=True))
            prev_filters = filters
            out_filters.append(prev_filters)
            models.append(model)
        elif t == 'maxpool':
            size, stride = int(block['size']), int(block['stride'])
            models.append(nn.MaxPool2d(size, stride) if stride > 1 else MaxPoolStride1())
            out_filters.append(prev_filters)
        elif t == 'reorg':
            stride = int(block['stride'])
            prev_filters = stride * stride * prev_filters
            out_filters.append(prev_filters)
            models.append(Reorg(stride))
        elif t == 'route':
            layers = [int(i) for i in block['layers'].split(',')]
            ind = len(models)
            layers = [i if i > 0 else i + ind for i in layers]
            if len(layers) == 1:
                prev_filters = out_filters[layers[0]]
            else:
                prev_filters = out_filters[layers[0]] + out_filters[layers[1]]
            out_filters.append(prev_filters)
            models.append(EmptyModule())
        elif t == 'globalmax':
            out_filters.append(prev_filters)
            models.append(GlobalMaxPool2d())
        elif t == 'region':
            out_filters.append(prev_filters)
            models.append(EmptyModule())
        else:
            raise NotImplementedError('oracle: block type %s' % t)
    return models


def _run_blocks(blocks, models, x, dynamic_weights=None):
    """darknet_meta.py:130-195 / darknet.py:80-129."""
    ind = -2
    dyn_cnt = 0
    outputs = {}
    for block in blocks:
        ind += 1
        t = block['type']
        if t in ('net', 'learnet'):
            continue
        elif t in ('convolutional', 'maxpool', 'reorg', 'globalmax'):
            if is_dynamic(block):
                x = models[ind]((x, dynamic_weights[dyn_cnt]))
                dyn_cnt += 1
            else:
                x = models[ind](x)
            outputs[ind] = x
        elif t == 'route':
            layers = [int(i) for i in block['layers'].split(',')]
            layers = [i if i > 0 else i + ind for i in layers]
            if len(layers) == 1:
                x = outputs[layers[0]]
            else:
                x = torch.cat((outputs[layers[0]], outputs[layers[1]]), 1)
            outputs[ind] = x
        elif t == 'region':
            continue
        else:
            raise NotImplementedError(t)
    return x


class MetaDarknet(nn.Module):
    """darknet_meta.Darknet restated (metain_type=2: support input = RGB + mask)."""

    def __init__(self, det_blocks, learnet_blocks):
        super().__init__()
        self.blocks = det_blocks
        self.learnet_blocks = learnet_blocks
        self.models = create_network(det_blocks)
        self.learnet_models = create_network(learnet_blocks)
        r = det_blocks[-1]
        self.anchors = [float(a) for a in r['anchors'].split(',')]
        self.num_anchors = int(r['num'])
        self.num_classes = int(r['classes'])

    def meta_forward(self, metax, mask):  # darknet_meta.py:107-128 (feat_layer=0)
        metax = torch.cat([metax, mask], dim=1)
        for model in self.learnet_models:
            metax = model(metax)
        return [metax]

    def detect_forward(self, x, dynamic_weights):
        return _run_blocks(self.blocks, self.models, x, dynamic_weights)

    def forward(self, x, metax, mask):
        return self.detect_forward(x, self.meta_forward(metax, mask))


class PlainDarknet(nn.Module):
    """darknet.Darknet restated."""

    def __init__(self, blocks):
        super().__init__()
        self.blocks = blocks
        self.models = create_network(blocks)
        r = blocks[-1]
        self.anchors = [float(a) for a in r['anchors'].split(',')]
        self.num_anchors = int(r['num'])
        self.num_classes = int(r['classes'])

    def forward(self, x):
        return _run_blocks(self.blocks, self.models, x)
