"""`Darknet(darknet_file, learnet_file)`: the meta-detector of the reference
(darknet_meta.py:86-482) on the B200-native engine.

Kept from the reference so that train_meta.py / valid_ensemble.py drop in:
constructor arguments (cfg path or parsed block list, darknet_meta.py:87-90),
attributes `.blocks .learnet_blocks .models .learnet_models .loss .width .height
.anchors .num_anchors .anchor_step .num_classes .header .seen`, methods
`forward(x, metax, mask, ids=None)`, `meta_forward`, `detect_forward`,
`load_weights`, `save_weights(outfile, cutoff=0)`, `print_network`,
`is_dynamic`, the ModuleList / Sequential(conv{i}, bn{i}, leaky{i}) tree (so
`parameters()` order, `state_dict()` keys and the Darknet weight stream are the
same), and `nn.Module` behaviour (.cuda(), .train(), .eval()).

New underneath: nn.Conv2d / nn.BatchNorm2d are only parameter containers here;
the forward and backward passes are executed by engine.NetRunner with
hand-written sm_100a kernels (libfsdet.so).  CUDA only - no CPU fallback.
"""
import numpy as np
import torch
import torch.nn as nn

from .cfg import cfg, parse_cfg, load_conv, load_conv_bn, save_conv, save_conv_bn
from .dynamic_conv import dynamic_conv2d
from .engine import NetRunner, run_network
from .pooling import GlobalMaxPool2d, MaxPoolStride1, MaxPool2x2, Reorg, EmptyModule
from .region_loss import RegionLossV2, RegionLoss


def create_network(owner, blocks, loss_cls):
    """Block list -> nn.ModuleList (darknet_meta.py:208-353 / darknet.py:134-245)."""
    models = nn.ModuleList()
    prev_filters = 3
    out_filters = []
    conv_id = 0
    dynamic_count = 0
    for block in blocks:
        t = block['type']
        if t == 'net' or t == 'learnet':
            prev_filters = int(block['channels'])
            continue
        elif t == 'convolutional':
            conv_id = conv_id + 1
            batch_normalize = int(block['batch_normalize'])
            filters = int(block['filters'])
            kernel_size = int(block['size'])
            stride = int(block['stride'])
            pad = (kernel_size - 1) // 2 if int(block['pad']) else 0
            activation = block['activation']
            bias = bool(int(block['bias'])) if 'bias' in block else True
            if owner.is_dynamic(block):
                partial = int(block['partial']) if 'partial' in block else None
                Conv2d = dynamic_conv2d(dynamic_count == 0, partial=partial)
                dynamic_count += 1
            else:
                Conv2d = nn.Conv2d
            if 'groups' in block and int(block['groups']) != 1:
                raise NotImplementedError('grouped convolution')
            model = nn.Sequential()
            if batch_normalize:
                model.add_module('conv{0}'.format(conv_id), Conv2d(prev_filters, filters, kernel_size, stride, pad, bias=False))
                model.add_module('bn{0}'.format(conv_id), nn.BatchNorm2d(filters))
            else:
                model.add_module('conv{0}'.format(conv_id), Conv2d(prev_filters, filters, kernel_size, stride, pad, bias=bias))
            if activation == 'leaky':
                model.add_module('leaky{0}'.format(conv_id), nn.LeakyReLU(0.1, inplace=True))
            elif activation == 'relu':
                raise NotImplementedError('relu activation')
            prev_filters = filters
            out_filters.append(prev_filters)
            models.append(model)
        elif t == 'maxpool':
            pool_size = int(block['size'])
            stride = int(block['stride'])
            models.append(MaxPool2x2(pool_size, stride) if stride > 1 else MaxPoolStride1())
            out_filters.append(prev_filters)
        elif t == 'reorg':
            stride = int(block['stride'])
            prev_filters = stride * stride * prev_filters
            out_filters.append(prev_filters)
            models.append(Reorg(stride))
        elif t == 'route':
            layers = block['layers'].split(',')
            ind = len(models)
            layers = [int(i) if int(i) > 0 else int(i) + ind for i in layers]
            if len(layers) == 1:
                prev_filters = out_filters[layers[0]]
            elif len(layers) == 2:
                assert layers[0] == ind - 1
                prev_filters = out_filters[layers[0]] + out_filters[layers[1]]
            out_filters.append(prev_filters)
            models.append(EmptyModule())
        elif t == 'region':
            loss = loss_cls()
            anchors = block['anchors'].split(',')
            loss.anchors = [float(i) for i in anchors]
            loss.num_classes = int(block['classes'])
            loss.num_anchors = int(block['num'])
            loss.anchor_step = len(loss.anchors) // loss.num_anchors
            loss.object_scale = float(block['object_scale'])
            loss.noobject_scale = float(block['noobject_scale'])
            loss.class_scale = float(block['class_scale'])
            loss.coord_scale = float(block['coord_scale'])
            out_filters.append(prev_filters)
            models.append(loss)
        elif t == 'globalmax':
            out_filters.append(prev_filters)
            models.append(GlobalMaxPool2d())
        else:
            raise NotImplementedError('block type %s is not on the supported hot path' % t)
    return models


def _copy_blocks(blocks):
    return [dict(b) for b in blocks]


class Darknet(nn.Module):
    def __init__(self, darknet_file, learnet_file):
        super(Darknet, self).__init__()
        self.blocks = darknet_file if isinstance(darknet_file, list) else parse_cfg(darknet_file)
        self.learnet_blocks = learnet_file if isinstance(learnet_file, list) else parse_cfg(learnet_file)
        self.models = create_network(self, self.blocks, RegionLossV2)
        self.learnet_models = create_network(self, self.learnet_blocks, RegionLossV2)
        self.loss = self.models[len(self.models) - 1]

        self.width = int(self.blocks[0]['width'])
        self.height = int(self.blocks[0]['height'])

        if self.blocks[(len(self.blocks) - 1)]['type'] == 'region':
            self.anchors = self.loss.anchors
            self.num_anchors = self.loss.num_anchors
            self.anchor_step = self.loss.anchor_step
            self.num_classes = self.loss.num_classes

        self.header = torch.IntTensor([0, 0, 0, 0])
        self.seen = 0
        if int(self.learnet_blocks[0].get('feat_layer', 0)) != 0:
            raise NotImplementedError('[learnet] feat_layer != 0 (shared trunk layers) is not used by the shipped cfgs')
        self._det = NetRunner(self.blocks, self.models)
        self._ler = NetRunner(self.learnet_blocks, self.learnet_models)

    # ------------------------------------------------------------------ forward
    def _params(self, models):
        return [p for p in models.parameters()]

    def meta_forward(self, metax, mask):
        """Support branch (darknet_meta.py:107-128): returns [ [n_cls, C, 1, 1] ]."""
        if cfg.metain_type in [2, 3]:
            inputs = [metax, mask]
        else:
            inputs = [metax]
        dw = run_network(self._ler, inputs, None, self._params(self.learnet_models), self.training)
        return [dw]

    def detect_forward(self, x, dynamic_weights):
        """Query branch (darknet_meta.py:130-195): [B,3,H,W] -> [B*n_cls, A*(5+nC), H/32, W/32]."""
        self.loss = None  # the reference resets it here (darknet_meta.py:134)
        dw = dynamic_weights[0] if isinstance(dynamic_weights, (list, tuple)) else dynamic_weights
        return run_network(self._det, [x], dw, self._params(self.models), self.training)

    def forward(self, x, metax, mask, ids=None):
        dynamic_weights = self.meta_forward(metax, mask)
        return self.detect_forward(x, dynamic_weights)

    def print_network(self):
        for name, blocks in (('detector', self.blocks), ('reweighting net', self.learnet_blocks)):
            print('--- %s' % name)
            for i, b in enumerate(blocks):
                print('%3d %-14s %s' % (i - 1, b['type'], ' '.join('%s=%s' % kv for kv in b.items() if kv[0] != 'type')))

    def is_dynamic(self, block):
        return 'dynamic' in block and int(block['dynamic']) == 1

    # --------------------------------------------------------------- weight IO
    def load_weights(self, weightfile):
        """Darknet weight stream (darknet_meta.py:355-411): detector blocks first,
        then the reweighting net; loading stops silently when the buffer is
        exhausted (that is how darknet19_448.conv.23 initialises only the trunk)."""
        with open(weightfile, 'rb') as fp:
            header = np.fromfile(fp, count=4, dtype=np.int32)
            self.header = torch.from_numpy(header)
            self.seen = int(self.header[3])
            buf = np.fromfile(fp, dtype=np.float32)
        start = 0
        for blocks, models in [(self.blocks, self.models), (self.learnet_blocks, self.learnet_models)]:
            ind = -2
            for block in blocks:
                if start >= buf.size:
                    break
                ind = ind + 1
                if block['type'] == 'convolutional':
                    model = models[ind]
                    if self.is_dynamic(block) and model[0].weight is None:
                        continue
                    if int(block['batch_normalize']):
                        start = load_conv_bn(buf, start, model[0], model[1])
                    else:
                        start = load_conv(buf, start, model[0])

    def save_weights(self, outfile, cutoff=0):
        """darknet_meta.py:413-479."""
        if cutoff <= 0:
            cutoff = len(self.blocks) - 1 + len(self.learnet_blocks)
        with open(outfile, 'wb') as fp:
            self.header[3] = int(self.seen)
            self.header.numpy().tofile(fp)
            ind = -1
            for blockId in range(1, cutoff + 1):
                if blockId >= len(self.blocks):
                    if blockId == len(self.blocks):
                        ind = -2
                    blockId = blockId - len(self.blocks)
                    blocks, models = self.learnet_blocks, self.learnet_models
                else:
                    blocks, models = self.blocks, self.models
                ind = ind + 1
                block = blocks[blockId]
                if block['type'] == 'convolutional':
                    model = models[ind]
                    if self.is_dynamic(block) and model[0].weight is None:
                        continue
                    if int(block['batch_normalize']):
                        save_conv_bn(fp, model[0], model[1])
                    else:
                        save_conv(fp, model[0])
