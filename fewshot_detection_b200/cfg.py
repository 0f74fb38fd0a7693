"""Config front-end: Darknet `.cfg` parser, the process-global `cfg` options
object and the Darknet weight-stream (de)serialisers.

Mirrors the reference's cfg.py surface that the hot path and its driver use:
`parse_cfg` (cfg.py:198-228), the `cfg` singleton with `config_data /
config_meta / config_net` (cfg.py:70-195) and `load_conv / load_conv_bn /
save_conv / save_conv_bn` (cfg.py:411-470).  `cfg.neg_ratio`, `cfg.metayolo`,
`cfg.metain_type`, `cfg.max_boxes` are read at call time by region_loss and
darknet_meta exactly as in the reference.
"""
import numpy as np
import torch


class _Options(dict):
    """Attribute-style dict (the reference uses easydict.EasyDict)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


cfg = _Options()
cfg.voc_classes = ["aeroplane", "bicycle", "bird", "boat", "bottle", "bus", "car", "cat", "chair", "cow",
                   "diningtable", "dog", "horse", "motorbike", "person", "pottedplant", "sheep", "sofa", "train",
                   "tvmonitor"]
cfg.max_boxes = 50        # maximum number of boxes per class row (cfg.py:29)
cfg.neg_ratio = 'full'    # cfg.py:31
cfg.tuning = False
cfg.metayolo = True
cfg.repeat = 1
cfg.save_interval = 10
cfg.multiscale = True
cfg.metain_type = 2       # 2 = support image + mask channel (cfg.py:37-38)
# defaults of the few-shot bookkeeping that image.fill_truth_detection(_meta) reads (cfg.py:103-145): all 20 VOC
# classes are base classes until a .data file says otherwise
cfg.classes = cfg.voc_classes
cfg.base_classes = list(cfg.voc_classes)
cfg.base_ids = list(range(len(cfg.voc_classes)))
cfg.novel_classes = []
cfg.novel_ids = []
cfg.yolo_joint = False
cfg.metaids = []


def _configure_net(netopt):
    cfg.height = int(netopt['height'])
    cfg.width = int(netopt['width'])
    cfg.batch_size = int(netopt['batch'])


def _configure_meta(metaopt):
    """cfg.py:155-190: mask size and the support net's input channel count."""
    cfg.meta_height = int(metaopt['height'])
    cfg.meta_width = int(metaopt['width'])
    factor = int(metaopt['feat_layer'])
    if factor == 0:
        cfg.mask_height, cfg.mask_width = cfg.meta_height, cfg.meta_width
        chans = {1: 3, 2: 4, 3: 7, 4: 6}
    elif factor == 4:
        cfg.mask_height, cfg.mask_width = cfg.meta_height // factor, cfg.meta_width // factor
        chans = {1: 64, 2: 65, 3: 129, 4: 128}
    else:
        raise NotImplementedError('Feat layer not found{}'.format(factor))
    if cfg.metain_type not in chans:
        raise NotImplementedError('Meta input type not found: {}'.format(cfg.metain_type))
    metaopt['channels'] = chans[cfg.metain_type]


def _configure_data(dataopt):
    """The subset of cfg.py:70-147 that feeds the hot path (class lists, neg
    ratio, metayolo / metain_type switches).  Dataset bookkeeping (few-shot id
    lists, backup dir naming) belongs to the out-of-scope input pipeline."""
    cfg.data = dataopt.get('data', 'voc')
    if 'scale' in dataopt:
        cfg.multiscale = int(dataopt['scale'])
    if 'metain_type' in dataopt:
        cfg.metain_type = int(dataopt['metain_type'])
    if 'tuning' in dataopt:
        cfg.tuning = bool(int(dataopt['tuning']))
    neg = dataopt['neg'] if 'neg' in dataopt else cfg.neg_ratio
    if isinstance(neg, str) and neg.isdigit():
        neg = float(neg)
        if neg.is_integer():
            neg = int(neg)
    cfg.neg_ratio = neg
    if 'metayolo' in dataopt:
        cfg.metayolo = bool(int(dataopt['metayolo']))
    if 'gpus' in dataopt:
        cfg.num_gpus = len(dataopt['gpus'].split(','))


cfg.config_data = _configure_data
cfg.config_meta = _configure_meta
cfg.config_net = _configure_net


def parse_cfg(cfgfile):
    """Darknet `.cfg` -> list of block dicts with string values (cfg.py:198-228):
    `[name]` opens a block with block['type'] = name; a `type=` key inside a
    block is stored as '_type'; convolutional blocks default
    batch_normalize to the int 0; blank lines and `#` lines are skipped."""
    blocks = []
    block = None
    with open(cfgfile, 'r') as fp:
        for raw in fp:
            line = raw.rstrip()
            if line == '' or line[0] == '#':
                continue
            if line[0] == '[':
                if block:
                    blocks.append(block)
                block = {'type': line.lstrip('[').rstrip(']')}
                if block['type'] == 'convolutional':
                    block['batch_normalize'] = 0
            else:
                key, value = line.split('=')
                key = key.strip()
                if key == 'type':
                    key = '_type'
                block[key] = value.strip()
    if block:
        blocks.append(block)
    return blocks


# ---- Darknet weight stream (cfg.py:411-470) --------------------------------
# header int32[4] = (major, minor, revision, seen), then a flat float32 stream;
# conv+BN: bn.bias, bn.weight, running_mean, running_var, conv.weight (OIHW);
# conv without BN: bias, weight.

def _take(buf, start, t):
    n = t.numel()
    src = torch.from_numpy(np.ascontiguousarray(buf[start:start + n])).view(t.shape)
    with torch.no_grad():
        t.copy_(src)  # strided (channels_last) destinations receive logical OIHW order
    return start + n


def load_conv(buf, start, conv_model):
    if conv_model.bias is not None:
        start = _take(buf, start, conv_model.bias.data)
    return _take(buf, start, conv_model.weight.data)


def load_conv_bn(buf, start, conv_model, bn_model):
    start = _take(buf, start, bn_model.bias.data)
    start = _take(buf, start, bn_model.weight.data)
    start = _take(buf, start, bn_model.running_mean)
    start = _take(buf, start, bn_model.running_var)
    return _take(buf, start, conv_model.weight.data)


def _put(fp, t):
    # .contiguous() yields logical (OIHW) order whatever the storage format
    t.detach().to('cpu', torch.float32).contiguous().numpy().tofile(fp)


def save_conv(fp, conv_model):
    if conv_model.bias is not None:
        _put(fp, conv_model.bias.data)
    _put(fp, conv_model.weight.data)


def save_conv_bn(fp, conv_model, bn_model):
    _put(fp, bn_model.bias.data)
    _put(fp, bn_model.weight.data)
    _put(fp, bn_model.running_mean)
    _put(fp, bn_model.running_var)
    _put(fp, conv_model.weight.data)
