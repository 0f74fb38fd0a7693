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


COCO_NAMES = [
    'person', 'bicycle', 'car', 'motorbike', 'aeroplane', 'bus', 'train', 'truck', 'boat', 'traffic light', 'fire hydrant',
    'stop sign', 'parking meter', 'bench', 'bird', 'cat', 'dog', 'horse', 'sheep', 'cow', 'elephant', 'bear', 'zebra',
    'giraffe', 'backpack', 'umbrella', 'handbag', 'tie', 'suitcase', 'frisbee', 'skis', 'snowboard', 'sports ball', 'kite',
    'baseball bat', 'baseball glove', 'skateboard', 'surfboard', 'tennis racket', 'bottle', 'wine glass', 'cup', 'fork',
    'knife', 'spoon', 'bowl', 'banana', 'apple', 'sandwich', 'orange', 'broccoli', 'carrot', 'hot dog', 'pizza', 'donut',
    'cake', 'chair', 'sofa', 'pottedplant', 'bed', 'diningtable', 'toilet', 'tvmonitor', 'laptop', 'mouse', 'remote',
    'keyboard', 'cell phone', 'microwave', 'oven', 'toaster', 'sink', 'refrigerator', 'book', 'clock', 'vase', 'scissors',
    'teddy bear', 'hair drier', 'toothbrush']


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
cfg.coco_classes = COCO_NAMES      # data/coco.names of the reference (cfg.py:25-27), VOC spellings for the shared classes
cfg.vocids_in_coco = [COCO_NAMES.index(c) for c in cfg.voc_classes]
cfg.cocoonly_ids = [i for i in range(len(COCO_NAMES)) if i not in cfg.vocids_in_coco]
cfg.randmeta = False
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


def read_names(path):
    """One class name per line (data/voc.names, data/coco.names; cfg.py:11-17)."""
    with open(path) as f:
        return [l.strip() for l in f.readlines()]


def novel_classes_of(spec, novelid):
    """cfg.py:55-63: `novel` is either a comma-separated class list or a .txt file whose line `novelid` is one;
    novelid 'None' selects no novel classes."""
    if not spec.endswith('txt'):
        return spec.split(',')
    if novelid == 'None':
        return []
    with open(spec) as f:
        return f.readlines()[int(novelid)].strip().split(',')


def fewshot_image_ids(metafile, base_classes):
    """cfg.py:41-53: image ids listed by the per-class files of a few-shot dict (`<class> <listfile>` lines), base
    classes only, sorted and de-duplicated."""
    with open(metafile) as f:
        rows = [l.rstrip().split() for l in f.readlines()]
    lines = []
    for row in rows:
        if row and row[0] in base_classes:
            with open(row[-1]) as g:
                lines.extend(g.readlines())
    return [l.split('/')[-1].split('.')[0] for l in sorted(set(lines))]


def _save_interval_for(max_epoch, repeat, data):
    """cfg.py:88-98: fine-tuning saves more often the fewer (epoch / repeat) passes it makes."""
    passes = max_epoch / repeat
    interval = 10
    for limit, every in ((20, 1), (50, 2), (100, 5)):
        if passes <= limit:
            interval = every
            break
    return 2 if data == 'coco' else interval


def _backup_dir(dataopt, novelid):
    """cfg.py:133-145: the backup directory name encodes the run's switches."""
    name = dataopt['backup']
    if not cfg.multiscale:
        name += 'fix'
    if cfg.metain_type != 2:
        head, _, tail = name.partition('_')
        name = head + 'in{}'.format(cfg.metain_type) + (('_' + tail) if tail else '')
    name += '_novel{}'.format(novelid)
    if cfg.metayolo:
        name += '_neg{}'.format(cfg.neg_ratio)
    if cfg.randmeta:
        name += '_rand'
    return name


def _configure_data(dataopt):
    """The `.data` file -> process-global options (cfg.py:70-147): class lists and the base / novel split that the
    few-shot protocol rests on (base training must NOT see the novel classes), negative-row ratio, fine-tuning
    schedule, backup directory name.  Keys the file does not have keep their defaults."""
    data = dataopt.get('data', 'voc')
    cfg.data = data
    if data == 'voc':
        cfg.classes = cfg.voc_classes
    elif data == 'coco':
        cfg.classes = cfg.coco_classes
        cfg.save_interval = 2
    else:
        raise NotImplementedError('Data type {} not found'.format(data))
    for key, conv, dst in (('scale', int, 'multiscale'), ('metain_type', int, 'metain_type')):
        if key in dataopt:
            cfg[dst] = conv(dataopt[key])
    if 'tuning' in dataopt:
        cfg.tuning = bool(int(dataopt['tuning']))
        cfg.max_epoch = int(dataopt.get('max_epoch', 500))
        cfg.repeat = int(dataopt.get('repeat', 100))
        cfg.save_interval = _save_interval_for(cfg.max_epoch, cfg.repeat, data)
        if 'meta' in dataopt:
            cfg.shot = int(dataopt['meta'].split('.')[0].split('_')[-1].replace('shot', ''))
    novelid = dataopt.get('novelid', 'None')
    cfg.novelid = novelid
    cfg.novel_classes = novel_classes_of(dataopt['novel'], novelid) if 'novel' in dataopt else []
    unknown = [c for c in cfg.novel_classes if c not in cfg.classes]
    if unknown:
        raise ValueError('novel classes %r are not classes of the %s data set' % (unknown, data))
    if cfg.tuning:
        cfg.base_classes = list(cfg.classes)          # fine-tuning sees every class (cfg.py:105-113)
    else:
        cfg.base_classes = [c for c in cfg.classes if c not in cfg.novel_classes]
    cfg.base_ids = [cfg.classes.index(c) for c in cfg.base_classes]
    cfg.novel_ids = [cfg.classes.index(c) for c in cfg.novel_classes]
    cfg._real_base_ids = [i for i in range(len(cfg.classes)) if i not in cfg.novel_ids]
    if 'gpus' in dataopt:
        cfg.num_gpus = len(dataopt['gpus'].split(','))
    neg = dataopt['neg'] if 'neg' in dataopt else cfg.neg_ratio
    if isinstance(neg, str) and neg.isdigit():
        neg = float(neg)
        if neg.is_integer():
            neg = int(neg)
    cfg.neg_ratio = neg
    cfg.randmeta = bool(int(dataopt['rand'])) if 'rand' in dataopt else False
    if 'metayolo' in dataopt:
        cfg.metayolo = bool(int(dataopt['metayolo']))
    if 'backup' in dataopt:
        cfg.backup = _backup_dir(dataopt, novelid)
    cfg.yolo_joint = int(dataopt['joint']) if 'joint' in dataopt else False
    if cfg.yolo_joint:
        cfg.metaids = fewshot_image_ids(dataopt['meta'], cfg.base_classes)
        cfg.backup = cfg.get('backup', '') + '_joint{}'.format(
            int(dataopt['meta'].split('.')[0].split('_')[-1].replace('shot', '')))


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
