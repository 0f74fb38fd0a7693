#!/usr/bin/env python
"""Command-line front end with the reference driver's arguments (train_meta.py:1-6):

    python tools/train_meta_b200.py datacfg darknetcfg learnetcfg weightfile            # one GPU
    torchrun --nproc-per-node 8 tools/train_meta_b200.py datacfg darknetcfg learnetcfg weightfile

Base training from plain image-list files (`train = <list>` in the .data file, one image path per line; labels
are found like listDataset.get_labpath) and a support dictionary (`meta = <file>` with `class list-file` lines).
What the reference's script does beyond that - few-shot list construction for fine-tuning (dataset.build_dataset /
build_fewset), the in-training test() pass - is not wired here; `fewshot_detection_b200.evaluate` / `valid` /
`voc_eval` are the evaluation entry points.

STATUS: written against tests/test_trainer_cpu.py (loop logic with a stub model) and the batcher tests; it has not
been run end to end (no dataset in the build container, no GPU minutes left in round 1).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def read_list(path):
    with open(path, 'r') as f:
        return [l.rstrip() for l in f.readlines() if l.strip()]


def read_metadict(path, classes):
    """`class list-file` lines (MetaDataset.__init__, dataset.py:316-336) -> per-class image lists."""
    pairs = {}
    for line in read_list(path):
        p = line.split()
        if len(p) == 4:
            p = [p[0] + ' ' + p[1], p[2] + ' ' + p[3]]
        pairs[p[0]] = p[1]
    return [read_list(pairs[c]) for c in classes]


def meta_inds(metalines, nbatch):
    """MetaDataset.inds for training (dataset.py:331-340): per class `nbatch` random picks, interleaved class by class."""
    per_class = [list(zip([i] * nbatch, np.random.choice(range(len(lines)), nbatch).tolist())) for i, lines in enumerate(metalines)]
    return sum(list(zip(*per_class)), ())


def main():
    if len(sys.argv) != 5:
        print('Usage:')
        print('python tools/train_meta_b200.py datacfg darknetcfg learnetcfg weightfile')
        return 1
    from fewshot_detection_b200.cfg import cfg, parse_cfg
    from fewshot_detection_b200.utils import read_data_cfg, logging
    from fewshot_detection_b200.darknet_meta import Darknet
    from fewshot_detection_b200.optim import FusedSGD
    from fewshot_detection_b200.distributed import GradAllReducer
    from fewshot_detection_b200.dataset import DetectionBatcher, MetaBatcher
    from fewshot_detection_b200 import trainer as T
    import torch.distributed as dist

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    data_options = read_data_cfg(sys.argv[1])
    darknetcfg, learnetcfg = parse_cfg(sys.argv[2]), parse_cfg(sys.argv[3])
    net_options, meta_options = darknetcfg[0], learnetcfg[0]
    cfg.config_data(data_options)
    cfg.config_meta(meta_options)
    cfg.config_net(net_options)
    batch_size = int(net_options['batch'])                      # GLOBAL batch, as in the reference
    per_rank = batch_size // world
    steps = [float(s) for s in net_options['steps'].split(',')]
    scales = [float(s) for s in net_options['scales'].split(',')]

    model = Darknet(darknetcfg, learnetcfg)
    model.load_weights(sys.argv[4])
    model = model.cuda()
    classes = cfg.base_classes
    factor = T.lr_factor(cfg.neg_ratio, len(classes))
    hp = T.sgd_hyper_parameters(float(net_options['learning_rate']), float(net_options['momentum']), float(net_options['decay']),
                                batch_size, factor)
    optimizer = FusedSGD(model.parameters(), **hp)
    reducer = GradAllReducer(model) if world > 1 else None

    trainlist = read_list(data_options['train'])
    nsamples = len(trainlist)
    metalines = read_metadict(data_options['meta'], classes)
    processed, init_epoch, max_epochs = T.epoch_plan(model.seen, nsamples, batch_size, int(net_options['max_batches']),
                                                     cfg.tuning, cfg.get('max_epoch'), cfg.repeat)
    backupdir = data_options.get('backup', 'backup')
    if rank == 0 and not os.path.exists(backupdir):
        os.makedirs(backupdir)

    def make_train_batcher(seen):
        # every rank walks the same shuffled list and takes its own slice of each global batch
        lines = trainlist if world == 1 else \
            [trainlist[i] for b in range(0, nsamples - batch_size + 1, batch_size) for i in range(b + rank * per_rank, b + (rank + 1) * per_rank)]
        return DetectionBatcher(lines, shape=(model.width, model.height), shuffle=False, train=True, seen=seen,
                                batch_size=per_rank, num_workers=int(data_options['num_workers']))

    def make_meta_batcher():
        nbatch = 500 * 64 // batch_size * (4 if cfg.get('data') == 'coco' else 1)
        return MetaBatcher(metalines, meta_inds(metalines, nbatch), classes=classes, train=True)

    tr = T.MetaTrainer(model, optimizer, float(net_options['learning_rate']) / factor, batch_size, steps, scales,
                       make_train_batcher, make_meta_batcher, backupdir=backupdir if rank == 0 else None,
                       save_interval=cfg.save_interval, reducer=reducer, world=world, processed_batches=processed,
                       log=logging if rank == 0 else (lambda *_: None))
    model.loss.verbose = rank == 0
    tr.fit(init_epoch, max_epochs)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
