#!/usr/bin/env python
"""Command-line front end with the reference driver's arguments (train_meta.py:1-6):

    python tools/train_meta_b200.py datacfg darknetcfg learnetcfg weightfile            # one GPU
    torchrun --nproc-per-node 8 tools/train_meta_b200.py datacfg darknetcfg learnetcfg weightfile

The `.data` file drives everything as in the reference: `train` (image list or class dict), `meta` (support dict),
`novel` / `novelid` (base / novel split), `neg`, `tuning` (+ `max_epoch`, `repeat`, `dynamic`) - the image list comes
from lists.build_dataset, the support index from lists.support_index, the step from trainer.MetaTrainer (CUDA-graph
replay, one graph per multi-scale input size).  Under torchrun every rank builds the same lists with the same seeds,
takes its slice of each global batch, and rank 0's parameters are broadcast once before the first step.
What the reference's script does beyond that - the in-training test() pass - is not wired here;
`fewshot_detection_b200.evaluate` / `valid` / `voc_eval` are the evaluation entry points.
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


def broadcast_parameters(model, src=0):
    """Identical replicas: the gradient all-reduce keeps replicas in sync only if they START in sync.  A weight file
    may initialise just the trunk (darknet19_448.conv.23 stops after 23 layers, darknet_meta.py:367-368), every other
    tensor is per-process random - so rank `src`'s parameters and BN buffers are broadcast once."""
    import torch.distributed as dist
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src)
        chk = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
        ref = chk.clone()
        dist.broadcast(ref, src)
        if not torch.equal(chk, ref):
            raise RuntimeError('replicas differ after the parameter broadcast')


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
    from fewshot_detection_b200 import trainer as T, lists as LS
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
    if os.path.exists(sys.argv[4]):
        model.load_weights(sys.argv[4])
    else:
        logging('weight file %s not found: training from the random initialisation' % sys.argv[4])
    model = model.cuda()
    if world > 1:
        broadcast_parameters(model)
    classes = cfg.base_classes
    factor = T.lr_factor(cfg.neg_ratio, len(classes))
    hp = T.sgd_hyper_parameters(float(net_options['learning_rate']), float(net_options['momentum']), float(net_options['decay']),
                                batch_size, factor)
    optimizer = FusedSGD(model.parameters(), **hp)
    reducer = GradAllReducer(model) if world > 1 else None

    seed = int(os.environ.get('FSDET_SEED', str(int.from_bytes(os.urandom(4), 'little')) if world == 1 else '0'))
    import random
    random.seed(seed)                 # every rank must build the same lists and draw the same sizes
    np.random.seed(seed % (2 ** 32))
    trainlist = LS.build_dataset(data_options)
    nsamples = len(trainlist)
    processed, init_epoch, max_epochs = T.epoch_plan(model.seen, nsamples, batch_size, int(net_options['max_batches']),
                                                     cfg.tuning, cfg.get('max_epoch'), cfg.repeat)
    backupdir = cfg.get('backup') or data_options.get('backup', 'backup')     # cfg.py:133-145 names it after the run's switches
    if rank == 0 and not os.path.exists(backupdir):
        os.makedirs(backupdir)

    def make_train_batcher(seen):
        # every rank walks the same shuffled list and takes its own slice of each global batch
        lines = trainlist if world == 1 else \
            [trainlist[i] for b in range(0, nsamples - batch_size + 1, batch_size) for i in range(b + rank * per_rank, b + (rank + 1) * per_rank)]
        return DetectionBatcher(lines, shape=(model.width, model.height), shuffle=False, train=True, seen=seen,
                                batch_size=per_rank, seen_step=world)

    def make_meta_batcher():
        cfg.num_gpus = 1              # one process per GPU: each rank draws its own n_cls support images per step
        metalines, inds = LS.support_index(data_options['meta'], classes, LS.support_batches_per_epoch(train=True),
                                           shuffle=cfg.randmeta)
        return MetaBatcher(metalines, inds, classes=classes, train=True)

    tr = T.MetaTrainer(model, optimizer, float(net_options['learning_rate']) / factor, batch_size, steps, scales,
                       make_train_batcher, make_meta_batcher, backupdir=backupdir if rank == 0 else None,
                       save_interval=cfg.save_interval, reducer=reducer, world=world, processed_batches=processed,
                       log=logging if rank == 0 else (lambda *_: None))
    model.loss.verbose = rank == 0
    tr.fit(init_epoch, max_epochs)
    if world > 1:
        if tr.graphed is not None:
            tr.graphed.entries.clear()      # graphs that captured NCCL work go before their communicator
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
