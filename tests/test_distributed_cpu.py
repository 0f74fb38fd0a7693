"""Host-side logic of the N>1 path on CPU: world_size-2 gloo processes exercise the flat-gradient bucket
all-reducer (the only collective of the hot path) - no GPU, no compute kernels."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, overlap, results):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location('fsdet_distributed', os.path.join(ROOT, 'fewshot_detection_b200', 'distributed.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        torch.manual_seed(0)
        m = nn.Sequential(nn.Conv2d(4, 8, 3), nn.BatchNorm2d(8), nn.Conv2d(8, 30, 1), nn.Conv2d(30, 7, 3))
        red = mod.GradAllReducer(m, bucket_mb=0.001)     # tiny buckets -> several collectives
        red.overlap = overlap
        assert len(red.buckets) > 1
        params = list(m.parameters())
        # every gradient is a view into the flat buffer, 128-byte aligned, OHWI storage for conv weights
        for p in params:
            assert p.grad.data_ptr() % 128 == red.flat.data_ptr() % 128
            assert p.grad.shape == p.shape
            if p.dim() == 4:
                assert p.grad.permute(0, 2, 3, 1).is_contiguous()
        for step in range(2):
            red.begin_step()
            for i, p in enumerate(params):
                p.grad.fill_(float((rank + 1) * (i + 1) + step))
            for p in reversed(params):           # backward produces gradients last layer first
                red.grad_ready(p)
            red.finish()
            for i, p in enumerate(params):
                want = sum((r + 1) * (i + 1) + step for r in range(world))
                assert torch.all(p.grad == want), (rank, i, p.grad.flatten()[:4], want)
        results[rank] = 'ok'
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('overlap', [True, False])
def test_grad_allreducer_gloo_world2(overlap):
    world = 2
    port = 29000 + (os.getpid() % 2000) + (1 if overlap else 0)
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, overlap, results), nprocs=world, join=True)
    assert dict(results) == {0: 'ok', 1: 'ok'}
