"""Worker of tests/test_gpu_multi.py (one process per GPU, launched with torch.distributed.run).  Checks, on 2+ GPUs:
  1. the bucketed, in-backward NCCL all-reduce leaves in the flat gradient buffer the SUM of the ranks' single-GPU
     gradients (each rank recomputes every shard locally without any collective and compares);
  2. the CUDA-graph step with the all-reduce captured INSIDE the graph follows the eager multi-rank loop (losses,
     parameters) and keeps the replicas identical.
Prints 'MULTI_OK' on every rank on success."""
import contextlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def relt(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    from fewshot_detection_b200 import netcfg
    from fewshot_detection_b200.darknet_meta import Darknet
    from fewshot_detection_b200.optim import FusedSGD
    from fewshot_detection_b200.distributed import GradAllReducer
    from fewshot_detection_b200.graph import GraphedTrainStep
    from seeding import seeded_init, synth_targets, synth_masks
    det, ler = netcfg.mini_dynamic_blocks(128, 8), netcfg.mini_reweighting_blocks(64, 8, 256)
    bs, cs = 4, 3

    def batch(it, r):
        g = torch.Generator().manual_seed(1000 * r + it)
        x = torch.rand(bs, 3, 128, 128, generator=g).to(dev)
        metax = torch.rand(cs, 3, 64, 64, generator=g).to(dev)
        return x, metax, torch.from_numpy(synth_masks(cs, 64, 50 * r + it)).to(dev), torch.from_numpy(synth_targets(bs, cs, 70 * r + it, max_gt=3))

    def model():
        with contextlib.redirect_stdout(sys.stderr):
            m = Darknet([dict(b) for b in det], [dict(b) for b in ler])
        seeded_init(m, 3)
        m = m.to(dev).train()
        L = m.models[len(m.models) - 1]
        L.verbose = False
        L.seen = 20000
        return m, L

    # ---- 1. all-reduced flat buffer == sum over ranks of the single-GPU gradients
    m, L = model()
    red = GradAllReducer(m, bucket_mb=0.05)          # many small buckets: several collectives launched during backward
    assert len(red.buckets) > 3
    red.begin_step()
    L(m(*batch(0, rank)[:3]), batch(0, rank)[3]).backward()
    red.finish()
    torch.cuda.synchronize()
    got = red.flat.clone()
    want = torch.zeros_like(got)
    m2, L2 = model()
    red2 = GradAllReducer(m2, bucket_mb=0.05)
    red2.world = 1                                   # same flat layout, no collective
    for r in range(world):
        red2.begin_step()
        L2(m2(*batch(0, r)[:3]), batch(0, r)[3]).backward()
        red2.finish()
        want += red2.flat
    torch.cuda.synchronize()
    e = relt(got, want)
    assert e < 1e-5, ('all-reduced gradients differ from the sum of the shards', e)

    # ---- 2. graph step (collective inside the graph) == eager multi-rank loop; replicas stay identical
    runs = []
    for graph in (False, True):
        m, L = model()
        opt = FusedSGD(m.parameters(), lr=1e-3 / world, momentum=0.9, dampening=0, weight_decay=5e-4)
        red = GradAllReducer(m, bucket_mb=0.05)
        gs = GraphedTrainStep(m, L, opt, red) if graph else None
        losses = []
        for it in range(5):
            x, metax, mask, tgt = batch(it, rank)
            L.seen += bs * world
            if graph:
                losses.append(gs(x, metax, mask, tgt).item())
            else:
                red.overlap = True
                red.begin_step()
                loss = L(m(x, metax, mask), tgt)
                loss.backward()
                red.finish()
                opt.step()
                losses.append(loss.item())
        if graph:
            sys.stderr.write('rank %d: in-graph all-reduce %s, %d capture(s)\n' % (rank, 'fell back to two graphs' if gs.in_graph_allreduce is False else 'captured', gs.captures))
        runs.append((losses, torch.cat([p.detach().reshape(-1) for p in m.parameters()])))
        if graph:
            gs.entries.clear()      # graphs that captured NCCL work must be gone before their communicator is destroyed
            del gs
    (l0, p0), (l1, p1) = runs
    for a, b in zip(l0, l1):
        assert abs(a - b) <= 1e-5 * abs(a), (l0, l1)
    assert relt(p1, p0) < 1e-5
    ref = p1.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(ref, p1), 'replicas diverged'
    dist.barrier()
    print('MULTI_OK rank %d all-reduce err %.1e' % (rank, e), flush=True)
    import gc
    import threading
    gc.collect()
    torch.cuda.synchronize()
    wd = threading.Timer(30.0, lambda: os._exit(0))      # never let a stalled teardown hang the test
    wd.daemon = True
    wd.start()
    dist.destroy_process_group()
    wd.cancel()


if __name__ == '__main__':
    main()
