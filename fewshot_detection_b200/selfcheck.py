"""smoke(): one tiny meta-training step on cuda:0 through the public API,
checked against the CPU oracle (the only place the product tree touches
`oracle/`, and only as the checker; see oracle/__init__.py)."""
import os
import sys

import torch


def smoke_check():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, 'tests', 'golden')):
        if p not in sys.path:
            sys.path.insert(0, p)
    from fewshot_detection_b200 import netcfg
    from fewshot_detection_b200.darknet_meta import Darknet
    from fewshot_detection_b200.optim import FusedSGD
    from oracle import darknet as ODK, region_loss as ORL
    from seeding import seeded_init, synth_targets, synth_masks
    assert torch.cuda.is_available(), 'smoke() needs cuda:0'
    torch.cuda.set_device(0)
    det, ler = netcfg.mini_dynamic_blocks(128, 4), netcfg.mini_reweighting_blocks(64, 4, 128)
    m = Darknet([dict(b) for b in det], [dict(b) for b in ler])
    seeded_init(m, 1)
    m = m.cuda().train()
    om = ODK.MetaDarknet([dict(b) for b in det], [dict(b) for b in ler])
    seeded_init(om, 1)
    om.train()
    g = torch.Generator().manual_seed(2)
    bs, cs = 2, 3
    x = torch.rand(bs, 3, 128, 128, generator=g)
    metax = torch.rand(cs, 3, 64, 64, generator=g)
    mask = torch.from_numpy(synth_masks(cs, 64, 3))
    tgt = torch.from_numpy(synth_targets(bs, cs, 4, max_gt=4))
    opt = FusedSGD(m.parameters(), lr=1e-3, momentum=0.9, weight_decay=5e-4)
    opt.zero_grad()
    L = m.models[len(m.models) - 1]
    L.seen = 20000
    out = m(x.cuda(), metax.cuda(), mask.cuda())
    loss = L(out, tgt)
    loss.backward()
    opt.step()
    torch.cuda.synchronize()
    lo = ORL.region_loss_v2(om(x, metax, mask), tgt, om.anchors, 5, 1, seen=20000)
    lo.backward()
    err = abs(loss.item() - lo.item()) / abs(lo.item())
    assert err < 1e-3, ('loss mismatch vs oracle', loss.item(), lo.item())
    worst = 0.0
    for (n, p), (_, q) in zip(m.named_parameters(), om.named_parameters()):
        d = (p.grad.detach().cpu().contiguous().double() - q.grad.double()).norm() / max(q.grad.double().norm().item(), 1e-30)
        worst = max(worst, d.item())
    assert worst < 1e-3, ('gradient mismatch vs oracle', worst)
    # evaluation decode + NMS of the same head output (csrc/detect.cu) vs the oracle's Python loops
    from fewshot_detection_b200 import utils as U
    from oracle import utils as OU
    kept = U.region_detections(out.detach(), 0.005, 1, m.anchors, 5, 0, 1, n_models=cs).kept_boxes(0.45)
    want = [OU.nms(r, 0.45) for r in OU.get_region_boxes_v2(out.detach().cpu(), cs, 0.005, 1, om.anchors, 5, 0, 1)]
    assert len(kept) == len(want) == bs * cs
    assert sum(abs(len(a) - len(b)) for a, b in zip(kept, want)) <= 1, ('NMS survivors differ from the oracle',
                                                                        [len(a) for a in kept], [len(b) for b in want])
    print('smoke ok: loss %.6f (oracle %.6f), worst grad rel err %.2e, %d NMS survivors'
          % (loss.item(), lo.item(), worst, sum(len(a) for a in kept)))
