"""Darknet weight-stream interoperability with the REFERENCE: tests/golden/weights.npz holds byte
streams written by the reference's own Darknet.save_weights and the state its own load_weights leaves behind
(tests/golden/make_golden_weights.py).  CPU only - weight IO does not touch the GPU."""
import os

import numpy as np
import pytest
import torch

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(G, 'weights.npz'))


def model(seed):
    from fewshot_detection_b200 import netcfg
    from fewshot_detection_b200.darknet_meta import Darknet
    from seeding import seeded_init
    m = Darknet(netcfg.mini_dynamic_blocks(128, 4), netcfg.mini_reweighting_blocks(64, 4, 128))
    return seeded_init(m, seed)


def check_digest(m, gold, prefix):
    for name, t in list(m.named_parameters()) + [(n, b) for n, b in m.named_buffers() if 'running' in n]:
        v = t.detach().double().contiguous().reshape(-1)
        assert np.array_equal(v[:8].numpy(), gold['%s/head/%s' % (prefix, name)]), name
        assert abs(v.sum().item() - float(gold['%s/sum/%s' % (prefix, name)])) <= 1e-9 * max(1.0, abs(float(gold['%s/sum/%s' % (prefix, name)]))), name


def test_save_weights_writes_the_reference_byte_stream(gold, tmp_path):
    m = model(int(gold['seed_written']))
    m.seen = int(gold['seen'])
    f = str(tmp_path / 'all.weights')
    m.save_weights(f)
    assert np.array_equal(np.fromfile(f, dtype=np.uint8), gold['stream_all'])
    f2 = str(tmp_path / 'cut.weights')
    m.save_weights(f2, cutoff=12)
    assert np.array_equal(np.fromfile(f2, dtype=np.uint8), gold['stream_cutoff12'])
    # the same with channels_last (OHWI) parameter storage, which is what the engine converts weights to
    for p in m.parameters():
        if p.dim() == 4:
            p.data = p.data.contiguous(memory_format=torch.channels_last)
    m.save_weights(f)
    assert np.array_equal(np.fromfile(f, dtype=np.uint8), gold['stream_all'])


def test_load_weights_reads_a_reference_written_stream(gold, tmp_path):
    f = str(tmp_path / 'ref.weights')
    gold['stream_all'].tofile(f)
    m = model(int(gold['seed_loaded_into']))
    for p in list(m.parameters())[::3]:           # mixed storage formats on the receiving side
        if p.dim() == 4:
            p.data = p.data.contiguous(memory_format=torch.channels_last)
    m.load_weights(f)
    assert m.seen == int(gold['loaded_seen']) == int(gold['seen'])
    check_digest(m, gold, 'all')
    src = model(int(gold['seed_written']))
    for (n, a), (_, b) in zip(src.named_parameters(), m.named_parameters()):
        assert torch.equal(a.detach(), b.detach().contiguous()), n


def test_truncated_stream_stops_silently_like_the_reference(gold, tmp_path):
    """darknet19_448.conv.23: the stream ends after the trunk; later tensors keep their initialisation."""
    n = int(gold['truncated_floats'])
    f = str(tmp_path / 'trunc.weights')
    gold['stream_all'][:16 + 4 * n].tofile(f)
    m = model(int(gold['seed_loaded_into']))
    m.load_weights(f)
    check_digest(m, gold, 'trunc')
    fresh = model(int(gold['seed_loaded_into']))
    src = model(int(gold['seed_written']))
    names = [n_ for n_, _ in m.named_parameters()]
    changed = [n_ for (n_, a), (_, b) in zip(m.named_parameters(), fresh.named_parameters()) if not torch.equal(a, b)]
    assert changed == names[:9]                   # 3 x (conv.weight, bn.weight, bn.bias) and nothing else
    for (n_, a), (_, b) in zip(m.named_parameters(), src.named_parameters()):
        if n_ in changed:
            assert torch.equal(a, b), n_


def test_cutoff_stream_loads_as_a_prefix(gold, tmp_path):
    f = str(tmp_path / 'cut.weights')
    gold['stream_cutoff12'].tofile(f)
    m = model(int(gold['seed_loaded_into']))
    m.load_weights(f)
    src = model(int(gold['seed_written']))
    got = [torch.equal(a, b) for (_, a), (_, b) in zip(m.named_parameters(), src.named_parameters())]
    k = got.index(False)
    assert k > 0 and not any(got[k:])             # a prefix of the parameters was replaced, the rest untouched
    positions = [pos for pos, _, _ in m._weight_stream()]
    assert positions == sorted(positions) and sum(1 for p in positions if p <= 12) * 3 == k
