"""Pin the augmentation oracle (oracle/image.py: a numpy restatement of Pillow's uint8 pipeline behind the
reference's image.data_augmentation) against tests/golden/augment.npz, minted from the reference itself.  CPU only,
bit-exact."""
import os
import random

import numpy as np
import pytest

from oracle import image as OI

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(G, 'augment.npz'), allow_pickle=False)


@pytest.mark.parametrize('tag', ['a', 'b', 'c', 'd', 'e', 'f', 'g'])
def test_data_augmentation_bit_exact(gold, tag):
    W, H, seed, flag = [int(v) for v in gold[tag + '/args']]
    random.seed(seed)
    img, flip, dx, dy, sx, sy = OI.data_augmentation(gold[tag + '/src'], (W, H), 0.2, 0.1, 1.5, 1.5, flag=bool(flag))
    assert np.array_equal(img, gold[tag + '/img'])
    assert [flip, dx, dy, sx, sy] == list(gold[tag + '/params'])
    random.seed(seed)
    img0 = OI.data_augmentation(gold[tag + '/src'], (W, H), 0.2, 0.1, 1.5, 1.5, flag=bool(flag), bicubic=False)[0]
    assert np.array_equal(img0, gold['nearest/' + tag])
    assert OI.to_tensor(img).shape == (3, H, W)


def test_crop_zero_fill_outside():
    a = np.arange(5 * 4 * 3, dtype=np.uint8).reshape(5, 4, 3)
    c = OI.crop(a, -2, -1, 3, 7)
    assert c.shape == (8, 5, 3)
    assert not c[0].any() and not c[:, :2].any() and not c[6:].any()
    assert np.array_equal(c[1:6, 2:5], a[:, :3])
