"""fewshot_detection_b200.lists (few-shot list construction) and cfg.config_data (base / novel split) against what the
REFERENCE's own dataset.py / cfg.py returned on the same files (tests/golden/lists.json, minted by
tests/golden/make_golden_lists.py which imports /root/reference/dataset.py).  CPU only."""
import json
import os
import random

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.fixture()
def world(tmp_path):
    d = json.load(open(os.path.join(G, 'lists.json')))
    root = str(tmp_path)
    for rel, text in d['files'].items():
        p = os.path.join(root, rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, 'w') as f:
            f.write(text.replace('<ROOT>', root))
    os.makedirs(os.path.join(root, 'JPEGImages'), exist_ok=True)
    from fewshot_detection_b200.cfg import cfg
    saved = dict(cfg)
    yield d, root, cfg
    cfg.clear()
    cfg.update(saved)


def configure(cfg, d, tuning, repeat=1, shot=2):
    classes, novel = d['classes'], d['novel']
    cfg.data, cfg.classes, cfg.tuning, cfg.repeat, cfg.shot = 'voc', classes, tuning, repeat, shot
    cfg.novel_classes = novel
    cfg.base_classes = list(classes) if tuning else [c for c in classes if c not in novel]
    cfg.base_ids = [classes.index(c) for c in cfg.base_classes]
    cfg.novel_ids = [classes.index(c) for c in novel]
    cfg.num_gpus, cfg.batch_size, cfg.randmeta = 1, 64, False


def test_list_builders_match_the_reference(world):
    from fewshot_detection_b200 import lists as LS
    d, root, cfg = world
    P = lambda rel: os.path.join(root, rel)
    want = lambda k: [l.replace('<ROOT>', root) for l in d['cases'][k]]
    configure(cfg, d, False)
    assert LS.load_lines(P('lists/train.txt')) == want('base_plain')
    assert LS.load_lines(P('lists/dict_full.txt')) == want('base_dict')
    assert LS.load_lines(P('lists/dict_full.txt'), checkvalid=False) == want('base_dict_nocheck')
    assert LS.build_dataset({'train': P('lists/train.txt'), 'meta': P('lists/dict_full.txt')}) == want('build_base')
    # base training never sees an image without a base-class object
    novel_ids = set(cfg.novel_ids)
    for l in want('build_base'):
        assert not set(LS.label_classes(l)) <= novel_ids
    configure(cfg, d, True, repeat=1)
    assert LS.build_dataset({'train': P('lists/train.txt'), 'meta': P('lists/dict_2shot.txt')}) == want('tune_repeat1')
    configure(cfg, d, True, repeat=3)
    assert LS.build_dataset({'train': P('lists/train.txt'), 'meta': P('lists/dict_2shot.txt'), 'dynamic': '0'}) == want('tune_repeat3')
    configure(cfg, d, True, repeat=2)
    ml, mc = LS.load_metadict(P('lists/dict_2shot.txt'), 2)
    assert sorted(ml) == sorted(want('metadict_list_sorted')) and mc == d['cases']['metadict_counts']


def test_dynamic_fewset_reaches_shot_boxes_per_class(world):
    """build_fewset draws with `random`: the reference's set iteration order of the novel images is hash-dependent, so
    the list is compared as a multiset with the reference's seeded run, plus the stopping rule: every class ends
    with at least shot * repeat boxes and no added base image shows a novel object or more than 3 boxes."""
    from fewshot_detection_b200 import lists as LS
    d, root, cfg = world
    P = lambda rel: os.path.join(root, rel)
    configure(cfg, d, True, repeat=2)
    random.seed(11)
    got = LS.build_dataset({'train': P('lists/train.txt'), 'meta': P('lists/dict_2shot.txt'), 'dynamic': '1'})
    want = [l.replace('<ROOT>', root) for l in d['cases']['tune_dynamic_seed11']]
    assert sorted(got) == sorted(want)          # same draws -> same images (the shuffle order depends on the set order)
    counts = {c: 0 for c in d['classes']}
    for l in got:
        for ci in LS.label_classes(l):
            counts[d['classes'][ci]] += 1
    assert min(counts.values()) >= cfg.shot * cfg.repeat, counts
    novel_imgs = set(l for l in [x.replace('<ROOT>', root) for x in d['cases']['metadict_list_sorted']])
    for l in got:
        if l not in novel_imgs:
            ids = LS.label_classes(l)
            assert len(ids) <= 3 and set(ids).isdisjoint(cfg.novel_ids)


def test_support_index_matches_the_reference(world):
    from fewshot_detection_b200 import lists as LS
    d, root, cfg = world
    configure(cfg, d, False)
    w = d['cases']['support_train_seed3']
    np.random.seed(3)
    nbatch = LS.support_batches_per_epoch(train=True)
    metalines, inds = LS.support_index(os.path.join(root, 'lists/dict_full.txt'), cfg.base_classes, nbatch)
    assert len(inds) == w['n'] and [len(m) for m in metalines] == w['meta_cnts']
    assert [list(map(int, t)) for t in inds[:600]] == w['inds']
    assert len(cfg.base_classes) * cfg.num_gpus == w['batch_size']
    # one support image per class per step: consecutive entries walk the classes in order
    assert [c for c, _ in inds[:len(cfg.base_classes)]] == list(range(len(cfg.base_classes)))
    ml2, inds2 = LS.support_index(os.path.join(root, 'lists/dict_full.txt'), cfg.base_classes, nbatch, ensemble=True)
    assert len(inds2) == sum(len(m) for m in ml2) and inds2[0] == (0, 0)


def test_config_data_base_novel_split(world, tmp_path):
    """cfg/metayolo.data and cfg/metatune.data of the reference (key lines restated here): base training must exclude
    the novel classes of split `novelid`; fine-tuning sees all 20 and sets the schedule keys."""
    d, root, cfg = world
    novels = tmp_path / 'voc_novels.txt'
    novels.write_text('bird,bus,cow,motorbike,sofa\naeroplane,bottle,cow,horse,sofa\nboat,cat,motorbike,sheep,sofa\n')
    base = {'metayolo': '1', 'metain_type': '2', 'data': 'voc', 'neg': '1', 'rand': '0', 'novel': str(novels), 'novelid': '0',
            'meta': 'data/voc_traindict_full.txt', 'backup': 'backup/metayolo', 'gpus': '1,2,3,4'}
    cfg.tuning = False
    cfg.config_data(base)
    assert cfg.novel_classes == ['bird', 'bus', 'cow', 'motorbike', 'sofa']
    assert cfg.base_ids == [0, 1, 3, 4, 6, 7, 8, 10, 11, 12, 14, 15, 16, 18, 19] and cfg.novel_ids == [2, 5, 9, 13, 17]
    assert cfg._real_base_ids == cfg.base_ids and len(cfg.base_classes) == 15
    assert cfg.neg_ratio == 1 and isinstance(cfg.neg_ratio, int) and cfg.num_gpus == 4 and cfg.metayolo is True
    assert cfg.backup == 'backup/metayolo_novel0_neg1'
    tune = dict(base, tuning='1', neg='0', max_epoch='2000', repeat='200', dynamic='0', scale='1',
                meta='data/voc_traindict_bbox_5shot.txt', backup='backup/metatunetest1', novelid='1')
    cfg.config_data(tune)
    assert cfg.tuning is True and cfg.base_classes == cfg.voc_classes and cfg.base_ids == list(range(20))
    assert cfg.novel_classes == ['aeroplane', 'bottle', 'cow', 'horse', 'sofa'] and cfg.novel_ids == [0, 4, 9, 12, 17]
    assert (cfg.max_epoch, cfg.repeat, cfg.shot, cfg.save_interval, cfg.neg_ratio) == (2000, 200, 5, 1, 0)
    assert cfg.backup == 'backup/metatunetest1_novel1_neg0'
    from fewshot_detection_b200 import trainer as T
    assert T.epoch_plan(0, 1000, 64, 80200, cfg.tuning, cfg.max_epoch, cfg.repeat) == (0, 0, 10)
    with pytest.raises(ValueError):
        cfg.config_data(dict(base, novel='bird,unicorn'))


def test_parse_cfg_matches_the_reference_parser(tmp_path):
    """cfg.parse_cfg on files: the three shipped architectures (serialised by netcfg.write_cfg, which make_golden.py
    checks against the reference's own cfg/*.cfg) and a file with comments / blank lines / spaces / a `type=` key,
    against what the reference's parser returned for the same text (tests/golden/cfg_parse.json)."""
    from fewshot_detection_b200 import netcfg
    from fewshot_detection_b200.cfg import parse_cfg
    gold = json.load(open(os.path.join(G, 'cfg_parse.json')))
    for name, blocks in (('tiny_yolo_voc', netcfg.tiny_yolo_voc_blocks()), ('darknet_dynamic', netcfg.darknet_dynamic_blocks()),
                         ('reweighting_net', netcfg.reweighting_net_blocks())):
        p = str(tmp_path / (name + '.cfg'))
        netcfg.write_cfg(blocks, p)
        assert parse_cfg(p) == gold[name], name
    p = str(tmp_path / 'odd.cfg')
    open(p, 'w').write(gold['odd_text'])
    got = parse_cfg(p)
    assert got == gold['odd']
    assert got[1]['batch_normalize'] == 0 and got[2]['_type'] == 'sse' and got[0]['width'] == '416'
