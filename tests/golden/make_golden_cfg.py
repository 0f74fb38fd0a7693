#!/usr/bin/env python
"""Mint tests/golden/cfg_parse.json: the REFERENCE's cfg.parse_cfg applied to the three shipped architectures
(serialised by netcfg.write_cfg; make_golden.py asserts that these parse to the same blocks as the reference's own
cfg/*.cfg files) plus a hand-written file with comments, blank lines, spaces around '=' and a `type=` key.
Build container only:  python tests/golden/make_golden_cfg.py"""
import io
import json
import os
import sys
import tempfile
from contextlib import redirect_stdout

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, '_shims'))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
with redirect_stdout(io.StringIO()):
    import cfg as RC
from fewshot_detection_b200 import netcfg

ODD = """# a comment
[net]
batch=64
width = 416

height=416
channels=3
[convolutional]
filters=8
size=3
stride=1
pad=1
activation=leaky
# batch_normalize left at its default

[cost]
type=sse
"""


def main():
    out = {'odd_text': ODD}
    with tempfile.TemporaryDirectory() as td:
        for name, blocks in (('tiny_yolo_voc', netcfg.tiny_yolo_voc_blocks()), ('darknet_dynamic', netcfg.darknet_dynamic_blocks()),
                             ('reweighting_net', netcfg.reweighting_net_blocks())):
            p = os.path.join(td, name + '.cfg')
            netcfg.write_cfg(blocks, p)
            out[name] = RC.parse_cfg(p)
            ref_file = os.path.join('/root/reference/cfg', name.replace('_', '-') + '.cfg' if name == 'tiny_yolo_voc' else name + '.cfg')
            assert RC.parse_cfg(ref_file)[1:] == out[name][1:], name
        p = os.path.join(td, 'odd.cfg')
        open(p, 'w').write(ODD)
        out['odd'] = RC.parse_cfg(p)
    json.dump(out, open(os.path.join(HERE, 'cfg_parse.json'), 'w'))
    print('wrote cfg_parse.json', {k: len(v) for k, v in out.items()})


if __name__ == '__main__':
    main()
