"""tcgen05 / TMEM / TMA mnemonics per kernel of libfsdet.so (cuobjdump -sass), plus the count of serialising issue loops
(BRA.U.ANY: the ELECT / R2UR / BRA.U.ANY loop ptxas wraps around a uniform-operand instruction when it cannot prove that
one thread executes it).  Usage: python tools/sass_table.py > profiles/sass_r02b.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'fewshot_detection_b200', 'libfsdet.so')
COLS = ['UTCHMMA', 'LDTM', 'UTMALDG', 'UTMASTG', 'UTMAREDG', 'UTCBAR', 'ELECT', 'BRA.U.ANY', 'R2UR', 'LDS', 'STS']


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], stdout=subprocess.PIPE, text=True).stdout
    fn = None
    counts = collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            fn = subprocess.run(['c++filt', m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            fn = re.sub(r'\(.*', '', fn).replace('fsdet::', '').replace('(int)', '').replace('(bool)', '')
            counts[fn] = collections.Counter()
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m and fn:
            op = m.group(1)
            for c in COLS:
                if op == c or op.startswith(c + '.') or (c == 'BRA.U.ANY' and op.startswith('BRA.U.ANY')):
                    counts[fn][c] += 1
            if op.startswith('UTMALDG') and 'IM2COL' in op:
                counts[fn]['IM2COL'] += 1
    print('# Round 2 (second session) - SASS evidence per kernel of fewshot_detection_b200/libfsdet.so (sm_100a)\n')
    print('`python tools/sass_table.py` (cuobjdump -sass).  UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor load '
          '(IM2COL = im2col mode), UTMASTG / UTMAREDG = TMA tensor store / reduce-add, UTCBAR = tcgen05.commit.  '
          '**BRA.U.ANY** counts the serialising loops (ELECT / R2UR / UTCHMMA / BRA.U.ANY) that ptxas wraps around a '
          'uniform-operand instruction issued under `if (lane == 0)`: the tensor-core kernels issue under `elect.sync` '
          'with the whole warp converged and have none; LDS / STS instead of generic LD / ST show that the epilogue '
          'staging stays in the shared address space.\n')
    cols = COLS[:6] + ['IM2COL'] + COLS[6:]
    print('| kernel | ' + ' | '.join(cols) + ' |')
    print('|---|' + '---:|' * len(cols))
    for fn, c in counts.items():
        if c['UTCHMMA'] or c['UTMALDG']:
            print('| `%s` | ' % fn + ' | '.join(str(c[k]) for k in cols) + ' |')


if __name__ == '__main__':
    main()
