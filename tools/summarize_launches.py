"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel share table (markdown).
Usage: python tools/summarize_launches.py launches.csv [skip_fraction] > launches.md
`skip_fraction` (default 0.5) drops the first part of the list (the first of two profiled steps)."""
import csv
import re
import sys


def rows(path):
    with open(path, newline='') as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        ns = v * {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(unit, 1.0)
        yield r['Kernel Name'], ns


def short(name):
    name = re.sub(r'^void ', '', name)
    m = re.match(r'([A-Za-z0-9_:]+(<[0-9, ]+>)?)', name)
    return (m.group(1) if m else name)[:70]


def main():
    path = sys.argv[1]
    skip = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
    all_rows = list(rows(path))
    sel = all_rows[int(len(all_rows) * skip):]
    agg = {}
    for k, ns in sel:
        a = agg.setdefault(short(k), [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    print('Total %.2f ms over %d launches.\n' % (tot / 1e6, len(sel)))
    print('| kernel | launches | ms | share |\n|---|---:|---:|---:|')
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| `%s` | %d | %.3f | %.1f%% |' % (k, n, ns / 1e6, 100 * ns / tot))


if __name__ == '__main__':
    main()
