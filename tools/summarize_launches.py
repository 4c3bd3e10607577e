"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel totals and shares.
usage: python tools/summarize_launches.py launches.csv [first_kernel_substring]"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get('Metric Name') == 'gpu__time_duration.sum':
            rows.append((r['Kernel Name'], float(r['Metric Value']) / 1e3))
    agg = OrderedDict()
    for name, us in rows:
        short = re.sub(r'\(.*', '', name)[:100]
        t, n = agg.get(short, (0.0, 0))
        agg[short] = (t + us, n + 1)
    total = sum(t for t, _ in agg.values())
    print(f'# {len(rows)} launches, total {total:.1f} us')
    for name, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f'{t:10.1f} us {100 * t / total:5.1f}%  n={n:4d}  {name}')


if __name__ == '__main__':
    main()
