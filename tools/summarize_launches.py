"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel totals and shares.
usage: python tools/summarize_launches.py launches.csv [step_marker]
With a marker (e.g. 'adam_rows_kernel<0>', the first kernel of a train step) only the launches from its first
occurrence up to (not including) its second one are summarised: exactly one step."""
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
    if len(sys.argv) > 2:
        hits = [i for i, (n, _) in enumerate(rows) if sys.argv[2] in n]
        if len(hits) >= 2:
            rows = rows[hits[0]:hits[1]]
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
