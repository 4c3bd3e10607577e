"""Top warp-stall locations of the kernels in an `ncu --page source --csv` dump (SASS view):
    ncu -i report.ncu-rep --page source --csv > src.csv ; python tools/ncu_hot.py src.csv [n_top] [kernel substring]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
only = sys.argv[3] if len(sys.argv) > 3 else ''
kernels, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'data': []}
        kernels.append(cur)
    elif cur is not None and cur['hdr'] is None:
        cur['hdr'] = r
    elif cur is not None and len(r) > 10:
        cur['data'].append(r)
for k in kernels:
    if only not in k['name']:
        continue
    hdr, data = k['hdr'], k['data']
    isrc, iss, iex = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
    tot = sum(int(r[iss]) for r in data)
    totx = sum(int(r[iex]) for r in data)
    print(f"== {k['name'][:90]}: {len(data)} SASS instr, {totx} warp-instr executed, {tot} samples")
    top = sorted(range(len(data)), key=lambda i: -int(data[i][iss]))[:n_top]
    for i in sorted(top):
        print(f'{i:6d} {100 * int(data[i][iss]) / tot:5.1f}% {int(data[i][iex]):>10d}  {data[i][isrc][:120]}')
