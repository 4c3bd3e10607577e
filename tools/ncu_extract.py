"""Pull the handful of numbers the roofline discussion needs out of an `ncu --set full` report.

usage (where ncu is installed, no GPU needed):
    python tools/ncu_extract.py gpurun_out/cin_all.ncu-rep [--json profiles/rN_cin_tc_traffic.json]

Per kernel launch: duration, DRAM bytes read/written (-> `roofline.traffic`), DRAM/L2 throughput %, tensor-pipe
and issue utilisation, registers, achieved occupancy.  With --json the per-kernel DRAM traffic of the first launch
of each CIN tensor-core kernel is written in the format bench.py reads (key `cin_tc_fwd_kernel_compact`)."""
import csv
import io
import json
import re
import subprocess
import sys

WANT = {
    'gpu__time_duration.sum': 'duration',
    'dram__bytes_read.sum': 'dram_read',
    'dram__bytes_write.sum': 'dram_write',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed': 'l2_pct',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm_pct',
    'sm__inst_executed_pipe_tensor.sum': 'tensor_inst',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_pct',
    'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_hmma_pct',
    'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_pct',
    'launch__registers_per_thread': 'regs',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'occupancy_pct',
    'launch__grid_size': 'grid',
    'launch__block_size': 'block',
}
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12,
         'nsecond': 1e-9, 'usecond': 1e-6, 'msecond': 1e-3, 'second': 1.0, 'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0}


def read_raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    lines = [l for l in out.splitlines() if l.startswith('"')]
    rows = list(csv.reader(io.StringIO('\n'.join(lines))))
    return rows[0], rows[1], rows[2:]


def extract(path):
    header, units, rows = read_raw(path)
    cols = {}
    for i, h in enumerate(header):
        for metric, short in WANT.items():
            if h == metric or h.endswith('.' + metric):
                cols.setdefault(short, []).append(i)       # the same metric can appear in several (partly empty) sections
    k_name = header.index('Kernel Name')
    launches = []
    for r in rows:
        rec = {'kernel': re.sub(r'\(.*', '', r[k_name]).replace('void ', '')}
        for short, idxs in cols.items():
            for i in idxs:
                try:
                    v = float(r[i].replace(',', ''))
                except ValueError:
                    continue
                rec[short] = v * SCALE.get(units[i], 1.0)
                break
        launches.append(rec)
    return launches


def main():
    path = sys.argv[1]
    launches = extract(path)
    for r in launches:
        dur = r.get('duration', 0.0)
        traffic = r.get('dram_read', 0.0) + r.get('dram_write', 0.0)
        print(f"{r['kernel'][:44]:44s} {dur * 1e6:9.1f} us  dram R {r.get('dram_read', 0) / 1e6:9.1f} MB  W {r.get('dram_write', 0) / 1e6:9.1f} MB"
              f"  ({traffic / dur / 1e9 if dur else 0:7.1f} GB/s, {r.get('dram_pct', 0):4.1f}% dram, {r.get('l2_pct', 0):4.1f}% L2)"
              f"  tensor {r.get('tensor_pct', r.get('tensor_hmma_pct', 0)):5.1f}%  issue {r.get('issue_pct', 0):5.1f}%"
              f"  regs {int(r.get('regs', 0))}  occ {r.get('occupancy_pct', 0):4.1f}%")
    if '--json' in sys.argv:
        dst = sys.argv[sys.argv.index('--json') + 1]
        out = {'source': f'ncu --set full --clock-control none ({path}); tools/ncu_extract.py'}
        # key = kernel name as bench.py looks it up; the first launch of each kernel in the report
        names = {'cin_tc2_fwd_kernel': 'cin_tc2_fwd_kernel', 'cin_tc2_dgrad_kernel': 'cin_tc2_dgrad_kernel',
                 'cin_tc2_wgrad_kernel': 'cin_tc2_wgrad_kernel_first_launch', 'cin_tc_fwd_kernel': 'cin_tc_fwd_kernel',
                 'cin_tc_dgrad_kernel': 'cin_tc_dgrad_kernel', 'cin_tc_wgrad_kernel': 'cin_tc_wgrad_kernel_first_launch'}
        for r in launches:
            for needle, key in names.items():
                if needle in r['kernel'] and key not in out and 'dram_read' in r:
                    out[key] = {'dram_bytes_read': r['dram_read'], 'dram_bytes_write': r['dram_write'],
                                'duration_s': r.get('duration')}
        try:
            with open(dst) as f:
                old = json.load(f)
        except (OSError, ValueError):
            old = {}
        old.update(out)
        with open(dst, 'w') as f:
            json.dump(old, f, indent=1)
        print(f'wrote {dst}')


if __name__ == '__main__':
    main()
