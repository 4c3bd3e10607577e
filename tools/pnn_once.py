"""Time the PNN product and AutoInt attention-core C-ABI calls at the bench shapes (CUDA events, warm).
python tools/pnn_once.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeptables_b200 import _native as N  # noqa: E402

P = lambda t: None if t is None else t.data_ptr()


def timeit(fn, reps=int(os.environ.get('REPS', '10'))):
    for _ in range(3 if reps > 1 else 0):      # REPS=1: one launch each (the ncu capture)
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    f, d, vocab = 26, 16, 100000
    for b in (16384, 65536):
        pairs = f * (f - 1) // 2
        g = torch.Generator(device='cuda').manual_seed(1)
        table = torch.randn(f * vocab, d, device='cuda', generator=g) * 0.05
        offs = torch.arange(0, (f + 1) * vocab, vocab, dtype=torch.int64, device='cuda')
        idx = torch.randint(0, vocab, (b, f), dtype=torch.int32, device='cuda', generator=g)
        kern = torch.randn(d, pairs, d, device='cuda', generator=g) / d ** 0.5
        ip = torch.empty(b, pairs, device='cuda')
        op = torch.empty(b, pairs, device='cuda')
        gip, gop = torch.randn_like(ip), torch.randn_like(op)
        gt = torch.zeros_like(table)
        dk = torch.zeros_like(kern)
        fw = lambda: N.check(N.lib.dtb_pnn_fwd(P(idx), P(table), P(offs), P(kern), P(ip), P(op), b, f, d, 0, None, None), 'f')
        bw = lambda: N.check(N.lib.dtb_pnn_bwd(P(idx), P(table), P(offs), P(kern), P(gip), P(gop), P(gt), P(dk), b, f, d, 0,
                                               None), 'b')
        print(f'pnn (mat) rows {b} x {f} fields x D {d}: fwd {timeit(fw):.3f} ms   bwd (dE + dK) {timeit(bw):.3f} ms', flush=True)
    for b, dd, heads in ((65536, 32, 4), (16384, 16, 1)):
        qkvr = torch.relu(torch.randn(b, f, 4 * dd, device='cuda'))
        y = torch.empty(b, f, dd, device='cuda')
        dy = torch.randn_like(y)
        dq = torch.empty_like(qkvr)
        fw = lambda: N.check(N.lib.dtb_attention_core_fwd(P(qkvr), P(y), b, f, dd, heads, 1, None), 'f')
        bw = lambda: N.check(N.lib.dtb_attention_core_bwd(P(qkvr), P(y), P(dy), P(dq), b, f, dd, heads, 1, 0, None), 'b')
        tf, tb = timeit(fw), timeit(bw)
        byf, byb = b * f * dd * 5 * 4, b * f * dd * (4 + 2 + 4) * 4
        print(f'attention core rows {b} x {f} x D {dd}, {heads} heads: fwd {tf:.3f} ms ({byf / tf / 1e6:.0f} GB/s)   '
              f'bwd {tb:.3f} ms ({byb / tb / 1e6:.0f} GB/s)', flush=True)


if __name__ == '__main__':
    main()
