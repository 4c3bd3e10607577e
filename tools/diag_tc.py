"""Layout diagnostics for the tcgen05 self-test GEMM (prints which rows / columns / k land where)."""
import ctypes
import sys
import os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeptables_b200 import _native as nat

P = lambda t: ctypes.c_void_p(t.data_ptr())


def run(a, bm, variant):
    n, k = bm.shape[1], bm.shape[0]
    c = torch.full((128, n), -777.0, device='cuda')
    ws = torch.zeros(4 * n * k, dtype=torch.uint8, device='cuda')
    A = torch.tensor(a, dtype=torch.float32).cuda()
    B = torch.tensor(bm, dtype=torch.float32).cuda()
    nat.check(nat.lib.dtb_tc_selftest(P(A), P(B), P(c), P(ws), n, k, variant, None))
    torch.cuda.synchronize()
    return c.cpu().numpy()


for variant in (1, 0):
    for n, k in ((32, 16), (64, 32), (128, 64)):
        print(f'=== variant {"TMEM" if variant else "SMEM"} N={n} K={k}')
        ones_a = np.ones((128, k), dtype=np.float32)
        # which n lands in which column: A=delta(k=0), B[k][n] = n
        a = np.zeros((128, k), dtype=np.float32); a[:, 0] = 1
        bm = np.tile(np.arange(n, dtype=np.float32)[None, :], (k, 1))
        c = run(a, bm, variant)
        print(' col map (row 0):', c[0].astype(int).tolist())
        print(' col map (row 77):', c[77].astype(int).tolist())
        # which k is used: A = delta(k=k0), B[k][n] = k+1
        for k0 in (0, 1, 7, 8, 15, k - 1):
            a = np.zeros((128, k), dtype=np.float32); a[:, k0] = 1
            bm = np.tile((np.arange(k, dtype=np.float32) + 1)[:, None], (1, n))
            c = run(a, bm, variant)
            print(f' k0={k0}: C[0,:8]={c[0, :8].astype(int).tolist()} C[100,-4:]={c[100, -4:].astype(int).tolist()} uniq={np.unique(c).astype(int).tolist()[:8]}')
        # row map: A[m][0] = m+1, B = 1
        a = np.zeros((128, k), dtype=np.float32); a[:, 0] = np.arange(128) + 1
        bm = np.ones((k, n), dtype=np.float32)
        c = run(a, bm, variant)
        print(' row map (col 0):', c[:, 0].astype(int).tolist())
        # full random check by blocks
        g = np.random.default_rng(0)
        a = g.normal(size=(128, k)).astype(np.float32); bm = g.normal(size=(k, n)).astype(np.float32)
        c = run(a, bm, variant)
        bf = lambda x: torch.tensor(x).to(torch.bfloat16).float().numpy().astype(np.float64)
        want = bf(a) @ bf(bm)
        bad = np.abs(c - want) > 1e-3 * (1 + np.abs(want))
        print(' bad rows by 8-block:', bad.reshape(16, 8, n).any(axis=(1, 2)).astype(int).tolist())
        print(' bad cols by 8-block:', bad.reshape(128, n // 8, 8).any(axis=(0, 2)).astype(int).tolist(), 'frac', bad.mean())
