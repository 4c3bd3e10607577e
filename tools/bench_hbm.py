"""Achieved HBM bandwidth of the bandwidth-bound kernels at the BASELINE shape (65 536 rows, 26 x 16 + 13),
timed alone with CUDA events, L2 flushed between launches.  Algorithmic bytes per row: SURVEY.md 8(d)."""
import ctypes
import json
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeptables_b200 import _native as nat

P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
B, F, D, C, V = int(os.environ.get('B', 65536)), 26, 16, 13, int(os.environ.get('V', 1_000_000))
W = F * D + C
g = torch.Generator(device='cuda').manual_seed(0)
table = (torch.rand(F * V, D, device='cuda', generator=g) - 0.5) * 0.1
gtab = torch.zeros_like(table)
offs = torch.arange(F + 1, dtype=torch.int64, device='cuda') * V
idx = torch.randint(0, V, (B, F), device='cuda', dtype=torch.int32, generator=g)
dense = torch.randn(B, C, device='cuda', generator=g)
wl = torch.randn(F + C, device='cuda', generator=g)
lin, fm = torch.empty(B, device='cuda'), torch.empty(B, device='cuda')
gl, gf = torch.randn(B, device='cuda', generator=g), torch.randn(B, device='cuda', generator=g)
gw = torch.zeros(F + C, device='cuda')
X, Y, dY = torch.empty(B, W, device='cuda'), torch.empty(B, W, device='cuda'), torch.randn(B, W, device='cuda', generator=g)
gamma, beta = torch.ones(W, device='cuda'), torch.zeros(W, device='cuda')
mm, mv, sm, sv = torch.zeros(W, device='cuda'), torch.ones(W, device='cuda'), torch.empty(W, device='cuda'), torch.empty(W, device='cuda')
ws = torch.empty(2 * W, dtype=torch.float64, device='cuda')
dg, db = torch.zeros(W, device='cuda'), torch.zeros(W, device='cuda')
ck, cb = torch.randn(6, W, device='cuda', generator=g) * 0.05, torch.zeros(6, W, device='cuda')
xw = torch.empty(B, 6, device='cuda')
dk, dbb = torch.zeros(6, W, device='cuda'), torch.zeros(6, W, device='cuda')
cws = torch.empty(nat.lib.dtb_cross_bwd_workspace_bytes(B, W, 6), dtype=torch.uint8, device='cuda')
flush = torch.zeros(512 << 20, dtype=torch.uint8, device='cuda')
peak = 6485.5
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
if os.path.exists(pk):
    peak = json.load(open(pk))['hbm_gbs']

cases = {
    'fm_linear_fwd (ids+rows+dense+2 out)': (lambda: nat.lib.dtb_fm_linear_fwd(P(idx), P(table), P(offs), P(dense), P(wl), P(lin), P(fm), B, F, D, C, None, None), 4 * F + 64 * F + 4 * C + 8),
    'fm_linear_bwd (re-gather + RED rows)': (lambda: nat.lib.dtb_fm_linear_bwd(P(idx), P(table), P(offs), P(dense), P(wl), P(gl), P(gf), P(gtab), P(gw), B, F, D, C, None), 4 * F + 3 * 64 * F + 4 * C + 8),
    'concat_emb_dense_fwd': (lambda: nat.lib.dtb_concat_emb_dense_fwd(P(idx), P(table), P(offs), P(dense), P(X), B, F, D, C, None, None), 4 * F + 64 * F + 4 * C + 4 * W),
    'concat_emb_dense_bwd (RED rows)': (lambda: nat.lib.dtb_concat_emb_dense_bwd(P(idx), P(offs), P(dY), P(gtab), B, F, D, C, None), 4 * F + 4 * F * D + 2 * 64 * F),
    'batchnorm_train_fwd (2 reads + 1 write)': (lambda: nat.lib.dtb_batchnorm_train_fwd(P(X), P(Y), P(gamma), P(beta), P(mm), P(mv), P(sm), P(sv), P(ws), B, W, 1e-3, 0.99, None), 3 * 4 * W),
    'batchnorm_bwd (4 reads + 1 write)': (lambda: nat.lib.dtb_batchnorm_bwd(P(X), P(dY), P(Y), P(gamma), P(sm), P(sv), P(dg), P(db), P(ws), B, W, 1e-3, None), 5 * 4 * W),
    'cross_fwd 6 layers': (lambda: nat.lib.dtb_cross_fwd(P(X), P(ck), P(cb), P(Y), P(xw), B, W, 6, None), 2 * 4 * W),
    'cross_bwd 6 layers': (lambda: nat.lib.dtb_cross_bwd(P(X), P(ck), P(cb), P(xw), P(dY), P(Y), P(dk), P(dbb), P(cws), cws.numel(), B, W, 6, None), 4 * 4 * W),
}
nat.lib.dtb_concat_emb_dense_fwd(P(idx), P(table), P(offs), P(dense), P(X), B, F, D, C, None, None)
only = os.environ.get('ONLY')
for name, (fn, bpr) in cases.items():
    if only and only not in name:
        continue
    for _ in range(2):
        nat.check(fn())
    ts = []
    for _ in range(7):
        flush.sum()      # evict with CLEAN lines (a written flush buffer would be written back during the timed kernel)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); nat.check(fn()); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[3]
    gbs = B * bpr / t / 1e6
    print(f'{name:42s} {t * 1e3:8.1f} us  {bpr:6d} B/row  {gbs:7.0f} GB/s  {100 * gbs / peak:5.1f}% of measured HBM peak ({peak:.0f})', flush=True)
