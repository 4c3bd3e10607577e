"""Time the AFM and FiBiNet (bilinear) C-ABI calls at the Criteo shape (CUDA events, warm): python tools/f3_once.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeptables_b200 import _native as N  # noqa: E402

P = lambda t: None if t is None else t.data_ptr()


def timeit(fn, reps=int(os.environ.get('REPS', '5'))):
    for _ in range(2 if reps > 1 else 0):
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
    f, d, h, vocab = 26, 16, 16, 100000
    pairs = f * (f - 1) // 2
    for b in [int(r) for r in os.environ.get('ROWS', '16384,65536').split(',')]:
        g = torch.Generator(device='cuda').manual_seed(1)
        table = torch.randn(f * vocab, d, device='cuda', generator=g) * 0.1
        offs = torch.arange(0, (f + 1) * vocab, vocab, dtype=torch.int64, device='cuda')
        idx = torch.randint(0, vocab, (b, f), dtype=torch.int32, device='cuda', generator=g)
        wa, ba, ph = torch.randn(d, h, device='cuda') * 0.3, torch.zeros(h, device='cuda'), torch.randn(h, device='cuda')
        pooled, gp = torch.empty(b, d, device='cuda'), torch.randn(b, d, device='cuda')
        gt, dwa, dba, dph = torch.zeros_like(table), torch.zeros_like(wa), torch.zeros_like(ba), torch.zeros_like(ph)
        nb = N.lib.dtb_afm_workspace_bytes(b, f, d, h)
        ws = torch.empty(nb, dtype=torch.uint8, device='cuda')
        fw = lambda: N.check(N.lib.dtb_afm_fwd(P(idx), P(table), P(offs), P(wa), P(ba), P(ph), P(pooled), b, f, d, h, 1, None, None), 'f')
        bw = lambda: N.check(N.lib.dtb_afm_bwd(P(idx), P(table), P(offs), P(wa), P(ba), P(ph), P(gp), P(gt), P(dwa), P(dba), P(dph), P(ws),
                                               nb, b, f, d, h, 1, None), 'b')
        print(f'afm rows {b} x {f} fields x D {d}, H {h}: fwd {timeit(fw):.3f} ms   bwd {timeit(bw):.3f} ms', flush=True)
        del table, gt, ws
        x = torch.randn(b, f, d, device='cuda')
        for bt, n_w in ((2, pairs), (1, f - 1), (0, 1)):
            w = torch.randn(n_w, d, d, device='cuda') / d ** 0.5
            out = torch.empty(b, pairs, d, device='cuda')
            go = torch.randn_like(out)
            dx, dw = torch.empty_like(x), torch.zeros_like(w)
            fw = lambda: N.check(N.lib.dtb_bilinear_fwd(P(x), P(w), P(out), b, f, d, bt, None), 'f')
            bw = lambda: N.check(N.lib.dtb_bilinear_bwd(P(x), P(w), P(go), P(dx), P(dw), b, f, d, bt, None), 'b')
            tf, tb = timeit(fw), timeit(bw)
            print(f'bilinear type {bt} rows {b}: fwd {tf:.3f} ms ({out.numel() * 4 / tf / 1e6:.0f} GB/s written)   bwd {tb:.3f} ms', flush=True)
            del out, go


def fgcnn():
    """The two FGCNN layers of the reference's defaults (14 / 16 filters, height 7, pool 2) at F = 26, D = 16."""
    for b in [int(r) for r in os.environ.get('ROWS', '16384,65536').split(',')]:
        h, w, cin = 26, 16, 1
        x = torch.randn(b, h, w, cin, device='cuda')
        for filters, kh, pool in ((14, 7, 2), (16, 7, 2)):
            k = torch.randn(kh, 1, cin, filters, device='cuda') / (kh * cin) ** 0.5
            bias = torch.zeros(filters, device='cuda')
            y = torch.empty(b, h, w, filters, device='cuda')
            ho = -(-h // pool)
            pooled = torch.empty(b, ho, w, filters, device='cuda')
            gy, gp = torch.randn_like(y), torch.randn_like(pooled)
            dx, dk, db, dy = torch.empty_like(x), torch.zeros_like(k), torch.zeros_like(bias), torch.empty_like(y)
            cf = lambda: N.check(N.lib.dtb_conv_fields_fwd(P(x), P(k), P(bias), P(y), b, h, w, cin, filters, kh, 2, None), 'cf')
            cb = lambda: N.check(N.lib.dtb_conv_fields_bwd(P(x), P(k), P(y), P(gy), P(dx), P(dk), P(db), b, h, w, cin, filters, kh, 2,
                                                           None), 'cb')
            pf = lambda: N.check(N.lib.dtb_maxpool_fields_fwd(P(y), P(pooled), b, h, w * filters, pool, None), 'pf')
            pb = lambda: N.check(N.lib.dtb_maxpool_fields_bwd(P(y), P(gp), P(dy), b, h, w * filters, pool, None), 'pb')
            print(f'fgcnn conv rows {b} [{h} x {w} x {cin}] -> {filters} filters, height {kh}: fwd {timeit(cf):.3f} ms   '
                  f'bwd {timeit(cb):.3f} ms   pool fwd {timeit(pf):.3f} ms   bwd {timeit(pb):.3f} ms', flush=True)
            x, h, cin = pooled.clone(), ho, filters
            del y, gy, dy


if __name__ == '__main__':
    if os.environ.get('ONLY', '') != 'fgcnn':
        main()
    fgcnn()
