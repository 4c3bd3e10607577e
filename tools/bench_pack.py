"""Isolate the cost of dtb_grad_rows_pack / unpack (row-wise table-gradient exchange kernels)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeptables_b200 import _native as N  # noqa: E402
from deeptables_b200._native import check, ptr, stream_ptr  # noqa: E402


def timed(fn, reps=1):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    B, F, D = 65536, 26, 16
    for vocab in (1000000, 1000):
        offs = torch.arange(F + 1, dtype=torch.int64, device='cuda') * vocab
        grad = torch.zeros(F * vocab, D, device='cuda')
        claim = torch.zeros(F * vocab, dtype=torch.int32, device='cuda')
        packed = torch.empty(B, F, D, device='cuda')
        g = torch.Generator(device='cuda').manual_seed(1)
        dy = torch.randn(B, F, D, device='cuda')
        step = 0
        res = {}
        for trial in range(4):
            idx = torch.randint(0, vocab, (B, F), generator=g, dtype=torch.int32, device='cuda')
            check(N.lib.dtb_embedding_scatter_add(ptr(idx), ptr(offs), ptr(dy), ptr(grad), B, F, D, stream_ptr()), 'scatter')
            step += 1
            s = step
            res['pack_after_scatter'] = timed(lambda: check(N.lib.dtb_grad_rows_pack(
                ptr(idx), ptr(offs), ptr(grad), ptr(claim), ptr(packed), s, B, F, D, stream_ptr()), 'pack'))
            res['pack_again_same_step(claims fail)'] = timed(lambda: check(N.lib.dtb_grad_rows_pack(
                ptr(idx), ptr(offs), ptr(grad), ptr(claim), ptr(packed), s, B, F, D, stream_ptr()), 'pack'))
            step += 1
            s = step
            res['pack_clean_rows'] = timed(lambda: check(N.lib.dtb_grad_rows_pack(
                ptr(idx), ptr(offs), ptr(grad), ptr(claim), ptr(packed), s, B, F, D, stream_ptr()), 'pack'))
            res['unpack'] = timed(lambda: check(N.lib.dtb_grad_rows_unpack(
                ptr(idx), ptr(offs), ptr(packed), ptr(grad), B, F, D, stream_ptr()), 'unpack'))
            res['scatter_add'] = timed(lambda: check(N.lib.dtb_embedding_scatter_add(
                ptr(idx), ptr(offs), ptr(dy), ptr(grad), B, F, D, stream_ptr()), 'scatter'))
            grad.zero_()
        print(f'vocab {vocab}: ' + '  '.join(f'{k}={v:.0f}us' for k, v in res.items()), flush=True)


if __name__ == '__main__':
    main()
