"""Debug aid: fp16 CIN backward at the BASELINE shape with an upstream gradient of configurable magnitude / sparsity;
prints the statistics words of the workspace and whether the filter gradient is finite."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeptables_b200 import _native as nat  # noqa: E402

P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())  # noqa: E731
B = int(os.environ.get('B', 16384))
F, D, sizes, V = 26, 16, (128, 128, 128), 100000
sizes_c = nat.int_array(sizes)
g = torch.Generator(device='cuda').manual_seed(0)
table = (torch.rand(F * V, D, device='cuda', generator=g) - 0.5) * 0.1
offs = torch.arange(F + 1, dtype=torch.int64, device='cuda') * V
idx = torch.randint(0, V, (B, F), device='cuda', dtype=torch.int32, generator=g)
K = [26 * 26, 26 * 64, 26 * 64]
w = torch.cat([(torch.rand(k * 128, device='cuda', generator=g) * 2 - 1) * (6.0 / k) ** 0.5 for k in K])
ws_bytes = nat.lib.dtb_cin_workspace_bytes(B, F, D, sizes_c, 3, 0, 1)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device='cuda')
saved = torch.empty(nat.lib.dtb_cin_saved_bytes(B, F, D, sizes_c, 3, 0), dtype=torch.uint8, device='cuda')
pooled = torch.empty(B, 256, device='cuda')
for name, dp in (('randn*1e-3', torch.randn(B, 256, device='cuda', generator=g) * 1e-3),
                 ('rank-1 (dz * w_out), like a Dense(1) head', (torch.randn(B, 1, device='cuda', generator=g) * 1e-4) *
                  torch.randn(1, 256, device='cuda', generator=g) * 0.1),
                 ('half the rows exactly zero', torch.randn(B, 256, device='cuda', generator=g) * 1e-3 *
                  (torch.arange(B, device='cuda') % 2).float().view(-1, 1))):
    for prec in (4, 2):
        grad = torch.zeros_like(table)
        dw = torch.zeros_like(w)
        nat.check(nat.lib.dtb_cin_fwd(P(idx), P(table), P(offs), P(w), None, P(pooled), P(saved), P(ws), ws_bytes, B, F, D,
                                      sizes_c, 3, 0, 1, prec, None, None), 'fwd')
        nat.check(nat.lib.dtb_cin_bwd(P(idx), P(table), P(offs), P(w), P(dp.contiguous()), P(saved), P(grad), P(dw), None, P(ws),
                                      ws_bytes, B, F, D, sizes_c, 3, 0, 1, prec, None), 'bwd')
        torch.cuda.synchronize()
        print(f'{name}: precision {prec}: dW finite {bool(torch.isfinite(dw).all())} max {float(dw.abs().max()):.3e}; '
              f'table grad finite {bool(torch.isfinite(grad).all())} max {float(grad.abs().max()):.3e}', flush=True)
        if prec == 4:
            sv = saved[:64].view(torch.float32)
            print('   forward maxima words:', [f'{float(v):.3e}' for v in sv[:4]], flush=True)
