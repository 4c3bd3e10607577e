"""Debug aid: train the five-net BASELINE config on one repeated batch and report, per step, the loss and which
gradient tensors hold non-finite values.  DTB_CIN_PRECISION selects the CIN mode (0 auto, 2 bf16x3, 4 fp16)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeptables_b200 import deeptable, engine as E  # noqa: E402
from deeptables_b200.deepmodel import DeepModel  # noqa: E402
from deeptables_b200.metainfo import CategoricalColumn, ContinuousColumn  # noqa: E402

F, C, V, dim, b = 26, 13, 1_000_000, 16, int(os.environ.get('B', 16384))
nets = os.environ.get('NETS', 'fm_nets,cin_nets,cross_nets,autoint_nets,pnn_nets').split(',')
conf = deeptable.ModelConfig(nets=nets, embeddings_output_dim=dim, embedding_dropout=0, metrics=['AUC'],
                             cin_params={'cross_layer_size': (128, 128, 128), 'activation': 'relu', 'use_residual': False,
                                         'use_bias': False, 'direct': False, 'reduce_D': False})
cats = [CategoricalColumn(f'C{i + 1}', V, dim) for i in range(F)]
conts = [ContinuousColumn('input_continuous_all', [f'I{i + 1}' for i in range(C)])]
model = DeepModel('binary', 2, conf, cats, conts, seed=21)
model._build_model()
g = torch.Generator().manual_seed(5)
idx = torch.randint(0, V, (b, F), generator=g, dtype=torch.int32)
cont = torch.randn(b, C, generator=g)
half = b // 2
idx[half:] = idx[:half]
cont[half:] = cont[:half]
y = (torch.rand(b, generator=g) < 0.25).float()
d_idx, d_cont, d_y = idx.cuda(), cont.cuda(), y.cuda().view(-1, 1)
scope, t = model._scope, model.table
_orig_bwd = E.CINFn.backward


def _spy(ctx, g_):
    ws, saved = ctx.ws, ctx.saved_buf
    out = _orig_bwd(ctx, g_)
    torch.cuda.synchronize()
    m_pad = ((b + 15) // 16) * 256
    wpack = 26 * 128 * (32 + 64 + 64) * 4
    dc = 3 * (m_pad // 16) * 64 * 128
    end = wpack + 1024 + dc + 1024
    st_i = ws[end - 256:end].view(torch.int32)
    st_f = ws[end - 256:end].view(torch.float32)
    print('   stats words: max|W_k|', [f'{float(v):.3e}' for v in st_f[0:3]], ' max|dC_k| bound', [f'{float(v):.3e}' for v in st_f[8:11]],
          ' max|x0|', f'{float(st_f[16]):.3e}', ' max|h_k|', [f'{float(v):.3e}' for v in st_f[24:27]], flush=True)
    print('   forward maxima (saved head):', [f'{float(v):.3e}' for v in saved[:16].view(torch.float32)], flush=True)
    print('   d_pooled: max', f'{float(g_.abs().max()):.3e}', 'finite', bool(torch.isfinite(g_).all()),
          ' dW finite per layer', [bool(torch.isfinite(x).all()) for x in out[1].split([676 * 128, 1664 * 128, 1664 * 128])], flush=True)
    dpm = ws[end:end + b * 8 * 4].view(torch.float32).view(b, 8)[:, :3]
    print('   dpmax: min', [f'{float(v):.3e}' for v in dpm.min(0).values], 'max', [f'{float(v):.3e}' for v in dpm.max(0).values], flush=True)
    return out


E.CINFn.backward = staticmethod(_spy)
orig_adam = E.N.lib.dtb_adam_dense
for step in range(4):
    model._loss_acc.zero_()
    # run the step but look at the gradients before the optimiser consumes them: hook by a manual forward/backward
    t.ensure_training_state()
    model._catch_up(d_idx, model._step)
    t.pending_bwd, t.on_grad_final = 0, None
    z = model._forward(d_idx, d_cont, training=True)
    prob, dz = E.loss_forward_backward(z, d_y, model.task, None, True, model._loss_acc)
    z.backward(dz)
    torch.cuda.synchronize()
    bad = [n for n, p in scope.params.items() if not bool(torch.isfinite(p.grad).all())]
    tg = t.grad
    print(f'step {step}: loss {float(model._loss_acc.item()) / b:.4f}  z finite {bool(torch.isfinite(z).all())}  '
          f'table grad finite {bool(torch.isfinite(tg).all())} max {float(tg.abs().max()):.3e}  bad dense grads {bad[:6]}', flush=True)
    for n, p in scope.params.items():
        if n.startswith('cin/') or 'exFM' in n:
            print(f'    {n}: |w| max {float(p.abs().max()):.3e}  |g| max {float(p.grad.abs().max()):.3e}', flush=True)
    # optimiser step exactly as train_step does
    s = model._step + 1
    alpha = E.adam_alpha(s)
    E.check(E.N.lib.dtb_adam_dense(E.ptr(scope.flat_p), E.ptr(scope.flat_m), E.ptr(scope.flat_v), E.ptr(scope.flat_g),
                                   scope.flat_p.numel(), alpha, E.ADAM_B1, E.ADAM_B2, E.ADAM_EPS, 1, E.stream_ptr()), 'adam')
    a = model._alpha_table(s)
    E.check(E.N.lib.dtb_adam_rows_apply(E.ptr(d_idx), E.ptr(t.row_offsets), E.ptr(t.weight), E.ptr(t.m), E.ptr(t.v), E.ptr(t.grad),
                                        E.ptr(t.last_step), E.ptr(a), s, E.ADAM_B1, E.ADAM_B2, E.ADAM_EPS, d_idx.shape[0],
                                        t.n_fields, t.dim, E.stream_ptr()), 'adam_rows')
    model._step = s
