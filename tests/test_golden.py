"""Golden vectors (tests/golden/hotpath_small.npz, made by tests/golden/make_golden.py from the
brute-force float64 definitions): the op-sequence oracle must reproduce them on CPU, the CUDA kernels
on the GPU (through the C ABI)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import layers_ref as L

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'hotpath_small.npz'))


def emb_list(dtype=torch.float64):
    x = torch.tensor(G['x'], dtype=dtype)
    return x, [x[:, i:i + 1, :] for i in range(x.shape[1])]


def test_oracle_reproduces_golden_vectors():
    x, emb = emb_list()
    dense = torch.tensor(G['dense'])
    np.testing.assert_allclose(L.linear(emb, dense, torch.tensor(G['w_lin'])).numpy(), G['linear_out'], rtol=1e-10)
    np.testing.assert_allclose(L.fm(x).numpy(), G['fm_out'], rtol=1e-10, atol=1e-12)
    sizes = tuple(int(s) for s in G['cin_sizes'])
    w = {f'f_{k}': torch.tensor(G[f'cin_f{k}'][None]) for k in range(len(sizes))}
    pw = G['cin_pooled'].shape[1]
    for col in (0, pw // 2, pw - 1):
        kern = torch.zeros(pw, 1, dtype=torch.float64)
        kern[col] = 1
        w['exFM_out/kernel'], w['exFM_out/bias'] = kern, torch.zeros(1, dtype=torch.float64)
        got = L.cin(x, dict(cross_layer_size=sizes, direct=False), w).numpy()[:, 0]
        np.testing.assert_allclose(got, G['cin_pooled'][:, col], rtol=1e-9, atol=1e-12)
    ks = [torch.tensor(k).reshape(-1, 1) for k in G['cross_k']]
    bs = [torch.tensor(b).reshape(-1, 1) for b in G['cross_b']]
    np.testing.assert_allclose(L.cross(torch.tensor(G['cross_in']), ks, bs).numpy(), G['cross_out'], rtol=1e-10)
    np.testing.assert_allclose(L.inner_product(emb).numpy(), G['pnn_ip'], rtol=1e-10)
    np.testing.assert_allclose(L.outer_product(emb, torch.tensor(G['pnn_kernel_mat']), 'mat').numpy(), G['pnn_op_mat'],
                               rtol=1e-9, atol=1e-12)
    y, _, _ = L.batch_norm(torch.tensor(G['cross_in']), torch.tensor(G['bn_gamma']), torch.tensor(G['bn_beta']),
                           torch.zeros(G['bn_mean'].shape[0], dtype=torch.float64),
                           torch.ones(G['bn_mean'].shape[0], dtype=torch.float64), True)
    np.testing.assert_allclose(y.numpy(), G['bn_out'], rtol=1e-9)


@pytest.mark.gpu
def test_cuda_kernels_reproduce_golden_vectors():
    from deeptables_b200 import _native as nat
    keep = []

    def dev(a, dtype=torch.float32):
        t = torch.as_tensor(np.ascontiguousarray(a)).to(dtype).cuda()
        keep.append(t)
        return t

    def P(t):
        return None if t is None else ctypes.c_void_p(t.data_ptr())

    vocab = G['vocab']
    b, f, d = G['x'].shape
    c = G['dense'].shape[1]
    offs = dev(np.concatenate([[0], np.cumsum(vocab)]), torch.int64)
    idx, tab, dense = dev(G['idx'], torch.int32), dev(G['table_flat']), dev(G['dense'])
    lin, fm = torch.empty(b, device='cuda'), torch.empty(b, device='cuda')
    nat.check(nat.lib.dtb_fm_linear_fwd(P(idx), P(tab), P(offs), P(dense), P(dev(G['w_lin'][:, 0])), P(lin), P(fm),
                                        b, f, d, c, None, None))
    np.testing.assert_allclose(lin.cpu().numpy(), G['linear_out'][:, 0], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(fm.cpu().numpy(), G['fm_out'][:, 0], rtol=1e-4, atol=1e-5)
    sizes = tuple(int(s) for s in G['cin_sizes'])
    sizes_c = nat.int_array(sizes)
    wcat = dev(np.concatenate([G[f'cin_f{k}'].reshape(-1) for k in range(len(sizes))]))
    pooled = torch.empty(b, G['cin_pooled'].shape[1], device='cuda')
    ws_bytes = nat.lib.dtb_cin_workspace_bytes(b, f, d, sizes_c, len(sizes), 0, 0)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device='cuda')
    nat.check(nat.lib.dtb_cin_fwd(P(idx), P(tab), P(offs), P(wcat), None, P(pooled), None, P(ws), ws_bytes, b, f, d,
                                  sizes_c, len(sizes), 0, 1, 0, None, None))
    np.testing.assert_allclose(pooled.cpu().numpy(), G['cin_pooled'], rtol=1e-3, atol=1e-5)
    w = G['cross_in'].shape[1]
    y = torch.empty(b, w, device='cuda')
    xw = torch.empty(b, 3, device='cuda')
    nat.check(nat.lib.dtb_cross_fwd(P(dev(G['cross_in'])), P(dev(G['cross_k'])), P(dev(G['cross_b'])), P(y), P(xw), b, w,
                                    3, None))
    np.testing.assert_allclose(y.cpu().numpy(), G['cross_out'], rtol=1e-4, atol=1e-5)
    pairs = f * (f - 1) // 2
    ip, op = torch.empty(b, pairs, device='cuda'), torch.empty(b, pairs, device='cuda')
    nat.check(nat.lib.dtb_pnn_fwd(P(idx), P(tab), P(offs), P(dev(G['pnn_kernel_mat'])), P(ip), P(op), b, f, d, 0, None, None))
    np.testing.assert_allclose(ip.cpu().numpy(), G['pnn_ip'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(op.cpu().numpy(), G['pnn_op_mat'], rtol=1e-4, atol=1e-5)
    # attention: projections via dtb_dense_fwd on [B*F, D] rows with the 4 kernels side by side
    wq = dev(np.concatenate(list(G['att_w']), axis=1))
    bq = dev(np.concatenate(list(G['att_b'])))
    qkvr = torch.empty(b * f, 4 * d, device='cuda')
    wsb = nat.lib.dtb_dense_workspace_bytes(d, 4 * d)
    wsd = torch.empty(max(wsb, 16), dtype=torch.uint8, device='cuda')
    nat.check(nat.lib.dtb_dense_fwd(P(dev(G['x'].reshape(b * f, d))), P(wq), P(bq), P(qkvr), P(wsd), wsb, b * f, d, 4 * d, 1,
                                    None))
    ya = torch.empty(b, f, d, device='cuda')
    nat.check(nat.lib.dtb_attention_core_fwd(P(qkvr), P(ya), b, f, d, 2, 1, None))
    # projections wider than 8 run as bf16x3 tensor-core GEMMs (2^-16 per product): 5e-5 absolute on O(1) values
    np.testing.assert_allclose(ya.cpu().numpy(), G['att_out'], rtol=1e-4, atol=5e-5)
    torch.cuda.synchronize()
