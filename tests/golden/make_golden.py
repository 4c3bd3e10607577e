"""Generate tests/golden/hotpath_small.npz.

The reference (DeepTables on TensorFlow/Keras 3) cannot be imported in this environment
(tensorflow / keras / hypernets are not installed and there is no network), and its own tests hold
no numeric fixtures for the hot path -- so these vectors are NOT reference outputs.  They are
produced by the independent brute-force float64 definitions in oracle/bruteforce.py (explicit
loops written from the papers' formulas), which pin both the op-sequence oracle
(oracle/layers_ref.py, CPU tests) and the CUDA kernels (GPU tests) from a second direction.
Anyone with TensorFlow can replay the same inputs through the reference layers and compare.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bruteforce as BF   # noqa: E402


def main():
    g = np.random.default_rng(20260922)
    out = {}
    b, f, d, c = 6, 5, 4, 3
    vocab = [7, 5, 9, 4, 6]
    tables = [g.uniform(-0.5, 0.5, size=(v, d)) for v in vocab]
    idx = np.stack([g.integers(0, v, size=b) for v in vocab], axis=1).astype(np.int32)
    x = np.stack([tables[i][idx[:, i]] for i in range(f)], axis=1)            # (B,F,D)
    dense = g.normal(size=(b, c))
    out['vocab'] = np.array(vocab)
    out['table_flat'] = np.concatenate(tables, axis=0)
    out['idx'] = idx
    out['dense'] = dense
    out['x'] = x
    # linear + FM
    w_lin = g.normal(size=(f + c, 1))
    out['w_lin'] = w_lin
    out['linear_out'] = BF.linear_def(x, dense, w_lin)
    out['fm_out'] = BF.fm_pairs(x)
    # CIN (6,4,4), direct False
    sizes = (6, 4, 4)
    hs = [f, 3, 2]
    filt = [g.normal(size=(f * hs[k], s)) * 0.5 for k, s in enumerate(sizes)]
    for k, w in enumerate(filt):
        out[f'cin_f{k}'] = w
    out['cin_sizes'] = np.array(sizes)
    out['cin_pooled'] = BF.cin_def(x, sizes, filt, direct=False)
    # Cross, 3 layers on a (B, F*D + C) input
    xin = np.concatenate([x.reshape(b, -1), dense], axis=1)
    ks = [g.normal(size=(xin.shape[1], 1)) * 0.2 for _ in range(3)]
    bs = [g.normal(size=(xin.shape[1], 1)) * 0.1 for _ in range(3)]
    out['cross_in'] = xin
    out['cross_k'] = np.stack([k_[:, 0] for k_ in ks])
    out['cross_b'] = np.stack([b_[:, 0] for b_ in bs])
    out['cross_out'] = BF.cross_def(xin, ks, bs)
    # attention (2 heads) before BN
    ws = [g.normal(size=(d, d)) * 0.7 for _ in range(4)]
    bb = [g.normal(size=d) * 0.1 for _ in range(4)]
    out['att_w'] = np.stack(ws)
    out['att_b'] = np.stack(bb)
    out['att_out'] = BF.attention_def(x, ws[0], bb[0], ws[1], bb[1], ws[2], bb[2], ws[3], bb[3], 2, True)
    # PNN
    pairs = f * (f - 1) // 2
    kern = g.normal(size=(d, pairs, d)) * 0.5
    out['pnn_kernel_mat'] = kern
    out['pnn_ip'] = BF.inner_product_def(x)
    out['pnn_op_mat'] = BF.outer_product_def(x, kern, 'mat')
    # BN (training mode)
    gamma, beta = g.normal(size=xin.shape[1]) + 2.0, g.normal(size=xin.shape[1])
    y, mean, var = BF.batch_norm_def(xin, gamma, beta)
    out['bn_gamma'], out['bn_beta'], out['bn_out'], out['bn_mean'], out['bn_var'] = gamma, beta, y, mean, var
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'hotpath_small.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
