"""Independent brute-force definitions (numpy float64, explicit loops) of the interaction layers,
written from the defining formulas of the papers the reference cites rather than from its TF op
sequence.  They exist to pin oracle/layers_ref.py (which mirrors the op sequence) from a second
direction, because the reference ships no golden vectors (see the parity note in layers_ref.py).

TEST INFRASTRUCTURE ONLY.  Small cases only -- these are O(everything) Python loops.
"""
import numpy as np


def fm_pairs(x):
    """Rendle FM second-order term: sum_{i<j} <v_i, v_j>  (what layers.py:53-62 computes via the
    square-of-sum trick).  x: (B,F,D) -> (B,1)."""
    b, f, _ = x.shape
    out = np.zeros((b, 1))
    for n in range(b):
        s = 0.0
        for i in range(f):
            for j in range(i + 1, f):
                s += float(np.dot(x[n, i], x[n, j]))
        out[n, 0] = s
    return out


def linear_def(emb, dense, kernel):
    """deepnets.py:43-66 from the definition: sum_f w_f * (sum_d e_fd) + sum_c w_{F+c} * x_c."""
    b = emb.shape[0] if emb is not None else dense.shape[0]
    out = np.zeros((b, 1))
    for n in range(b):
        s, k = 0.0, 0
        if emb is not None:
            for f in range(emb.shape[1]):
                s += kernel[k, 0] * emb[n, f].sum()
                k += 1
        if dense is not None:
            for c in range(dense.shape[1]):
                s += kernel[k, 0] * dense[n, c]
                k += 1
        out[n, 0] = s
    return out


def cin_def(x, sizes, filters, direct=False, biases=None, act='relu'):
    """xDeepFM eq.(6): X^k_{h,*} = sum_{i,j} W^{k,h}_{ij} (X^0_{i,*} o X^{k-1}_{j,*}), with the
    reference's flattening index i*H_k + j (layers.py:693-694), half-split (716-720) and sum
    pooling over D (726).  filters[k]: (F0*H_k, L_k).  Returns pooled (B, sum L')."""
    b, f0, d = x.shape
    pooled_all = []
    for n in range(b):
        h = x[n]
        pooled = []
        for k, size in enumerate(sizes):
            hk = h.shape[0]
            t = np.zeros((size, d))
            for l in range(size):
                for e in range(d):
                    s = 0.0
                    for i in range(f0):
                        for j in range(hk):
                            s += filters[k][i * hk + j, l] * x[n, i, e] * h[j, e]
                    if biases is not None:
                        s += biases[k][l]
                    t[l, e] = max(s, 0.0) if act == 'relu' else s
            if direct:
                h, dc = t, t
            elif k != len(sizes) - 1:
                h, dc = t[:size // 2], t[size // 2:]
            else:
                h, dc = None, t
            pooled.append(dc.sum(axis=1))
        pooled_all.append(np.concatenate(pooled))
    return np.stack(pooled_all)


def cross_def(x, kernels, biases):
    """DCN eq.(3), per row: x_{l+1} = x0 * (x_l . w_l) + b_l + x_l."""
    out = np.zeros_like(x)
    for n in range(x.shape[0]):
        x0 = x[n]
        xl = x0.copy()
        for w, bias in zip(kernels, biases):
            xl = x0 * float(np.dot(xl, w[:, 0])) + bias[:, 0] + xl
        out[n] = xl
    return out


def attention_def(x, wq, bq, wk, bk, wv, bv, wr, br, num_heads, use_residual=True):
    """AutoInt interacting layer per row and head, before BatchNormalization
    (layers.py:115-150; note the reference applies relu to Q, K, V and the residual projection)."""
    b, f, d = x.shape
    dh = d // num_heads
    out = np.zeros((b, f, d))
    relu = lambda a: np.maximum(a, 0.0)
    for n in range(b):
        q = relu(x[n] @ wq + bq)
        k = relu(x[n] @ wk + bk)
        v = relu(x[n] @ wv + bv)
        for h in range(num_heads):
            sl = slice(h * dh, (h + 1) * dh)
            for i in range(f):
                sc = np.array([np.dot(q[i, sl], k[j, sl]) for j in range(f)]) / np.sqrt(dh)
                sc = np.exp(sc - sc.max())
                sc = sc / sc.sum()
                acc = np.zeros(dh)
                for j in range(f):
                    acc += sc[j] * v[j, sl]
                out[n, i, sl] = acc
        if use_residual:
            out[n] += relu(x[n] @ wr + br)
        out[n] = relu(out[n])
    return out


def inner_product_def(x):
    """PNN inner products, pairs (i<j) in row-major order (layers.py:478-483). x: (B,F,D)."""
    b, f, _ = x.shape
    cols = []
    for i in range(f - 1):
        for j in range(i + 1, f):
            cols.append(np.einsum('bd,bd->b', x[:, i], x[:, j]))
    return np.stack(cols, axis=1)


def outer_product_def(x, kernel, kernel_type='mat'):
    """PNN outer products contracted with the kernel: mat: p^T K_p^T q with K indexed [k, pair, d]
    (layers.py:536,557-574) i.e. sum_{k,d} p_d K[k,p,d] q_k; vec: sum_d p_d q_d k_{p,d}; num."""
    b, f, d = x.shape
    out = []
    pair = 0
    for i in range(f - 1):
        for j in range(i + 1, f):
            col = np.zeros(b)
            for n in range(b):
                p, q = x[n, i], x[n, j]
                if kernel_type == 'mat':
                    s = 0.0
                    for k in range(d):
                        s += float(np.dot(p, kernel[k, pair])) * q[k]
                elif kernel_type == 'vec':
                    s = float(np.sum(p * q * kernel[pair]))
                else:
                    s = float(np.sum(p * q) * kernel[pair, 0])
                col[n] = s
            out.append(col)
            pair += 1
    return np.stack(out, axis=1)


def batch_norm_def(x, gamma, beta, eps=1e-3):
    """Training-mode BN over every axis but the last (biased variance)."""
    flat = x.reshape(-1, x.shape[-1])
    mean = flat.mean(axis=0)
    var = ((flat - mean) ** 2).mean(axis=0)
    return (x - mean) / np.sqrt(var + eps) * gamma + beta, mean, var
