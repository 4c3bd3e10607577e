"""Pin the CPU oracle (oracle/layers_ref.py, mirrors the reference's TF op sequence) against
independent brute-force definitions (oracle/bruteforce.py) and structural properties.
The reference has no golden vectors for this path -- these tests ARE the pin."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import layers_ref as L
from oracle import bruteforce as BF
from oracle import model_ref as M

torch.manual_seed(0)
T64 = dict(dtype=torch.float64)


def rnd(*shape, seed=0, scale=1.0):
    g = np.random.default_rng(seed)
    return g.normal(size=shape) * scale


def as_list(x):  # (B,F,D) -> list of (B,1,D)
    return [x[:, i:i + 1, :] for i in range(x.shape[1])]


@pytest.mark.parametrize('b,f,d', [(3, 5, 4), (2, 26, 16), (1, 1, 3)])
def test_fm_matches_pairwise_definition(b, f, d):
    x = rnd(b, f, d, seed=1)
    got = L.fm(torch.tensor(x)).numpy()
    np.testing.assert_allclose(got, BF.fm_pairs(x), rtol=1e-10, atol=1e-12)


def test_fm_custom_copy_in_reference_tests():
    # deeptables/tests/models/nets_test.py:192-206 (CustomFM) is a second copy of FM.call
    x = torch.tensor(rnd(4, 6, 3, seed=2))
    sq = torch.square(x.sum(dim=1, keepdim=True))
    ss = (x * x).sum(dim=1, keepdim=True)
    ref = 0.5 * (sq - ss).sum(dim=2)
    torch.testing.assert_close(L.fm(x), ref)


@settings(max_examples=20, deadline=None)
@given(st.integers(1, 4), st.integers(2, 7), st.integers(1, 5), st.integers(0, 10 ** 6))
def test_fm_permutation_invariant(b, f, d, seed):
    x = torch.tensor(rnd(b, f, d, seed=seed))
    perm = torch.randperm(f)
    torch.testing.assert_close(L.fm(x), L.fm(x[:, perm]), rtol=1e-9, atol=1e-10)


def test_linear_definition():
    emb = rnd(4, 5, 3, seed=3)
    dense = rnd(4, 2, seed=4)
    k = rnd(7, 1, seed=5)
    got = L.linear(as_list(torch.tensor(emb)), torch.tensor(dense), torch.tensor(k)).numpy()
    np.testing.assert_allclose(got, BF.linear_def(emb, dense, k), rtol=1e-10)
    got = L.linear(as_list(torch.tensor(emb)), None, torch.tensor(k[:5])).numpy()
    np.testing.assert_allclose(got, BF.linear_def(emb, None, k[:5]), rtol=1e-10)
    got = L.linear([], torch.tensor(dense), torch.tensor(k[:2])).numpy()
    np.testing.assert_allclose(got, BF.linear_def(None, dense, k[:2]), rtol=1e-10)


@pytest.mark.parametrize('direct', [False, True])
@pytest.mark.parametrize('sizes', [(4, 6), (6, 4, 5), (3,)])
def test_cin_matches_paper_definition(direct, sizes):
    b, f0, d = 2, 3, 4
    x = rnd(b, f0, d, seed=6)
    params = dict(cross_layer_size=sizes, direct=direct, activation='relu')
    fns = L.cin_field_nums(f0, sizes, direct)
    filt = [rnd(f0 * fns[k], s, seed=10 + k, scale=0.5) for k, s in enumerate(sizes)]
    pooled = BF.cin_def(x, sizes, filt, direct=direct)
    width = L.cin_pooled_width(f0, params)
    assert pooled.shape == (b, width)
    wk = rnd(width, 1, seed=20)
    w = {f'f_{k}': torch.tensor(filt[k][None]) for k in range(len(sizes))}
    w['exFM_out/kernel'] = torch.tensor(wk)
    w['exFM_out/bias'] = torch.tensor([0.25])
    got = L.cin(torch.tensor(x), params, w).numpy()
    np.testing.assert_allclose(got, pooled @ wk + 0.25, rtol=1e-9, atol=1e-11)


def test_cin_bias_residual_reduce_d_shapes_and_values():
    b, f0, d = 3, 4, 2
    sizes = (4, 2)
    x = torch.tensor(rnd(b, f0, d, seed=7))
    params = dict(cross_layer_size=sizes, direct=False, use_bias=True, use_residual=True,
                  reduce_D=True, activation='relu')
    cfg = dict(nets=['cin_nets'], cin_params=params, embeddings_initializer='uniform')
    state = M.init_state(cfg, [5] * f0, [d] * f0, 0, seed=1)
    w = {k[4:]: v.double() for k, v in state.items() if k.startswith('cin/')}
    w['bias0'] = torch.tensor(rnd(4, seed=8))
    out = L.cin(x, params, w)
    assert out.shape == (b, 1)
    # reduce_D is only a reparametrisation of the filter: f = transpose(reshape(f0_ @ f__))
    p2 = dict(params, reduce_D=False)
    w2 = dict(w)
    fns = L.cin_field_nums(f0, sizes, False)
    for k, s in enumerate(sizes):
        f_m = w[f'f0_{k}'] @ w[f'f__{k}']
        w2[f'f_{k}'] = f_m.reshape(1, s, f0 * fns[k]).permute(0, 2, 1)
    torch.testing.assert_close(out, L.cin(x, p2, w2))


def test_cin_odd_layer_rejected():
    with pytest.raises(ValueError):
        L.cin_field_nums(5, (3, 4), direct=False)
    assert L.cin_field_nums(5, (4, 3), direct=False) == [5, 2, 1]   # last layer may be odd
    assert L.cin_field_nums(5, (3, 4), direct=True) == [5, 3, 4]


def test_cross_definition_and_zero_kernel_property():
    x = rnd(5, 7, seed=9)
    ks = [rnd(7, 1, seed=30 + i) for i in range(3)]
    bs = [rnd(7, 1, seed=40 + i) for i in range(3)]
    got = L.cross(torch.tensor(x), [torch.tensor(k) for k in ks], [torch.tensor(b) for b in bs]).numpy()
    np.testing.assert_allclose(got, BF.cross_def(x, ks, bs), rtol=1e-10)
    zero = [torch.zeros(7, 1, **T64)] * 3
    got0 = L.cross(torch.tensor(x), zero, [torch.tensor(b) for b in bs]).numpy()
    np.testing.assert_allclose(got0, x + sum(b[:, 0] for b in bs), rtol=1e-12)


@pytest.mark.parametrize('heads,res', [(1, True), (2, True), (4, False)])
def test_attention_definition(heads, res):
    b, f, d = 2, 5, 4
    x = rnd(b, f, d, seed=11)
    ws = {n: rnd(d, d, seed=50 + i, scale=0.7) for i, n in enumerate('QKVR')}
    bs = {n: rnd(d, seed=60 + i, scale=0.1) for i, n in enumerate('QKVR')}
    pre = BF.attention_def(x, ws['Q'], bs['Q'], ws['K'], bs['K'], ws['V'], bs['V'], ws['R'], bs['R'],
                           heads, use_residual=res)
    gamma, beta = rnd(d, seed=70) + 2.0, rnd(d, seed=71)
    want, _, _ = BF.batch_norm_def(pre, gamma, beta)
    w = {}
    for n, key in zip('QKVR', ['dense_Q', 'dense_K', 'dense_V', 'dense_residual']):
        w[f'{key}/kernel'] = torch.tensor(ws[n])
        w[f'{key}/bias'] = torch.tensor(bs[n])
    w['batch_normalize/gamma'] = torch.tensor(gamma)
    w['batch_normalize/beta'] = torch.tensor(beta)
    st_ = {'moving_mean': torch.zeros(d, **T64), 'moving_variance': torch.ones(d, **T64)}
    got, ns = L.multihead_attention(torch.tensor(x), dict(num_heads=heads, use_residual=res), w, st_, True)
    np.testing.assert_allclose(got.numpy(), want, rtol=1e-9, atol=1e-10)
    assert ns['moving_mean'].shape == (d,)


def test_inner_outer_product_definitions_and_pair_order():
    b, f, d = 3, 5, 4
    x = rnd(b, f, d, seed=12)
    emb = as_list(torch.tensor(x))
    row, col = L.pair_lists(f)
    assert (row[:4], col[:4]) == ([0, 0, 0, 0], [1, 2, 3, 4]) and len(row) == 10
    np.testing.assert_allclose(L.inner_product(emb).numpy(), BF.inner_product_def(x), rtol=1e-10)
    for kt, shape in (('mat', (d, 10, d)), ('vec', (10, d)), ('num', (10, 1))):
        k = rnd(*shape, seed=13)
        got = L.outer_product(emb, torch.tensor(k), kt).numpy()
        np.testing.assert_allclose(got, BF.outer_product_def(x, k, kt), rtol=1e-9, atol=1e-12)


def test_batch_norm_train_and_inference():
    x = torch.tensor(rnd(16, 6, seed=14) * 3 + 1)
    g, bta = torch.tensor(rnd(6, seed=15)), torch.tensor(rnd(6, seed=16))
    y, nm, nv = L.batch_norm(x, g, bta, torch.zeros(6, **T64), torch.ones(6, **T64), True)
    want, mean, var = BF.batch_norm_def(x.numpy(), g.numpy(), bta.numpy())
    np.testing.assert_allclose(y.numpy(), want, rtol=1e-10)
    np.testing.assert_allclose(nm.numpy(), 0.01 * mean, rtol=1e-10)
    np.testing.assert_allclose(nv.numpy(), 0.99 + 0.01 * var, rtol=1e-10)
    y2, _, _ = L.batch_norm(x, g, bta, nm, nv, False)
    np.testing.assert_allclose(y2.numpy(), (x.numpy() - nm.numpy()) / np.sqrt(nv.numpy() + 1e-3)
                               * g.numpy() + bta.numpy(), rtol=1e-10)


def test_embedding_lookup_float_ids_and_range():
    tabs = [torch.arange(12.).reshape(4, 3), torch.arange(10.).reshape(5, 2)]
    idx = torch.tensor([[3.0, 0.0], [1.0, 4.0]])           # float32-encoded ids (dataset_generator.py:41)
    out = L.embedding_lookup(tabs, idx)
    assert out[0].shape == (2, 1, 3) and out[1].shape == (2, 1, 2)
    torch.testing.assert_close(out[0][:, 0], tabs[0][[3, 1]])
    torch.testing.assert_close(out[1][:, 0], tabs[1][[0, 4]])
    with pytest.raises(IndexError):
        L.embedding_lookup(tabs, torch.tensor([[4, 0]]))
    flat = L.flatten_embeddings(out)
    assert flat.shape == (2, 5)                              # field-major, dim-minor
    torch.testing.assert_close(flat[0], torch.cat([tabs[0][3], tabs[1][0]]))


def test_bce_and_adam_closed_forms():
    y = torch.tensor([[1.0], [0.0], [1.0]])
    p = torch.tensor([[0.9], [0.2], [1.0]])
    want = -(np.log(0.9) + np.log(0.8) + np.log(1 - 1e-7)) / 3
    assert abs(float(L.binary_crossentropy(y, p)) - want) < 1e-6
    pr = torch.tensor([1.0]); g = torch.tensor([0.5]); m = torch.zeros(1); v = torch.zeros(1)
    L.adam_step(pr, g, m, v, 1)
    # first Adam step moves by ~lr regardless of gradient scale
    assert abs(float(pr) - (1.0 - 1e-3)) < 1e-6
    assert abs(float(m) - 0.05) < 1e-7 and abs(float(v) - 0.00025) < 1e-9


ALL_NETS = ['linear', 'fm_nets', 'cin_nets', 'dnn_nets', 'cross_nets', 'dcn_nets', 'cross_dnn_nets',
            'autoint_nets', 'pnn_nets', 'ipnn_nets', 'opnn_nets']


def small_config(nets, **kw):
    cfg = dict(nets=nets, stacking_op='add', output_use_bias=True, embeddings_initializer='uniform',
               dnn_params={'hidden_units': ((8, 0, False), (4, 0, True)), 'activation': 'relu'},
               autoint_params={'num_attention': 2, 'num_heads': 2, 'dropout_rate': 0, 'use_residual': True},
               cross_params={'num_cross_layer': 3}, pnn_params={'outer_product_kernel_type': 'mat'},
               cin_params={'cross_layer_size': (6, 4), 'activation': 'relu', 'use_residual': False,
                           'use_bias': False, 'direct': False, 'reduce_D': False})
    cfg.update(kw)
    return cfg


@pytest.mark.parametrize('nets', [[n] for n in ALL_NETS] + [['linear', 'cin_nets', 'dnn_nets'], ALL_NETS])
def test_model_oracle_runs_and_trains(nets):
    vocab, dims, n_cont, b = [7, 5, 9, 4], [4] * 4, 3, 16
    cfg = small_config(nets)
    state = M.init_state(cfg, vocab, dims, n_cont, seed=3)
    g = np.random.default_rng(0)
    idx = torch.tensor(np.stack([g.integers(0, v, size=b) for v in vocab], axis=1).astype(np.int32))
    cont = torch.tensor(g.normal(size=(b, n_cont)).astype(np.float32))
    y = torch.tensor((g.random(b) < 0.4).astype(np.float32))
    tr = M.RefTrainer(state, cfg, len(vocab))
    p0 = tr.predict(idx, cont)
    assert p0.shape == (b, 1) and bool(((p0 > 0) & (p0 < 1)).all())
    losses = [tr.train_step(idx, cont, y) for _ in range(30)]
    assert losses[-1] < losses[0]


def test_model_oracle_single_categorical_and_no_continuous():
    # reference edge cases: deeptables/tests/models/nets_test.py:166-189, model_input_test.py
    for nets in (['linear', 'fm_nets', 'cin_nets', 'autoint_nets', 'dnn_nets', 'pnn_nets'],):
        cfg = small_config(nets)
        state = M.init_state(cfg, [6], [4], 0, seed=1)
        idx = torch.tensor([[1], [5], [0]], dtype=torch.int32)
        out, _ = M.forward(state, cfg, idx, None, 1, False)
        assert out.shape == (3, 1)      # pnn_nets skipped: needs >= 2 embeddings
    cfg = small_config(['linear', 'dnn_nets', 'cross_nets', 'fm_nets'])
    state = M.init_state(cfg, [], [], 5, seed=1)
    out, _ = M.forward(state, cfg, None, torch.randn(4, 5), 0, False)
    assert out.shape == (4, 1)          # fm_nets skipped: no embeddings


def test_concat_stacking_and_multiclass_head():
    cfg = small_config(['linear', 'dnn_nets', 'fm_nets'], stacking_op='concat')
    state = M.init_state(cfg, [5, 6], [3, 3], 2, task='multiclass', num_classes=3, seed=2)
    assert state['task_output/kernel'].shape == (3, 3)
    idx = torch.tensor([[1, 2], [4, 5]], dtype=torch.int32)
    out, _ = M.forward(state, cfg, idx, torch.randn(2, 2), 2, False, task='multiclass')
    torch.testing.assert_close(out.sum(dim=-1), torch.ones(2))
