"""GPU parity of the assembled model (DeepModel) against the CPU oracle's model restatement:
forward, gradients-through-training (Adam trajectories) and the public fit/predict surface."""
import os

import numpy as np
import pandas as pd
import pytest
import torch

from oracle import model_ref as M

pytestmark = pytest.mark.gpu


def build(nets, vocab, dim, n_cont, seed=5, **cfg_kw):
    from deeptables_b200 import deeptable
    from deeptables_b200.deepmodel import DeepModel
    from deeptables_b200.metainfo import CategoricalColumn, ContinuousColumn
    kw = dict(nets=nets, embeddings_output_dim=dim, embedding_dropout=0, metrics=['AUC'],
              dnn_params={'hidden_units': ((16, 0, False), (8, 0, True)), 'activation': 'relu'},
              cross_params={'num_cross_layer': 3},
              autoint_params={'num_attention': 2, 'num_heads': 2, 'dropout_rate': 0, 'use_residual': True},
              cin_params={'cross_layer_size': (8, 6), 'activation': 'relu', 'use_residual': False,
                          'use_bias': False, 'direct': False, 'reduce_D': False})
    kw.update(cfg_kw)
    conf = deeptable.ModelConfig(**kw)
    cats = [CategoricalColumn(f'c{i}', v, dim) for i, v in enumerate(vocab)]
    conts = [ContinuousColumn('input_continuous_all', [f'n{i}' for i in range(n_cont)])] if n_cont else []
    model = DeepModel('binary', 2, conf, cats, conts, seed=seed)
    model._build_model()
    return model, conf


def batch(vocab, n_cont, b, seed=0):
    g = np.random.default_rng(seed)
    idx = np.stack([g.integers(0, v, size=b) for v in vocab], axis=1).astype(np.int32) if vocab else None
    cont = g.normal(size=(b, n_cont)).astype(np.float32) if n_cont else None
    y = (g.random(b) < 0.35).astype(np.float32)
    return idx, cont, y


NET_SETS = [['linear'], ['fm_nets'], ['dnn_nets'], ['cin_nets'], ['cross_nets'], ['dcn_nets'], ['cross_dnn_nets'],
            ['linear', 'fm_nets', 'dnn_nets'], ['linear', 'cin_nets', 'dnn_nets'], ['autoint_nets'], ['pnn_nets'],
            ['ipnn_nets'], ['opnn_nets'], ['fm_nets', 'cin_nets', 'cross_nets', 'autoint_nets', 'pnn_nets'],
            ['afm_nets'], ['linear', 'afm_nets', 'dnn_nets'], ['fibi_dnn_nets'], ['fm_nets', 'fibi_nets'],
            ['fgcnn_dnn_nets'], ['linear', 'fgcnn_fm_nets'], ['fgcnn_cin_nets'], ['fgcnn_afm_nets'], ['fgcnn_ipnn_nets'], ['fg_nets']]


@pytest.mark.parametrize('nets', NET_SETS)
def test_forward_and_training_match_oracle(nets):
    _forward_and_training_match_oracle(nets)


@pytest.mark.parametrize('fibinet_params', [
    {'senet_pooling_op': 'max', 'senet_reduction_ratio': 2, 'bilinear_type': 'field_all'},
    {'senet_pooling_op': 'mean', 'senet_reduction_ratio': 3, 'bilinear_type': 'field_each'}])
def test_fibinet_pooling_and_weight_sharing_variants(fibinet_params):
    _forward_and_training_match_oracle(['fibi_dnn_nets'], fibinet_params=fibinet_params)


def test_fgcnn_small_kernels_and_uneven_pooling():
    _forward_and_training_match_oracle(['fgcnn_dnn_nets'], fgcnn_params={'fg_filters': (3, 4), 'fg_heights': (3, 2),
                                                                         'fg_pool_heights': (2, 3), 'fg_new_feat_filters': (2, 1)})


def test_binary_focal_loss_training_matches_oracle():
    from deeptables_b200 import layers
    _forward_and_training_match_oracle(['linear', 'dnn_nets'], loss=layers.BinaryFocalLoss(gamma=2.0, alpha=0.25))


def test_afm_hidden_factor_and_linear_attention():
    _forward_and_training_match_oracle(['afm_nets', 'dnn_nets'], afm_params={'hidden_factor': 5, 'activation': 'linear', 'dropout_rate': 0})


def _forward_and_training_match_oracle(nets, **cfg_kw):
    vocab, dim, n_cont, b = [11, 7, 13, 5, 9], 4, 3, 48
    model, conf = build(nets, vocab, dim, n_cont, **cfg_kw)
    idx, cont, y = batch(vocab, n_cont, b)
    state = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    ref = M.RefTrainer(state, conf, len(vocab))
    t_idx, t_cont, t_y = torch.tensor(idx), torch.tensor(cont), torch.tensor(y)
    got = model.predict_step(t_idx.cuda(), t_cont.cuda()).cpu()
    want = ref.predict(t_idx, t_cont)
    torch.testing.assert_close(got, want, rtol=1e-3, atol=1e-5)       # north_star: 1e-3 relative fp32
    losses_g, losses_r = [], []
    for step in range(8):
        idx, cont, y = batch(vocab, n_cont, b, seed=step)
        losses_g.append(model.train_on_batch(idx, cont, y))
        losses_r.append(ref.train_step(torch.tensor(idx), torch.tensor(cont), torch.tensor(y)))
    np.testing.assert_allclose(losses_g, losses_r, rtol=2e-3, atol=1e-5)
    new_state = model.state_dict()
    for k, v in ref.state.items():
        # Adam normalises the step size, so weights drift by O(lr) per step regardless of gradient
        # scale: compare against lr * steps
        np.testing.assert_allclose(new_state[k].cpu().numpy(), v.numpy(), rtol=1e-2, atol=2e-4, err_msg=k)
    got = model.predict_step(t_idx.cuda(), t_cont.cuda()).cpu()
    torch.testing.assert_close(got, ref.predict(t_idx, t_cont), rtol=5e-3, atol=1e-4)


def test_only_categorical_only_continuous_single_column():
    # reference edge cases: model_input_test.py, nets_test.py:166-189
    for vocab, n_cont, nets in (([9, 4], 0, ['linear', 'fm_nets', 'dnn_nets', 'cin_nets']),
                                ([], 4, ['linear', 'dnn_nets', 'cross_nets', 'fm_nets']),
                                ([12], 2, ['linear', 'fm_nets', 'cin_nets', 'dnn_nets'])):
        model, conf = build(nets, vocab, 4, n_cont)
        idx, cont, y = batch(vocab, n_cont, 20)
        state = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        want, _ = M.forward(state, conf, torch.tensor(idx) if idx is not None else None,
                            torch.tensor(cont) if cont is not None else None, len(vocab), False)
        got = model.predict_step(torch.tensor(idx).cuda() if idx is not None else None,
                                 torch.tensor(cont).cuda() if cont is not None else None).cpu()
        torch.testing.assert_close(got, want, rtol=1e-3, atol=1e-5)
        l0 = model.train_on_batch(idx, cont, y)
        assert np.isfinite(l0)


def test_concat_stacking_order_and_custom_net():
    from deeptables_b200 import deepnets, layers

    def my_net(embeddings, flatten_emb_layer, dense_layer, concat_emb_dense, config, model_desc):
        x = layers.Dense(5, activation='relu', name='my_dense')(concat_emb_dense)
        return layers.FM(name='my_fm')(embeddings)   # custom net may call the fused layers too

    assert deepnets.register_nets(my_net) == 'my_net'
    with pytest.raises(ValueError):
        deepnets.register_nets(lambda a, b: None)
    model, conf = build(['linear', 'dnn_nets', my_net], [6, 7, 8], 4, 2, stacking_op='concat')
    assert conf.nets == ['linear', 'dnn_nets', 'my_net']
    assert model.state_dict()['task_output/kernel'].shape == (3, 1)
    idx, cont, y = batch([6, 7, 8], 2, 16)
    assert np.isfinite(model.train_on_batch(idx, cont, y))


def test_out_of_range_id_raises():
    model, conf = build(['linear', 'dnn_nets'], [5, 6], 4, 0)
    idx = np.array([[1, 2], [5, 0]], dtype=np.int32)
    model.predict_step(torch.tensor(idx).cuda(), None)
    with pytest.raises(IndexError):
        model.table.check_status()


def test_deeptable_fit_predict_evaluate_save_load(tmp_path):
    # config (1) stand-in: bank-like schema (hypernets' dsutils.load_bank is absent), DeepFM, bs 512 eval
    from deeptables_b200 import deeptable, deepnets
    g = np.random.default_rng(0)
    n = 3000
    df = pd.DataFrame({
        'job': g.choice(list('abcdefghijkl'), size=n), 'marital': g.choice(['m', 's', 'd'], size=n),
        'education': g.choice(['p', 's', 't', 'u'], size=n), 'default': g.choice(['yes', 'no'], size=n),
        'housing': g.choice(['yes', 'no'], size=n), 'contact': g.choice(['c', 't', 'u'], size=n),
        'age': g.integers(18, 90, size=n).astype(float), 'balance': g.normal(1000, 500, size=n),
        'duration': g.exponential(200, size=n), 'campaign': g.integers(1, 10, size=n).astype(float),
    })
    logit = (df['housing'] == 'yes') * 1.5 + (df['duration'] - 200) / 150 + (df['job'] == 'a') * 2 - 1
    y = np.where(g.random(n) < 1 / (1 + np.exp(-logit)), 'yes', 'no')
    conf = deeptable.ModelConfig(nets=deepnets.DeepFM, embedding_dropout=0, metrics=['AUC'], auto_scale=True,
                                 earlystopping_patience=5)
    dt = deeptable.DeepTable(config=conf)
    model, history = dt.fit(df, y, batch_size=128, epochs=6, verbose=0)
    assert 'val_auc' in history.history and len(history.history['loss']) >= 1
    result = dt.evaluate(df, y, batch_size=512, verbose=0)
    assert result["AUC"] > 0.62
    proba = dt.predict_proba(df.head(100))
    assert proba.shape == (100, 2) and np.allclose(proba.sum(axis=1), 1.0, atol=1e-6)
    preds = dt.predict(df.head(100))
    assert set(preds) <= {'yes', 'no'}
    assert (dt.proba2predict(proba) == preds).all()
    # unseen categories are mapped to the reserved slot (reference deeptable_test.py:178-201)
    odd = df.head(5).copy()
    odd['job'] = 'never-seen'
    assert dt.predict_proba(odd).shape == (5, 2)
    # apply(): intermediate layer outputs by name (reference deeptable_test.py:79-86)
    outs = dt.apply(df.head(7), output_layers=['flatten_embeddings', 'dnn_dense_1', 'dnn_dense_2'])
    assert [o.shape for o in outs] == [(7, 6 * 4), (7, 128), (7, 64)]
    dt.save(str(tmp_path / 'm'))
    dt2 = deeptable.DeepTable.load(str(tmp_path / 'm'))
    np.testing.assert_allclose(dt2.predict_proba(df.head(100)), proba, rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        bad = df.copy()
        bad.columns = list(bad.columns[:-1]) + [bad.columns[0]]
        deeptable.DeepTable(config=conf).fit(bad, y, epochs=1, verbose=0)


def test_fit_cross_validation_oof_and_model_set(tmp_path):
    """reference deeptable.py:373-517 / tests/models/deeptable_cv_test.py: out-of-fold matrix complete, fold means for
    X_eval / X_test, one registered model per fold, model_selector='all' = mean of the folds, oof_metrics scores."""
    from deeptables_b200 import deeptable, deepnets
    g = np.random.default_rng(3)
    n = 1200
    df = pd.DataFrame({'a': g.choice(list('abcdef'), size=n), 'b': g.choice(['x', 'y', 'z'], size=n),
                       'u': g.normal(size=n), 'v': g.exponential(size=n)})
    logit = (df['a'] == 'a') * 2.0 + df['u'] * 1.5 - 0.5
    y = np.where(g.random(n) < 1 / (1 + np.exp(-logit)), 1, 0)
    conf = deeptable.ModelConfig(nets=deepnets.DeepFM, embedding_dropout=0, metrics=['AUC'], auto_scale=True,
                                 earlystopping_patience=3, home_dir=str(tmp_path / 'out'))
    dt = deeptable.DeepTable(config=conf)
    oof, ev, te, scores = dt.fit_cross_validation(df, y, X_eval=df.head(50), X_test=df.tail(40), num_folds=3, stratified=True,
                                                  batch_size=64, epochs=10, verbose=0, oof_metrics=['auc', 'accuracy'])
    assert oof.shape == (n, 2) and not np.isnan(oof).any() and np.allclose(oof.sum(1), 1.0, atol=1e-6)
    assert ev.shape == (50, 2) and te.shape == (40, 2)
    assert len(scores) == 3 and all(0.5 < sc['auc'] <= 1.0 for sc in scores)
    from sklearn.metrics import roc_auc_score
    assert roc_auc_score(y, oof[:, 1]) > 0.65
    models = dt.get_model('all')
    assert len(models) == 3
    p_all = dt.predict_proba(df.head(30), model_selector='all')
    p_each = [dt.predict_proba(df.head(30), model_selector=f'{"+".join(conf.nets)}-kfold-{k + 1}') for k in range(3)]
    np.testing.assert_allclose(p_all, sum(p_each) / 3, rtol=1e-5, atol=1e-6)
    assert len([f for f in os.listdir(dt.output_path) if f.endswith('.npz')]) == 3


def test_fit_with_rows_in_pinned_host_memory(monkeypatch):
    """DTB_DATA_ON_HOST=1: the encoded rows stay in pinned host memory and reach the GPU through the double-buffered
    loader (the replacement of utils/dataset_generator.py for data sets beyond HBM); same steps arithmetic, it learns."""
    model, conf = build(['linear', 'fm_nets', 'dnn_nets'], [7, 5, 9], 4, 2)
    g = np.random.default_rng(11)
    n = 2000
    df = pd.DataFrame({'c0': g.integers(0, 7, n), 'c1': g.integers(0, 5, n), 'c2': g.integers(0, 9, n),
                       'n0': g.normal(size=n), 'n1': g.normal(size=n)})
    y = ((df['c0'] == 3) * 2.0 + df['n0'] > 0.5).astype(np.float32).values
    monkeypatch.setenv('DTB_DATA_ON_HOST', '1')
    hist = model.fit(df, y, batch_size=128, epochs=40, verbose=0, validation_split=0.2,
                     sample_weight=np.ones(n, dtype=np.float32))
    assert len(hist.history['loss']) == 40 and hist.history['loss'][-1] < 0.6 * hist.history['loss'][0]
    assert hist.history['val_AUC'][-1] > 0.9            # rows, labels and weights stay paired through the loader


def test_fit_steps_arithmetic_and_history_keys():
    model, conf = build(['dnn_nets'], [5, 6], 4, 2)
    g = np.random.default_rng(1)
    n = 100
    df = pd.DataFrame({'c0': g.integers(0, 5, n), 'c1': g.integers(0, 6, n), 'n0': g.normal(size=n),
                       'n1': g.normal(size=n)})
    y = (g.random(n) < 0.5).astype(np.float32)
    hist = model.fit(df, y, batch_size=32, epochs=2, verbose=0, validation_split=0.2)
    assert set(hist.history.keys()) >= {'loss', 'auc', 'val_loss', 'val_auc'}
    assert hist.history['AUC'] == hist.history['auc']           # IgnoreCaseDict
    assert model._step == 2 * (80 // 32)                          # steps_per_epoch = len(X)//bs
    ev = model.evaluate(df, y, batch_size=64)
    assert 'loss' in ev and 'auc' in ev


def test_embedding_and_dense_dropout_train_and_are_off_at_inference():
    # reference defaults: embedding_dropout=0.3 (config.py:84); dropout must only act in training
    vocab, n_cont, b = [30, 20, 10, 25], 3, 512
    model, conf = build(['linear', 'fm_nets', 'cin_nets', 'dnn_nets'], vocab, 4, n_cont, embedding_dropout=0.3,
                        dense_dropout=0.2)
    idx, cont, y = batch(vocab, n_cont, b)
    ti, tc = torch.tensor(idx).cuda(), torch.tensor(cont).cuda()
    p1, p2 = model.predict_step(ti, tc), model.predict_step(ti, tc)
    assert torch.equal(p1, p2)                                   # inference: deterministic, no mask
    losses = [model.train_on_batch(idx, cont, y) for _ in range(40)]
    assert np.isfinite(losses).all() and np.mean(losses[-5:]) < np.mean(losses[:5])
    # the mask really drops ~30 %: compare a training-mode forward against inference on fresh weights
    from deeptables_b200 import engine as E
    x = torch.ones(1 << 16, device='cuda')
    yk = E.DropoutFn.apply(x, 0.3, 123)
    kept = float((yk > 0).float().mean())
    assert abs(kept - 0.7) < 0.01 and abs(float(yk.max()) - 1 / 0.7) < 1e-5
    assert torch.equal(yk, E.DropoutFn.apply(x, 0.3, 123)) and not torch.equal(yk, E.DropoutFn.apply(x, 0.3, 124))


def test_lazy_and_dense_table_optimizer_switch_is_transparent():
    """The table optimiser runs either as the row-wise exact-lazy Adam or as the dense sweep, chosen by
    cost; both are the same Keras Adam, so forcing either form, or flipping between them mid-training,
    must not change the trajectory."""
    vocab, n_cont, b = [400, 300, 500], 2, 32
    results = {}
    for name, schedule in (('lazy', ['lazy'] * 8), ('dense', ['dense'] * 8),
                           ('mixed', ['lazy', 'lazy', 'dense', 'dense', 'lazy', 'dense', 'lazy', 'lazy'])):
        model, conf = build(['linear', 'fm_nets', 'dnn_nets'], vocab, 4, n_cont, seed=9)
        modes = []
        for step, mode in enumerate(schedule):
            model._table_mode_override = mode
            model.train_on_batch(*batch(vocab, n_cont, b, seed=step))
            modes.append(model.table.lazy_active)
        assert modes == [m == 'lazy' for m in schedule]
        results[name] = {k: v.clone() for k, v in model.state_dict().items()}
    for k in results['lazy']:
        for other in ('dense', 'mixed'):
            torch.testing.assert_close(results['lazy'][k], results[other][k], rtol=1e-5, atol=1e-7, msg=f'{other}:{k}')


def test_table_gradient_is_final_before_cin_weight_gradient_kernels():
    """Data-parallel overlap hook: the table gradient must be declared final exactly once per step, after
    every embedding-gradient producer ran and BEFORE the CIN weight-gradient launches (so the row exchange
    can run under them)."""
    from deeptables_b200 import _native as nat
    vocab, n_cont, b = [40, 30, 20, 50], 3, 64
    model, conf = build(['linear', 'fm_nets', 'cin_nets', 'dnn_nets', 'pnn_nets'], vocab, 8, n_cont,
                        cin_params={'cross_layer_size': (16, 16), 'activation': 'relu', 'use_residual': False,
                                    'use_bias': False, 'direct': False, 'reduce_D': False})
    idx, cont, y = batch(vocab, n_cont, b)
    model.train_on_batch(idx, cont, y)                      # allocate training state
    t = model.table
    ti, tc, ty = torch.tensor(idx).cuda(), torch.tensor(cont).cuda(), torch.tensor(y).cuda().view(-1, 1)
    seen = []
    t.pending_bwd = 0
    snapshot = {}

    def hook():
        seen.append(nat.lib.dtb_launch_count())
        snapshot['grad'] = t.grad.clone()

    z = model._forward(ti, tc, training=True)
    t.on_grad_final = hook
    n_consumers = t.pending_bwd
    assert n_consumers == 5                                  # concat, linear, fm, cin, pnn
    from deeptables_b200 import engine as E
    prob, dz = E.loss_forward_backward(z, ty, 'binary', None, True, None)
    z.backward(dz)
    end = nat.lib.dtb_launch_count()
    t.on_grad_final = None
    assert len(seen) == 1 and t.pending_bwd == 0
    assert end - seen[0] >= 2                                # the CIN wgrad launches came after the hook
    assert torch.equal(snapshot['grad'], t.grad)             # nothing touched the table gradient afterwards
    t.grad.zero_()
    model._scope.flat_g.zero_()
