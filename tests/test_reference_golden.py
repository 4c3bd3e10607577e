"""The CPU oracle against golden vectors produced by the REFERENCE'S OWN code.

tests/golden/reference_layers.npz / reference_models.npz were generated in the build container by
tests/golden/make_reference_golden.py: the reference's layer classes, net builders and
DeepModel.__build_model are imported unmodified from the reference checkout and executed eagerly in float64
over tests/golden/tf_shim.py (a stand-in for the TensorFlow/Keras primitives they call; TensorFlow itself is not
installable here).  Nothing in this file reads the reference checkout.  Tolerance: float64 round-off.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import layers_ref as L
from oracle import model_ref as M

HERE = os.path.dirname(os.path.abspath(__file__))
TOL = dict(rtol=1e-9, atol=1e-11)


def _load(name):
    z = np.load(os.path.join(HERE, 'golden', name))
    manifest = json.loads(str(z['__manifest__']))
    return z, manifest


def _t(a):
    return torch.tensor(np.asarray(a), dtype=torch.float64)


def _weights(z, case):
    pre = f'{case}/w/'
    return {k[len(pre):]: _t(z[k]) for k in z.files if k.startswith(pre)}


LAYERS_Z, LAYER_CASES = _load('reference_layers.npz')
MODELS_Z, MODEL_CASES = _load('reference_models.npz')


def test_fixture_inventory():
    kinds = {m['kind'] for m in LAYER_CASES}
    assert kinds == {'fm', 'cin', 'cross', 'mha', 'inner', 'outer', 'embedding', 'dnn'}
    assert len(MODEL_CASES) >= 14


@pytest.mark.parametrize('meta', LAYER_CASES, ids=[m['case'] for m in LAYER_CASES])
def test_oracle_layer_reproduces_reference_layer(meta):
    z, case, p = LAYERS_Z, meta['case'], meta['params']
    want = z[f'{case}/out'] if f'{case}/out' in z.files else None
    kind = meta['kind']
    if kind == 'fm':
        got = L.fm(_t(z[f'{case}/x']))
    elif kind == 'cin':
        params = dict(p, cross_layer_size=tuple(p['cross_layer_size']))
        got = L.cin(_t(z[f'{case}/x']), params, _weights(z, case))
    elif kind == 'cross':
        w = _weights(z, case)
        n = p['num_cross_layer']
        got = L.cross(_t(z[f'{case}/x']), [w[f'kernels_{i}'] for i in range(n)], [w[f'bias_{i}'] for i in range(n)])
    elif kind == 'mha':
        w = _weights(z, case)
        bn = {'moving_mean': w['batch_normalize/moving_mean'], 'moving_variance': w['batch_normalize/moving_variance']}
        if not p['use_residual']:
            assert 'dense_residual/kernel' not in w          # the reference never builds the unused projection
            w['dense_residual/kernel'] = w['dense_residual/bias'] = None
        got, _ = L.multihead_attention(_t(z[f'{case}/x']), p, w, bn, p['training'])
    elif kind in ('inner', 'outer'):
        embs = [_t(z[f'{case}/e{i}']) for i in range(4)]
        got = L.inner_product(embs) if kind == 'inner' else \
            L.outer_product(embs, _t(z[f'{case}/kernel']), p['outer_product_kernel_type'])
    elif kind == 'embedding':
        tables = [_t(z[f'{case}/table{i}']) for i in range(len(p['vocab']))]
        outs = L.embedding_lookup(tables, torch.tensor(z[f'{case}/ids']))
        assert len(outs) == len(tables)
        for i, o in enumerate(outs):
            np.testing.assert_allclose(o.numpy(), z[f'{case}/out{i}'], **TOL)
        return
    elif kind == 'dnn':
        w = _weights(z, case)
        params = {'hidden_units': tuple(tuple(h) for h in p['hidden_units']), 'activation': p['activation']}
        got, _ = L.dnn(_t(z[f'{case}/x']), params, w, w, p['training'], cellname='dnn')
    else:
        raise AssertionError(kind)
    assert tuple(got.shape) == tuple(want.shape)
    np.testing.assert_allclose(got.numpy(), want, **TOL)


F3_Z, F3_CASES = _load('reference_f3.npz')
F3_LAYER_CASES = [m for m in F3_CASES if m['kind'] != 'model']
F3_MODEL_CASES = [m for m in F3_CASES if m['kind'] == 'model']


@pytest.mark.parametrize('meta', F3_LAYER_CASES, ids=[m['case'] for m in F3_LAYER_CASES])
def test_oracle_f3_layer_reproduces_reference_layer(meta):
    """AFM, SENET, BilinearInteraction against the reference's own classes (layers.py:742-812, 245-382)."""
    z, case, p, kind = F3_Z, meta['case'], meta['params'], meta['kind']
    want = z[f'{case}/out']
    if kind == 'afm':
        embs = [_t(z[f'{case}/e{i}']) for i in range(p['n_fields'])]
        got = L.afm(embs, _t(z[f'{case}/att_kernel']), _t(z[f'{case}/att_bias']), _t(z[f'{case}/projection_h']),
                    _t(z[f'{case}/out_kernel']), p.get('activation', 'relu'))
        pooled = L.afm_pooled(embs, _t(z[f'{case}/att_kernel']), _t(z[f'{case}/att_bias']), _t(z[f'{case}/projection_h']),
                              p.get('activation', 'relu'))
        np.testing.assert_allclose((pooled @ _t(z[f'{case}/out_kernel'])).numpy(), want, **TOL)
    elif kind == 'senet':
        w = _weights(z, case)
        got = L.senet(_t(z[f'{case}/x']), w['dense_att1/kernel'], w['dense_att1/bias'], w['dense_att2/kernel'],
                      w['dense_att2/bias'], p['pooling_op'])
    elif kind == 'focal_binary':
        got = L.binary_focal_loss(_t(z[f'{case}/y_true']), _t(z[f'{case}/y_pred']), p['gamma'], p['alpha'])
    elif kind == 'focal_categorical':
        got = L.categorical_focal_loss(_t(z[f'{case}/y_true']), _t(z[f'{case}/y_pred']), p['gamma'], p['alpha'])
    elif kind == 'fgcnn':
        pooled, got = L.fgcnn(_t(z[f'{case}/x']), _t(z[f'{case}/conv_kernel']), _t(z[f'{case}/conv_bias']),
                              _t(z[f'{case}/dense_kernel']), _t(z[f'{case}/dense_bias']), p['pool_height'], p['new_filters'])
        np.testing.assert_allclose(pooled.numpy(), z[f'{case}/pooled'], **TOL)
    elif kind == 'bilinear':
        n = len([k for k in z.files if k.startswith(f'{case}/w')])
        got = L.bilinear_interaction(_t(z[f'{case}/x']), [_t(z[f'{case}/w{i}']) for i in range(n)], p['bilinear_type'])
    else:
        raise AssertionError(kind)
    assert tuple(got.shape) == tuple(want.shape)
    np.testing.assert_allclose(got.numpy(), want, **TOL)


def _model_config(p):
    from deeptables_b200 import deeptable
    kw = dict(p['config'])
    for key in ('dnn_params', 'cin_params'):
        if key in kw:
            kw[key] = dict(kw[key])
    if 'dnn_params' in kw:
        kw['dnn_params']['hidden_units'] = tuple(tuple(h) for h in kw['dnn_params']['hidden_units'])
    if 'cin_params' in kw:
        kw['cin_params']['cross_layer_size'] = tuple(kw['cin_params']['cross_layer_size'])
    if 'fgcnn_params' in kw:
        kw['fgcnn_params'] = {k: tuple(v) for k, v in kw['fgcnn_params'].items()}
    return deeptable.ModelConfig(embedding_dropout=0, dense_dropout=0, embeddings_output_dim=p['dim'], **kw)


ALL_MODEL_CASES = [(MODELS_Z, m) for m in MODEL_CASES] + [(F3_Z, m) for m in F3_MODEL_CASES]


@pytest.mark.parametrize('zm', ALL_MODEL_CASES, ids=[m['case'] for _, m in ALL_MODEL_CASES])
def test_oracle_model_reproduces_reference_build_model(zm):
    """Same weights (by the reference's layer/weight names), same batch -> same task_output as the graph that
    DeepModel.__build_model (deepmodel.py:259-317) assembled, in inference and in training mode (batch-statistics
    BatchNormalization)."""
    z, meta = zm
    case, p = meta['case'], meta['params']
    conf = _model_config(p)
    state = _weights(z, case)
    spec, _ = M.param_spec(conf, p['vocab'], [p['dim']] * len(p['vocab']), p['n_cont'], p['task'], p['num_classes'] or 2)
    want_names = {name for name, _, _ in spec}
    for name in want_names:                                    # BN moving statistics ride along in the same dict
        assert name in state, f'oracle expects {name}, the reference model has {sorted(state)}'
    extra = {n for n in state if n not in want_names and not n.endswith(('moving_mean', 'moving_variance'))}
    assert not extra, f'reference weights the oracle does not know: {sorted(extra)}'
    for name, shape, _ in spec:
        assert tuple(state[name].shape) == tuple(shape), name
    ids = torch.tensor(z[f'{case}/ids']) if p['vocab'] else None
    cont = _t(z[f'{case}/cont']) if p['n_cont'] else None
    for training, key in ((False, 'out_infer'), (True, 'out_train')):
        got, _ = M.forward(state, conf, ids, cont, len(p['vocab']), training, task=p['task'])
        np.testing.assert_allclose(got.numpy(), z[f'{case}/{key}'], **TOL, err_msg=f'{case} training={training}')


def test_modelconfig_mirror_matches_reference_defaults():
    """Field order, defaults, presets and the nets plug-in signature of the reference (config.py:8-151,
    deepnets.py:12-20,43), dumped from the reference's own ModelConfig()."""
    import inspect
    from deeptables_b200 import deeptable, deepnets
    with open(os.path.join(HERE, 'golden', 'reference_modelconfig.json')) as f:
        ref = json.load(f)
    mine = deeptable.ModelConfig()
    assert list(mine._fields) == ref['fields']
    norm = json.loads(json.dumps({k: v for k, v in mine._asdict().items() if k != 'home_dir'}, default=list))
    for k, v in ref['defaults'].items():
        assert norm[k] == v, f'ModelConfig default {k}: {norm[k]!r} != reference {v!r}'
    for name, nets in ref['presets'].items():
        assert getattr(deepnets, name) == nets
    assert list(inspect.signature(deepnets.linear).parameters) == ref['net_signature']


# ---------------------------------------------------------------------------------------------------------------
# GPU: the CUDA engine against the same reference-code vectors (through DeepModel -> C ABI)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('zm', ALL_MODEL_CASES, ids=[m['case'] for _, m in ALL_MODEL_CASES])
def test_cuda_model_reproduces_reference_build_model(zm):
    """Load the reference model's weights by name into the CUDA DeepModel and compare task_output with the
    reference's own graph: inference (moving statistics) and, for binary/regression tasks, the training-mode
    forward (batch statistics) that train_step returns.  fp32 kernels, bf16x3 CIN: north_star tolerance 1e-3."""
    from deeptables_b200.deepmodel import DeepModel
    from deeptables_b200.metainfo import CategoricalColumn, ContinuousColumn
    z, meta = zm
    case, p = meta['case'], meta['params']
    conf = _model_config(p)
    cats = [CategoricalColumn(f'c{i}', v, p['dim']) for i, v in enumerate(p['vocab'])]
    conts = [ContinuousColumn('input_continuous_all', [f'n{i}' for i in range(p['n_cont'])])] if p['n_cont'] else []
    model = DeepModel(p['task'], p['num_classes'], conf, cats, conts, seed=3)
    model._build_model()
    model.load_state_dict({k: v.to(torch.float32) for k, v in _weights(z, case).items()}, strict=True)
    ids = torch.tensor(z[f'{case}/ids'].astype(np.int32)).cuda() if p['vocab'] else None
    cont = torch.tensor(z[f'{case}/cont'].astype(np.float32)).cuda() if p['n_cont'] else None
    got = model.predict_step(ids, cont).cpu().double().numpy()
    np.testing.assert_allclose(got, z[f'{case}/out_infer'], rtol=1e-3, atol=1e-5)
    if p['task'] in ('binary', 'regression'):
        y = torch.zeros(z[f'{case}/out_train'].shape[0], 1, device='cuda')
        got = model.train_step(ids, cont, y).detach().cpu().double().numpy()
        np.testing.assert_allclose(got, z[f'{case}/out_train'], rtol=1e-3, atol=1e-5)
