"""Generate golden vectors by EXECUTING THE REFERENCE'S OWN hot-path code (run HERE; /root/reference is read-only
and does not exist on the GPU box, so the output is committed):

    python tests/golden/make_reference_golden.py        ->  tests/golden/reference_layers.npz
                                                            tests/golden/reference_models.npz
                                                            tests/golden/reference_modelconfig.json

The reference's layer classes (deeptables/models/layers.py), net builders (deepnets.py) and model assembly
(DeepModel.__build_model, deepmodel.py:259-317) are imported unmodified from /root/reference and run eagerly in
float64 on top of tests/golden/tf_shim.py, a stand-in for the handful of TensorFlow/Keras primitives they call
(TensorFlow itself cannot be installed in this image).  tests/test_reference_golden.py then requires the CPU
oracle (oracle/) to reproduce every vector: this pins the oracle's restatement of the reference's layer logic
(index order, reshapes/transposes, head splitting, pair order, half-split, stacking, weight names) against the
reference's code itself.  What it cannot pin is TensorFlow's own arithmetic (the shim restates those
primitives), the losses and the optimiser, which run inside Keras.
"""
import importlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_shim  # noqa: E402

REFERENCE_ROOT = '/root/reference'


def np64(t):
    return t.detach().cpu().numpy().astype(np.float64)


class Recorder:
    def __init__(self):
        self.arrays = {}
        self.manifest = []

    def add(self, case, kind, params, **arrays):
        for k, v in arrays.items():
            self.arrays[f'{case}/{k}'] = v
        self.manifest.append({'case': case, 'kind': kind, 'params': params, 'arrays': sorted(arrays)})

    def save(self, path):
        np.savez_compressed(path, __manifest__=np.array(json.dumps(self.manifest)), **self.arrays)


def rand(*shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float64)


def layer_weights(layer, prefix=''):
    return {f'{prefix}{k}': np64(v) for k, v in layer.weights_by_name.items()}


# ------------------------------------------------------------------------------------------------------------
# A. layers called directly
# ------------------------------------------------------------------------------------------------------------
def make_layer_cases(layers, deepnets, rec):
    # FM (layers.py:53-62)
    x = rand(5, 4, 3, seed=1)
    rec.add('fm', 'fm', {}, x=np64(x), out=np64(layers.FM()(x)))

    # CIN (layers.py:638-734): every switch
    cin_cases = {
        'cin_half': dict(cross_layer_size=(6, 4), activation='relu', use_residual=False, use_bias=False, direct=False, reduce_D=False),
        'cin_direct_bias_linear': dict(cross_layer_size=(5, 3), activation='linear', use_residual=False, use_bias=True, direct=True, reduce_D=False),
        'cin_residual': dict(cross_layer_size=(4, 4), activation='relu', use_residual=True, use_bias=True, direct=False, reduce_D=False),
        'cin_reduce_d': dict(cross_layer_size=(6, 2), activation='relu', use_residual=False, use_bias=False, direct=False, reduce_D=True),
        'cin_odd_last': dict(cross_layer_size=(6, 5), activation='relu', use_residual=False, use_bias=False, direct=False, reduce_D=False),
        'cin_three': dict(cross_layer_size=(8, 8, 4), activation='relu', use_residual=False, use_bias=False, direct=False, reduce_D=False),
    }
    for i, (case, params) in enumerate(cin_cases.items()):
        x = rand(6, 5, 4, seed=10 + i)
        lyr = layers.CIN(params=params)
        out = lyr(x)
        w = layer_weights(lyr)
        w.update(layer_weights(lyr.exFM_out, 'exFM_out/'))
        if params['use_residual']:
            w.update(layer_weights(lyr.exFM_out0, 'exFM_out0/'))
        rec.add(case, 'cin', {**params, 'cross_layer_size': list(params['cross_layer_size'])}, x=np64(x), out=np64(out),
                **{f'w/{k}': v for k, v in w.items()})

    # Cross (layers.py:417-436)
    x = rand(6, 7, seed=20)
    lyr = layers.Cross(params={'num_cross_layer': 3})
    out = lyr(x)
    rec.add('cross3', 'cross', {'num_cross_layer': 3}, x=np64(x), out=np64(out),
            **{f'w/{k}': v for k, v in layer_weights(lyr).items()})

    # MultiheadAttention (layers.py:104-153), inference and training-mode BatchNormalization
    for case, params, training in (('mha_1head', dict(num_heads=1, dropout_rate=0, use_residual=True), False),
                                   ('mha_2head_nores', dict(num_heads=2, dropout_rate=0, use_residual=False), False),
                                   ('mha_2head_train', dict(num_heads=2, dropout_rate=0, use_residual=True), True)):
        x = rand(7, 5, 4, seed=30)
        tf_shim.set_training(training)
        lyr = layers.MultiheadAttention(params=params)
        out = lyr(x)
        tf_shim.set_training(False)
        w = {}
        for sub in ('dense_Q', 'dense_K', 'dense_V', 'dense_residual', 'batch_normalize'):
            w.update(layer_weights(getattr(lyr, sub), f'{sub}/'))
        rec.add(case, 'mha', {**params, 'training': training}, x=np64(x), out=np64(out), **{f'w/{k}': v for k, v in w.items()})

    # InnerProduct / OuterProduct (layers.py:473-487, 531-581): list of F tensors (B,1,D)
    embs = [rand(5, 1, 3, seed=40 + i) for i in range(4)]
    rec.add('inner', 'inner', {}, out=np64(layers.InnerProduct()(embs)), **{f'e{i}': np64(e) for i, e in enumerate(embs)})
    for kt in ('mat', 'vec', 'num'):
        lyr = layers.OuterProduct(params={'outer_product_kernel_type': kt})
        out = lyr(embs)
        rec.add(f'outer_{kt}', 'outer', {'outer_product_kernel_type': kt}, out=np64(out), kernel=np64(lyr.kernel),
                **{f'e{i}': np64(e) for i, e in enumerate(embs)})

    # MultiColumnEmbedding (layers.py:853-904): float32-encoded ids in, list of (B,1,D) out
    vocab, dims = [5, 7, 3], [4, 4, 4]
    g = np.random.default_rng(50)
    ids = np.stack([g.integers(0, v, size=6) for v in vocab], axis=1)
    lyr = layers.MultiColumnEmbedding(vocab, dims, 0.0, name='emb_categorical_vars_all')
    outs = lyr(torch.tensor(ids.astype(np.float32)))
    rec.add('embedding', 'embedding', {'vocab': vocab, 'dims': dims}, ids=ids.astype(np.int64),
            **{f'table{i}': np64(t) for i, t in enumerate(lyr.embeddings)},
            **{f'out{i}': np64(o) for i, o in enumerate(outs)})

    # deepnets.dnn (deepnets.py:401-427) with and without BatchNormalization, inference and training
    for case, hidden, training in (('dnn_plain', ((6, 0, False), (3, 0, False)), False),
                                   ('dnn_bn', ((6, 0, True), (3, 0, False)), False),
                                   ('dnn_bn_train', ((6, 0, True), (4, 0, True)), True)):
        tf_shim.reset_layers()
        tf_shim.set_training(training)
        x = rand(8, 5, seed=60)
        params = {'hidden_units': hidden, 'activation': 'relu'}
        out = deepnets.dnn(x, params, cellname='dnn')
        tf_shim.set_training(False)
        w = {}
        for lyr in tf_shim.created_layers():
            w.update(layer_weights(lyr, f'{lyr.name}/'))
        rec.add(case, 'dnn', {'hidden_units': [list(h) for h in hidden], 'activation': 'relu', 'training': training},
                x=np64(x), out=np64(out), **{f'w/{k}': v for k, v in w.items()})


# ------------------------------------------------------------------------------------------------------------
# B. whole models through the reference's DeepModel.__build_model
# ------------------------------------------------------------------------------------------------------------
CIN_SMALL = {'cross_layer_size': (8, 8, 4), 'activation': 'relu', 'use_residual': False, 'use_bias': False,
             'direct': False, 'reduce_D': False}

MODEL_CASES = [
    # name, config kwargs, vocab, emb dim, n_cont, task, num_classes
    ('xdeepfm', dict(nets=['linear', 'cin_nets', 'dnn_nets'], cin_params=CIN_SMALL), [7, 5, 9, 4], 4, 3, 'binary', 2),
    ('deepfm', dict(nets=['linear', 'fm_nets', 'dnn_nets']), [7, 5, 9, 4], 4, 3, 'binary', 2),
    ('dcn', dict(nets=['dcn_nets'], cross_params={'num_cross_layer': 3}), [7, 5, 9], 4, 2, 'binary', 2),
    ('cross_only', dict(nets=['cross_nets'], cross_params={'num_cross_layer': 2}), [7, 5, 9], 4, 2, 'regression', None),
    ('cross_dnn', dict(nets=['cross_dnn_nets'], cross_params={'num_cross_layer': 2},
                       dnn_params={'hidden_units': ((8, 0, True), (4, 0, False)), 'activation': 'relu'}), [7, 5, 9], 4, 2, 'binary', 2),
    ('autoint', dict(nets=['autoint_nets'], autoint_params={'num_attention': 2, 'num_heads': 2, 'dropout_rate': 0,
                                                            'use_residual': True}), [7, 5, 9, 4], 4, 0, 'binary', 2),
    ('pnn', dict(nets=['pnn_nets']), [7, 5, 9, 4], 4, 3, 'binary', 2),
    ('ipnn_opnn_vec', dict(nets=['ipnn_nets', 'opnn_nets'], pnn_params={'outer_product_kernel_type': 'vec'}), [7, 5, 9], 4, 1, 'binary', 2),
    ('five_nets_add', dict(nets=['fm_nets', 'cin_nets', 'cross_nets', 'autoint_nets', 'pnn_nets'], cin_params=CIN_SMALL,
                           autoint_params={'num_attention': 1, 'num_heads': 1, 'dropout_rate': 0, 'use_residual': True}),
     [7, 5, 9, 4], 4, 3, 'binary', 2),
    ('five_nets_concat', dict(nets=['fm_nets', 'cin_nets', 'cross_nets', 'autoint_nets', 'pnn_nets'], cin_params=CIN_SMALL,
                              stacking_op='concat', output_use_bias=False,
                              autoint_params={'num_attention': 1, 'num_heads': 1, 'dropout_rate': 0, 'use_residual': True}),
     [7, 5, 9, 4], 4, 3, 'binary', 2),
    ('linear_no_cont', dict(nets=['linear']), [7, 5, 9], 4, 0, 'binary', 2),
    ('dnn_no_cat', dict(nets=['dnn_nets']), [], 4, 5, 'binary', 2),
    ('multiclass', dict(nets=['linear', 'fm_nets', 'dnn_nets']), [7, 5, 9], 4, 2, 'multiclass', 3),
    ('single_column', dict(nets=['linear', 'fm_nets', 'cin_nets', 'dnn_nets'], cin_params={**CIN_SMALL, 'cross_layer_size': (4, 2)}),
     [6], 4, 0, 'binary', 2),
]

CHILD_ATTRS = ('dense_Q', 'dense_K', 'dense_V', 'dense_residual', 'batch_normalize', 'dropout_weights', 'exFM_out',
               'exFM_out0')


def collect_state(created):
    """Flatten the built reference model into '<layer>/<weight>' names (the names Keras would give: explicit
    names from the builders, 'cin' / 'multihead_attention[_n]' for the auto-named layers, attribute names for
    the layers those own)."""
    children = set()
    state = {}
    for lyr in created:
        for attr in CHILD_ATTRS:
            child = getattr(lyr, attr, None)
            if isinstance(child, tf_shim.Layer):
                children.add(id(child))
                for k, v in child.weights_by_name.items():
                    state[f'{lyr.name}/{attr}/{k}'] = np64(v)
        for child in getattr(lyr, 'activation_layers', []) or []:
            children.add(id(child))
        for child in getattr(lyr, 'dropouts', []) or []:
            children.add(id(child))
    for lyr in created:
        if id(lyr) in children:
            continue
        for k, v in lyr.weights_by_name.items():
            state[f'{lyr.name}/{k}'] = np64(v)
    return state


def make_model_cases(deepmodel, config_mod, metainfo, rec):
    for ci, (case, cfg_kwargs, vocab, dim, n_cont, task, num_classes) in enumerate(MODEL_CASES):
        conf = config_mod.ModelConfig(embedding_dropout=0, dense_dropout=0, embeddings_output_dim=dim, **cfg_kwargs)
        cats = [metainfo.CategoricalColumn(f'c{i}', v, dim) for i, v in enumerate(vocab)]
        conts = [metainfo.ContinuousColumn('input_continuous_all', [f'n{i}' for i in range(n_cont)])] if n_cont else []
        b = 9
        g = np.random.default_rng(100 + ci)
        ids = np.stack([g.integers(0, v, size=b) for v in vocab], axis=1) if vocab else np.zeros((b, 0), np.int64)
        cont = g.normal(size=(b, n_cont))
        outs = {}
        for training in (False, True):
            tf_shim.reset_layers()
            tf_shim.seed(1000 + ci)                      # same weights in both passes
            tf_shim.set_training(training)
            if vocab:
                tf_shim.feed('input_categorical_vars_all', torch.tensor(ids.astype(np.float32)))
            if n_cont:
                tf_shim.feed('input_continuous_all', torch.tensor(cont, dtype=torch.float64))
            dm = deepmodel.DeepModel(task, num_classes, conf, cats, conts)
            model = dm._DeepModel__build_model(task=task, num_classes=num_classes, nets=conf.nets, categorical_columns=cats,
                                               continuous_columns=conts, var_len_categorical_columns=None, config=conf)
            outs[training] = np64(model.output)
            state = collect_state(tf_shim.created_layers())
            tf_shim.set_training(False)
        params = {k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg_kwargs.items()}
        # ModelConfig.__new__ passes nets through deepnets.get_nets, whose set() loses the user's order
        # (deepnets.py:486): record the order this run of the reference actually built (it decides the row order of
        # task_output/kernel under stacking_op='concat')
        params['nets'] = list(conf.nets)
        rec.add(case, 'model', {'config': json.loads(json.dumps(params)), 'vocab': vocab, 'dim': dim, 'n_cont': n_cont,
                                'task': task, 'num_classes': num_classes},
                ids=ids.astype(np.int64), cont=cont.astype(np.float64), out_infer=outs[False], out_train=outs[True],
                **{f'w/{k}': v for k, v in state.items()})


# ------------------------------------------------------------------------------------------------------------
# C. the SURVEY 8f-3 layers (AFM, SENET + BilinearInteraction = FiBiNet) and their nets, in a file of their own so that the
#    vectors above stay bit-identical to the ones committed in round 1
# ------------------------------------------------------------------------------------------------------------
F3_CHILD_ATTRS = ('dense_attention', 'dense_out', 'dense_att1', 'dense_att2', 'conv2d', 'dense_output', 'exFM_out', 'exFM_out0')

F3_MODEL_CASES = [
    ('afm', dict(nets=['afm_nets'], afm_params={'hidden_factor': 5, 'dropout_rate': 0}), [7, 5, 9, 4], 4, 3, 'binary', 2),
    ('afm_linear_dnn', dict(nets=['linear', 'afm_nets', 'dnn_nets']), [7, 5, 9], 4, 2, 'binary', 2),
    ('fibi_dnn', dict(nets=['fibi_dnn_nets']), [7, 5, 9, 4], 4, 3, 'binary', 2),
    ('fibi_all_max', dict(nets=['fibi_dnn_nets'], fibinet_params={'senet_pooling_op': 'max', 'senet_reduction_ratio': 2,
                                                                  'bilinear_type': 'field_all'}), [7, 5, 9], 4, 2, 'regression', None),
    ('fgcnn_dnn', dict(nets=['fgcnn_dnn_nets'], fgcnn_params={'fg_filters': (3, 4), 'fg_heights': (3, 2), 'fg_pool_heights': (2, 2), 'fg_new_feat_filters': (2, 1)}), [7, 5, 9, 4, 6], 4, 3, 'binary', 2),
    ('fgcnn_dnn_no_cont', dict(nets=['fgcnn_dnn_nets'], fgcnn_params={'fg_filters': (3, 4), 'fg_heights': (3, 2), 'fg_pool_heights': (2, 2), 'fg_new_feat_filters': (2, 1)}), [7, 5, 9], 4, 0, 'regression', None),
    ('fgcnn_fm_plus_linear', dict(nets=['linear', 'fgcnn_fm_nets'], fgcnn_params={'fg_filters': (3, 4), 'fg_heights': (3, 2), 'fg_pool_heights': (2, 2), 'fg_new_feat_filters': (2, 1)}), [7, 5, 9, 4], 4, 2, 'binary', 2),
    ('fgcnn_cin', dict(nets=['fgcnn_cin_nets'], fgcnn_params={'fg_filters': (3, 4), 'fg_heights': (3, 2), 'fg_pool_heights': (2, 2), 'fg_new_feat_filters': (2, 1)}, cin_params=CIN_SMALL), [7, 5, 9, 4], 4, 0, 'binary', 2),
    ('fgcnn_afm', dict(nets=['fgcnn_afm_nets'], fgcnn_params={'fg_filters': (3, 4), 'fg_heights': (3, 2), 'fg_pool_heights': (2, 2), 'fg_new_feat_filters': (2, 1)}, afm_params={'hidden_factor': 4}), [7, 5, 9], 4, 1, 'binary', 2),
    ('fgcnn_ipnn', dict(nets=['fgcnn_ipnn_nets'], fgcnn_params={'fg_filters': (3, 4), 'fg_heights': (3, 2), 'fg_pool_heights': (2, 2), 'fg_new_feat_filters': (2, 1)}), [7, 5, 9, 4], 4, 2, 'binary', 2),
    ('fg_only', dict(nets=['fg_nets'], fgcnn_params={'fg_filters': (3, 4), 'fg_heights': (3, 2), 'fg_pool_heights': (2, 2), 'fg_new_feat_filters': (2, 1)}), [7, 5, 9], 4, 1, 'binary', 2),
    ('fibi_each_plus_fm', dict(nets=['fm_nets', 'fibi_nets'], fibinet_params={'senet_pooling_op': 'mean', 'senet_reduction_ratio': 3,
                                                                              'bilinear_type': 'field_each'}), [7, 5, 9, 4], 4, 1, 'binary', 2),
]


def collect_state_f3(created):
    children, state = set(), {}
    for lyr in created:
        for attr in F3_CHILD_ATTRS:
            child = getattr(lyr, attr, None)
            if isinstance(child, tf_shim.Layer):
                children.add(id(child))
                for k, v in child.weights_by_name.items():
                    state[f'{lyr.name}/{attr}/{k}'] = np64(v)
    for lyr in created:
        if id(lyr) not in children:
            for k, v in lyr.weights_by_name.items():
                state[f'{lyr.name}/{k}'] = np64(v)
    return state


def make_f3_cases(layers, deepmodel, config_mod, metainfo, counter, rec):
    # AFM (layers.py:742-812): list of F tensors (B,1,D) -> (B,1)
    for case, params, n_f in (('afm_relu', {'hidden_factor': 6}, 4), ('afm_linear', {'hidden_factor': 3, 'activation': 'linear'}, 3)):
        embs = [rand(5, 1, 3, seed=70 + i) for i in range(n_f)]
        lyr = layers.AFM(params=params, name='afm_layer')
        out = lyr(embs)
        rec.add(case, 'afm', {**params, 'n_fields': n_f}, out=np64(out), att_kernel=np64(lyr.dense_attention.weights_by_name['kernel']),
                att_bias=np64(lyr.dense_attention.weights_by_name['bias']), projection_h=np64(lyr.weights_by_name['projection_h']),
                out_kernel=np64(lyr.dense_out.weights_by_name['kernel']), **{f'e{i}': np64(e) for i, e in enumerate(embs)})
    # SENET (layers.py:245-311)
    for case, op, ratio in (('senet_mean', 'mean', 3), ('senet_max', 'max', 2)):
        x = rand(6, 7, 4, seed=80)
        lyr = layers.SENET(pooling_op=op, reduction_ratio=ratio)
        out = lyr(x)
        rec.add(case, 'senet', {'pooling_op': op, 'reduction_ratio': ratio}, x=np64(x), out=np64(out),
                **{f'w/{sub}/{k}': np64(v) for sub in ('dense_att1', 'dense_att2')
                   for k, v in getattr(lyr, sub).weights_by_name.items()})
    # BilinearInteraction (layers.py:314-382): the three weight-sharing types
    for bt in ('field_all', 'field_each', 'field_interaction'):
        x = rand(6, 5, 3, seed=90)
        lyr = layers.BilinearInteraction(bilinear_type=bt)
        out = lyr(x)
        ws = [lyr.W] if bt == 'field_all' else list(lyr.W_list)
        rec.add(f'bilinear_{bt}', 'bilinear', {'bilinear_type': bt}, x=np64(x), out=np64(out),
                **{f'w{i}': np64(w) for i, w in enumerate(ws)})
    # FGCNN (layers.py:161-242): two stacked layers, odd / even kernel heights, pool heights that do and do not divide F
    x = rand(5, 7, 4, 1, seed=95)
    lyr1 = layers.FGCNN(filters=3, kernel_height=3, new_filters=2, pool_height=2)
    po1, nf1 = lyr1(x)
    lyr2 = layers.FGCNN(filters=4, kernel_height=4, new_filters=1, pool_height=3)
    po2, nf2 = lyr2(po1)
    for case, lyr, xin, po, nf, p_ in (('fgcnn_first', lyr1, x, po1, nf1, dict(filters=3, kernel_height=3, new_filters=2, pool_height=2)),
                                       ('fgcnn_second', lyr2, po1, po2, nf2, dict(filters=4, kernel_height=4, new_filters=1, pool_height=3))):
        rec.add(case, 'fgcnn', p_, x=np64(xin), out=np64(nf), pooled=np64(po),
                conv_kernel=np64(lyr.conv2d.weights_by_name['kernel']), conv_bias=np64(lyr.conv2d.weights_by_name['bias']),
                dense_kernel=np64(lyr.dense_output.weights_by_name['kernel']), dense_bias=np64(lyr.dense_output.weights_by_name['bias']))
    # focal losses (layers.py:983-1083) on probabilities: binary (scalar) and categorical (per sample)
    g = np.random.default_rng(97)
    for case, cols, gamma, alpha in (('focal_binary', 1, 2.0, 0.25), ('focal_multilabel', 3, 1.5, 0.6)):
        p_ = torch.tensor(g.random((8, cols)))
        p_[0, 0], p_[1, 0] = 1e-9, 1.0 - 1e-9                    # inside the clipped range's edges
        y_ = torch.tensor((g.random((8, cols)) < 0.4).astype(np.float64))
        rec.add(case, 'focal_binary', {'gamma': gamma, 'alpha': alpha}, y_true=np64(y_), y_pred=np64(p_),
                out=np64(layers.BinaryFocalLoss(gamma=gamma, alpha=alpha).call(y_, p_.clone())))
    p_ = torch.softmax(torch.tensor(g.normal(size=(6, 4)) * 2), dim=-1)
    y_ = torch.nn.functional.one_hot(torch.tensor(g.integers(0, 4, size=6)), 4).double()
    rec.add('focal_categorical', 'focal_categorical', {'gamma': 2.0, 'alpha': 0.25}, y_true=np64(y_), y_pred=np64(p_),
            out=np64(layers.CategoricalFocalLoss(gamma=2.0, alpha=0.25).call(y_, p_.clone())))
    # whole models with afm_nets / fibi_nets / fibi_dnn_nets / fgcnn_* through DeepModel.__build_model
    for ci, (case, cfg_kwargs, vocab, dim, n_cont, task, num_classes) in enumerate(F3_MODEL_CASES):
        conf = config_mod.ModelConfig(embedding_dropout=0, dense_dropout=0, embeddings_output_dim=dim, **cfg_kwargs)
        cats = [metainfo.CategoricalColumn(f'c{i}', v, dim) for i, v in enumerate(vocab)]
        conts = [metainfo.ContinuousColumn('input_continuous_all', [f'n{i}' for i in range(n_cont)])] if n_cont else []
        b = 9
        g = np.random.default_rng(300 + ci)
        ids = np.stack([g.integers(0, v, size=b) for v in vocab], axis=1)
        cont = g.normal(size=(b, n_cont))
        outs = {}
        for training in (False, True):
            tf_shim.reset_layers()
            tf_shim.seed(3000 + ci)
            tf_shim.set_training(training)
            counter._data_.clear()                       # senet_layer_<n>: n counts fibi_nets calls of the PROCESS (counter.py)
            tf_shim.feed('input_categorical_vars_all', torch.tensor(ids.astype(np.float32)))
            if n_cont:
                tf_shim.feed('input_continuous_all', torch.tensor(cont, dtype=torch.float64))
            dm = deepmodel.DeepModel(task, num_classes, conf, cats, conts)
            model = dm._DeepModel__build_model(task=task, num_classes=num_classes, nets=conf.nets, categorical_columns=cats,
                                               continuous_columns=conts, var_len_categorical_columns=None, config=conf)
            outs[training] = np64(model.output)
            state = collect_state_f3(tf_shim.created_layers())
            tf_shim.set_training(False)
        params = {k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg_kwargs.items()}
        params['nets'] = list(conf.nets)
        rec.add(case, 'model', {'config': json.loads(json.dumps(params)), 'vocab': vocab, 'dim': dim, 'n_cont': n_cont,
                                'task': task, 'num_classes': num_classes},
                ids=ids.astype(np.int64), cont=cont.astype(np.float64), out_infer=outs[False], out_train=outs[True],
                **{f'w/{k}': v for k, v in state.items()})


def dump_modelconfig(config_mod, deepnets, path):
    conf = config_mod.ModelConfig()
    d = conf._asdict()
    d.pop('home_dir')                                     # machine-dependent default
    out = {'fields': list(conf._fields), 'defaults': json.loads(json.dumps(d, default=list)),
           'presets': {k: getattr(deepnets, k) for k in ('WideDeep', 'DeepFM', 'xDeepFM', 'AutoInt', 'DCN', 'FGCNN', 'FiBiNet',
                                                         'PNN', 'AFM')},
           'net_signature': list(__import__('inspect').signature(deepnets.linear).parameters)}
    with open(path, 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)


def main():
    tf_shim.install(REFERENCE_ROOT)
    layers = importlib.import_module('deeptables.models.layers')
    deepnets = importlib.import_module('deeptables.models.deepnets')
    config_mod = importlib.import_module('deeptables.models.config')
    metainfo = importlib.import_module('deeptables.models.metainfo')
    deepmodel = importlib.import_module('deeptables.models.deepmodel')

    rec = Recorder()
    make_layer_cases(layers, deepnets, rec)
    rec.save(os.path.join(HERE, 'reference_layers.npz'))
    print(f'reference_layers.npz: {len(rec.manifest)} cases')

    rec = Recorder()
    make_model_cases(deepmodel, config_mod, metainfo, rec)
    rec.save(os.path.join(HERE, 'reference_models.npz'))
    print(f'reference_models.npz: {len(rec.manifest)} cases')

    rec = Recorder()
    counter = importlib.import_module('deeptables.utils.counter')
    make_f3_cases(layers, deepmodel, config_mod, metainfo, counter, rec)
    rec.save(os.path.join(HERE, 'reference_f3.npz'))
    print(f'reference_f3.npz: {len(rec.manifest)} cases')

    dump_modelconfig(config_mod, deepnets, os.path.join(HERE, 'reference_modelconfig.json'))
    print('reference_modelconfig.json written')


if __name__ == '__main__':
    main()
