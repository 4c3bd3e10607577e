"""CPU oracle: the reference's model assembly (DeepModel.__build_model, deepmodel.py:259-317),
net builders (deepnets.py:43-224, 401-427), loss (deepmodel.py:319-346) and Keras Adam, restated
over a flat state dict keyed by the reference's layer / weight names.

TEST INFRASTRUCTURE ONLY (see oracle/layers_ref.py header).  Parity: ``forward`` is pinned against the
reference's own ``DeepModel.__build_model`` executed over tests/golden/tf_shim.py
(tests/test_reference_golden.py); losses / Adam / BN moving statistics are restated from the Keras
documentation and stay UNPINNED (TF/Keras absent).

The state dict maps ``'<layer>/<weight>'`` -> torch CPU tensor, e.g.
``emb_categorical_vars_all/embeddings_3``, ``bn_concat_emb_dense/gamma``, ``linear_logit/kernel``,
``cin/f_0``, ``cin/exFM_out/kernel``, ``dnn_dense_1/kernel``, ``cross_layer/kernels_0``,
``multihead_attention_1/dense_Q/kernel``, ``pnn_outer_product_layer/kernel``,
``dense_logit_dnn_nets/kernel``, ``task_output/kernel``.  Non-trainable BN statistics use
``.../moving_mean`` and ``.../moving_variance``.
"""
import math
import numpy as np
import torch

from . import layers_ref as L

SUPPORTED_NETS = ('linear', 'cin_nets', 'fm_nets', 'opnn_nets', 'ipnn_nets', 'pnn_nets', 'dnn_nets',
                  'cross_nets', 'cross_dnn_nets', 'dcn_nets', 'autoint_nets', 'afm_nets', 'fibi_nets', 'fibi_dnn_nets', 'fg_nets',
                  'fgcnn_cin_nets', 'fgcnn_fm_nets', 'fgcnn_afm_nets', 'fgcnn_ipnn_nets', 'fgcnn_dnn_nets')


def bilinear_weight_names(layer, f, bilinear_type):
    """BilinearInteraction.build (layers.py:343-356): weight names per sharing type, in the order call() uses them."""
    if bilinear_type == 'field_all':
        return [f'{layer}/bilinear_weight']
    if bilinear_type == 'field_each':
        return [f'{layer}/bilinear_weight{i}' for i in range(f - 1)]
    return [f'{layer}/bilinear_weight{i}_{j}' for i in range(f) for j in range(i + 1, f)]


FG_NETS = ('fg_nets', 'fgcnn_cin_nets', 'fgcnn_fm_nets', 'fgcnn_afm_nets', 'fgcnn_ipnn_nets', 'fgcnn_dnn_nets')


def _fg_levels(f, params):
    """fg_nets (deepnets.py:245-258): per FGCNN layer (filters, kernel height, pool height, new filters, fields in, fields out)."""
    levels, h = [], f
    for filters, kh, pool, new in zip(params.get('fg_filters', (14, 16)), params.get('fg_heights', (7, 7)),
                                      params.get('fg_pool_heights', (2, 2)), params.get('fg_new_feat_filters', (2, 2))):
        h_out = -(-h // pool)
        levels.append((filters, kh, pool, new, h, h_out))
        h = h_out
    return levels


def _fg_entries(f, d0, params, first_layer_index):
    """Weights of the FGCNN layers of one fg_nets call (auto-named fgcnn, fgcnn_1, ... in creation order), the number of
    fields of its output (new features + the original embeddings) and the next free layer index."""
    ents, cin, total, k = [], 1, f, first_layer_index
    for filters, kh, pool, new, h, h_out in _fg_levels(f, params):
        name = 'fgcnn' if k == 0 else f'fgcnn_{k}'
        ents += [(f'{name}/conv2d/kernel', (kh, 1, cin, filters), 'glorot_uniform'), (f'{name}/conv2d/bias', (filters,), 'zeros'),
                 (f'{name}/dense_output/kernel', (h_out * d0 * filters, h * d0 * new), 'glorot_uniform'),
                 (f'{name}/dense_output/bias', (h * d0 * new,), 'zeros')]
        total += h * new
        cin = filters
        k += 1
    return ents, total, k


def fg_forward(state, cat, params, first_layer_index):
    """fg_nets (deepnets.py:227-261): (B, F, D) -> (B, F_fg, D) = [new features of every FGCNN layer ..., the embeddings]."""
    x = cat.unsqueeze(-1)
    feats, k = [], first_layer_index
    for filters, kh, pool, new, h, h_out in _fg_levels(cat.shape[1], params):
        name = 'fgcnn' if k == 0 else f'fgcnn_{k}'
        x, nf = L.fgcnn(x, state[f'{name}/conv2d/kernel'], state[f'{name}/conv2d/bias'], state[f'{name}/dense_output/kernel'],
                        state[f'{name}/dense_output/bias'], pool, new)
        feats.append(nf)
        k += 1
    return torch.cat(feats + [cat], dim=1), k


def _cin_entries(f, d0, p, name):
    ents = []
    sizes = tuple(p.get('cross_layer_size', (128, 128)))
    fns = L.cin_field_nums(f, sizes, p.get('direct', False))
    for k, size in enumerate(sizes):
        if p.get('reduce_D', False):
            ents.append((f'{name}/f0_{k}', (1, size, fns[0], d0), 'he_uniform'))
            ents.append((f'{name}/f__{k}', (1, size, d0, fns[k]), 'he_uniform'))
        else:
            ents.append((f'{name}/f_{k}', (1, fns[k] * fns[0], size), 'he_uniform'))
        if p.get('use_bias', False):
            ents.append((f'{name}/bias{k}', (size,), 'zeros'))
    pooled = L.cin_pooled_width(f, p)
    if p.get('use_residual', False):
        ents.append((f'{name}/exFM_out0/kernel', (pooled, sizes[-1]), 'he_uniform'))
        ents.append((f'{name}/exFM_out0/bias', (sizes[-1],), 'zeros'))
        ents.append((f'{name}/exFM_out/kernel', (pooled + sizes[-1], 1), 'glorot_uniform'))
    else:
        ents.append((f'{name}/exFM_out/kernel', (pooled, 1), 'glorot_uniform'))
    ents.append((f'{name}/exFM_out/bias', (1,), 'zeros'))
    return ents


def _fibi_entries(f, d0, params, index):
    """fibi_nets (deepnets.py:344-371): SENET + two BilinearInteraction layers; `index` = senet_index."""
    ratio = params.get('senet_reduction_ratio', 3)
    bt = params.get('bilinear_type', 'field_interaction')
    red = max(f // ratio, 1)
    ents = [(f'senet_layer_{index}/dense_att1/kernel', (f, red), 'he_uniform'),
            (f'senet_layer_{index}/dense_att1/bias', (red,), 'zeros'),
            (f'senet_layer_{index}/dense_att2/kernel', (red, f), 'he_uniform'),
            (f'senet_layer_{index}/dense_att2/bias', (f,), 'zeros')]
    for lname in (f'senet_bilinear_layer_{index}', f'embedding_bilinear_layer_{index}'):
        ents += [(n, (d0, d0), 'glorot_uniform') for n in bilinear_weight_names(lname, f, bt)]
    return ents


def _get(cfg, name, default=None):
    if isinstance(cfg, dict):
        return cfg.get(name, default)
    return getattr(cfg, name, default)


# ---------------------------------------------------------------------------------------------
# Keras initialisers
# ---------------------------------------------------------------------------------------------
def _fans(shape):
    if len(shape) < 1:
        return 1, 1
    if len(shape) == 1:
        return shape[0], shape[0]
    if len(shape) == 2:
        return shape[0], shape[1]
    rf = int(np.prod(shape[:-2]))
    return shape[-2] * rf, shape[-1] * rf


def init_weight(rng, shape, kind):
    shape = tuple(int(s) for s in shape)
    fan_in, fan_out = _fans(shape)
    if kind == 'uniform':
        lim = 0.05
    elif kind == 'glorot_uniform':
        lim = math.sqrt(6.0 / (fan_in + fan_out))
    elif kind == 'he_uniform':
        lim = math.sqrt(6.0 / fan_in)
    elif kind == 'glorot_normal':
        # keras GlorotNormal: truncated normal (2 sigma), stddev sqrt(2 / (fan_in + fan_out)) / 0.87962566
        std = math.sqrt(2.0 / (fan_in + fan_out)) / 0.87962566103423978
        out = rng.normal(0.0, std, size=shape)
        bad = np.abs(out) > 2 * std
        while bad.any():
            out[bad] = rng.normal(0.0, std, size=int(bad.sum()))
            bad = np.abs(out) > 2 * std
        return torch.from_numpy(out.astype(np.float32))
    elif kind == 'zeros':
        return torch.zeros(shape, dtype=torch.float32)
    elif kind == 'ones':
        return torch.ones(shape, dtype=torch.float32)
    else:
        raise ValueError(kind)
    return torch.from_numpy(rng.uniform(-lim, lim, size=shape).astype(np.float32))


def _bn_entries(name, width):
    return [(f'{name}/gamma', (width,), 'ones'), (f'{name}/beta', (width,), 'zeros'),
            (f'{name}/moving_mean', (width,), 'zeros'), (f'{name}/moving_variance', (width,), 'ones')]


def _dnn_entries(width, params, cellname):
    ents = []
    hidden = params.get('hidden_units', ((128, 0, True), (64, 0, False)))
    kinit = params.get('kernel_initializer', 'he_uniform')
    for i, (units, _drop, use_bn) in enumerate(hidden, start=1):
        ents.append((f'{cellname}_dense_{i}/kernel', (width, units), kinit))
        if use_bn:
            ents += _bn_entries(f'{cellname}_bn_{i}', units)
        else:
            ents.append((f'{cellname}_dense_{i}/bias', (units,), 'zeros'))
        width = units
    return ents, width


def param_spec(config, vocab_sizes, emb_dims, n_cont, task='binary', num_classes=2):
    """[(name, shape, initializer)] in build order + {net: output width} (deepmodel.py:259-317)."""
    nets = list(_get(config, 'nets'))
    f = len(vocab_sizes)
    ents = []
    for i, (v, d) in enumerate(zip(vocab_sizes, emb_dims)):
        ents.append((f'emb_categorical_vars_all/embeddings_{i}', (v, d),
                     _get(config, 'embeddings_initializer', 'uniform')))
    sum_d = int(sum(emb_dims))
    w = sum_d + n_cont
    ents += _bn_entries('bn_concat_emb_dense', w)
    widths = {}
    d0 = emb_dims[0] if f else 0
    n_fibi = 0
    n_fg = 0
    for net in nets:
        if net == 'linear':
            ents.append(('linear_logit/kernel', (f + n_cont, 1), 'glorot_uniform'))
            widths[net] = 1
        elif net == 'fm_nets':
            if f:
                widths[net] = 1
        elif net == 'cin_nets':
            if not f:
                continue
            ents += _cin_entries(f, d0, _get(config, 'cin_params'), 'cin')
            widths[net] = 1
        elif net in FG_NETS:
            if not f:
                continue
            fg_ents, f_fg, n_fg = _fg_entries(f, d0, _get(config, 'fgcnn_params'), n_fg)
            ents += fg_ents
            if net == 'fg_nets':
                widths[net] = f_fg * d0
            elif net == 'fgcnn_cin_nets':
                ents += _cin_entries(f_fg, d0, _get(config, 'cin_params'), 'cin')
                widths[net] = 1
            elif net == 'fgcnn_fm_nets':
                widths[net] = 1
            elif net == 'fgcnn_afm_nets':
                h = _get(config, 'afm_params').get('hidden_factor', 16)
                ents += [('afm/dense_attention/kernel', (d0, h), 'glorot_normal'), ('afm/dense_attention/bias', (h,), 'zeros'),
                         ('afm/dense_out/kernel', (d0, 1), 'glorot_uniform'), ('afm/projection_h', (h, 1), 'glorot_uniform')]
                widths[net] = 1
            else:
                cell = 'fgcnn_ipnn' if net == 'fgcnn_ipnn_nets' else 'fgcnn_dnn'
                width_in = f_fg * d0 + n_cont + (f_fg * (f_fg - 1) // 2 if net == 'fgcnn_ipnn_nets' else 0)
                e, width = _dnn_entries(width_in, _get(config, 'dnn_params'), cell)
                ents += e
                widths[net] = width
        elif net in ('opnn_nets', 'ipnn_nets', 'pnn_nets'):
            if f < 2:
                continue
            cell = net[:-5]
            pairs = f * (f - 1) // 2
            extra = 0
            if net in ('ipnn_nets', 'pnn_nets'):
                extra += pairs
            if net in ('opnn_nets', 'pnn_nets'):
                kt = _get(config, 'pnn_params').get('outer_product_kernel_type', 'mat')
                shape = {'mat': (d0, pairs, d0), 'vec': (pairs, d0), 'num': (pairs, 1)}[kt]
                lname = 'pnn_outer_product_layer' if net == 'pnn_nets' else 'outer_product_layer'
                ents.append((f'{lname}/kernel', shape, 'glorot_uniform'))
                extra += pairs
            e, width = _dnn_entries(w + extra, _get(config, 'dnn_params'), cell)
            ents += e
            widths[net] = width
        elif net == 'dnn_nets':
            e, width = _dnn_entries(w, _get(config, 'dnn_params'), 'dnn')
            ents += e
            widths[net] = width
        elif net in ('cross_nets', 'cross_dnn_nets', 'dcn_nets'):
            lname = {'cross_nets': 'cross_layer', 'cross_dnn_nets': 'cross_dnn_layer',
                     'dcn_nets': 'dcn_cross_layer'}[net]
            n_layers = _get(config, 'cross_params').get('num_cross_layer', 2)
            for i in range(n_layers):
                ents.append((f'{lname}/kernels_{i}', (w, 1), 'glorot_uniform'))
                ents.append((f'{lname}/bias_{i}', (w, 1), 'zeros'))
            if net == 'cross_nets':
                widths[net] = w
            elif net == 'cross_dnn_nets':
                e, width = _dnn_entries(w, _get(config, 'dnn_params'), 'cross_dnn')
                ents += e
                widths[net] = width
            else:
                e, width = _dnn_entries(w, _get(config, 'dnn_params'), 'dcn')
                ents += e
                widths[net] = w + width
        elif net == 'afm_nets':
            if f < 2:
                continue
            h = _get(config, 'afm_params').get('hidden_factor', 16)      # the layer reads 'hidden_factor' (layers.py:773)
            ents += [('afm_layer/dense_attention/kernel', (d0, h), 'glorot_normal'),
                     ('afm_layer/dense_attention/bias', (h,), 'zeros'),
                     ('afm_layer/dense_out/kernel', (d0, 1), 'glorot_uniform'),
                     ('afm_layer/projection_h', (h, 1), 'glorot_uniform')]
            widths[net] = 1
        elif net in ('fibi_nets', 'fibi_dnn_nets'):
            if f < 1 or (net == 'fibi_dnn_nets' and f < 2):
                continue
            ents += _fibi_entries(f, d0, _get(config, 'fibinet_params'), n_fibi)
            n_fibi += 1
            pairs = f * (f - 1) // 2
            if net == 'fibi_nets':
                widths[net] = 2 * pairs * d0
            else:
                e, width = _dnn_entries(2 * pairs * d0 + n_cont, _get(config, 'dnn_params'), 'fibi_dnn')
                ents += e
                widths[net] = width
        elif net == 'autoint_nets':
            if not f:
                continue
            p = _get(config, 'autoint_params')
            for i in range(p['num_attention']):
                lname = 'multihead_attention' if i == 0 else f'multihead_attention_{i}'
                for proj in ('dense_Q', 'dense_K', 'dense_V', 'dense_residual'):
                    ents.append((f'{lname}/{proj}/kernel', (d0, d0), 'he_uniform'))
                    ents.append((f'{lname}/{proj}/bias', (d0,), 'zeros'))
                ents += _bn_entries(f'{lname}/batch_normalize', d0)
            widths[net] = f * d0
        else:
            raise NotImplementedError(f'oracle: net {net!r} is outside the hot path')
    live = [n for n in nets if n in widths]
    if len(live) > 1:
        for n in live:
            if widths[n] > 1:
                ents.append((f'dense_logit_{n}/kernel', (widths[n], 1), 'glorot_uniform'))
        head_in = 1 if _get(config, 'stacking_op', 'add') == 'add' else len(live)
    elif len(live) == 1:
        head_in = widths[live[0]]
    else:
        raise ValueError('Unexpected logit output.')
    out_dim = 1 if task in ('binary', 'regression') else num_classes
    ents.append(('task_output/kernel', (head_in, out_dim), 'glorot_uniform'))
    if _get(config, 'output_use_bias', True):
        ents.append(('task_output/bias', (out_dim,), 'zeros'))
    return ents, widths


def init_state(config, vocab_sizes, emb_dims, n_cont, task='binary', num_classes=2, seed=0):
    rng = np.random.default_rng(seed)
    ents, _ = param_spec(config, vocab_sizes, emb_dims, n_cont, task, num_classes)
    return {name: init_weight(rng, shape, kind) for name, shape, kind in ents}


def is_trainable(name):
    return not (name.endswith('/moving_mean') or name.endswith('/moving_variance'))


def _sub(state, prefix):
    plen = len(prefix) + 1
    return {k[plen:]: v for k, v in state.items() if k.startswith(prefix + '/')}


# ---------------------------------------------------------------------------------------------
# Forward (deepmodel.py:259-317 + deepnets.py builders)
# ---------------------------------------------------------------------------------------------
def forward(state, config, cat_idx, cont, n_fields, training, task='binary', return_parts=False):
    """cat_idx: (B,F) integer/float tensor or None; cont: (B,C) float tensor or None.
    Returns (output, new_bn_state[, parts])."""
    nets = list(_get(config, 'nets'))
    new_bn = {}
    parts = {}
    embeddings = []
    if n_fields:
        tables = [state[f'emb_categorical_vars_all/embeddings_{i}'] for i in range(n_fields)]
        embeddings = L.embedding_lookup(tables, cat_idx)
    dense_layer = cont
    flat = L.flatten_embeddings(embeddings) if embeddings else None
    if flat is not None and dense_layer is not None:
        x = torch.cat([flat, dense_layer], dim=-1)
    elif flat is not None:
        x = flat
    elif dense_layer is not None:
        x = dense_layer
    else:
        raise ValueError('No input layer exists.')
    ced, nm, nv = L.batch_norm(x, state['bn_concat_emb_dense/gamma'], state['bn_concat_emb_dense/beta'],
                               state['bn_concat_emb_dense/moving_mean'],
                               state['bn_concat_emb_dense/moving_variance'], training)
    new_bn['bn_concat_emb_dense/moving_mean'] = nm
    new_bn['bn_concat_emb_dense/moving_variance'] = nv
    parts['concat_emb_dense'] = ced
    bn_state = {k: v for k, v in state.items() if not is_trainable(k)}

    def run_dnn(inp, cell):
        y, ns = L.dnn(inp, _get(config, 'dnn_params'), state, bn_state, training, cellname=cell)
        new_bn.update(ns)
        return y

    outs = {}
    n_fibi = 0
    n_fg = 0
    for net in nets:
        if net == 'linear':
            outs[net] = L.linear(embeddings, dense_layer, state['linear_logit/kernel'])
        elif net == 'fm_nets':
            cat = L.concat_embeddings(embeddings)
            if cat is not None:
                outs[net] = L.fm(cat)
        elif net == 'cin_nets':
            cat = L.concat_embeddings(embeddings)
            if cat is not None:
                outs[net] = L.cin(cat, _get(config, 'cin_params'), _sub(state, 'cin'))
        elif net in ('opnn_nets', 'ipnn_nets', 'pnn_nets'):
            if len(embeddings) < 2:
                continue
            feats = []
            if net in ('ipnn_nets', 'pnn_nets'):
                feats.append(L.inner_product(embeddings))
            if net in ('opnn_nets', 'pnn_nets'):
                lname = 'pnn_outer_product_layer' if net == 'pnn_nets' else 'outer_product_layer'
                kt = _get(config, 'pnn_params').get('outer_product_kernel_type', 'mat')
                feats.append(L.outer_product(embeddings, state[f'{lname}/kernel'], kt))
            outs[net] = run_dnn(torch.cat(feats + [ced], dim=-1), net[:-5])
        elif net == 'dnn_nets':
            outs[net] = run_dnn(ced, 'dnn')
        elif net in ('cross_nets', 'cross_dnn_nets', 'dcn_nets'):
            lname = {'cross_nets': 'cross_layer', 'cross_dnn_nets': 'cross_dnn_layer',
                     'dcn_nets': 'dcn_cross_layer'}[net]
            n_layers = _get(config, 'cross_params').get('num_cross_layer', 2)
            ks = [state[f'{lname}/kernels_{i}'] for i in range(n_layers)]
            bs = [state[f'{lname}/bias_{i}'] for i in range(n_layers)]
            c = L.cross(ced, ks, bs)
            if net == 'cross_nets':
                outs[net] = c
            elif net == 'cross_dnn_nets':
                outs[net] = run_dnn(c, 'cross_dnn')
            else:
                outs[net] = torch.cat([c, run_dnn(ced, 'dcn')], dim=-1)
        elif net in FG_NETS:
            cat = L.concat_embeddings(embeddings)
            if cat is None:
                continue
            fg, n_fg = fg_forward(state, cat, _get(config, 'fgcnn_params'), n_fg)
            split = [fg[:, i:i + 1, :] for i in range(fg.shape[1])]
            if net == 'fg_nets':
                outs[net] = fg
            elif net == 'fgcnn_cin_nets':
                outs[net] = L.cin(fg, _get(config, 'cin_params'), _sub(state, 'cin'))
            elif net == 'fgcnn_fm_nets':
                outs[net] = L.fm(fg)
            elif net == 'fgcnn_afm_nets':
                outs[net] = L.afm(split, state['afm/dense_attention/kernel'], state['afm/dense_attention/bias'],
                                  state['afm/projection_h'], state['afm/dense_out/kernel'],
                                  _get(config, 'afm_params').get('activation', 'relu'))
            else:
                parts_in = [fg.reshape(fg.shape[0], -1)]
                if net == 'fgcnn_ipnn_nets':
                    parts_in.append(L.inner_product(split))
                if dense_layer is not None:
                    parts_in.append(dense_layer)
                outs[net] = run_dnn(torch.cat(parts_in, dim=-1), 'fgcnn_ipnn' if net == 'fgcnn_ipnn_nets' else 'fgcnn_dnn')
        elif net == 'afm_nets':
            if len(embeddings) < 2:
                continue
            act = _get(config, 'afm_params').get('activation', 'relu')
            outs[net] = L.afm(embeddings, state['afm_layer/dense_attention/kernel'], state['afm_layer/dense_attention/bias'],
                              state['afm_layer/projection_h'], state['afm_layer/dense_out/kernel'], act)
        elif net in ('fibi_nets', 'fibi_dnn_nets'):
            if len(embeddings) < 1 or (net == 'fibi_dnn_nets' and len(embeddings) < 2):
                continue
            p = _get(config, 'fibinet_params')
            bt = p.get('bilinear_type', 'field_interaction')
            cat = L.concat_embeddings(embeddings)
            sl = f'senet_layer_{n_fibi}'
            sen = L.senet(cat, state[f'{sl}/dense_att1/kernel'], state[f'{sl}/dense_att1/bias'],
                          state[f'{sl}/dense_att2/kernel'], state[f'{sl}/dense_att2/bias'], p.get('senet_pooling_op', 'mean'))
            f_ = cat.shape[1]
            w_s = [state[n] for n in bilinear_weight_names(f'senet_bilinear_layer_{n_fibi}', f_, bt)]
            w_e = [state[n] for n in bilinear_weight_names(f'embedding_bilinear_layer_{n_fibi}', f_, bt)]
            n_fibi += 1
            fibi = torch.cat([L.bilinear_interaction(sen, w_s, bt), L.bilinear_interaction(cat, w_e, bt)], dim=1)
            if net == 'fibi_nets':
                outs[net] = fibi
            else:
                outs[net] = run_dnn(torch.cat([fibi.reshape(fibi.shape[0], -1), dense_layer], dim=-1), 'fibi_dnn')
        elif net == 'autoint_nets':
            cat = L.concat_embeddings(embeddings)
            if cat is None:
                continue
            p = _get(config, 'autoint_params')
            out = cat
            for i in range(p['num_attention']):
                lname = 'multihead_attention' if i == 0 else f'multihead_attention_{i}'
                st = {'moving_mean': state[f'{lname}/batch_normalize/moving_mean'],
                      'moving_variance': state[f'{lname}/batch_normalize/moving_variance']}
                out, ns = L.multihead_attention(out, p, _sub(state, lname), st, training)
                new_bn[f'{lname}/batch_normalize/moving_mean'] = ns['moving_mean']
                new_bn[f'{lname}/batch_normalize/moving_variance'] = ns['moving_variance']
            outs[net] = out.reshape(out.shape[0], -1)
        else:
            raise NotImplementedError(net)
    parts['outs'] = outs
    if len(outs) > 1:
        logits = []
        for name, out in outs.items():
            if out.dim() > 2:
                out = out.reshape(out.shape[0], -1)
            if out.shape[-1] > 1:
                out = out @ state[f'dense_logit_{name}/kernel']
            logits.append(out)
        if _get(config, 'stacking_op', 'add') == 'add':
            x = sum(logits[1:], logits[0])
        elif _get(config, 'stacking_op') == 'concat':
            x = torch.cat(logits, dim=-1)
        else:
            raise ValueError(f"Unsupported stacking_op:{_get(config, 'stacking_op')}.")
    elif len(outs) == 1:
        x = list(outs.values())[0]
        if x.dim() > 2:
            x = x.reshape(x.shape[0], -1)
    else:
        raise ValueError(f'Unexpected logit output.{outs}')
    parts['stack'] = x
    z = x @ state['task_output/kernel']
    if 'task_output/bias' in state:
        z = z + state['task_output/bias']
    parts['z'] = z
    if task in ('binary', 'multilabel'):
        y = torch.sigmoid(z)
    elif task == 'regression':
        y = z
    elif task == 'multiclass':
        y = torch.softmax(z, dim=-1)
    else:
        raise ValueError(f'Unknown task type:{task}')
    if return_parts:
        return y, new_bn, parts
    return y, new_bn


def loss_fn(task, y_true, y_pred, loss='auto'):
    """ModelConfig.loss: 'auto' (deepmodel.py:327-336) or one of the reference's focal-loss objects (recognised by class
    name and their gamma / alpha attributes, so that this file imports nothing of the product)."""
    kind = type(loss).__name__
    if kind == 'BinaryFocalLoss':
        return L.binary_focal_loss(y_true, y_pred, loss.gamma, loss.alpha)
    if kind == 'CategoricalFocalLoss':
        return L.categorical_focal_loss(y_true, y_pred, loss.gamma, loss.alpha).mean()
    if task in ('binary', 'multilabel'):
        return L.binary_crossentropy(y_true, y_pred)
    if task == 'regression':
        return L.mean_squared_error(y_true, y_pred)
    if task == 'multiclass':
        return L.categorical_crossentropy(y_true, y_pred)
    raise RuntimeError(f'unseen task "{task}"')


class RefTrainer:
    """One Keras-style train step at a time: forward (training=True) -> loss -> autograd ->
    dense Adam on every trainable weight (embedding tables included, as Keras does)."""

    def __init__(self, state, config, n_fields, task='binary', dtype=torch.float32):
        self.state = {k: v.detach().clone().to(dtype) for k, v in state.items()}
        self.config = config
        self.n_fields = n_fields
        self.task = task
        self.step = 0
        self.m = {k: torch.zeros_like(v) for k, v in self.state.items() if is_trainable(k)}
        self.v = {k: torch.zeros_like(v) for k, v in self.state.items() if is_trainable(k)}

    def loss_and_grads(self, cat_idx, cont, y):
        params = {k: v.detach().clone().requires_grad_(is_trainable(k)) for k, v in self.state.items()}
        out, new_bn = forward(params, self.config, cat_idx, cont, self.n_fields, True, self.task)
        loss = loss_fn(self.task, y.to(out.dtype).reshape(out.shape[0], -1), out, _get(self.config, 'loss', 'auto'))
        names = [k for k in params if is_trainable(k)]
        grads = torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)
        gd = {k: (g if g is not None else torch.zeros_like(params[k])) for k, g in zip(names, grads)}
        return loss.detach(), gd, {k: v.detach() for k, v in new_bn.items()}, out.detach()

    def train_step(self, cat_idx, cont, y):
        loss, grads, new_bn, _ = self.loss_and_grads(cat_idx, cont, y)
        self.step += 1
        for k, g in grads.items():
            L.adam_step(self.state[k], g, self.m[k], self.v[k], self.step)
        self.state.update(new_bn)
        return float(loss)

    def predict(self, cat_idx, cont):
        with torch.no_grad():
            out, _ = forward(self.state, self.config, cat_idx, cont, self.n_fields, False, self.task)
        return out
