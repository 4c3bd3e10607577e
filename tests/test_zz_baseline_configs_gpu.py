"""BASELINE.json configs [1], [3], [4] at their FULL shapes (Criteo: 26 sparse x vocab 1M + 13 dense), as parity-test
cases: the oracle cannot hold 1.66 GB tables comfortably, so each case checks size-independent properties on the
whole batch and compares a sample of rows against the CPU oracle run on a compacted copy of the weights
(only the embedding rows those sample rows reference).

Written after the round-1 GPU budget was spent, hence non-strict xfail until it has run once on a B200 (the file
name sorts last on purpose)."""
import numpy as np
import pytest
import torch

from oracle import model_ref as M

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason='new: not yet executed on a B200 (round-1 GPU budget spent)')]

F, C, V = 26, 13, 1_000_000

CASES = {
    # config [1]: DeepFM, embed_dim 16, bs 8192
    'deepfm_bs8192': dict(nets=['linear', 'fm_nets', 'dnn_nets'], dim=16, batch=8192, kw={}),
    # config [3]: DCN CrossNet depth 6 stacked with AutoInt 4-head d=32, bs 65536
    'dcn6_autoint4x32_bs65536': dict(nets=['dcn_nets', 'autoint_nets'], dim=32, batch=65536,
                                     kw=dict(cross_params={'num_cross_layer': 6},
                                             autoint_params={'num_attention': 3, 'num_heads': 4, 'dropout_rate': 0,
                                                             'use_residual': True})),
    # config [4]: the five-net mix; 131072 global rows over 8 GPUs = 16384 per GPU
    'five_nets_bs16384': dict(nets=['fm_nets', 'cin_nets', 'cross_nets', 'autoint_nets', 'pnn_nets'], dim=16, batch=16384,
                              kw=dict(cin_params={'cross_layer_size': (128, 128, 128), 'activation': 'relu',
                                                  'use_residual': False, 'use_bias': False, 'direct': False,
                                                  'reduce_D': False})),
}


def _build(nets, dim, kw):
    from deeptables_b200 import deeptable
    from deeptables_b200.deepmodel import DeepModel
    from deeptables_b200.metainfo import CategoricalColumn, ContinuousColumn
    conf = deeptable.ModelConfig(nets=nets, embeddings_output_dim=dim, embedding_dropout=0, metrics=['AUC'], **kw)
    cats = [CategoricalColumn(f'C{i + 1}', V, dim) for i in range(F)]
    conts = [ContinuousColumn('input_continuous_all', [f'I{i + 1}' for i in range(C)])]
    model = DeepModel('binary', 2, conf, cats, conts, seed=21)
    model._build_model()
    return model, conf


@pytest.mark.parametrize('case', sorted(CASES))
def test_baseline_config_full_shape(case):
    spec = CASES[case]
    model, conf = _build(spec['nets'], spec['dim'], spec['kw'])
    b, dim = spec['batch'], spec['dim']
    g = torch.Generator().manual_seed(5)
    idx = torch.randint(0, V, (b, F), generator=g, dtype=torch.int32)
    cont = torch.randn(b, C, generator=g)
    half = b // 2
    idx[half:] = idx[:half]                                   # duplicated rows
    cont[half:] = cont[:half]
    d_idx, d_cont = idx.cuda(), cont.cuda()
    out = model.predict_step(d_idx, d_cont)
    assert out.shape == (b, 1) and bool(torch.isfinite(out).all())
    # (1) duplicated rows give identical outputs (no cross-row coupling in inference)
    torch.testing.assert_close(out[:half], out[half:], rtol=1e-6, atol=1e-7)
    # (2) a row permutation permutes the output
    perm = torch.randperm(b, generator=g)
    out_p = model.predict_step(d_idx[perm.cuda()], d_cont[perm.cuda()])
    torch.testing.assert_close(out_p, out[perm.cuda()], rtol=1e-5, atol=1e-6)
    # (3) a sample of rows against the oracle, on weights compacted to the rows the sample references
    sample = torch.arange(0, half, max(1, half // 48))[:48]
    s_idx = idx[sample]
    state = {}
    sd = model.state_dict()
    for i in range(F):
        rows = s_idx[:, i].long().cuda()
        state[f'emb_categorical_vars_all/embeddings_{i}'] = sd[f'emb_categorical_vars_all/embeddings_{i}'][rows].cpu()
    for k, v in sd.items():
        if not k.startswith('emb_categorical_vars_all/'):
            state[k] = v.detach().cpu().clone()
    local_ids = torch.arange(len(sample), dtype=torch.int64).unsqueeze(1).repeat(1, F)   # row r of every compact table
    want, _ = M.forward(state, conf, local_ids, cont[sample], F, False)
    torch.testing.assert_close(out[sample.cuda()].cpu(), want, rtol=1e-3, atol=1e-5)      # north_star: 1e-3 relative fp32
    # (4) a few optimiser steps at the full shape stay finite and reduce the loss on a fixed batch
    y = (torch.rand(b, generator=g) < 0.25).float().numpy()
    losses = [model.train_on_batch(idx.numpy(), cont.numpy(), y) for _ in range(4)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0]
    model.release()
