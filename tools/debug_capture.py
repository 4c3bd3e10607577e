"""Debug aid: which net configurations capture their train step in a CUDA graph, and, with DTB_DEBUG_CAPTURE=1, the first
native call after which the capture is invalid."""
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeptables_b200 import deeptable  # noqa: E402
from deeptables_b200.deepmodel import DeepModel  # noqa: E402
from deeptables_b200.metainfo import CategoricalColumn, ContinuousColumn  # noqa: E402

vocab, dim, n_cont, b = [11, 7, 13, 5, 9], 8, 3, 256
for nets in (['linear'], ['fm_nets'], ['dnn_nets'], ['cross_nets'], ['autoint_nets'], ['pnn_nets'], ['cin_nets']):
    conf = deeptable.ModelConfig(nets=nets, embeddings_output_dim=dim, embedding_dropout=0, metrics=['AUC'],
                                 autoint_params={'num_attention': 2, 'num_heads': 2, 'dropout_rate': 0, 'use_residual': True},
                                 cin_params={'cross_layer_size': (32, 32), 'activation': 'relu', 'use_residual': False,
                                             'use_bias': False, 'direct': False, 'reduce_D': False})
    cats = [CategoricalColumn(f'c{i}', v, dim) for i, v in enumerate(vocab)]
    conts = [ContinuousColumn('input_continuous_all', [f'n{i}' for i in range(n_cont)])]
    model = DeepModel('binary', 2, conf, cats, conts, seed=3)
    model._build_model()
    g = np.random.default_rng(0)
    idx = np.stack([g.integers(0, v, size=b) for v in vocab], axis=1).astype(np.int32)
    cont = g.normal(size=(b, n_cont)).astype(np.float32)
    y = (g.random(b) < 0.4).astype(np.float32)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        try:
            losses = [model.train_on_batch(idx, cont, y) for _ in range(4)]
            msg = [str(x.message)[:300] for x in w if 'capture' in str(x.message)]
            print(f'{nets}: graph {bool(model._graphs) and not model._graph_failed}  losses {[round(l, 4) for l in losses]}  {msg}', flush=True)
        except Exception as exc:
            print(f'{nets}: EXCEPTION {type(exc).__name__}: {str(exc)[:500]}', flush=True)
            break
