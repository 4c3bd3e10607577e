"""Debug aid: find the layer / autograd node of an autoint model after which a CUDA-graph capture is no longer valid."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeptables_b200 import deeptable, layers as L, engine as E  # noqa: E402
from deeptables_b200.deepmodel import DeepModel  # noqa: E402
from deeptables_b200.metainfo import CategoricalColumn, ContinuousColumn  # noqa: E402

vocab, dim, n_cont, b = [11, 7, 13, 5, 9], 8, 3, 256
conf = deeptable.ModelConfig(nets=['autoint_nets'], embeddings_output_dim=dim, embedding_dropout=0, metrics=['AUC'],
                             autoint_params={'num_attention': 1, 'num_heads': 2, 'dropout_rate': 0, 'use_residual': True})
cats = [CategoricalColumn(f'c{i}', v, dim) for i, v in enumerate(vocab)]
conts = [ContinuousColumn('input_continuous_all', [f'n{i}' for i in range(n_cont)])]
model = DeepModel('binary', 2, conf, cats, conts, seed=3)
model._build_model()
g = np.random.default_rng(0)
idx = torch.tensor(np.stack([g.integers(0, v, size=b) for v in vocab], axis=1).astype(np.int32)).cuda()
cont = torch.tensor(g.normal(size=(b, n_cont)).astype(np.float32)).cuda()
y = torch.tensor((g.random(b) < 0.4).astype(np.float32)).cuda().view(-1, 1)
model.train_step(idx, cont, y)          # eager step: everything allocated / loaded


def ok():
    try:
        torch.cuda.is_current_stream_capturing()
        return True
    except Exception as exc:
        return f'{type(exc).__name__}: {str(exc)[:120]}'


orig_call = L.Layer.__call__


def spy_call(self, *a, **k):
    out = orig_call(self, *a, **k)
    st = ok()
    if st is not True:
        print(f'capture INVALID after forward of layer {self.name} ({type(self).__name__}): {st}', flush=True)
    return out


L.Layer.__call__ = spy_call
for cls in (E.GatherFn, E.DenseFn, E.AttentionCoreFn, E.BatchNormFn, E.ConcatEmbDenseFn):
    orig_b = cls.backward

    def make(orig, name):
        def spy(ctx, *g_):
            st0 = ok()
            out = orig(ctx, *g_)
            st = ok()
            print(f'backward {name}: before {st0 is True}, after {st is True}', flush=True)
            return out
        return staticmethod(spy)
    cls.backward = make(orig_b, cls.__name__)

model._step_dev = torch.zeros(1, dtype=torch.int32, device='cuda')
model._alpha_table(200000)
model._step_dev.fill_(model._step)
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(gr):
        print('capturing: start ok', ok(), flush=True)
        model._train_step_body(idx.clone(), cont.clone(), y.clone(), None, dev_step=True)
        print('capturing: end ok', ok(), flush=True)
    print('capture finished fine')
except Exception as exc:
    print('capture raised', type(exc).__name__, str(exc)[:200])
