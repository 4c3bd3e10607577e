"""Replay of a TensorFlow dump of the UNMODIFIED reference (tools/dump_tf_reference.py) through the CPU oracle and,
with `-m gpu`, through the CUDA engine: forward in inference and training mode, the loss of one train_on_batch and every
weight after the Adam step.  Skipped unless DTB_TF_DUMP points at a dump -- TensorFlow is not installable in this
repository's environment, so the dump has to come from elsewhere; the weight names it carries are the
reference's own (Keras 3 paths), which is what `state_dict()` / `load_state_dict()` of both the oracle and the engine use."""
import json
import os

import numpy as np
import pytest
import torch

DUMP = os.environ.get('DTB_TF_DUMP')
pytestmark = pytest.mark.skipif(not DUMP or not os.path.exists(DUMP or ''), reason='no TensorFlow dump (set DTB_TF_DUMP)')


def _cases():
    if not DUMP or not os.path.exists(DUMP):
        return {}, None
    z = np.load(DUMP)
    return json.loads(bytes(z['__meta__']).decode()), z


META, Z = _cases()


def _conf(spec):
    with open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_modelconfig.json')) as f:
        conf = dict(json.load(f)['defaults'])
    conf.update(nets=spec['nets'], embeddings_output_dim=spec['dim'], embedding_dropout=0, dense_dropout=0)
    conf.update(spec['kw'])
    return conf


def _weights(case, which):
    pre = f'{case}/{which}/'
    return {k[len(pre):]: torch.tensor(Z[k]) for k in Z.files if k.startswith(pre)}


@pytest.mark.parametrize('case', sorted(META))
def test_oracle_reproduces_tensorflow_step(case):
    from oracle import model_ref as M
    spec = META[case]
    conf = _conf(spec)
    state = {k: v.double() for k, v in _weights(case, 'w0').items()}
    ids = torch.tensor(Z[f'{case}/ids'].astype(np.int64))
    cont = torch.tensor(Z[f'{case}/cont']).double()
    f = len(spec['vocab'])
    for training, key in ((False, 'out_infer'), (True, 'out_train')):
        got, _ = M.forward(state, conf, ids, cont, f, training, task=spec['task'])
        np.testing.assert_allclose(got.numpy(), Z[f'{case}/{key}'], rtol=1e-4, atol=1e-6, err_msg=f'{case} {key}')
    tr = M.RefTrainer(state, conf, f, task=spec['task'], dtype=torch.float64)
    loss = tr.train_step(ids, cont, torch.tensor(Z[f'{case}/y']).double())
    np.testing.assert_allclose(float(loss), float(Z[f'{case}/loss'].reshape(-1)[0]), rtol=1e-4)
    for k, v in _weights(case, 'w1').items():
        np.testing.assert_allclose(tr.state[k].numpy(), v.numpy(), rtol=1e-3, atol=1e-6, err_msg=f'{case} after Adam: {k}')


@pytest.mark.gpu
@pytest.mark.parametrize('case', sorted(META))
def test_engine_reproduces_tensorflow_step(case):
    from deeptables_b200 import deeptable
    from deeptables_b200.deepmodel import DeepModel
    from deeptables_b200.metainfo import CategoricalColumn, ContinuousColumn
    spec = META[case]
    conf = deeptable.ModelConfig(nets=spec['nets'], embeddings_output_dim=spec['dim'], embedding_dropout=0, dense_dropout=0,
                                 **spec['kw'])
    cats = [CategoricalColumn(f'c{i}', v, spec['dim']) for i, v in enumerate(spec['vocab'])]
    conts = [ContinuousColumn('input_continuous_all', [f'n{i}' for i in range(spec['n_cont'])])]
    model = DeepModel(spec['task'], 2, conf, cats, conts, seed=1)
    model._build_model()
    model.load_state_dict(_weights(case, 'w0'), strict=True)
    ids = torch.tensor(Z[f'{case}/ids']).cuda()
    cont = torch.tensor(Z[f'{case}/cont']).cuda()
    got = model.predict_step(ids, cont).cpu().numpy()
    np.testing.assert_allclose(got, Z[f'{case}/out_infer'], rtol=1e-3, atol=1e-5)
    loss = model.train_on_batch(Z[f'{case}/ids'], Z[f'{case}/cont'], Z[f'{case}/y'])
    np.testing.assert_allclose(loss, float(Z[f'{case}/loss'].reshape(-1)[0]), rtol=1e-3)
    sd = model.state_dict()
    for k, v in _weights(case, 'w1').items():
        np.testing.assert_allclose(sd[k].cpu().numpy(), v.numpy(), rtol=1e-3, atol=1e-5, err_msg=f'{case} after Adam: {k}')
