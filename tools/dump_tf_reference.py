#!/usr/bin/env python
"""Close the unpinned half of parity on a machine that HAS TensorFlow (this image does not: SURVEY.md 8c).

Runs the UNMODIFIED reference (`pip install deeptables` or PYTHONPATH=/path/to/DeepTables) on real TensorFlow / Keras 3:
builds `DeepModel.__build_model` for a few configurations, takes ONE `train_on_batch` step (Adam 1e-3, the task's 'auto'
loss) and dumps everything the replay test needs to `.npz`:

    <case>/ids, cont, y              the batch (categorical ids int32 [B,F], continuous float32 [B,C], labels)
    <case>/w0/<weight name>          every model weight before the step   (Keras weight paths, e.g.
    <case>/w1/<weight name>          ... and after it                      emb_categorical_vars_all/embeddings_0)
    <case>/out_infer, out_train      model(x, training=False / True) before the step
    <case>/loss                      the value train_on_batch returned

    python tools/dump_tf_reference.py --out tf_reference_dump.npz
    DTB_TF_DUMP=tf_reference_dump.npz python -m pytest tests/test_tf_replay.py        # CPU oracle; with -m gpu the engine too

That pins what tests/golden/ cannot without TensorFlow: the Keras losses, the Adam update (incl. eps placement and bias
correction), the BatchNormalization moving-statistics update and TensorFlow's own op arithmetic.
NOT executed in this repository's environment (no TensorFlow): the script follows the public Keras 3 API only."""
import argparse
import json

import numpy as np

CASES = {
    'xdeepfm': dict(nets=['linear', 'cin_nets', 'dnn_nets'], dim=16, vocab=[50, 40, 30, 20, 10, 60], n_cont=5, task='binary',
                    kw={'cin_params': {'cross_layer_size': (32, 32, 16), 'activation': 'relu', 'use_residual': False,
                                       'use_bias': False, 'direct': False, 'reduce_D': False}}),
    'deepfm': dict(nets=['linear', 'fm_nets', 'dnn_nets'], dim=8, vocab=[30, 20, 10, 40], n_cont=3, task='binary', kw={}),
    'dcn_autoint': dict(nets=['dcn_nets', 'autoint_nets'], dim=8, vocab=[30, 20, 10, 40, 25], n_cont=4, task='binary',
                        kw={'cross_params': {'num_cross_layer': 3},
                            'autoint_params': {'num_attention': 2, 'num_heads': 2, 'dropout_rate': 0, 'use_residual': True}}),
    'pnn_regression': dict(nets=['pnn_nets'], dim=4, vocab=[12, 9, 7], n_cont=2, task='regression', kw={}),
}


def weight_dict(model):
    out = {}
    for w in model.weights:
        name = getattr(w, 'path', None) or w.name           # Keras 3: `path` is layer/weight
        out[name.replace(':0', '')] = np.asarray(w)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='tf_reference_dump.npz')
    ap.add_argument('--batch', type=int, default=64)
    args = ap.parse_args()
    import tensorflow as tf                                    # noqa: F401  (the reference imports it itself)
    from deeptables.models import deeptable, deepmodel
    from deeptables.utils import consts
    from deeptables.models.metainfo import CategoricalColumn, ContinuousColumn
    rec, meta = {}, {}
    for case, spec in CASES.items():
        rng = np.random.default_rng(abs(hash(case)) % (2 ** 31))
        conf = deeptable.ModelConfig(nets=spec['nets'], embeddings_output_dim=spec['dim'], embedding_dropout=0,
                                     dense_dropout=0, **spec['kw'])
        cats = [CategoricalColumn(f'c{i}', v, spec['dim']) for i, v in enumerate(spec['vocab'])]
        conts = [ContinuousColumn('input_continuous_all', [f'n{i}' for i in range(spec['n_cont'])])]
        dm = deepmodel.DeepModel(spec['task'], 2, conf, cats, conts)
        model = dm._DeepModel__build_model(spec['task'], 2, conf.nets, cats, conts, None, conf)
        b = args.batch
        ids = np.stack([rng.integers(0, v, size=b) for v in spec['vocab']], axis=1).astype(np.int32)
        cont = rng.normal(size=(b, spec['n_cont'])).astype(np.float32)
        y = (rng.random(b) < 0.4).astype(np.float32) if spec['task'] == 'binary' else rng.normal(size=b).astype(np.float32)
        x = {'all_categorical_vars': ids.astype(np.float32), 'input_continuous_all': cont}   # ids travel as float32
        rec[f'{case}/ids'], rec[f'{case}/cont'], rec[f'{case}/y'] = ids, cont, y
        rec[f'{case}/out_infer'] = np.asarray(model(x, training=False))
        rec[f'{case}/out_train'] = np.asarray(model(x, training=True))
        for k, v in weight_dict(model).items():
            rec[f'{case}/w0/{k}'] = v
        rec[f'{case}/loss'] = np.asarray(model.train_on_batch(x, y.reshape(-1, 1)))
        for k, v in weight_dict(model).items():
            rec[f'{case}/w1/{k}'] = v
        meta[case] = {k: (v if k != 'kw' else {a: {c: list(d) if isinstance(d, tuple) else d for c, d in bb.items()}
                                               for a, bb in v.items()}) for k, v in spec.items()}
    rec['__meta__'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(args.out, **rec)
    print(f'wrote {args.out}: {len(rec)} arrays, cases {sorted(CASES)}')


if __name__ == '__main__':
    main()
