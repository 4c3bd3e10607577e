"""ModelConfig -- the reference's 45-field configuration record (deeptables/models/config.py:9-216),
field-for-field and default-for-default, so user code written against
``deeptable.ModelConfig(nets=..., cin_params=..., ...)`` drops in unchanged.

Only the fields that shape the train/score hot path are interpreted by this build (nets, *_params,
embeddings_output_dim, embedding/dense dropout, stacking_op, output_use_bias, optimizer, loss,
metrics, task, earlystopping_*, distribute_strategy); the preprocessing switches are carried for
API parity and consumed by the minimal preprocessor in deeptable.py.
"""
import collections
import copy
import os

from . import consts
from . import deepnets

# (field, default) in the reference's positional order
_FIELDS = (
    ('name', 'conf-1'),
    ('nets', ['dnn_nets']),
    ('categorical_columns', 'auto'),
    ('exclude_columns', []),
    ('task', consts.TASK_AUTO),
    ('pos_label', None),
    ('metrics', ['accuracy']),
    ('auto_categorize', False),
    ('cat_exponent', 0.5),
    ('cat_remain_numeric', True),
    ('auto_encode_label', True),
    ('auto_imputation', True),
    ('auto_scale', False),
    ('auto_discrete', False),
    ('auto_discard_unique', True),
    ('apply_gbm_features', False),
    ('gbm_params', {}),
    ('gbm_feature_type', consts.GBM_FEATURE_TYPE_EMB),
    ('fixed_embedding_dim', True),
    ('embeddings_output_dim', 4),
    ('embeddings_initializer', 'uniform'),
    ('embeddings_regularizer', None),
    ('embeddings_activity_regularizer', None),
    ('dense_dropout', 0),
    ('embedding_dropout', 0.3),
    ('stacking_op', consts.STACKING_OP_ADD),
    ('output_use_bias', True),
    ('apply_class_weight', False),
    ('optimizer', 'auto'),
    ('loss', 'auto'),
    ('dnn_params', {'hidden_units': ((128, 0, False), (64, 0, False)), 'activation': 'relu'}),
    ('autoint_params', {'num_attention': 3, 'num_heads': 1, 'dropout_rate': 0, 'use_residual': True}),
    ('fgcnn_params', {'fg_filters': (14, 16), 'fg_heights': (7, 7), 'fg_pool_heights': (2, 2),
                      'fg_new_feat_filters': (2, 2)}),
    ('fibinet_params', {'senet_pooling_op': 'mean', 'senet_reduction_ratio': 3,
                        'bilinear_type': 'field_interaction'}),
    ('cross_params', {'num_cross_layer': 4}),
    ('pnn_params', {'outer_product_kernel_type': 'mat'}),
    ('afm_params', {'attention_factor': 4, 'dropout_rate': 0}),
    ('cin_params', {'cross_layer_size': (128, 128), 'activation': 'relu', 'use_residual': False,
                    'use_bias': False, 'direct': False, 'reduce_D': False}),
    ('home_dir', None),
    ('monitor_metric', None),
    ('earlystopping_patience', 1),
    ('earlystopping_mode', 'auto'),
    ('gpu_usage_strategy', consts.GPU_USAGE_STRATEGY_GROWTH),
    ('distribute_strategy', None),
    ('var_len_categorical_columns', None),
)
_NAMES = tuple(n for n, _ in _FIELDS)
_Base = collections.namedtuple('ModelConfig', _NAMES)


class ModelConfig(_Base):
    __slots__ = ()

    def __hash__(self):
        return self.name.__hash__()

    def __new__(cls, *args, **kwargs):
        if len(args) > len(_NAMES):
            raise TypeError(f'ModelConfig takes at most {len(_NAMES)} positional arguments')
        values = {n: copy.deepcopy(d) for n, d in _FIELDS}
        for n, a in zip(_NAMES, args):
            values[n] = a
        for k, v in kwargs.items():
            if k not in values:
                raise TypeError(f'ModelConfig got an unexpected keyword argument {k!r}')
            if k in _NAMES[:len(args)]:
                raise TypeError(f'ModelConfig got multiple values for argument {k!r}')
            values[k] = v

        vl = values['var_len_categorical_columns']
        if vl is not None and len(vl) > 0:
            for v in vl:       # same checks as the reference (config.py:137-149)
                if not isinstance(v, (tuple, list)) or len(v) != 3:
                    raise ValueError('Var len column config should be a tuple 3.')
                if values['exclude_columns'] is not None and v[0] in values['exclude_columns']:
                    raise ValueError(f"Var len column {v[0]} can not put in 'exclude_columns' ")
                cc = values['categorical_columns']
                if cc is not None and isinstance(cc, list) and v[0] in cc:
                    raise ValueError(f"Var len column {v[0]} can not put in 'categorical_columns' ")

        values['nets'] = deepnets.get_nets(values['nets'])
        if values['home_dir'] is None and os.environ.get(consts.ENV_DEEPTABLES_HOME) is not None:
            values['home_dir'] = os.environ.get(consts.ENV_DEEPTABLES_HOME)
        return super().__new__(cls, *(values[n] for n in _NAMES))

    @property
    def first_metric_name(self):
        if self.metrics is None or len(self.metrics) <= 0:
            raise ValueError('`metrics` is none or empty.')
        first = self.metrics[0]
        if isinstance(first, str):
            return first
        if hasattr(first, 'name'):
            return first.name
        if callable(first):
            return first.__name__
        raise ValueError('`metric` must be string or callable object.')
