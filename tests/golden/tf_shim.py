"""Stand-in for the TensorFlow / Keras 3 / hypernets names that the reference's hot-path modules import, so
that the reference's OWN code (deeptables/models/layers.py, deepnets.py, deepmodel.py, config.py) can be
executed in this container, where TensorFlow cannot be installed.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_reference_golden.py, here, to generate golden vectors;
nothing under deeptables_b200/ and no test run imports it (the fixtures are committed; /root/reference does
not exist on the GPU box).

What is real and what is restated:
  * REAL, unmodified, imported from /root/reference: every layer class (`FM`, `CIN`, `Cross`,
    `MultiheadAttention`, `InnerProduct`, `OuterProduct`, `MultiColumnEmbedding`), every net builder of
    `deepnets.py`, `DeepModel.__build_model / __concat_emb_dense / __output_layer`, `ModelConfig`.
  * RESTATED here from their public documentation (float64, eager, torch-CPU): the ~25 TensorFlow primitives
    those modules call (`tf.matmul`, `tf.split`, `tf.nn.conv1d`, ...) and the stock Keras layers they
    instantiate (`Dense`, `BatchNormalization`, `Activation`, `Concatenate`, `Flatten`, `Add`, `Dropout`).
    Each is a few lines; see the docstrings.  The optimiser and the losses are NOT exercised through this shim
    (Keras runs them inside `Model.fit`); those stay restated in oracle/model_ref.py.

Tensors are plain ``torch.Tensor`` (float64 / int64) with ``get_shape()`` patched on.
"""
import contextlib
import logging
import sys
import types

import torch

DTYPE = torch.float64
_STATE = {'training': False, 'feeds': {}, 'layers': [], 'names': {}, 'gen': None}


# ------------------------------------------------------------------------------------------------
# tensors
# ------------------------------------------------------------------------------------------------
class TensorShape(tuple):
    def as_list(self):
        return list(self)


def _as_tensor(x):
    """tf.convert_to_tensor: python lists of tensors are stacked on a new leading axis."""
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (list, tuple)) and x and isinstance(x[0], torch.Tensor):
        return torch.stack(list(x), dim=0)
    return torch.as_tensor(x, dtype=DTYPE)


def _axes(axis):
    if axis is None:
        return None
    return tuple(axis) if isinstance(axis, (list, tuple)) else (axis,)


# ------------------------------------------------------------------------------------------------
# tensorflow primitives (eager)
# ------------------------------------------------------------------------------------------------
def tf_reduce_sum(x, axis=None, keepdims=False, name=None):
    x = _as_tensor(x)
    return x.sum() if axis is None else x.sum(dim=_axes(axis), keepdim=keepdims)


def tf_split(value, num_or_size_splits, axis=0, name=None):
    """tf.split: an int -> that many equal pieces; a list -> pieces of those sizes."""
    value = _as_tensor(value)
    if isinstance(num_or_size_splits, int):
        assert value.shape[axis] % num_or_size_splits == 0
        return list(torch.split(value, value.shape[axis] // num_or_size_splits, dim=axis))
    return list(torch.split(value, list(num_or_size_splits), dim=axis))


def tf_matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    a, b = _as_tensor(a), _as_tensor(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return torch.matmul(a, b)


def tf_tensordot(a, b, axes, name=None):
    a, b = _as_tensor(a), _as_tensor(b)
    if isinstance(axes, int):
        return torch.tensordot(a, b, dims=axes)
    ax_a, ax_b = axes
    ax_a = list(ax_a) if isinstance(ax_a, (list, tuple)) else [ax_a]
    ax_b = list(ax_b) if isinstance(ax_b, (list, tuple)) else [ax_b]
    return torch.tensordot(a, b, dims=(ax_a, ax_b))


def tf_conv1d(input, filters, stride=1, padding='VALID', data_format='NWC', name=None):
    """tf.nn.conv1d, NWC: input [batch, width, in_ch], filters [filter_width, in_ch, out_ch]."""
    assert data_format == 'NWC' and padding == 'VALID' and stride == 1
    x = _as_tensor(input).permute(0, 2, 1)                 # [batch, in_ch, width]
    w = _as_tensor(filters).permute(2, 1, 0)               # [out_ch, in_ch, filter_width]
    return torch.nn.functional.conv1d(x, w).permute(0, 2, 1)


def tf_bias_add(value, bias, name=None):
    return _as_tensor(value) + _as_tensor(bias)


def _make_tf():
    tf = types.ModuleType('tensorflow')
    tf.float32, tf.float64, tf.int32, tf.int64 = 'float32', 'float64', 'int32', 'int64'
    tf.square = lambda x, name=None: _as_tensor(x) ** 2
    tf.reduce_sum = tf_reduce_sum
    tf.reduce_mean = lambda x, axis=None, keepdims=False, name=None: \
        _as_tensor(x).mean() if axis is None else _as_tensor(x).mean(dim=_axes(axis), keepdim=keepdims)
    tf.reduce_max = lambda x, axis=None, keepdims=False, name=None: \
        _as_tensor(x).max() if axis is None else _as_tensor(x).amax(dim=_axes(axis), keepdim=keepdims)
    tf.concat = lambda values, axis, name=None: torch.cat([_as_tensor(v) for v in values], dim=axis)
    tf.split = tf_split
    tf.transpose = lambda a, perm=None, name=None: _as_tensor(a).permute(*perm) if perm is not None else _as_tensor(a).t()
    tf.matmul = tf_matmul
    tf.reshape = lambda tensor, shape, name=None: _as_tensor(tensor).reshape(*[int(s) for s in shape])
    tf.expand_dims = lambda input, axis, name=None: _as_tensor(input).unsqueeze(axis)
    tf.multiply = lambda x, y, name=None: _as_tensor(x) * _as_tensor(y)
    tf.tensordot = tf_tensordot
    tf.where = lambda cond, x, y, name=None: torch.where(cond, _as_tensor(x), _as_tensor(y))
    tf.equal = lambda x, y, name=None: _as_tensor(x) == y
    tf.ones_like = lambda x, name=None: torch.ones_like(_as_tensor(x))
    tf.zeros_like = lambda x, name=None: torch.zeros_like(_as_tensor(x))
    nn = types.ModuleType('tensorflow.nn')
    nn.softmax = lambda logits, axis=-1, name=None: torch.softmax(_as_tensor(logits), dim=axis)
    nn.relu = lambda features, name=None: torch.relu(_as_tensor(features))
    nn.conv1d = tf_conv1d
    nn.bias_add = tf_bias_add
    tf.nn = nn
    tf.keras = types.ModuleType('tensorflow.keras')
    return tf


# ------------------------------------------------------------------------------------------------
# keras: Layer protocol + the stock layers the reference instantiates
# ------------------------------------------------------------------------------------------------
def _unique_name(base):
    n = _STATE['names'].get(base, 0)
    _STATE['names'][base] = n + 1
    return base if n == 0 else f'{base}_{n}'


def _snake(name):
    out = []
    for i, ch in enumerate(name):
        if ch.isupper() and i and (not name[i - 1].isupper() or (i + 1 < len(name) and name[i + 1].islower())):
            out.append('_')
        out.append(ch.lower())
    return ''.join(out)


def _shape_of(x):
    if isinstance(x, (list, tuple)):
        return [_shape_of(v) for v in x]
    return tuple(x.shape)


class Layer:
    """keras.layers.Layer, eager: the first __call__ builds with the input shape(s), then runs call()."""

    def __init__(self, name=None, dtype=None, trainable=True, **kwargs):
        self.name = name if name is not None else _unique_name(_snake(type(self).__name__))
        self.built = False
        self.weights_by_name = {}
        self.trainable = trainable
        _STATE['layers'].append(self)

    def add_weight(self, name=None, shape=None, dtype=None, initializer=None, trainable=True, regularizer=None,
                   constraint=None, **kwargs):
        shape = tuple(int(s) for s in shape)
        w = torch.randn(shape, generator=_STATE['gen'], dtype=DTYPE) * 0.3     # values are re-drawn by the generator script
        self.weights_by_name[name] = w
        return w

    def build(self, input_shape):
        self.built = True

    def __call__(self, inputs, *args, **kwargs):
        if not self.built:
            self.build(_shape_of(inputs))
            self.built = True
        kwargs.pop('training', None)
        return self.call(inputs, *args, **kwargs)

    def call(self, inputs, **kwargs):
        return inputs

    def get_config(self):
        return {'name': self.name}


_ACTIVATIONS = {None: lambda x: x, 'linear': lambda x: x, 'relu': torch.relu, 'sigmoid': torch.sigmoid,
                'tanh': torch.tanh, 'softmax': lambda x: torch.softmax(x, dim=-1)}


class Dense(Layer):
    """keras Dense: activation(inputs @ kernel + bias), kernel [in, units]."""

    def __init__(self, units, activation=None, use_bias=True, kernel_initializer='glorot_uniform', **kwargs):
        for k in ('kernel_regularizer', 'activity_regularizer', 'bias_initializer', 'bias_regularizer'):
            kwargs.pop(k, None)
        super().__init__(**kwargs)
        self.units, self.activation, self.use_bias = int(units), activation, use_bias

    def build(self, input_shape):
        self.kernel = self.add_weight('kernel', (input_shape[-1], self.units))
        self.bias = self.add_weight('bias', (self.units,)) if self.use_bias else None

    def call(self, x, **kwargs):
        y = torch.matmul(x, self.weights_by_name['kernel'])
        if self.use_bias:
            y = y + self.weights_by_name['bias']
        return _ACTIVATIONS[self.activation](y)


class BatchNormalization(Layer):
    """keras BatchNormalization(axis=-1, momentum=0.99, epsilon=1e-3): inference uses the moving statistics;
    training uses the batch mean and the biased batch variance over every axis but the last."""

    def __init__(self, axis=-1, momentum=0.99, epsilon=1e-3, **kwargs):
        super().__init__(**kwargs)
        assert axis == -1
        self.momentum, self.epsilon = momentum, epsilon

    def build(self, input_shape):
        c = input_shape[-1]
        for n in ('gamma', 'beta', 'moving_mean', 'moving_variance'):
            self.add_weight(n, (c,))
        self.weights_by_name['moving_variance'] = self.weights_by_name['moving_variance'].abs() + 0.5
        self.weights_by_name['gamma'] = self.weights_by_name['gamma'] + 1.0

    def call(self, x, **kwargs):
        w = self.weights_by_name
        if _STATE['training']:
            dims = tuple(range(x.dim() - 1))
            mean = x.mean(dim=dims)
            var = ((x - mean) ** 2).mean(dim=dims)
        else:
            mean, var = w['moving_mean'], w['moving_variance']
        return (x - mean) / torch.sqrt(var + self.epsilon) * w['gamma'] + w['beta']


class Activation(Layer):
    def __init__(self, activation, **kwargs):
        super().__init__(**kwargs)
        self.activation = activation

    def call(self, x, **kwargs):
        return _ACTIVATIONS[self.activation](x)


class Dropout(Layer):
    """Inference behaviour (identity); the golden vectors are generated with every dropout rate at 0."""

    def __init__(self, rate=0.0, **kwargs):
        super().__init__(**kwargs)
        self.rate = rate
        assert not rate or not _STATE['training'], 'golden vectors are generated with dropout 0'


class SpatialDropout1D(Dropout):
    pass


class Concatenate(Layer):
    def __init__(self, axis=-1, **kwargs):
        super().__init__(**kwargs)
        self.axis = axis

    def call(self, xs, **kwargs):
        return torch.cat(list(xs), dim=self.axis)


class Flatten(Layer):
    def call(self, x, **kwargs):
        return x.reshape(x.shape[0], -1)


class Add(Layer):
    def call(self, xs, **kwargs):
        out = xs[0]
        for v in xs[1:]:
            out = out + v
        return out


def _same_pad(size, k, stride):
    """TensorFlow 'SAME' padding along one axis: (before, after)."""
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return total // 2, total - total // 2


class Conv2D(Layer):
    """keras Conv2D, channels_last, stride 1: kernel [kh, kw, in, filters]; 'same' = TensorFlow's asymmetric padding."""

    def __init__(self, filters, kernel_size, strides=(1, 1), padding='valid', activation=None, use_bias=True,
                 kernel_initializer='glorot_uniform', **kwargs):
        super().__init__(**kwargs)
        assert tuple(strides) == (1, 1)
        self.filters, self.kernel_size, self.padding = int(filters), tuple(kernel_size), padding
        self.activation, self.use_bias = activation, use_bias

    def build(self, input_shape):
        kh, kw = self.kernel_size
        self.kernel = self.add_weight('kernel', (kh, kw, input_shape[-1], self.filters))
        self.bias = self.add_weight('bias', (self.filters,)) if self.use_bias else None

    def call(self, x, **kwargs):
        kh, kw = self.kernel_size
        xc = x.permute(0, 3, 1, 2)                                    # NHWC -> NCHW
        if self.padding == 'same':
            (t, b), (l, r) = _same_pad(x.shape[1], kh, 1), _same_pad(x.shape[2], kw, 1)
            xc = torch.nn.functional.pad(xc, (l, r, t, b))
        w = self.weights_by_name['kernel'].permute(3, 2, 0, 1)        # [filters, in, kh, kw]
        y = torch.nn.functional.conv2d(xc, w, self.weights_by_name['bias'] if self.use_bias else None)
        return _ACTIVATIONS[self.activation](y.permute(0, 2, 3, 1))


class MaxPooling2D(Layer):
    """keras MaxPooling2D, channels_last, strides = pool_size; 'same' pads with -inf the TensorFlow way."""

    def __init__(self, pool_size=(2, 2), strides=None, padding='valid', **kwargs):
        super().__init__(**kwargs)
        self.pool_size, self.padding = tuple(pool_size), padding
        self.strides = tuple(strides) if strides is not None else self.pool_size

    def call(self, x, **kwargs):
        (ph, pw), (sh, sw) = self.pool_size, self.strides
        xc = x.permute(0, 3, 1, 2)
        if self.padding == 'same':
            (t, b), (l, r) = _same_pad(x.shape[1], ph, sh), _same_pad(x.shape[2], pw, sw)
            xc = torch.nn.functional.pad(xc, (l, r, t, b), value=float('-inf'))
        return torch.nn.functional.max_pool2d(xc, (ph, pw), (sh, sw)).permute(0, 2, 3, 1)


class _LossBase:
    """keras.losses.Loss: the reference's focal losses only use the constructor and call()."""

    def __init__(self, reduction=None, name=None, **kwargs):
        self.reduction, self.name = reduction, name

    def get_config(self):
        return {'name': self.name}


class _Unused(Layer):
    def __init__(self, *a, **k):
        raise NotImplementedError(f'{type(self).__name__} is not on the hot path; the shim does not restate it')


def Input(shape=None, name=None, dtype=None, **kwargs):
    """Eager stand-in: returns the concrete batch registered under this input name (feed(name, tensor))."""
    return _STATE['feeds'][name]


class Model:
    def __init__(self, inputs=None, outputs=None, **kwargs):
        self.inputs, self.output = inputs, outputs

    def compile(self, *a, **k):
        pass


# ------------------------------------------------------------------------------------------------
# install
# ------------------------------------------------------------------------------------------------
def _module(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Anything:
    """Placeholder object for names that are imported but never used on the hot path."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, item):
        return _Anything()


def install(reference_root='/root/reference'):
    """Register the stand-in modules and the `deeptables` package skeleton (package __init__ files are NOT
    executed: they import the whole framework; the hot-path modules are loaded from their real files)."""
    import os
    torch.Tensor.get_shape = lambda self: TensorShape(self.shape)
    _STATE['gen'] = torch.Generator().manual_seed(20260923)

    tf = _make_tf()
    sys.modules['tensorflow'] = tf
    py = _module('tensorflow.python')
    _module('tensorflow.python.eager')
    ctx = _module('tensorflow.python.eager.context', executing_eagerly=lambda: True,
                  context=lambda: types.SimpleNamespace(num_gpus=lambda: 0))
    _module('tensorflow.python.framework')
    fops = _module('tensorflow.python.framework.ops', device=lambda name: contextlib.nullcontext())
    _module('tensorflow.python.keras')
    _module('tensorflow.python.keras.utils')
    _module('tensorflow.python.keras.utils.tf_utils')
    _module('tensorflow.python.keras.utils.generic_utils', deserialize_keras_object=_deserialize,
            serialize_keras_object=lambda obj: getattr(obj, '__name__', str(obj)))
    _module('tensorflow.python.ops')
    _module('tensorflow.python.ops.embedding_ops',
            embedding_lookup=lambda params, ids, **k: _as_tensor(params)[_as_tensor(ids).long()])
    _module('tensorflow.python.ops.math_ops')
    tf.python = py

    layer_names = dict(Layer=Layer, Dense=Dense, Dropout=Dropout, BatchNormalization=BatchNormalization,
                       Activation=Activation, Concatenate=Concatenate, Flatten=Flatten, Input=Input, Add=Add,
                       SpatialDropout1D=SpatialDropout1D, Conv2D=Conv2D, MaxPooling2D=MaxPooling2D)
    for unused in ('Embedding', 'Lambda'):
        layer_names[unused] = type(unused, (_Unused,), {})
    keras = _module('keras')
    _module('keras.api')
    keras.layers = _module('keras.api.layers', **layer_names)
    tf.keras.layers = keras.layers                # layers.AFM builds tf.keras.layers.Dropout (layers.py:788)
    _module('keras.api.metrics', RootMeanSquaredError=_Anything)
    _module('keras.api.models', Model=Model, load_model=_Anything(), save_model=_Anything())
    _module('keras.src')
    _module('keras.src.legacy')
    _module('keras.src.legacy.losses', Reduction=_Anything())
    keras.backend = _module('keras.backend', floatx=lambda: 'float32', epsilon=lambda: 1e-7)
    ops = _module('keras.ops')
    ops.ndim = lambda x: _as_tensor(x).dim()
    ops.sum = lambda x, axis=None, keepdims=False: tf_reduce_sum(x, axis, keepdims)
    ops.cast = lambda x, dtype: _as_tensor(x).to(torch.int64 if 'int' in str(dtype) else DTYPE)
    ops.not_equal = lambda a, b: _as_tensor(a) != b
    ops.expand_dims = lambda x, axis: _as_tensor(x).unsqueeze(axis)
    ops.clip = lambda x, lo, hi: torch.clamp(_as_tensor(x), lo, hi)
    ops.mean = lambda x, axis=None, keepdims=False: _as_tensor(x).mean() if axis is None else _as_tensor(x).mean(dim=_axes(axis), keepdim=keepdims)
    ops.power = lambda x, y: torch.pow(_as_tensor(x), y)
    ops.log = lambda x: torch.log(_as_tensor(x))
    ops.split = lambda x, indices_or_sections, axis=0: tf_split(x, indices_or_sections, axis)
    keras.ops = ops
    getter = types.SimpleNamespace(get=lambda ident: ident, serialize=lambda obj: obj)
    for sub in ('initializers', 'regularizers', 'constraints'):
        setattr(keras, sub, _module(f'keras.{sub}', get=getter.get, serialize=getter.serialize))
    keras.losses = _module('keras.losses', Loss=_LossBase, BinaryCrossentropy=_Anything, MeanSquaredError=_Anything,
                           CategoricalCrossentropy=_Anything)
    keras.optimizers = _module('keras.optimizers', Adam=_Anything)

    # hypernets: constants and logging only
    _module('hypernets')
    _module('hypernets.utils', logging=types.SimpleNamespace(get_logger=_get_logger, getLogger=_get_logger), fs=_Anything(),
            isnotebook=lambda: False)
    _module('hypernets.utils.const', TASK_AUTO='auto', TASK_BINARY='binary', TASK_MULTICLASS='multiclass',
            TASK_REGRESSION='regression', TASK_MULTILABEL='multilabel')
    _module('hypernets.tabular', get_tool_box=_Anything())

    # deeptables package skeleton over the real source tree
    pkg = _module('deeptables')
    pkg.__path__ = [os.path.join(reference_root, 'deeptables')]
    utils = _module('deeptables.utils', to_dataset=_Anything())
    utils.__path__ = [os.path.join(reference_root, 'deeptables', 'utils')]
    _module('deeptables.utils.gpu', set_memory_growth=lambda: None)
    utils.gpu = sys.modules['deeptables.utils.gpu']
    models = _module('deeptables.models')
    models.__path__ = [os.path.join(reference_root, 'deeptables', 'models')]
    return tf, keras


def _get_logger(name):
    logger = logging.getLogger(name)
    logger.is_info_enabled = lambda: False
    return logger


def _deserialize(identifier, module_objects=None, custom_objects=None, printable_module_name='object'):
    if custom_objects and identifier in custom_objects:
        return custom_objects[identifier]
    return (module_objects or {}).get(identifier)


# ------------------------------------------------------------------------------------------------
# controls used by the generator script
# ------------------------------------------------------------------------------------------------
def set_training(flag):
    _STATE['training'] = bool(flag)


def feed(name, tensor):
    _STATE['feeds'][name] = tensor


def reset_layers():
    _STATE['layers'].clear()
    _STATE['names'].clear()


def created_layers():
    return list(_STATE['layers'])


def seed(k):
    _STATE['gen'].manual_seed(int(k))
