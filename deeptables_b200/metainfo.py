"""Column schema the hot path is built from: which inputs are categorical ids (and how wide their
embedding is) and which are continuous.  Same three record types, field order and defaults as the
reference (deeptables/models/metainfo.py:33-86) so preprocessors written for it keep working."""
from collections import namedtuple

from . import consts


def _auto_dim(vocabulary_size, dim):
    # reference rule: a zero width means "fourth root of the vocabulary" (metainfo.py:44-45)
    return int(round(vocabulary_size ** 0.25)) if dim == 0 else dim


def _record(typename, fields):
    base = namedtuple(typename, fields)

    class _Hashable(base):
        __slots__ = ()

        def __hash__(self):            # columns are looked up by name
            return hash(self.name)

    _Hashable.__name__ = _Hashable.__qualname__ = typename
    return _Hashable


_Cat = _record('CategoricalColumn', 'name vocabulary_size embeddings_output_dim dtype input_name')
_VarLen = _record('VarLenCategoricalColumn', 'name vocabulary_size embeddings_output_dim dtype input_name sep')
_Cont = _record('ContinuousColumn', 'name column_names input_dim dtype input_name')


class CategoricalColumn(_Cat):
    __slots__ = ()

    def __new__(cls, name, vocabulary_size, embeddings_output_dim=10, dtype='int32', input_name=None):
        return super().__new__(cls, name, vocabulary_size, _auto_dim(vocabulary_size, embeddings_output_dim),
                               dtype, input_name or consts.INPUT_PREFIX_CAT + name)

    __hash__ = _Cat.__hash__


class VarLenCategoricalColumn(_VarLen):
    __slots__ = ()

    def __new__(cls, name, vocabulary_size, embeddings_output_dim=10, dtype='int32', input_name=None, sep='|'):
        return super().__new__(cls, name, vocabulary_size, _auto_dim(vocabulary_size, embeddings_output_dim),
                               dtype, input_name or consts.INPUT_PREFIX_CAT + name, sep)

    __hash__ = _VarLen.__hash__


class ContinuousColumn(_Cont):
    """A group of continuous columns fed as one (B, input_dim) float input."""
    __slots__ = ()

    def __new__(cls, name, column_names, input_dim=0, dtype='float32', input_name=None):
        return super().__new__(cls, name, column_names, len(column_names), dtype, input_name)

    __hash__ = _Cont.__hash__
