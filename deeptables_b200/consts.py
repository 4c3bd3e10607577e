"""Names shared across the host layer.

The values are part of the reference's user-visible surface (task names, stacking ops, layer and
input name prefixes that `apply()` and weight interchange key on -- reference
deeptables/utils/consts.py plus the task names it re-exports from hypernets), so they are kept
value-for-value; everything unrelated to the train/score hot path is left out.
"""

# learning tasks ------------------------------------------------------------------------------
TASK_AUTO, TASK_BINARY, TASK_MULTICLASS = 'auto', 'binary', 'multiclass'
TASK_REGRESSION, TASK_MULTILABEL = 'regression', 'multilabel'
ALL_TASKS = (TASK_BINARY, TASK_MULTICLASS, TASK_REGRESSION, TASK_MULTILABEL)

# how per-net logits are combined (deepmodel.py:296-301) ----------------------------------------
STACKING_OP_ADD, STACKING_OP_CONCAT = 'add', 'concat'

# layer / input naming ---------------------------------------------------------------------------
LAYER_PREFIX_EMBEDDING = 'emb_'
INPUT_PREFIX_CAT, INPUT_PREFIX_NUM, INPUT_PREFIX_SEQ = 'cat_', 'input_continuous_', 'seq_'
LAYER_NAME_CONCAT_CONT_INPUTS = 'concat_continuous_inputs'
LAYER_NAME_BN_DENSE_ALL = 'bn_dense_all'

# dtypes of the tensors crossing the host boundary -------------------------------------------------
DATATYPE_TENSOR_FLOAT, DATATYPE_PREDICT_CLASS = 'float32', 'int32'

# model selection / metrics --------------------------------------------------------------------------
MODEL_SELECT_MODE_MIN, MODEL_SELECT_MODE_MAX, MODEL_SELECT_MODE_AUTO = 'min', 'max', 'auto'
MODEL_SELECTOR_BEST, MODEL_SELECTOR_CURRENT, MODEL_SELECTOR_ALL = 'best', 'current', 'all'
METRIC_NAME_AUC, METRIC_NAME_ACCURACY, METRIC_NAME_MSE = 'AUC', 'accuracy', 'mse'

# misc ModelConfig defaults ----------------------------------------------------------------------------
EMBEDDING_OUT_DIM_DEFAULT = 4
GBM_FEATURE_TYPE_EMB, GBM_FEATURE_TYPE_DENSE = 'embedding', 'dense'
GPU_USAGE_STRATEGY_GROWTH = 'memory_growth'
ENV_DEEPTABLES_HOME = 'DEEPTABLES_HOME'
PROJECT_NAME = 'deeptables'
