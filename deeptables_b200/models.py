"""Alias so reference-style imports keep working: ``from deeptables_b200.models import deeptable, deepnets``."""
from . import deeptable, deepnets, deepmodel, layers, config, metainfo   # noqa: F401
from .config import ModelConfig                                           # noqa: F401
