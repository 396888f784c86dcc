"""shadowing_b200 -- B200-native drop-in for the path-shadowing scan of RudyMorel/shadowing.

Exports the names the reference's `shadowing` package exposes for this path
(shadowing/__init__.py:1-4 star-imports): PathShadowing, PathEmbedding, Identity, Foveal,
PathDistance, RelativeMSE, ContextManagerBase, PredictionContext, ArrayType,
realized_variance, and (leaking through path_shadowing.py:9 in the reference) Softmax,
Uniform, DiscreteProba.
"""
from .averaging import DiscreteProba, Softmax, Uniform, softmax_weights
from .dataset import TimeSeriesDataset
from .path_distance import PathDistance, RelativeMSE
from .path_embedding import (ArrayType, ContextManagerBase, CrossChannelContext, Foveal, Identity, ImputationContext,
                             PathEmbedding, PredictionContext)
from .path_shadowing import PathShadowing, select_cartesian_product
from .statistics import RealizedVariance, realized_variance

__version__ = "0.1.0"

__all__ = [
    "PathShadowing", "PathEmbedding", "Identity", "Foveal", "PathDistance", "RelativeMSE",
    "ContextManagerBase", "PredictionContext", "ImputationContext", "CrossChannelContext", "ArrayType", "realized_variance",
    "RealizedVariance", "TimeSeriesDataset", "Softmax", "Uniform", "DiscreteProba", "softmax_weights",
    "select_cartesian_product",
]
