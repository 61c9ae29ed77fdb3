"""learning_b200 -- B200-native drop-in for the hot path of ``learning`` (optimal-py-learning).

Same public names as learning/__init__.py:26-45 for everything on the MLP objective/gradient +
quasi-Newton path: transfers, error functions, Model, MLP and the ``optimize`` plugins.  All
arithmetic runs in hand-written sm_100a CUDA kernels behind the C-ABI of include/opl_b200.h;
importing this package fails loudly if that library cannot be loaded (no CPU fallback).
"""
from . import _native  # noqa: F401  (loads / builds libopl_b200.so or raises ImportError)

from .transfer import (Transfer, LinearTransfer, TanhTransfer, ReluTransfer, GaussianTransfer,
                       SoftmaxTransfer)
from .error import ErrorFunc, MeanSquaredError, CrossEntropyError, L1Penalty, L2Penalty
from .base import Model
from . import optimize
from . import validation
from .architecture.mlp import MLP, DropoutMLP
from .architecture.regression import LinearRegressionModel, LogisticRegressionModel
from .architecture.som import SOM
from .architecture.rbf import RBF

__all__ = ['Transfer', 'LinearTransfer', 'TanhTransfer', 'ReluTransfer', 'GaussianTransfer',
           'SoftmaxTransfer', 'ErrorFunc', 'MeanSquaredError', 'CrossEntropyError', 'L1Penalty', 'L2Penalty', 'Model',
           'MLP', 'DropoutMLP', 'LinearRegressionModel', 'LogisticRegressionModel', 'SOM', 'RBF',
           'optimize', 'validation']
