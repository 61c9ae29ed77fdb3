"""Model architectures on the B200 hot path."""
from .mlp import MLP, DropoutMLP
from .regression import LinearRegressionModel, LogisticRegressionModel
from .som import SOM
from .rbf import RBF
