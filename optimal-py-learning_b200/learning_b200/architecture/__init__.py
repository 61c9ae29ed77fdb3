"""Model architectures on the B200 hot path."""
from .mlp import MLP
