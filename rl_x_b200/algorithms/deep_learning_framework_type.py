from enum import Enum


class DeepLearningFrameworkType(Enum):
    """Same members as rl_x/algorithms/deep_learning_framework_type.py:4-6."""
    TORCH = 0
    JAX = 1
