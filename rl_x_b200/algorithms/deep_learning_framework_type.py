"""What a plugin's models are built on; the runner only reads it to decide which global seeds to set.  Member names follow the reference
(rl_x/algorithms/deep_learning_framework_type.py); members equal the reference's same-named members (NamedEnum), which is what its runner
compares them with (runner.py:100-101)."""
from rl_x_b200.environments.types import NamedEnum


class DeepLearningFrameworkType(NamedEnum):
    TORCH = 0
    JAX = 1
