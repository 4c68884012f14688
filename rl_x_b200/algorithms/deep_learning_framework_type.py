"""What a plugin's models are built on; the runner only reads it to decide which global seeds to set.  Member names follow the reference
(rl_x/algorithms/deep_learning_framework_type.py) because properties are compared by member name across packages."""
import enum

DeepLearningFrameworkType = enum.Enum("DeepLearningFrameworkType", ["TORCH", "JAX"], start=0)
