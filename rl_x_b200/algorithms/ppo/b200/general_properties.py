from rl_x_b200.environments.types import ActionSpaceType, ObservationSpaceType, DataInterfaceType
from rl_x_b200.algorithms.deep_learning_framework_type import DeepLearningFrameworkType


class GeneralProperties:
    """What this plugin supports; the runner's compatibility check rejects everything else instead of silently falling
    back (reference check: runner.py:86-91; the reference plugin also lists IMAGES / DISCRETE, general_properties.py:7-12)."""
    observation_space_types = [ObservationSpaceType.FLAT_VALUES]
    action_space_types = [ActionSpaceType.CONTINUOUS]
    data_interface_types = [DataInterfaceType.NUMPY, DataInterfaceType.TORCH]

    deep_learning_framework_type = DeepLearningFrameworkType.TORCH
