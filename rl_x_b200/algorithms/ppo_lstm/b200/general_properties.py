from rl_x_b200.environments.types import ActionSpaceType, ObservationSpaceType, DataInterfaceType
from rl_x_b200.algorithms.deep_learning_framework_type import DeepLearningFrameworkType


class GeneralProperties:
    """ref: rl_x/algorithms/ppo_lstm/flax/general_properties.py (FLAT_VALUES x CONTINUOUS x NUMPY there; TORCH-interface envs work too)."""
    observation_space_types = [ObservationSpaceType.FLAT_VALUES]
    action_space_types = [ActionSpaceType.CONTINUOUS]
    data_interface_types = [DataInterfaceType.NUMPY, DataInterfaceType.TORCH]

    deep_learning_framework_type = DeepLearningFrameworkType.TORCH
