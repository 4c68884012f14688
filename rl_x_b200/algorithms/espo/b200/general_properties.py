from rl_x_b200.environments.types import ActionSpaceType, ObservationSpaceType, DataInterfaceType
from rl_x_b200.algorithms.deep_learning_framework_type import DeepLearningFrameworkType


class GeneralProperties:
    """ref: rl_x/algorithms/espo/pytorch/general_properties.py:7-12 (NUMPY only there; the rollout code shared with the PPO plugin also
    serves TORCH-interface environments)."""
    observation_space_types = [ObservationSpaceType.FLAT_VALUES]
    action_space_types = [ActionSpaceType.CONTINUOUS]
    data_interface_types = [DataInterfaceType.NUMPY, DataInterfaceType.TORCH]

    deep_learning_framework_type = DeepLearningFrameworkType.TORCH
