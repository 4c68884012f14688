from rl_x_b200.environments.types import ActionSpaceType, ObservationSpaceType, DataInterfaceType
from rl_x_b200.algorithms.deep_learning_framework_type import DeepLearningFrameworkType


class GeneralProperties:
    """ref: rl_x/algorithms/fastsac/pytorch/general_properties.py (device-resident TORCH-interface environments only)."""
    observation_space_types = [ObservationSpaceType.FLAT_VALUES]
    action_space_types = [ActionSpaceType.CONTINUOUS]
    data_interface_types = [DataInterfaceType.TORCH]

    deep_learning_framework_type = DeepLearningFrameworkType.TORCH
