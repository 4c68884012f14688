from rl_x_b200.environments.types import ActionSpaceType, ObservationSpaceType, DataInterfaceType, SimulationType


class GeneralProperties:
    """Same attribute set as the reference's env GeneralProperties (custom_mujoco/ant/warp_torch/general_properties.py:7-12)."""
    observation_space_type = ObservationSpaceType.FLAT_VALUES
    action_space_type = ActionSpaceType.CONTINUOUS
    data_interface_type = DataInterfaceType.TORCH

    simulation_type = SimulationType.WARP  # "torch tensors on the algorithm's device" class of simulators (runner.py:105)


class GeneralPropertiesNumpy(GeneralProperties):
    data_interface_type = DataInterfaceType.NUMPY
    simulation_type = SimulationType.DEFAULT
