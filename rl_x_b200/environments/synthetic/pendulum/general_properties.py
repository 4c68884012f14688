from rl_x_b200.environments.types import ActionSpaceType, ObservationSpaceType, DataInterfaceType, SimulationType


class GeneralProperties:
    """Same attribute set as the reference's gym env GeneralProperties (gym/mujoco/humanoid_v4/general_properties.py:7-12)."""
    observation_space_type = ObservationSpaceType.FLAT_VALUES
    action_space_type = ActionSpaceType.CONTINUOUS
    data_interface_type = DataInterfaceType.NUMPY

    simulation_type = SimulationType.DEFAULT
