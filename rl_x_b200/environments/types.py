"""The four environment enums of the reference (rl_x/environments/{action_space,observation_space,data_interface,simulation}_type.py),
same member names and values so that GeneralProperties of reference plugins and of this package compare by name."""
from enum import Enum


class ActionSpaceType(Enum):
    CONTINUOUS = 0
    DISCRETE = 1


class ObservationSpaceType(Enum):
    FLAT_VALUES = 0
    IMAGES = 1


class DataInterfaceType(Enum):
    LIST = 0
    NUMPY = 1
    TORCH = 2
    JAX = 3


class SimulationType(Enum):
    DEFAULT = 0
    JAX_BASED = 1
    ISAAC_LAB = 2
    MANISKILL = 3
    WARP = 4


def same_member(a, b):
    """Enum members of this package and of an unmodified reference checkout are different classes; compare by name."""
    return getattr(a, "name", a) == getattr(b, "name", b)
