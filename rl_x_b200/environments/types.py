"""The four environment enums of the reference (rl_x/environments/{action_space,observation_space,data_interface,simulation}_type.py),
same member names and values so that GeneralProperties of reference plugins and of this package compare by name."""
from enum import Enum


class NamedEnum(Enum):
    """Enum whose members equal the same-named members of a same-named Enum class of ANOTHER package.  The reference's runner checks
    plugin compatibility with `env_props.action_space_type not in algo_props.action_space_types` and
    `DeepLearningFrameworkType.TORCH == algo_props.deep_learning_framework_type` (rl_x/runner/runner.py:86-101): with the reference's enum
    classes on one side and this package's on the other, identity comparison would reject every pairing."""

    def __eq__(self, other):
        if isinstance(other, Enum) and type(other).__name__ == type(self).__name__:
            return self.name == other.name
        return NotImplemented

    def __hash__(self):
        return hash((type(self).__name__, self.name))


class ActionSpaceType(NamedEnum):
    CONTINUOUS = 0
    DISCRETE = 1


class ObservationSpaceType(NamedEnum):
    FLAT_VALUES = 0
    IMAGES = 1


class DataInterfaceType(NamedEnum):
    LIST = 0
    NUMPY = 1
    TORCH = 2
    JAX = 3


class SimulationType(NamedEnum):
    DEFAULT = 0
    JAX_BASED = 1
    ISAAC_LAB = 2
    MANISKILL = 3
    WARP = 4


def require_identity_observation_indices(env, who):
    """The reference's networks index-select their input with `env.policy_observation_indices` / `critic_observation_indices`
    (ppo/pytorch/policy.py:14,62, critic.py:10,45; the same lines in sac, fastsac and ppo_lstm).  Every environment of the benchmark
    configs leaves them at arange(obs_dim); this build's kernels read the whole observation row, so anything else is rejected here
    rather than silently ignored."""
    import numpy as np
    obs_dim = int(env.single_observation_space.shape[0])
    for attr in ("policy_observation_indices", "critic_observation_indices"):
        ind = getattr(env, attr, None)
        if ind is not None and not np.array_equal(np.asarray(ind).reshape(-1), np.arange(obs_dim)):
            raise ValueError(f"rl_x_b200 {who} does not implement a non-identity {attr}.")


def same_member(a, b):
    """Enum members of this package and of an unmodified reference checkout are different classes; compare by name."""
    return getattr(a, "name", a) == getattr(b, "name", b)
