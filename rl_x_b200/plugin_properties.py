"""`GeneralProperties` classes of the plugins, built from member names.  The runner's compatibility check (ref: rl_x/runner/runner.py:86-91)
reads class attributes: an algorithm lists the space / interface types it accepts, an environment states the single type it has."""
from rl_x_b200.algorithms.deep_learning_framework_type import DeepLearningFrameworkType
from rl_x_b200.environments.types import ActionSpaceType, DataInterfaceType, ObservationSpaceType


def algorithm_properties(doc, observations, actions, interfaces, framework="TORCH"):
    return type("GeneralProperties", (), {
        "__doc__": doc,
        "observation_space_types": [ObservationSpaceType[n] for n in observations],
        "action_space_types": [ActionSpaceType[n] for n in actions],
        "data_interface_types": [DataInterfaceType[n] for n in interfaces],
        "deep_learning_framework_type": DeepLearningFrameworkType[framework],
    })


def environment_properties(doc, observation, action, interface, **extra):
    attrs = {"__doc__": doc, "observation_space_type": ObservationSpaceType[observation], "action_space_type": ActionSpaceType[action],
             "data_interface_type": DataInterfaceType[interface]}
    attrs.update(extra)
    return type("GeneralProperties", (), attrs)
