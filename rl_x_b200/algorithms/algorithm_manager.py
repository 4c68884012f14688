"""Algorithm registry with the reference's registration contract (rl_x/algorithms/algorithm_manager.py:8-25):
a plugin package's __init__.py calls register_algorithm(name, get_config, ModelClass, GeneralProperties)."""
from os import sep

_registry = {}


class Algorithm:
    def __init__(self, name, get_default_config, get_model_class, general_properties):
        self.name = name
        self.get_default_config = get_default_config
        self.get_model_class = get_model_class
        self.general_properties = general_properties


def extract_algorithm_name_from_file(file_name):
    # ".../algorithms/ppo/b200/__init__.py" -> "ppo.b200"
    tail = file_name.split(f"algorithms{sep}")[-1]
    return tail.split(f"{sep}__init__.py")[0].replace(sep, ".")


def register_algorithm(name, get_default_config, get_model_class, general_properties):
    _registry[name] = Algorithm(name, get_default_config, get_model_class, general_properties)


def get_algorithm_config(algorithm_name):
    return _registry[algorithm_name].get_default_config(algorithm_name)


def get_algorithm_model_class(algorithm_name):
    return _registry[algorithm_name].get_model_class


def get_algorithm_general_properties(algorithm_name):
    return _registry[algorithm_name].general_properties
