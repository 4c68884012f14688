"""Environment registry with the reference's contract (rl_x/environments/environment_manager.py:8-25)."""
from os import sep

_registry = {}


class Environment:
    def __init__(self, name, get_default_config, create_train_and_eval_env, general_properties):
        self.name = name
        self.get_default_config = get_default_config
        self.create_train_and_eval_env = create_train_and_eval_env
        self.general_properties = general_properties


def extract_environment_name_from_file(file_name):
    tail = file_name.split(f"environments{sep}")[-1]
    return tail.split(f"{sep}__init__.py")[0].replace(sep, ".")


def register_environment(name, get_default_config, create_train_and_eval_env, general_properties):
    _registry[name] = Environment(name, get_default_config, create_train_and_eval_env, general_properties)


def get_environment_config(environment_name):
    return _registry[environment_name].get_default_config(environment_name)


def get_environment_create_train_and_eval_env(environment_name):
    return _registry[environment_name].create_train_and_eval_env


def get_environment_general_properties(environment_name):
    return _registry[environment_name].general_properties
