"""Environment side of the plugin surface; function names and argument order are the reference's (rl_x/environments/environment_manager.py)."""
from rl_x_b200.plugin_registry import PluginRegistry

_environments = PluginRegistry("environments", ("get_default_config", "create_train_and_eval_env", "general_properties"))

extract_environment_name_from_file = _environments.name_from_file


def register_environment(name, get_default_config, create_train_and_eval_env, general_properties):
    _environments.register(name, get_default_config, create_train_and_eval_env, general_properties)


def get_environment_config(environment_name):
    return _environments.lookup(environment_name).get_default_config(environment_name)


def get_environment_create_train_and_eval_env(environment_name):
    return _environments.lookup(environment_name).create_train_and_eval_env


def get_environment_general_properties(environment_name):
    return _environments.lookup(environment_name).general_properties
