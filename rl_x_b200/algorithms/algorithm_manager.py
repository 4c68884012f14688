"""Algorithm side of the plugin surface; function names and argument order are the reference's (rl_x/algorithms/algorithm_manager.py)."""
from rl_x_b200.plugin_registry import PluginRegistry

_algorithms = PluginRegistry("algorithms", ("get_default_config", "model_class", "general_properties"))

extract_algorithm_name_from_file = _algorithms.name_from_file


def register_algorithm(name, get_default_config, get_model_class, general_properties):
    _algorithms.register(name, get_default_config, get_model_class, general_properties)


def get_algorithm_config(algorithm_name):
    return _algorithms.lookup(algorithm_name).get_default_config(algorithm_name)


def get_algorithm_model_class(algorithm_name):
    return _algorithms.lookup(algorithm_name).model_class


def get_algorithm_general_properties(algorithm_name):
    return _algorithms.lookup(algorithm_name).general_properties


def register_algorithm_package(init_file, module, class_name):
    """Registration by convention for a plugin package `<...>/algorithms/<algo>/<variant>/`: `default_config.get_config`,
    `general_properties.GeneralProperties` and `<module>.<class_name>` of that package, under the name derived from its path."""
    import importlib
    name = extract_algorithm_name_from_file(init_file)
    package = "rl_x_b200.algorithms." + name
    get_config = importlib.import_module(package + ".default_config").get_config
    properties = importlib.import_module(package + ".general_properties").GeneralProperties
    model_class = getattr(importlib.import_module(f"{package}.{module}"), class_name)
    register_algorithm(name, get_config, model_class, properties)
    return name
