from rl_x_b200.algorithms.algorithm_manager import extract_algorithm_name_from_file, register_algorithm
from rl_x_b200.algorithms.fastsac.b200.fastsac import FastSAC
from rl_x_b200.algorithms.fastsac.b200.default_config import get_config
from rl_x_b200.algorithms.fastsac.b200.general_properties import GeneralProperties


FASTSAC_B200 = extract_algorithm_name_from_file(__file__)
register_algorithm(FASTSAC_B200, get_config, FastSAC, GeneralProperties)
