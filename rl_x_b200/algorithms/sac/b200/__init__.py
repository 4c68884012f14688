from rl_x_b200.algorithms.algorithm_manager import extract_algorithm_name_from_file, register_algorithm
from rl_x_b200.algorithms.sac.b200.sac import SAC
from rl_x_b200.algorithms.sac.b200.default_config import get_config
from rl_x_b200.algorithms.sac.b200.general_properties import GeneralProperties


SAC_B200 = extract_algorithm_name_from_file(__file__)
register_algorithm(SAC_B200, get_config, SAC, GeneralProperties)
