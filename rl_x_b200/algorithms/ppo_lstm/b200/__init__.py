from rl_x_b200.algorithms.algorithm_manager import extract_algorithm_name_from_file, register_algorithm
from rl_x_b200.algorithms.ppo_lstm.b200.ppo_lstm import PPO_LSTM
from rl_x_b200.algorithms.ppo_lstm.b200.default_config import get_config
from rl_x_b200.algorithms.ppo_lstm.b200.general_properties import GeneralProperties


PPO_LSTM_B200 = extract_algorithm_name_from_file(__file__)
register_algorithm(PPO_LSTM_B200, get_config, PPO_LSTM, GeneralProperties)
