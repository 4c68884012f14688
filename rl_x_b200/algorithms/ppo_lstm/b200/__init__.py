"""Plugin package `ppo_lstm.b200`: importing it registers the algorithm (the reference's registration contract, rl_x/algorithms/algorithm_manager.py)."""
from rl_x_b200.algorithms.algorithm_manager import register_algorithm_package

NAME = register_algorithm_package(__file__, "ppo_lstm", "PPO_LSTM")
