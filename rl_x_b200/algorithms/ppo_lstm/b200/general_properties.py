from rl_x_b200.plugin_properties import algorithm_properties

GeneralProperties = algorithm_properties(
    'ref: rl_x/algorithms/ppo_lstm/flax/general_properties.py (NUMPY there; TORCH-interface environments work too).',
    observations=("FLAT_VALUES",), actions=("CONTINUOUS",), interfaces=('NUMPY', 'TORCH'))
