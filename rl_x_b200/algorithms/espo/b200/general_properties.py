from rl_x_b200.plugin_properties import algorithm_properties

GeneralProperties = algorithm_properties(
    'ref: rl_x/algorithms/espo/pytorch/general_properties.py (NUMPY only there; the rollout code shared with the PPO plugin also serves TORCH-interface environments).',
    observations=("FLAT_VALUES",), actions=("CONTINUOUS",), interfaces=('NUMPY', 'TORCH'))
