from rl_x_b200.plugin_properties import algorithm_properties

GeneralProperties = algorithm_properties(
    'ref: rl_x/algorithms/sac/pytorch/general_properties.py.',
    observations=("FLAT_VALUES",), actions=("CONTINUOUS",), interfaces=('NUMPY', 'TORCH'))
