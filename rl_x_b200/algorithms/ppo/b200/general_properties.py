from rl_x_b200.plugin_properties import algorithm_properties

GeneralProperties = algorithm_properties(
    "What this plugin supports; the runner's compatibility check rejects everything else instead of silently falling back (the reference plugin also lists IMAGES / DISCRETE, rl_x/algorithms/ppo/pytorch/general_properties.py).",
    observations=("FLAT_VALUES",), actions=("CONTINUOUS",), interfaces=('NUMPY', 'TORCH'))
