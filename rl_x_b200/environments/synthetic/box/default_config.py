from rl_x_b200.config_dict import config_tree_class


def get_config(environment_name):
    """Keys follow the reference's torch-interface env (custom_mujoco/ant/warp_torch/default_config.py:4-17) plus the
    synthetic-stream parameters of BASELINE config 2 (SURVEY.md §8 d)."""
    config = config_tree_class()()   # ml_collections.ConfigDict where installed (the reference runner requires it)

    config.name = environment_name

    config.seed = 1
    config.nr_envs = 4096
    config.render = False
    config.device = "gpu"
    config.horizon = 1000                 # truncation horizon; <= 0 disables truncation
    config.copy_train_env_for_eval = True

    config.obs_dim = 376                  # Humanoid-v4-like
    config.act_dim = 17
    config.action_low = -1.0
    config.action_high = 1.0
    config.termination_probability = 0.01
    config.data_interface = "torch"       # torch (device tensors, zero copy) | numpy (host arrays, pinned staging)
    config.stream = "fresh"               # fresh: new N(0,1) draws every step | ring: cycle through `ring_length` pre-drawn steps
    config.ring_length = 8

    return config
