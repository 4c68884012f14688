from rl_x_b200.config_dict import config_tree_class


def get_config(environment_name):
    """Keys of the reference's gym environments (gym/mujoco/humanoid_v4/default_config.py:4-15); `type` is fixed: gymnasium is not in
    this image, so Pendulum-v1 is restated in NumPy (BASELINE.json configs[0]: nr_envs=4, CPU-side env, plumbing check)."""
    config = config_tree_class()()   # ml_collections.ConfigDict where installed (the reference runner requires it)

    config.name = environment_name

    config.type = "Pendulum-v1"
    config.seed = 1
    config.nr_envs = 4
    config.async_skip_percentage = 0.0
    config.render = False
    config.copy_train_env_for_eval = True

    return config
