from rl_x_b200.environments.synthetic.box.environment import SyntheticBoxEnv
from rl_x_b200.environments.synthetic.box.general_properties import GeneralProperties, GeneralPropertiesNumpy
from rl_x_b200.environments.synthetic.box.wrappers import RLXInfo


def create_train_and_eval_env(config):
    """Same contract as the reference's create_train_and_eval_env (custom_mujoco/ant/warp_torch/create_env.py:6-17)."""
    props = GeneralPropertiesNumpy if config.environment.data_interface == "numpy" else GeneralProperties
    train_env = RLXInfo(SyntheticBoxEnv(config.environment))
    train_env.general_properties = props

    if config.environment.copy_train_env_for_eval:
        return train_env, train_env

    eval_env = RLXInfo(SyntheticBoxEnv(config.environment, seed_offset=1))
    eval_env.general_properties = props

    return train_env, eval_env
