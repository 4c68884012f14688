from rl_x_b200.environments.synthetic.pendulum.environment import PendulumVecEnv
from rl_x_b200.environments.synthetic.pendulum.general_properties import GeneralProperties
from rl_x_b200.environments.synthetic.pendulum.wrappers import RLXInfo


def create_train_and_eval_env(config):
    """Same contract as the reference's gym create_train_and_eval_env (gym/mujoco/humanoid_v4/create_env.py:8-39)."""
    if config.environment.type != "Pendulum-v1":
        raise ValueError("synthetic.pendulum restates Pendulum-v1 only (gymnasium is not available in this image)")
    train_env = RLXInfo(PendulumVecEnv(config.environment.nr_envs, config.environment.seed))
    train_env.general_properties = GeneralProperties
    train_env.reset(seed=config.environment.seed)

    if config.environment.copy_train_env_for_eval:
        return train_env, train_env

    eval_env = RLXInfo(PendulumVecEnv(config.environment.nr_envs, config.environment.seed))
    eval_env.general_properties = GeneralProperties
    eval_env.reset(seed=config.environment.seed)

    return train_env, eval_env
