import time

from rl_x_b200.config_dict import ConfigDict


def get_config(runner_mode):
    """Runner flags of the reference (rl_x/runner/default_config.py:5-33) minus the JAX-only ones."""
    config = ConfigDict()

    config.mode = runner_mode

    config.track_console = False
    config.track_tb = False
    config.track_wandb = False
    config.wandb_entity = "placeholder"
    config.project_name = "placeholder"
    config.exp_name = "placeholder"
    config.run_name = f"{int(time.time())}"
    config.notes = "placeholder"

    config.save_model = False
    config.load_model = ""

    config.nr_test_episodes = 10  # if runner mode = test

    return config
