"""Runner config tree: the flags of rl_x/runner/default_config.py minus the JAX-only ones; `run_name` defaults to the start time."""
import time

from rl_x_b200.config_dict import config_from_defaults

_DEFAULTS = (
    ("track_console", False),
    ("track_tb", False),
    ("track_wandb", False),
    ("wandb_entity", "placeholder"),
    ("project_name", "placeholder"),
    ("exp_name", "placeholder"),
    ("notes", "placeholder"),
    ("save_model", False),
    ("load_model", ""),
    ("nr_test_episodes", 10),   # runner mode = test
)


def get_config(runner_mode):
    config = config_from_defaults(None, (("mode", runner_mode),) + _DEFAULTS + (("run_name", f"{int(time.time())}"),))
    del config["name"]
    return config
