class RunnerMode:
    """rl_x/runner/runner_mode.py:1-4"""
    TRAIN = "train"
    TEST = "test"
    SHOW_CONFIG = "show_config"
