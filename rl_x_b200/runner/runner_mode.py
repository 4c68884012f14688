"""Run modes; the values are what `--runner.mode=` takes (ref: rl_x/runner/runner_mode.py).  str-valued so that the parsed flag compares equal."""
import enum


class RunnerMode(str, enum.Enum):
    TRAIN = "train"
    TEST = "test"
    SHOW_CONFIG = "show_config"

    def __str__(self):
        return self.value
