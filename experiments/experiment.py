"""CLI entry, same shape as the reference's experiments/experiment.py:1-6.

    python experiments/experiment.py --algorithm.name=ppo.b200 --environment.name=synthetic.box --algorithm.nr_steps=128 ...
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rl_x_b200.runner.runner import Runner

if __name__ == "__main__":
    Runner().run()
