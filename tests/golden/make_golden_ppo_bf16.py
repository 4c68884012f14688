#!/usr/bin/env python
"""Golden vectors of the reference PPO in its bf16 mixed-precision mode (`bf16_mixed_precision_training=True`), made by EXECUTING the
reference on CPU.  Build container only:

    python tests/golden/make_golden_ppo_bf16.py

The reference pins autocast to device_type="cuda" and refuses the mode on other devices (ppo.py:66-67, :100,123,155,208,253).  To run it
here (i) TorchScript is switched off (PYTORCH_JIT=0) so that the scripted mixed-precision GAE (ppo.py:98-107) runs as plain Python,
(ii) the module's `autocast` name is rebound to the CPU autocast context with the same dtype, (iii) the constructor is called in fp32
mode and the flag the training loop reads is set afterwards.  Nothing else is touched; capture is tests/golden/make_golden_ppo.py's.
What the run shows (and oracle/ppo_oracle.py restates): Linear layers produce bf16, tanh stays bf16, Normal(bf16 mean, fp32 std).sample()
is a bf16 tensor and the rollout's log-prob is evaluated on it, `gamma * next_values` is a bf16 product, everything else promotes to fp32.
Output: tests/golden/ppo_small_bf16.npz
"""
import os
import sys

os.environ["TORCHDYNAMO_DISABLE"] = "1"
os.environ["PYTORCH_JIT"] = "0"
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402

import make_golden_ppo as M  # noqa: E402

ref = M.refppo
ref.autocast = lambda device_type, dtype, enabled=True: torch.autocast("cpu", dtype=dtype, enabled=enabled)
_init = ref.PPO.__init__


def init(self, config, *args, **kwargs):
    config.algorithm.bf16_mixed_precision_training = False
    _init(self, config, *args, **kwargs)
    self.bf16_mixed_precision_training = True


ref.PPO.__init__ = init

if __name__ == "__main__":
    M.run("small_bf16", N=12, T=9, obs_dim=11, act_dim=3, hidden=64, mb=40, epochs=2, iterations=2, seed=3, act_low=-2.0, act_high=0.5, std_dev=0.7,
          entropy_coef=0.01, anneal=True)
