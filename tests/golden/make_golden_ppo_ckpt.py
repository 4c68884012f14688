#!/usr/bin/env python
"""Generate tests/golden/ppo_ref_checkpoint.model: a `best.model` written by the UNMODIFIED reference's own PPO.save()
(rl_x/algorithms/ppo/pytorch/ppo.py:426-436) after two training iterations, with torch.compile wrappers in place (so the module keys
carry "_orig_mod." and the optimizer states use the reference's parameter numbering).  Build container only:

    python tests/golden/make_golden_ppo_ckpt.py

The config tree class pickled into the file is rl_x_b200.config_dict.ConfigDict (installed as the `ml_collections` stub), so the file
unpickles wherever this repository is importable.  A side file ppo_ref_checkpoint_expect.npz holds the same tensors by NAME (taken from
named_parameters() / the optimizer's param -> state mapping) for the tests to compare against.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
os.environ["CXX"], os.environ["CC"] = "/usr/bin/g++", "/usr/bin/gcc"

import numpy as np
import torch

from rl_x_b200.config_dict import ConfigDict

mc, cd = types.ModuleType("ml_collections"), types.ModuleType("ml_collections.config_dict")
cd.ConfigDict = ConfigDict
mc.config_dict = cd
sys.modules["ml_collections"], sys.modules["ml_collections.config_dict"] = mc, cd
sys.path.insert(0, "/root/reference")
import rl_x.algorithms.ppo.pytorch.ppo as refppo  # noqa: E402
from rl_x.algorithms.ppo.pytorch.default_config import get_config  # noqa: E402
from make_golden_ppo import SyntheticTorchEnv  # noqa: E402  (scaffolding env of the PPO goldens)

N, T, OBS, ACT, HID, MB, E = 8, 4, 12, 3, 32, 16, 2


def main():
    a = get_config("ppo.pytorch")
    a.device, a.bf16_mixed_precision_training, a.compile_mode = "cpu", False, "default"
    a.nr_steps, a.nr_epochs, a.minibatch_size, a.nr_hidden_units, a.total_timesteps = T, E, MB, HID, float(2 * N * T)
    cfg = ConfigDict(algorithm=a, environment=ConfigDict(seed=3, nr_envs=N),
                     runner=ConfigDict(save_model=True, track_console=False, track_tb=False, track_wandb=False))
    env = SyntheticTorchEnv(N, OBS, ACT, seed=3)
    run_path = "/tmp/rlx_golden_ckpt"
    import shutil
    shutil.rmtree(run_path, ignore_errors=True)
    model = refppo.PPO(cfg, env, env, run_path, None)
    model.log = lambda *a_, **k_: None
    model.start_logging = model.end_logging = lambda *a_, **k_: None
    model.train()
    model.save()
    src = os.path.join(run_path, "models", "best.model")
    dst = os.path.join(HERE, "ppo_ref_checkpoint.model")
    shutil.copyfile(src, dst)
    expect = {"meta": np.asarray([N, T, OBS, ACT, HID, MB, E], dtype=np.int64)}
    for net, opt, tag in ((model.policy, model.policy_optimizer, "policy"), (model.critic, model.critic_optimizer, "critic")):
        for name, p in net.named_parameters():
            name = name.replace("_orig_mod.", "")
            st = opt.state[p]
            expect[f"{tag}/{name}/param"] = p.detach().numpy().copy()
            expect[f"{tag}/{name}/exp_avg"] = st["exp_avg"].numpy().copy()
            expect[f"{tag}/{name}/exp_avg_sq"] = st["exp_avg_sq"].numpy().copy()
            expect[f"{tag}/{name}/step"] = np.asarray(float(st["step"]))
    np.savez_compressed(os.path.join(HERE, "ppo_ref_checkpoint_expect.npz"), **expect)
    ck = torch.load(dst, weights_only=False)
    print("wrote", dst, os.path.getsize(dst), "bytes; policy keys:", list(ck["policy_state_dict"]))
    print("optimizer index 0 exp_avg shape:", tuple(ck["policy_optimizer_state_dict"]["state"][0]["exp_avg"].shape))


if __name__ == "__main__":
    main()
