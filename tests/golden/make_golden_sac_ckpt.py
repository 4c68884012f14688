#!/usr/bin/env python
"""Generate tests/golden/sac_ref_checkpoint.model: a `best.model` written by the UNMODIFIED reference's own SAC.save()
(rl_x/algorithms/sac/pytorch/sac.py:381-396) after a few updates, with the torch.compile wrappers in place (module keys carry
"_orig_mod.", optimizer states use the reference's parameter numbering: policy; q1 then q2; log_alpha).  Build container only:

    python tests/golden/make_golden_sac_ckpt.py

The config tree class pickled into the file is rl_x_b200.config_dict.ConfigDict (installed as the `ml_collections` stub), so the file
unpickles wherever this repository is importable.  The side file sac_ref_checkpoint_expect.npz holds the same tensors by NAME (taken
from named_parameters() and each optimizer's param -> state mapping) for the tests to compare against.
"""
import os
import shutil
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
sys.path.insert(0, HERE)
import rl_x.algorithms.sac.pytorch.sac as refsac  # noqa: E402
from rl_x.algorithms.sac.pytorch.default_config import get_config  # noqa: E402
from make_golden_sac import SyntheticNumpyEnv  # noqa: E402  (scaffolding env of the SAC goldens; installs its own stub first - ours wins below)

sys.modules["ml_collections"], sys.modules["ml_collections.config_dict"] = mc, cd
N, OBS, ACT, HID, BATCH = 2, 5, 3, 16, 8


def main():
    torch.set_num_threads(1)
    a = ConfigDict(dict(get_config("sac.pytorch")))
    a.device, a.bf16_mixed_precision_training, a.compile_mode = "cpu", False, "default"
    a.nr_hidden_units, a.batch_size, a.learning_starts, a.total_timesteps = HID, BATCH, 8, 40
    a.buffer_size, a.logging_frequency = 256, 4
    cfg = ConfigDict(algorithm=a, environment=ConfigDict(seed=5, nr_envs=N),
                     runner=ConfigDict(save_model=True, track_console=False, track_tb=False, track_wandb=False))
    env = SyntheticNumpyEnv(N, OBS, ACT, 105, -2.0, 1.0)
    run_path = "/tmp/rlx_golden_sac_ckpt"
    shutil.rmtree(run_path, ignore_errors=True)
    model = refsac.SAC(cfg, env, env, run_path, None)
    model.log = lambda *a_, **k_: None
    model.start_logging = model.end_logging = lambda *a_, **k_: None
    model.train()
    model.save()
    src = os.path.join(run_path, "models", "best.model")
    dst = os.path.join(HERE, "sac_ref_checkpoint.model")
    shutil.copyfile(src, dst)
    expect = {"meta": np.asarray([N, OBS, ACT, HID, BATCH], dtype=np.int64)}
    nets = [("policy", model.policy, model.policy_optimizer), ("q1", model.critic.q1, model.q_optimizer), ("q2", model.critic.q2, model.q_optimizer)]
    for tag, net, opt in nets:
        for name, p in net.named_parameters():
            name = name.replace("_orig_mod.", "")
            st = opt.state[p]
            expect[f"{tag}/{name}/param"] = p.detach().numpy().copy()
            expect[f"{tag}/{name}/exp_avg"] = st["exp_avg"].numpy().copy()
            expect[f"{tag}/{name}/exp_avg_sq"] = st["exp_avg_sq"].numpy().copy()
            expect[f"{tag}/{name}/step"] = np.asarray(float(st["step"]))
    for tag, net in (("q1_target", model.critic.q1_target), ("q2_target", model.critic.q2_target)):
        for name, p in net.named_parameters():
            expect[f"{tag}/{name.replace('_orig_mod.', '')}/param"] = p.detach().numpy().copy()
    la = model.entropy_coefficient.log_alpha
    st = model.entropy_optimizer.state[la]
    expect["log_alpha/param"] = la.detach().numpy().copy()
    expect["log_alpha/exp_avg"], expect["log_alpha/exp_avg_sq"] = st["exp_avg"].numpy().copy(), st["exp_avg_sq"].numpy().copy()
    expect["log_alpha/step"] = np.asarray(float(st["step"]))
    np.savez_compressed(os.path.join(HERE, "sac_ref_checkpoint_expect.npz"), **expect)
    ck = torch.load(dst, weights_only=False)
    print("wrote", dst, os.path.getsize(dst), "bytes; policy keys:", list(ck["policy_state_dict"]))
    print("q optimizer entries:", len(ck["q_optimizer_state_dict"]["state"]), "policy:", len(ck["policy_optimizer_state_dict"]["state"]))


if __name__ == "__main__":
    main()
