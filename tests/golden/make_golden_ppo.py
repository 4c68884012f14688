#!/usr/bin/env python
"""Generate PPO golden vectors by EXECUTING the unmodified reference (RL-X @ /root/reference).

Run in the build container only (the reference does not travel to the GPU box):

    TORCHDYNAMO_DISABLE=1 python tests/golden/make_golden_ppo.py

What is captured (SURVEY.md §4 / Appendix A recipe):
  * the rollout tensors of `Batch` (rl_x/algorithms/ppo/pytorch/batch.py:1-11) incl. the
    `advantages` / `returns` attributes assigned at ppo.py:256-258,
  * every index array produced by `self.rng.shuffle` (ppo.py:276),
  * policy / critic `state_dict()` and both Adam states before and after `train()`,
  * every metric passed to `PPO.log` (ppo.py:396-402).
The synthetic environment below is test scaffolding (TORCH data interface on CPU tensors), not
reference code.  Output: tests/golden/ppo_<tag>.npz (fp32 / int64 arrays only).
"""
import copy
import os
import sys
import types

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _install_ml_collections_stub():
    class ConfigDict(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        __setattr__ = dict.__setitem__

    mc = types.ModuleType("ml_collections")
    cd = types.ModuleType("ml_collections.config_dict")
    cd.ConfigDict = ConfigDict
    mc.config_dict = cd
    sys.modules["ml_collections"] = mc
    sys.modules["ml_collections.config_dict"] = cd
    return ConfigDict


ConfigDict = _install_ml_collections_stub()
sys.path.insert(0, "/root/reference")
import rl_x.algorithms.ppo.pytorch.ppo as refppo  # noqa: E402
from rl_x.algorithms.ppo.pytorch.default_config import get_config  # noqa: E402
from rl_x.environments.action_space_type import ActionSpaceType  # noqa: E402
from rl_x.environments.observation_space_type import ObservationSpaceType  # noqa: E402
from rl_x.environments.data_interface_type import DataInterfaceType  # noqa: E402


class _Space:
    def __init__(self, shape, low=None, high=None):
        self.shape = shape
        self.low = low
        self.high = high


class _Props:
    observation_space_type = ObservationSpaceType.FLAT_VALUES
    action_space_type = ActionSpaceType.CONTINUOUS
    data_interface_type = DataInterfaceType.TORCH


class SyntheticTorchEnv:
    """obs ~ N(0,1), reward ~ N(0,1), terminated ~ Bernoulli(p), truncated every `horizon` steps."""

    general_properties = _Props

    def __init__(self, nr_envs, obs_dim, act_dim, seed, p_term=0.05, horizon=11, act_low=-1.0, act_high=1.0):
        self.nr_envs, self.obs_dim, self.act_dim = nr_envs, obs_dim, act_dim
        self.single_observation_space = _Space((obs_dim,))
        self.single_action_space = _Space((act_dim,), np.full(act_dim, act_low, np.float32), np.full(act_dim, act_high, np.float32))
        self.gen = torch.Generator().manual_seed(seed)
        self.p_term, self.horizon, self.t = p_term, horizon, 0
        self.received_actions = []

    def reset(self):
        return torch.randn(self.nr_envs, self.obs_dim, generator=self.gen), {}

    def step(self, action):
        self.received_actions.append(action.clone())
        self.t += 1
        obs = torch.randn(self.nr_envs, self.obs_dim, generator=self.gen)
        rew = torch.randn(self.nr_envs, generator=self.gen)
        term = torch.rand(self.nr_envs, generator=self.gen) < self.p_term
        trunc = torch.full((self.nr_envs,), self.t % self.horizon == 0)
        return obs, rew, term, trunc, {}

    def get_logging_info_dict(self, info):
        return {}

    def close(self):
        pass


class RngSpy:
    def __init__(self, rng):
        self.rng, self.perms = rng, []

    def shuffle(self, a):
        self.rng.shuffle(a)
        self.perms.append(a.copy())


def sd_to_np(prefix, sd):
    return {f"{prefix}/{k}": v.detach().cpu().numpy().copy() for k, v in sd.items()}


def opt_to_np(prefix, opt, module):
    out = {}
    names = [n for n, _ in module.named_parameters()]
    st = opt.state_dict()["state"]
    for i, n in enumerate(names):
        if i in st:
            out[f"{prefix}/{n}/exp_avg"] = st[i]["exp_avg"].numpy().copy()
            out[f"{prefix}/{n}/exp_avg_sq"] = st[i]["exp_avg_sq"].numpy().copy()
            out[f"{prefix}/{n}/step"] = np.array(float(st[i]["step"]))
    return out


def run(tag, N, T, obs_dim, act_dim, hidden, mb, epochs, iterations, seed, act_low=-1.0, act_high=1.0, std_dev=1.0,
        entropy_coef=0.0, anneal=False, keep_moments=True):
    torch.set_num_threads(1)  # deterministic reduction order in the captured reference numbers
    captured = {}

    class CapBatch(refppo.Batch):
        def __init__(self, **kw):
            super().__init__(**kw)
            captured["batch"] = self

    refppo.Batch = CapBatch
    cfg = ConfigDict(
        algorithm=get_config("ppo.pytorch"),
        environment=ConfigDict(seed=seed, nr_envs=N),
        runner=ConfigDict(save_model=False, track_console=False, track_tb=False, track_wandb=False),
    )
    a = cfg.algorithm
    a.device, a.bf16_mixed_precision_training = "cpu", False
    a.nr_steps, a.minibatch_size, a.nr_epochs, a.nr_hidden_units = T, mb, epochs, hidden
    a.total_timesteps = N * T * iterations
    a.std_dev, a.entropy_coef, a.anneal_learning_rate = std_dev, entropy_coef, anneal
    env = SyntheticTorchEnv(N, obs_dim, act_dim, seed=seed + 1000, act_low=act_low, act_high=act_high)
    model = refppo.PPO(cfg, env, env, "/tmp/golden_run", None)
    out = {}
    out.update(sd_to_np("init/policy", model.policy.state_dict()))
    out.update(sd_to_np("init/critic", model.critic.state_dict()))
    # next_values is a local of train() (ppo.py:253-254): capture the critic output of the one call that sees the 3-D next_states
    next_values_log = []
    orig_get_value = model.critic.get_value

    def get_value(x):
        out = orig_get_value(x)
        if x.dim() == 3:
            next_values_log.append(out.detach().float().squeeze(-1).numpy().copy())
        return out

    model.critic.get_value = get_value
    spy = RngSpy(model.rng)
    model.rng = spy
    metrics = []
    model.log = lambda name, value, step: metrics.append((name, float(value), int(step)))

    # snapshot after each iteration: wrap start_logging (called once per iteration, ppo.py:367)
    per_iter = []
    orig_start = model.start_logging

    def start_logging(step):
        b = captured["batch"]
        snap = {k: getattr(b, k).detach().numpy().copy() for k in
                ["states", "next_states", "actions", "rewards", "values", "terminations", "log_probs", "advantages", "returns"]}
        snap.update(sd_to_np("policy", model.policy.state_dict()))
        snap.update(sd_to_np("critic", model.critic.state_dict()))
        if keep_moments:
            snap.update(opt_to_np("policy_opt", model.policy_optimizer, model.policy))
            snap.update(opt_to_np("critic_opt", model.critic_optimizer, model.critic))
        per_iter.append(snap)
        orig_start(step)

    model.start_logging = start_logging
    model.train()

    for it, snap in enumerate(per_iter):
        for k, v in snap.items():
            out[f"iter{it}/{k}"] = v
    assert len(next_values_log) == iterations
    for it, nv in enumerate(next_values_log):
        out[f"iter{it}/next_values"] = nv
    for i, p in enumerate(spy.perms):
        out[f"perm/{i}"] = p.astype(np.int64)
    out["env_actions"] = torch.stack(env.received_actions).numpy()
    mnames = sorted({m[0] for m in metrics})
    for n in mnames:
        if n.startswith("time/"):
            continue
        out[f"metric/{n}"] = np.array([m[1] for m in metrics if m[0] == n], dtype=np.float64)
    out["meta"] = np.array([N, T, obs_dim, act_dim, hidden, mb, epochs, iterations, seed], dtype=np.int64)
    out["meta_f"] = np.array([a.gamma, a.gae_lambda, a.clip_range, a.entropy_coef, a.critic_coef, a.max_grad_norm,
                              a.learning_rate, a.std_dev, act_low, act_high, float(anneal)], dtype=np.float64)
    path = os.path.join(HERE, f"ppo_{tag}.npz")
    np.savez_compressed(path, **out)
    print(tag, "->", path, os.path.getsize(path) // 1024, "KiB; perms:", len(spy.perms), "metrics:", len(mnames))


if __name__ == "__main__":
    # small odd dims (exercise unaligned paths), short last minibatch (B % mb != 0), non-trivial action bounds, entropy term
    run("small", N=12, T=9, obs_dim=11, act_dim=3, hidden=64, mb=40, epochs=2, iterations=2, seed=3,
        act_low=-2.0, act_high=0.5, std_dev=0.7, entropy_coef=0.01, anneal=True)
    # Humanoid-like dims (BASELINE config 2 network), reduced N/T; Adam moments omitted to keep the fixture small
    run("humanoid", N=16, T=8, obs_dim=376, act_dim=17, hidden=256, mb=32, epochs=2, iterations=1, seed=1, keep_moments=False)
