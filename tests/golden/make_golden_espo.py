#!/usr/bin/env python
"""Generate ESPO golden vectors by EXECUTING the unmodified reference (rl_x/algorithms/espo/pytorch/espo.py).

Run in the build container only (the reference does not travel to the GPU box):

    TORCHDYNAMO_DISABLE=1 python tests/golden/make_golden_espo.py

Captured: the rollout `Batch` of every iteration (incl. advantages / returns, espo.py:238-243), every index array drawn by
`self.rng.choice(batch_size, size=minibatch_size, replace=False)` (espo.py:256), the number of update steps each iteration made
before `ratio_delta > max_ratio_delta` stopped it (espo.py:273-274), policy / critic weights and Adam moments after each
iteration, every logged metric.  The synthetic NUMPY-interface environment is test scaffolding.
Output: tests/golden/espo_<tag>.npz.
"""
import os
import sys

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_ppo import ConfigDict, _Space, sd_to_np, opt_to_np  # noqa: E402  (installs the ml_collections stub, adds the reference to sys.path)

import rl_x.algorithms.espo.pytorch.espo as refespo  # noqa: E402
from rl_x.algorithms.espo.pytorch.default_config import get_config  # noqa: E402
from rl_x.environments.action_space_type import ActionSpaceType  # noqa: E402
from rl_x.environments.observation_space_type import ObservationSpaceType  # noqa: E402
from rl_x.environments.data_interface_type import DataInterfaceType  # noqa: E402


class _Props:
    observation_space_type = ObservationSpaceType.FLAT_VALUES
    action_space_type = ActionSpaceType.CONTINUOUS
    data_interface_type = DataInterfaceType.NUMPY


class SyntheticNumpyEnv:
    """obs ~ N(0,1), reward ~ N(0,1), terminated ~ Bernoulli(p), truncated every 11 steps; the final observation of a finished
    episode is the returned row itself (no separate reset observation)."""

    general_properties = _Props

    def __init__(self, nr_envs, obs_dim, act_dim, seed, p_term=0.05, horizon=11, act_low=-1.0, act_high=1.0):
        self.nr_envs, self.obs_dim, self.act_dim = nr_envs, obs_dim, act_dim
        self.single_observation_space = _Space((obs_dim,))
        self.single_action_space = _Space((act_dim,), np.full(act_dim, act_low, np.float32), np.full(act_dim, act_high, np.float32))
        self.gen = np.random.default_rng(seed)
        self.p_term, self.horizon, self.t = p_term, horizon, 0
        self.received_actions = []

    def reset(self):
        return self.gen.standard_normal((self.nr_envs, self.obs_dim)).astype(np.float32), {}

    def step(self, action):
        self.received_actions.append(np.array(action, dtype=np.float32, copy=True))
        self.t += 1
        self.obs = self.gen.standard_normal((self.nr_envs, self.obs_dim)).astype(np.float32)
        rew = self.gen.standard_normal(self.nr_envs).astype(np.float32)
        term = self.gen.random(self.nr_envs) < self.p_term
        trunc = np.full(self.nr_envs, self.t % self.horizon == 0)
        return self.obs, rew, term, trunc, {}

    def get_logging_info_dict(self, info):
        return {}

    def get_final_observation_at_index(self, info, i):
        return self.obs[i]

    def get_final_info_value_at_index(self, info, key, i):
        return 0.0

    def close(self):
        pass


class RngSpy:
    def __init__(self, rng):
        self.rng, self.draws = rng, []

    def choice(self, a, size=None, replace=True):
        out = self.rng.choice(a, size=size, replace=replace)
        self.draws.append(np.asarray(out, dtype=np.int64).copy())
        return out


def run(tag, N, T, obs_dim, act_dim, hidden, mb, max_epochs, max_ratio_delta, iterations, seed, lr, act_low=-1.0, act_high=1.0,
        std_dev=1.0, entropy_coef=0.0, anneal=False):
    torch.set_num_threads(1)
    captured = {}

    class CapBatch(refespo.Batch):
        def __init__(self, **kw):
            super().__init__(**kw)
            captured["batch"] = self

    refespo.Batch = CapBatch
    cfg = ConfigDict(algorithm=get_config("espo.pytorch"), environment=ConfigDict(seed=seed, nr_envs=N),
                     runner=ConfigDict(save_model=False, track_console=False, track_tb=False, track_wandb=False))
    a = cfg.algorithm
    a.device, a.bf16_mixed_precision_training = "cpu", False
    a.nr_steps, a.minibatch_size, a.max_epochs, a.nr_hidden_units = T, mb, max_epochs, hidden
    a.max_ratio_delta, a.learning_rate = max_ratio_delta, lr
    a.total_timesteps = N * T * iterations
    a.std_dev, a.entropy_coef, a.anneal_learning_rate = std_dev, entropy_coef, anneal
    env = SyntheticNumpyEnv(N, obs_dim, act_dim, seed=seed + 1000, act_low=act_low, act_high=act_high)
    model = refespo.ESPO(cfg, env, env, "/tmp/golden_run", None)
    out = {}
    out.update(sd_to_np("init/policy", model.policy.state_dict()))
    out.update(sd_to_np("init/critic", model.critic.state_dict()))
    next_values_log = []
    orig_get_value = model.critic.get_value

    def get_value(x):
        o = orig_get_value(x)
        if x.dim() == 3:
            next_values_log.append(o.detach().squeeze(-1).numpy().copy())
        return o

    model.critic.get_value = get_value
    spy = RngSpy(model.rng)
    model.rng = spy
    metrics = []
    model.log = lambda name, value, step: metrics.append((name, float(value), int(step)))
    per_iter, draws_before = [], [0]
    orig_start = model.start_logging

    def start_logging(step):
        b = captured["batch"]
        snap = {k: getattr(b, k).detach().numpy().copy() for k in
                ["states", "next_states", "actions", "rewards", "values", "terminations", "log_probs", "advantages", "returns"]}
        snap.update(sd_to_np("policy", model.policy.state_dict()))
        snap.update(sd_to_np("critic", model.critic.state_dict()))
        snap.update(opt_to_np("policy_opt", model.policy_optimizer, model.policy))
        snap.update(opt_to_np("critic_opt", model.critic_optimizer, model.critic))
        snap["nr_epochs"] = np.array(len(spy.draws) - draws_before[0], dtype=np.int64)
        draws_before[0] = len(spy.draws)
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
    for i, d in enumerate(spy.draws):
        out[f"choice/{i}"] = d
    out["env_actions"] = np.stack(env.received_actions)
    mnames = sorted({m[0] for m in metrics})
    for n in mnames:
        if n.startswith("time/"):
            continue
        out[f"metric/{n}"] = np.array([m[1] for m in metrics if m[0] == n], dtype=np.float64)
    out["meta"] = np.array([N, T, obs_dim, act_dim, hidden, mb, max_epochs, iterations, seed], dtype=np.int64)
    out["meta_f"] = np.array([a.gamma, a.gae_lambda, max_ratio_delta, a.entropy_coef, a.critic_coef, a.max_grad_norm, a.learning_rate,
                              a.std_dev, act_low, act_high, float(anneal)], dtype=np.float64)
    path = os.path.join(HERE, f"espo_{tag}.npz")
    np.savez_compressed(path, **out)
    print(tag, "->", path, os.path.getsize(path) // 1024, "KiB; update steps per iteration:", [int(s["nr_epochs"]) for s in per_iter])


if __name__ == "__main__":
    # the stop rule fires after a different number of steps in each iteration; odd dims, non-trivial action bounds, entropy term
    run("small", N=6, T=20, obs_dim=11, act_dim=3, hidden=64, mb=32, max_epochs=12, max_ratio_delta=0.02, iterations=3, seed=5, lr=1e-3,
        act_low=-2.0, act_high=0.5, std_dev=0.7, entropy_coef=0.01, anneal=True)
