#!/usr/bin/env python
"""Generate SAC golden vectors by EXECUTING the unmodified reference (rl_x/algorithms/sac/pytorch) on CPU.

    TORCHDYNAMO_DISABLE=1 python tests/golden/make_golden_sac.py

Captured: initial / final policy, q1, q2, q-target and log_alpha parameters, every sampled replay batch (via a ReplayBuffer
subclass), every standard-normal draw of Normal.rsample() (torch.distributions.utils._standard_normal is wrapped), and the
per-update metrics (logging_frequency = nr_envs makes every logged mean a single update).  Output: tests/golden/sac_small.npz
"""
import os
import sys
import types

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)
    __setattr__ = dict.__setitem__


mc = types.ModuleType("ml_collections"); cd = types.ModuleType("ml_collections.config_dict")
cd.ConfigDict = ConfigDict; mc.config_dict = cd
sys.modules["ml_collections"] = mc; sys.modules["ml_collections.config_dict"] = cd
sys.path.insert(0, "/root/reference")
import rl_x.algorithms.sac.pytorch.sac as refsac  # noqa: E402
import rl_x.algorithms.sac.pytorch.replay_buffer as refrb  # noqa: E402
from rl_x.algorithms.sac.pytorch.default_config import get_config  # noqa: E402
from rl_x.environments.action_space_type import ActionSpaceType  # noqa: E402
from rl_x.environments.observation_space_type import ObservationSpaceType  # noqa: E402
from rl_x.environments.data_interface_type import DataInterfaceType  # noqa: E402
import torch.distributions.normal as tnormal  # noqa: E402


class _Space:
    def __init__(self, shape, low=None, high=None, rng=None):
        self.shape, self.low, self.high, self.rng = shape, low, high, rng

    def sample(self):
        return self.rng.uniform(self.low, self.high).astype(np.float32)


class _Props:
    observation_space_type = ObservationSpaceType.FLAT_VALUES
    action_space_type = ActionSpaceType.CONTINUOUS
    data_interface_type = DataInterfaceType.NUMPY


class SyntheticNumpyEnv:
    general_properties = _Props

    def __init__(self, n, obs, act, seed, low, high):
        self.n, self.obs_dim = n, obs
        self.rng = np.random.default_rng(seed)
        self.single_observation_space = _Space((obs,))
        self.single_action_space = _Space((act,), np.full(act, low, np.float32), np.full(act, high, np.float32), np.random.default_rng(seed + 1))
        self.t = 0

    def reset(self):
        return self.rng.standard_normal((self.n, self.obs_dim)).astype(np.float32), {}

    def step(self, action):
        self.t += 1
        obs = self.rng.standard_normal((self.n, self.obs_dim)).astype(np.float32)
        rew = self.rng.standard_normal(self.n).astype(np.float32)
        term = self.rng.random(self.n) < 0.1
        trunc = np.full(self.n, self.t % 7 == 0)
        self.final = self.rng.standard_normal((self.n, self.obs_dim)).astype(np.float32)
        return obs, rew, term, trunc, {}

    def get_logging_info_dict(self, info):
        return {}

    def get_final_observation_at_index(self, info, i):
        return self.final[i]

    def get_final_info_value_at_index(self, info, key, i):
        return 0.0

    def close(self):
        pass


def run(tag, N, obs, act, hidden, batch, learning_starts, total_steps, seed, low=-2.0, high=1.0):
    torch.set_num_threads(1)
    batches, normals = [], []

    class SpyRB(refrb.ReplayBuffer):
        def sample(self, n):
            out = super().sample(n)
            batches.append([t.numpy().copy() for t in out])
            return out

    refsac.ReplayBuffer = SpyRB
    orig_sn = tnormal._standard_normal

    def spy_sn(shape, dtype, device):
        x = orig_sn(shape, dtype, device)
        normals.append(x.numpy().copy())
        return x

    tnormal._standard_normal = spy_sn
    cfg = ConfigDict(algorithm=get_config("sac.pytorch"), environment=ConfigDict(seed=seed, nr_envs=N),
                     runner=ConfigDict(save_model=False, track_console=False, track_tb=False, track_wandb=False))
    a = cfg.algorithm
    a.device, a.bf16_mixed_precision_training, a.compile_mode = "cpu", False, "default"
    a.nr_hidden_units, a.batch_size, a.learning_starts, a.total_timesteps = hidden, batch, learning_starts, total_steps
    a.buffer_size, a.logging_frequency = 4096, N
    env = SyntheticNumpyEnv(N, obs, act, seed + 100, low, high)
    model = refsac.SAC(cfg, env, env, "/tmp/golden_sac", None)
    out = {}

    def snap(prefix):
        for name, mod in [("policy", model.policy), ("q1", model.critic.q1), ("q2", model.critic.q2), ("q1_target", model.critic.q1_target),
                          ("q2_target", model.critic.q2_target)]:
            for k, v in mod.state_dict().items():
                out[f"{prefix}/{name}/{k.replace('_orig_mod.', '')}"] = v.detach().numpy().copy()
        out[f"{prefix}/log_alpha"] = model.entropy_coefficient.log_alpha.detach().numpy().copy()

    snap("init")
    metrics = []
    model.log = lambda name, value, step: metrics.append((name, float(value), int(step)))
    model.train()
    snap("final")
    tnormal._standard_normal = orig_sn
    nupd = len(batches)
    for u, b in enumerate(batches):
        for name, arr in zip(["states", "next_states", "actions", "rewards", "terminations"], b):
            out[f"batch{u}/{name}"] = arr
    # rsample draws: per update two [batch, act] draws (next-state policy, then current-state policy); acting draws are [N, act]
    upd_normals = [x for x in normals if x.shape == (batch, act)]
    assert len(upd_normals) == 2 * nupd, (len(upd_normals), nupd)
    for u in range(nupd):
        out[f"batch{u}/eps_next"], out[f"batch{u}/eps_cur"] = upd_normals[2 * u], upd_normals[2 * u + 1]
    for n in sorted({m[0] for m in metrics}):
        if n.startswith("time/"):
            continue
        out[f"metric/{n}"] = np.array([m[1] for m in metrics if m[0] == n])
    out["meta"] = np.array([N, obs, act, hidden, batch, nupd, seed], dtype=np.int64)
    out["meta_f"] = np.array([a.gamma, a.tau, a.learning_rate, a.log_std_min, a.log_std_max, -float(act), low, high], dtype=np.float64)
    path = os.path.join(HERE, f"sac_{tag}.npz")
    np.savez_compressed(path, **out)
    print(tag, "->", path, os.path.getsize(path) // 1024, "KiB; updates:", nupd, "metrics:", sorted({m[0] for m in metrics if not m[0].startswith('time/')}))


if __name__ == "__main__":
    run("small", N=4, obs=17, act=6, hidden=64, batch=32, learning_starts=40, total_steps=64, seed=2)
