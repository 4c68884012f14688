"""Synthetic Box(obs)/Box(act) vector environment (BASELINE.json configs 2/3): obs ~ N(0,1), reward ~ N(0,1),
terminated ~ Bernoulli(p), optional fixed-horizon truncation, auto-reset semantics of the reference's torch-interface
envs (custom_mujoco/ant/warp_torch/environment.py:142-186: `step` returns the post-reset observation for done envs).
The simulator cost is ~0 by construction, so the PPO hot path is what gets measured."""
import numpy as np
import torch


class BoxSpace:
    def __init__(self, low, high, shape, dtype=np.float32):
        self.shape = tuple(shape)
        self.dtype = dtype
        self.low = np.full(self.shape, low, dtype=np.float32)
        self.high = np.full(self.shape, high, dtype=np.float32)

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return np.random.uniform(lo, hi).astype(np.float32)


class SyntheticBoxEnv:
    def __init__(self, env_config, seed_offset=0):
        c = env_config
        self.nr_envs, self.obs_dim, self.act_dim = int(c.nr_envs), int(c.obs_dim), int(c.act_dim)
        self.horizon, self.p_term = int(c.horizon), float(c.termination_probability)
        self.numpy_interface = c.data_interface == "numpy"
        use_cuda = (c.device == "gpu") and torch.cuda.is_available() and not self.numpy_interface
        self.device = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
        self.single_observation_space = BoxSpace(-np.inf, np.inf, (self.obs_dim,))
        self.single_action_space = BoxSpace(c.action_low, c.action_high, (self.act_dim,))
        self.gen = torch.Generator(device=self.device).manual_seed(int(c.seed) + 7919 * seed_offset)
        self.ring = c.stream == "ring"
        self.ring_length = int(c.ring_length) if self.ring else 1
        pin = self.numpy_interface and torch.cuda.is_available()
        kw = dict(device=self.device)
        self._obs = torch.empty(self.ring_length, self.nr_envs, self.obs_dim, **kw)
        self._rew = torch.empty(self.ring_length, self.nr_envs, **kw)
        self._term = torch.empty(self.ring_length, self.nr_envs, dtype=torch.bool, **kw)
        self._u = torch.empty(self.nr_envs, **kw)
        if pin:
            self._obs, self._rew, self._term = self._obs.pin_memory(), self._rew.pin_memory(), self._term.pin_memory()
        if self.ring:
            for i in range(self.ring_length):
                self._draw(i)
        self.t = 0
        self.last_action = None

    def _draw(self, slot):
        self._obs[slot].normal_(generator=self.gen)
        self._rew[slot].normal_(generator=self.gen)
        self._u.uniform_(generator=self.gen)
        torch.lt(self._u, self.p_term, out=self._term[slot])

    def _out(self, t):
        return t.numpy() if self.numpy_interface else t

    def reset(self, seed=None):
        self.t = 0
        if not self.ring:
            self._obs[0].normal_(generator=self.gen)
        return self._out(self._obs[0]), {}

    def step(self, action):
        self.last_action = action
        self.t += 1
        slot = self.t % self.ring_length
        if not self.ring:
            self._draw(0)
        truncated = self.horizon > 0 and (self.t % self.horizon == 0)
        if self.numpy_interface:
            trunc = np.full(self.nr_envs, truncated, dtype=bool)
        else:
            trunc = self._trunc_true if truncated else self._trunc_false
        return self._out(self._obs[slot]), self._out(self._rew[slot]), self._out(self._term[slot]), trunc, {}

    @property
    def _trunc_false(self):
        if not hasattr(self, "_tf"):
            self._tf = torch.zeros(self.nr_envs, dtype=torch.bool, device=self.device)
        return self._tf

    @property
    def _trunc_true(self):
        if not hasattr(self, "_tt"):
            self._tt = torch.ones(self.nr_envs, dtype=torch.bool, device=self.device)
        return self._tt

    def close(self):
        pass
