"""FastSAC's device-resident n-step replay buffer — mirrors rl_x/algorithms/fastsac/pytorch/replay_buffer.py (same constructor, `add`,
`sample` and return tuple).  The ring lives in HBM as in the reference; `sample` is ONE native launch (`rlx_replay_sample_nstep_f32`:
n-step reward accumulation, episode-boundary search and the three row gathers) instead of the reference's ~40 indexing kernels.
The index draws stay `torch.randint` on the device, as in the reference (replay_buffer.py:37-38, 63-64).

Status: the FastSAC update itself (distributional critic, fastsac.py) is not built; this module is the SURVEY §8 f4 buffer row only.
"""
import ctypes as C

import torch

from rl_x_b200 import _native as nt


class ReplayBuffer:
    def __init__(self, buffer_size_per_env, nr_envs, os_shape, as_shape, n_steps, gamma, device):
        if not torch.cuda.is_available():
            raise RuntimeError("rl_x_b200 ReplayBuffer needs a CUDA device; there is no CPU fallback.")
        if len(os_shape) != 1 or len(as_shape) != 1:
            raise ValueError("flat observations / actions only")
        if not 1 <= int(n_steps) <= 32:
            raise ValueError("n_steps must be in [1, 32]")
        self.lib = nt.load()
        self.os_shape, self.as_shape = tuple(os_shape), tuple(as_shape)
        self.capacity, self.nr_envs, self.n_steps, self.gamma, self.device = int(buffer_size_per_env), int(nr_envs), int(n_steps), gamma, device
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=device)
        self.states = z(self.capacity, self.nr_envs, *self.os_shape)
        self.next_states = z(self.capacity, self.nr_envs, *self.os_shape)
        self.actions = z(self.capacity, self.nr_envs, *self.as_shape)
        self.rewards, self.dones, self.truncations = z(self.capacity, self.nr_envs), z(self.capacity, self.nr_envs), z(self.capacity, self.nr_envs)
        self.discounts = (self.gamma ** torch.arange(self.n_steps, dtype=torch.float32, device=device)).contiguous()  # replay_buffer.py:78
        self.pos = 0
        self.size = 0

    def add(self, states, next_states, actions, rewards, dones, truncations):
        """ref: replay_buffer.py:23-31."""
        self.states[self.pos] = states
        self.next_states[self.pos] = next_states
        self.actions[self.pos] = actions
        self.rewards[self.pos] = rewards
        self.dones[self.pos] = dones
        self.truncations[self.pos] = truncations
        self.pos = (self.pos + 1) % self.capacity
        self.size = min(self.size + 1, self.capacity)

    def sample_indices(self, nr_samples):
        """the two draws of replay_buffer.py:37-38 (n_steps == 1) / :59-64 (n_steps > 1)."""
        if self.n_steps == 1:
            high = self.size
        elif self.size >= self.capacity:
            high = self.capacity
        else:
            high = max(1, self.size - self.n_steps + 1)
        idx_t = torch.randint(0, high, (nr_samples,), device=self.device)
        idx_e = torch.randint(0, self.nr_envs, (nr_samples,), device=self.device)
        return idx_t, idx_e

    def gather(self, idx_t, idx_e):
        """everything of `sample` after the index draws, in one launch."""
        n = int(idx_t.shape[0])
        obs, act = self.os_shape[0], self.as_shape[0]
        e = lambda *s: torch.empty(s, dtype=torch.float32, device=self.device)
        out = (e(n, obs), e(n, obs), e(n, act), e(n), e(n), e(n), e(n))
        scratch = torch.empty(n, dtype=torch.int64, device=self.device)  # bootstrap row of every sample (stays alive until the call is queued)
        if not (idx_t.dtype == torch.int64 and idx_e.dtype == torch.int64 and idx_t.is_contiguous() and idx_e.is_contiguous()):
            raise TypeError("indices: contiguous int64 CUDA tensors expected")
        nt.check(self.lib.rlx_replay_sample_nstep_f32(
            idx_t.data_ptr(), idx_e.data_ptr(), n, self.capacity, self.nr_envs, obs, act, self.n_steps, self.discounts.data_ptr(), self.size, self.pos,
            self.states.data_ptr(), self.next_states.data_ptr(), self.actions.data_ptr(), self.rewards.data_ptr(), self.dones.data_ptr(),
            self.truncations.data_ptr(), *[t.data_ptr() for t in out], scratch.data_ptr(),
            C.c_void_p(torch.cuda.current_stream().cuda_stream)),
            "rlx_replay_sample_nstep_f32")
        return out

    def sample(self, nr_samples):
        """ref: replay_buffer.py:34-96 -> (states, next_states, actions, rewards, dones, truncations, effective_n_steps)."""
        return self.gather(*self.sample_indices(nr_samples))
