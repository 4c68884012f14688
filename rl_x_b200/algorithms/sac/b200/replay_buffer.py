import numpy as np
import torch

from rl_x_b200 import _native as nt


class ReplayBuffer:
    """Device-resident replay ring with the reference's layout and sampling stream (rl_x/algorithms/sac/pytorch/replay_buffer.py:5-40):
    arrays are [capacity // nr_envs, nr_envs, dim]; `sample` draws idx1 = rng.integers(size, n) and idx2 = rng.integers(nr_envs, n) from
    the numpy-compatible PCG64 stream on the host (bit-exact), uploads the 2 x n int64 indices and gathers the rows with one kernel
    (the reference fancy-indexes host arrays and does five H2D copies per update)."""

    def __init__(self, capacity, nr_envs, os_shape, as_shape, rng, device):
        self.os_shape, self.as_shape = os_shape, as_shape
        self.capacity = capacity // nr_envs
        self.nr_envs, self.rng, self.device = nr_envs, rng, device
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=device)
        self.states = z(self.capacity, nr_envs, *os_shape)
        self.next_states = z(self.capacity, nr_envs, *os_shape)
        self.actions = z(self.capacity, nr_envs, *as_shape)
        self.rewards = z(self.capacity, nr_envs)
        self.terminations = z(self.capacity, nr_envs)
        self.pos = 0
        self.size = 0
        self._lib = nt.load()
        self._idx_host = None
        self._out = None

    def _dev(self, x, dtype=torch.float32):
        if torch.is_tensor(x):
            return x.to(self.device, dtype)
        return torch.from_numpy(np.ascontiguousarray(x)).to(self.device, dtype)

    def add(self, states, next_states, actions, rewards, terminations):
        self.states[self.pos].copy_(self._dev(states))
        self.next_states[self.pos].copy_(self._dev(next_states))
        self.actions[self.pos].copy_(self._dev(actions))
        self.rewards[self.pos].copy_(self._dev(rewards))
        self.terminations[self.pos].copy_(self._dev(terminations))
        self.pos = (self.pos + 1) % self.capacity
        self.size = min(self.size + 1, self.capacity)

    def sample(self, nr_samples):
        n = int(nr_samples)
        if self._out is None or self._out[0].shape[0] != n:
            z = lambda *s: torch.empty(*s, dtype=torch.float32, device=self.device)
            self._out = (z(n, *self.os_shape), z(n, *self.os_shape), z(n, *self.as_shape), z(n), z(n))
            # two pinned staging slots, each guarded by the event of its last H2D copy: the host never rewrites a slot whose copy is
            # still queued (back-to-back sample() calls without a sync in between, e.g. CUDA-graph replays of the update)
            self._idx_host = [torch.empty(2, n, dtype=torch.int64).pin_memory() for _ in range(2)]
            self._idx_event = [torch.cuda.Event(), torch.cuda.Event()]
            self._idx_pending = [False, False]
            self._idx_slot = 0
            self._idx_dev = torch.empty(2, n, dtype=torch.int64, device=self.device)
        slot = self._idx_slot
        self._idx_slot ^= 1
        if self._idx_pending[slot]:
            self._idx_event[slot].synchronize()
        host = self._idx_host[slot]
        host[0].numpy()[:] = self.rng.integers(self.size, n)
        host[1].numpy()[:] = self.rng.integers(self.nr_envs, n)
        self._idx_dev.copy_(host, non_blocking=True)
        self._idx_event[slot].record()
        self._idx_pending[slot] = True
        s, ns, a, r, t = self._out
        obs_dim, act_dim = int(np.prod(self.os_shape)), int(np.prod(self.as_shape))
        nt.check(self._lib.rlx_replay_sample_gather_f32(self._idx_dev[0].data_ptr(), self._idx_dev[1].data_ptr(), n, self.nr_envs, obs_dim, act_dim,
                                                        self.states.data_ptr(), self.next_states.data_ptr(), self.actions.data_ptr(),
                                                        self.rewards.data_ptr(), self.terminations.data_ptr(), s.data_ptr(), ns.data_ptr(), a.data_ptr(),
                                                        r.data_ptr(), t.data_ptr(), torch.cuda.current_stream().cuda_stream), "rlx_replay_sample_gather_f32")
        return s, ns, a, r, t
