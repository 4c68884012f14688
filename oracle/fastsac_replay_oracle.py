"""CPU oracle for FastSAC's n-step replay sampling — TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as oracle/ppo_oracle.py).

Restates `ReplayBuffer.sample` of rl_x/algorithms/fastsac/pytorch/replay_buffer.py:34-96 (nico-bohlinger/RL-X @ 46d8e26) with the two
`torch.randint` draws (:37-38, :63-64) passed in as arguments.  Written as explicit per-sample loops (numpy float32), i.e. NOT the
reference's vectorised formulation, so that agreement with the reference is a real check.

Parity status: PINNED against `tests/golden/fastsac_replay.npz`, captured by executing the reference class on CPU
(`tests/golden/make_golden_fastsac_replay.py`; checked by tests/test_oracle_vs_reference.py::test_fastsac_replay_sample_matches_reference).
"""
import numpy as np


def sample(ring, idx_t, idx_e, n_steps, gamma_discounts, size, pos):
    """ring: dict of numpy arrays states/next_states/actions [cap, nr_envs, dim], rewards/dones/truncations [cap, nr_envs].
    gamma_discounts: float32 [n_steps] = gamma ** arange(n_steps) as torch computes it.  Returns the reference's 7-tuple."""
    cap = ring["rewards"].shape[0]
    n = len(idx_t)
    f32 = np.float32
    out_s = ring["states"][idx_t, idx_e].copy()
    out_a = ring["actions"][idx_t, idx_e].copy()
    out_ns = np.empty_like(out_s)
    rew, done, trunc, eff = (np.empty(n, f32) for _ in range(4))
    full = size >= cap
    last_idx = (pos - 1) % cap
    for i in range(n):
        t0, e = int(idx_t[i]), int(idx_e[i])
        if n_steps == 1:  # replay_buffer.py:36-47
            out_ns[i] = ring["next_states"][t0, e]
            rew[i], done[i], trunc[i], eff[i] = ring["rewards"][t0, e], ring["dones"][t0, e], ring["truncations"][t0, e], 1.0
            continue
        mask, acc, count = f32(1.0), f32(0.0), f32(0.0)
        first_done = first_trunc = n_steps - 1
        seen_d = seen_t = False
        trs = []
        for j in range(n_steps):
            t = (t0 + j) % cap
            d, tr = ring["dones"][t, e], ring["truncations"][t, e]
            if full and t == last_idx:  # :50-57: the newest row of a full ring counts as truncated unless it is a done
                tr = tr if d > 0 else f32(1.0)
            trs.append(tr)
            acc = f32(acc + f32(f32(ring["rewards"][t, e] * mask) * gamma_discounts[j]))
            count = f32(count + mask)
            if not seen_d and d > 0:
                first_done, seen_d = j, True
            if not seen_t and tr > 0:
                first_trunc, seen_t = j, True
            mask = f32(mask * f32(1.0 - d))
        off = min(first_done, first_trunc)
        tf = (t0 + off) % cap
        out_ns[i] = ring["next_states"][tf, e]
        rew[i], done[i], trunc[i], eff[i] = acc, ring["dones"][tf, e], trs[off], count
    return out_s, out_ns, out_a, rew, done, trunc, eff
