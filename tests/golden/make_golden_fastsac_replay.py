#!/usr/bin/env python
"""Golden vectors of FastSAC's n-step replay sampling, made by EXECUTING the unmodified reference class
(rl_x/algorithms/fastsac/pytorch/replay_buffer.py) on CPU.  Build container only:

    python tests/golden/make_golden_fastsac_replay.py

Cases: n_steps 1 / 3 / 5, ring partially filled and wrapped-around full (the newest-row truncation patch, :50-57).  torch.randint is
wrapped to record the index draws.  Output: tests/golden/fastsac_replay.npz.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
import importlib.util

spec = importlib.util.spec_from_file_location("ref_fastsac_replay", "/root/reference/rl_x/algorithms/fastsac/pytorch/replay_buffer.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


def run_case(out, tag, cap, nr_envs, obs, act, n_steps, gamma, adds, nr_samples, seed):
    g = torch.Generator().manual_seed(seed)
    buf = ref.ReplayBuffer(cap, nr_envs, (obs,), (act,), n_steps, gamma, "cpu")
    for _ in range(adds):
        done = (torch.rand(nr_envs, generator=g) < 0.15).float()
        trunc = (torch.rand(nr_envs, generator=g) < 0.10).float()
        buf.add(torch.randn(nr_envs, obs, generator=g), torch.randn(nr_envs, obs, generator=g), torch.randn(nr_envs, act, generator=g),
                torch.randn(nr_envs, generator=g), done, trunc)
    draws = []
    orig = torch.randint

    def randint(*a, **k):
        k.pop("device", None)
        t = orig(*a, **k, generator=g)
        draws.append(t.clone())
        return t

    torch.randint = randint
    try:
        res = buf.sample(nr_samples)
    finally:
        torch.randint = orig
    assert len(draws) == 2
    for name in ["states", "next_states", "actions", "rewards", "dones", "truncations"]:
        out[f"{tag}/ring/{name}"] = getattr(buf, name).numpy().copy()
    out[f"{tag}/idx_t"], out[f"{tag}/idx_e"] = draws[0].numpy().astype(np.int64), draws[1].numpy().astype(np.int64)
    for name, v in zip(["states", "next_states", "actions", "rewards", "dones", "truncations", "effective_n_steps"], res):
        out[f"{tag}/out/{name}"] = v.numpy().copy()
    out[f"{tag}/discounts"] = (gamma ** torch.arange(n_steps, dtype=torch.float32)).numpy()
    out[f"{tag}/meta"] = np.array([cap, nr_envs, obs, act, n_steps, buf.size, buf.pos, nr_samples], dtype=np.int64)
    print(tag, "size", buf.size, "pos", buf.pos, "mean effective n", float(res[6].mean()))


if __name__ == "__main__":
    out = {}
    run_case(out, "n1_partial", 16, 3, 5, 2, 1, 0.97, adds=9, nr_samples=64, seed=1)
    run_case(out, "n1_full", 16, 3, 5, 2, 1, 0.97, adds=21, nr_samples=64, seed=2)
    run_case(out, "n3_partial", 16, 3, 5, 2, 3, 0.97, adds=9, nr_samples=128, seed=3)
    run_case(out, "n3_full", 16, 3, 5, 2, 3, 0.97, adds=37, nr_samples=256, seed=4)
    run_case(out, "n5_full", 32, 4, 7, 3, 5, 0.9, adds=70, nr_samples=256, seed=5)
    run_case(out, "n3_short", 16, 2, 4, 2, 3, 0.99, adds=2, nr_samples=16, seed=6)  # fewer rows than n_steps: reads unwritten (zero) rows
    names = sorted({k.split("/")[0] for k in out})
    out["cases"] = np.array(names)
    path = os.path.join(HERE, "fastsac_replay.npz")
    np.savez_compressed(path, **out)
    print("->", path, os.path.getsize(path) // 1024, "KiB")
