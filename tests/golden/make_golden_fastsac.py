#!/usr/bin/env python
"""Golden vectors of the FastSAC update, made by EXECUTING the unmodified reference (rl_x/algorithms/fastsac/pytorch/fastsac.py) on CPU,
fp32 (bf16 autocast off).  Build container only:

    TORCHDYNAMO_DISABLE=1 python tests/golden/make_golden_fastsac.py

Captured per environment step that optimises: the sampled batch (ReplayBuffer.sample is wrapped), every standard-normal draw of
Normal.rsample() ([batch, act]: per critic update one for the next-state action, per policy update one for the current-state action),
the observation-normaliser statistics before the step, the logged metrics (logging_frequency = nr_envs: every logged value is the mean
over that step's policy updates) and, at the end, all small parameter tensors plus every 61st element of the large matrices.
Initial parameters are NOT stored (1.4 M floats): the test rebuilds them by constructing the same torch modules in the same order
under torch.manual_seed(seed), as FastSAC.__init__ does (fastsac.py:77-84).  Output: tests/golden/fastsac_update.npz, fastsac_update_clipped.npz
"""
import os
import sys

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_ppo import ConfigDict  # noqa: E402  (installs the ml_collections stub, adds the reference to sys.path)

import rl_x.algorithms.fastsac.pytorch.fastsac as ref  # noqa: E402
import rl_x.algorithms.fastsac.pytorch.replay_buffer as refrb  # noqa: E402
from rl_x.algorithms.fastsac.pytorch.default_config import get_config  # noqa: E402
from rl_x.environments.action_space_type import ActionSpaceType  # noqa: E402
from rl_x.environments.observation_space_type import ObservationSpaceType  # noqa: E402
from rl_x.environments.data_interface_type import DataInterfaceType  # noqa: E402
import torch.distributions.normal as tnormal  # noqa: E402

STRIDE = 61


class _Space:
    def __init__(self, shape, **kw):
        self.shape = shape
        self.__dict__.update(kw)


class _Props:
    observation_space_type = ObservationSpaceType.FLAT_VALUES
    action_space_type = ActionSpaceType.CONTINUOUS
    data_interface_type = DataInterfaceType.TORCH


class SyntheticTorchEnv:
    general_properties = _Props
    horizon = 5

    def __init__(self, n, obs, act, seed):
        self.n, self.obs = n, obs
        self.g = torch.Generator().manual_seed(seed)
        self.single_observation_space = _Space((obs,))
        low, high = np.full(act, -1.5, np.float32), np.full(act, 0.5, np.float32)
        self.single_action_space = _Space((act,), low=low, high=high, center=(low + high) / 2, scale=np.full(act, 0.8, np.float32))
        self.t = 0

    def reset(self):
        return torch.randn(self.n, self.obs, generator=self.g) * 2 + 1, {}

    def step(self, action):
        self.t += 1
        return (torch.randn(self.n, self.obs, generator=self.g) * 2 + 1, torch.randn(self.n, generator=self.g),
                torch.rand(self.n, generator=self.g) < 0.2, torch.full((self.n,), self.t % 4 == 0), {})

    def get_logging_info_dict(self, info):
        return {}

    def close(self):
        pass


def run(tag="update", N=4, obs=7, act=3, batch=16, n_steps=3, steps=9, seed=4, ncu=2, npu=2, clipped=False, max_grad_norm=-1.0):
    torch.set_num_threads(1)
    batches, normals, norm_states = [], [], []

    class SpyRB(refrb.ReplayBuffer):
        def sample(self, n):
            out = super().sample(n)
            batches.append([t.numpy().copy() for t in out])
            return out

    ref.ReplayBuffer = SpyRB
    orig_sn = tnormal._standard_normal

    def spy_sn(shape, dtype, device):
        x = orig_sn(shape, dtype, device)
        normals.append(x.numpy().copy())
        return x

    tnormal._standard_normal = spy_sn
    cfg = ConfigDict(algorithm=get_config("fastsac.pytorch"), environment=ConfigDict(seed=seed, nr_envs=N),
                     runner=ConfigDict(save_model=False, track_console=False, track_tb=False, track_wandb=False))
    a = cfg.algorithm
    a.device, a.bf16_mixed_precision_training, a.compile_mode = "cpu", False, "default"
    a.batch_size, a.buffer_size_per_env, a.learning_starts, a.total_timesteps, a.n_steps = batch, 8, 3, N * steps, n_steps
    a.nr_critic_updates_per_policy_update, a.nr_policy_updates_per_step, a.logging_frequency, a.save_frequency = ncu, npu, N, -1
    a.learning_rate, a.target_entropy = 1e-3, -float(act)
    a.clipped_double_q_learning, a.max_grad_norm = clipped, max_grad_norm
    env = SyntheticTorchEnv(N, obs, act, seed + 100)
    model = ref.FastSAC(cfg, env, env, "/tmp/golden_fastsac", None)
    nrm = model.observation_normalizer
    orig_normalize = nrm.normalize

    def normalize(observations, update=True):
        if update:
            norm_states.append([nrm.running_mean.numpy().copy(), nrm.running_var.numpy().copy(), int(nrm.count)])
        return orig_normalize(observations, update=update)

    nrm.normalize = normalize
    metrics = []
    model.log = lambda name, value, step: metrics.append((name, float(value), int(step)))
    model.train()
    tnormal._standard_normal = orig_sn
    out = {}
    nopt = len(batches)
    for u, b in enumerate(batches):
        for name, arr in zip(["states", "next_states", "actions", "rewards", "dones", "truncations", "effective_n_steps"], b):
            out[f"step{u}/{name}"] = arr
        out[f"step{u}/norm_mean"], out[f"step{u}/norm_var"], out[f"step{u}/norm_count"] = norm_states[2 * u][0], norm_states[2 * u][1], np.array(norm_states[2 * u][2])
    upd = [x for x in normals if x.shape == (batch, act)]
    per_step = npu * ncu + npu
    assert len(upd) == nopt * per_step, (len(upd), nopt, per_step)
    for u in range(nopt):
        out[f"step{u}/normals"] = np.stack(upd[u * per_step:(u + 1) * per_step])   # order of use: (critic x ncu, policy) x npu
    for n in sorted({m[0] for m in metrics}):
        if not n.startswith("time/"):
            out[f"metric/{n}"] = np.array([m[1] for m in metrics if m[0] == n])

    def put(prefix, sd):
        for k, v in sd.items():
            v = v.detach().numpy().reshape(-1)
            out[f"{prefix}/{k.replace('_orig_mod.', '')}"] = v.copy() if v.size <= 2048 else v[::STRIDE].copy()

    for name, mod in [("policy", model.policy), ("q1", model.critic.q1), ("q2", model.critic.q2), ("q1_target", model.critic.q1_target),
                      ("q2_target", model.critic.q2_target)]:
        put(f"final/{name}", mod.state_dict())
    out["final/log_alpha"] = model.entropy_coefficient.log_alpha.detach().numpy().copy()
    out["final/norm_mean"], out["final/norm_var"], out["final/norm_count"] = nrm.running_mean.numpy().copy(), nrm.running_var.numpy().copy(), np.array(int(nrm.count))
    out["meta"] = np.array([N, obs, act, batch, n_steps, nopt, seed, ncu, npu, a.nr_atoms, STRIDE], dtype=np.int64)
    out["meta_f"] = np.array([a.gamma, a.tau, a.learning_rate, a.log_std_min, a.log_std_max, a.target_entropy, a.v_min, a.v_max, a.weight_decay,
                              a.adam_beta1, a.adam_beta2, a.alpha_init, -1.5, 0.5, 0.8, float(clipped), max_grad_norm], dtype=np.float64)
    path = os.path.join(HERE, f"fastsac_{tag}.npz")
    np.savez_compressed(path, **out)
    print("->", path, os.path.getsize(path) // 1024, "KiB; optimising steps:", nopt, "metrics:", sorted({m[0] for m in metrics if not m[0].startswith("time/")}))
    print({n: out[f"metric/{n}"][:3] for n in ["loss/q_loss", "loss/policy_loss", "entropy/alpha"]})


if __name__ == "__main__":
    run()
    # the two non-default options: clipped double-Q targets / min-Q actor loss, and clip_grad_norm_ on both optimisers
    run("update_clipped", n_steps=1, seed=6, clipped=True, max_grad_norm=0.5)
