"""ESPO (early-stopping policy optimisation) on the PPO kernels — mirrors rl_x/algorithms/espo/pytorch/espo.py.

The reference's ESPO is its PPO with a different update (espo.py:123-166, 246-299): every "epoch" is ONE minibatch drawn without
replacement from the rollout (`rng.choice`, espo.py:256), the surrogate is unclipped (espo.py:138), and the loop stops as soon as
`ratio_delta = mean|ratio - 1|` of the minibatch just applied exceeds `max_ratio_delta` (espo.py:273-274).  Acting, GAE, evaluation,
checkpoints and logging are the PPO class's (espo.py:100-118,203-243 are the same code as ppo.py).  Here:

  * indices: `rlx_pcg64_choice_i64`, bit-exact with numpy's Generator.choice,
  * gather + advantage statistics of just that minibatch (`rlx_gather_minibatch_f32`, `rlx_advantage_stats_f32`),
  * forward/backward + clip/Adam: the PPO entry points with clip_range = +inf (max(a, a) with torch's tie rule has the unclipped
    gradient) and the clip-fraction slot switched to sum|ratio - 1| (`rlx_ppo_hparams.ratio_delta_metric`),
  * one 4-byte device->host read per step for the stop test — the reference does seven `.item()` calls there (espo.py:262-270).
"""
import logging

import numpy as np
import torch

from rl_x_b200 import _native as nt
from rl_x_b200.algorithms.ppo.b200.kernels import make_hparams
from rl_x_b200.algorithms.ppo.b200.ppo import PPO

rlx_logger = logging.getLogger("rl_x")


class _AlgorithmView:
    """config.algorithm of an ESPO run, answering the two PPO-only keys the shared constructor reads."""

    def __init__(self, base, **extra):
        object.__setattr__(self, "_base", base)
        object.__setattr__(self, "_extra", extra)

    def __getattr__(self, key):
        extra = object.__getattribute__(self, "_extra")
        if key in extra:
            return extra[key]
        return getattr(object.__getattribute__(self, "_base"), key)

    def get(self, key, default=None):
        try:
            return getattr(self, key)
        except (AttributeError, KeyError):
            return default


class _ConfigView:
    def __init__(self, config, algorithm):
        self.algorithm, self.environment, self.runner = algorithm, config.environment, config.runner


class ESPO(PPO):
    def __init__(self, config, train_env, eval_env, run_path, writer):
        a = config.algorithm
        self.max_epochs = int(a.max_epochs)              # espo.py:40
        self.max_ratio_delta = float(a.max_ratio_delta)  # espo.py:44
        if a.delta_calc_operator not in ("mean", "median"):
            raise ValueError("Unknown delta_calc_operator")  # espo.py:57-63
        self.delta_calc_operator = a.delta_calc_operator  # mean: accumulated inside the loss kernel; median: radix-select kernel (torch.median = lower median)
        view = _ConfigView(config, _AlgorithmView(a, nr_epochs=self.max_epochs, clip_range=float("inf")))
        super().__init__(view, train_env, eval_env, run_path, writer)
        self.config = config  # what save() stores (espo.py:396-404)
        if self.world_size > 1:
            raise NotImplementedError("rl_x_b200 ESPO is single-GPU (the reference configuration is one small environment batch).")
        if self.minibatch_size > self.batch_size:
            raise ValueError("Cannot take a larger sample than population when 'replace=False'")  # numpy's error at espo.py:256
        self._idx_host = torch.zeros(self.minibatch_size, dtype=torch.int64).pin_memory()
        self._espo_steps = 0

    # ------------------------------------------------------------------------------------------------ hooks of the shared class
    def _make_hparams(self):
        return make_hparams(float("inf"), self.entropy_coef, self.critic_coef, self.max_grad_norm, ratio_delta_metric=self.delta_calc_operator)

    def _make_index_stream(self):
        return None  # indices are drawn step by step: whether another draw happens depends on the data (the stop rule)

    def _optimize(self):
        """ref: espo.py:246-274."""
        b, k = self.batch, self.kernels
        T, N, obs, act = self.nr_steps, self.nr_envs, k.obs_dim, k.act_dim
        flat_states = b.states[:T].view(T * N, obs)
        flat_actions = b.actions.view(T * N, act)
        lp, adv, ret = b.log_probs.view(-1), b.advantages.view(-1), b.returns.view(-1)
        mb = self.minibatch_size
        idx_dev = self.perm_dev[:mb]
        self._espo_steps = 0
        for epoch in range(self.max_epochs):
            self._idx_host.numpy()[:] = self.rng.choice(self.batch_size, mb, replace=False)  # espo.py:256
            idx_dev.copy_(self._idx_host, non_blocking=True)
            k.gather(idx_dev, flat_states, flat_actions, lp, adv, ret, self.g_states, self.g_actions, self.g_log_probs, self.g_advantages,
                     self.g_returns, count=mb, out_states_ld=self.ldx)
            k.advantage_stats(self.g_advantages, mb, mb, self.adv_stats)
            args = k.minibatch_args(
                m=mb, m_global=mb, states=self.g_states, actions=self.g_actions, log_probs=self.g_log_probs, advantages=self.g_advantages,
                returns=self.g_returns, adv_stats=self.adv_stats, params=self.params.flat, grads=self.grads, exp_avg=self.exp_avg,
                exp_avg_sq=self.exp_avg_sq, lr=self.lr_dev, step_count=self.adam_step, hp=self.hp, metrics=self.metrics_dev[epoch],
                workspace=self.train_ws, states_ld=self.ldx, states_ones_col=True)
            k.fwdbwd(args)
            k.clip_adam(args)
            self._espo_steps = epoch + 1
            ratio_delta = float(self.metrics_dev[epoch, 4])  # synchronises: the pinned index buffer is free again, too
            if ratio_delta > self.max_ratio_delta:           # espo.py:273-274
                break

    def _optimization_metrics(self, m, ev_host):
        """ref: espo.py:259-296: means over the update steps that were made."""
        m = m[:self._espo_steps]
        optimization_metrics = {
            "loss/policy_gradient_loss": m[:, 0].mean(),
            "loss/critic_loss": m[:, 1].mean(),
            "loss/entropy_loss": m[:, 2].mean(),
            "policy_ratio/ratio_delta": m[:, 4].mean(),
            "policy_ratio/approx_kl": m[:, 3].mean(),
            "gradients/policy_grad_norm": m[:, 5].mean(),
            "gradients/critic_grad_norm": m[:, 6].mean(),
        }
        optimization_metrics["optim/nr_epochs"] = self._espo_steps
        optimization_metrics["lr/learning_rate"] = self.current_learning_rate()
        optimization_metrics["v_value/explained_variance"] = np.nan if float(ev_host[0]) == 0 else float(ev_host[1])
        optimization_metrics["policy/std_dev"] = float(np.mean(np.exp(self.params.view(self.params.flat, "logstd").cpu().numpy())))
        self.nr_updates += self._espo_steps  # espo.py:298
        return optimization_metrics
