"""CPU oracle for the PPO hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu-baseline / `--impl reference` legs may import this module;
nothing under `rl_x_b200/` does (the product path fails loudly without its CUDA library).

It is a plain-PyTorch (CPU, fp32, autograd) restatement of the reference algorithm, function by function, each citing the
reference file:line it follows (paths relative to the RL-X repository root):

    rl_x/algorithms/ppo/pytorch/policy.py, critic.py, ppo.py   (nico-bohlinger/RL-X @ 46d8e26)

Parity status: PINNED.  The reference ships no tests or golden vectors of its own, so the oracle is pinned
against outputs of the reference itself, executed in the build container by `tests/golden/make_golden_ppo.py`
(fixtures `tests/golden/ppo_*.npz`; checked by `tests/test_oracle_vs_reference.py`): GAE advantages/returns bit-exact,
teacher-forced log-probs/values, every shuffled index array bit-exact, and post-update weights / Adam moments / logged losses.

Third-party arithmetic (not under the reference tree): torch (Linear/Tanh/Normal/autograd/clip_grad_norm_/optim.Adam,
unpinned by the reference, torch 2.11.0 here) and numpy (default_rng/PCG64/Generator.shuffle, numpy>=2.2.6, 2.3.5 here).
The oracle calls the same torch primitives the reference calls, and restates numpy's shuffle in `pcg64_shuffle_py`.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

POLICY_KEYS = ["policy_mean.0.weight", "policy_mean.0.bias", "policy_mean.2.weight", "policy_mean.2.bias",
               "policy_mean.4.weight", "policy_mean.4.bias", "policy_logstd"]
CRITIC_KEYS = ["critic.0.weight", "critic.0.bias", "critic.2.weight", "critic.2.bias", "critic.4.weight", "critic.4.bias"]


# ----------------------------------------------------------------------------------------------------- networks
def init_params(obs_dim, act_dim, hidden, std_dev=1.0, seed=0):
    """Orthogonal init, gains sqrt(2), sqrt(2), 0.01 (policy) / 1.0 (critic), zero bias, logstd = log(std_dev).
    ref: policy.py:45-58, critic.py:29-41."""
    g = torch.Generator().manual_seed(seed)

    def lin(out_f, in_f, gain):
        w = torch.empty(out_f, in_f)
        torch.nn.init.orthogonal_(w, gain, generator=g)
        return w, torch.zeros(out_f)

    pol, cri = {}, {}
    for i, (o, inp, gain) in zip((0, 2, 4), ((hidden, obs_dim, math.sqrt(2)), (hidden, hidden, math.sqrt(2)), (act_dim, hidden, 0.01))):
        pol[f"policy_mean.{i}.weight"], pol[f"policy_mean.{i}.bias"] = lin(o, inp, gain)
    pol["policy_logstd"] = torch.full((1, act_dim), math.log(std_dev))
    for i, (o, inp, gain) in zip((0, 2, 4), ((hidden, obs_dim, math.sqrt(2)), (hidden, hidden, math.sqrt(2)), (1, hidden, 1.0))):
        cri[f"critic.{i}.weight"], cri[f"critic.{i}.bias"] = lin(o, inp, gain)
    return pol, cri


def policy_mean(pol, x):
    """ref: policy.py:45-51 (nn.Sequential Linear-Tanh-Linear-Tanh-Linear)."""
    h = torch.tanh(F.linear(x, pol["policy_mean.0.weight"], pol["policy_mean.0.bias"]))
    h = torch.tanh(F.linear(h, pol["policy_mean.2.weight"], pol["policy_mean.2.bias"]))
    return F.linear(h, pol["policy_mean.4.weight"], pol["policy_mean.4.bias"])


def critic_value(cri, x):
    """ref: critic.py:29-35,44-46."""
    h = torch.tanh(F.linear(x, cri["critic.0.weight"], cri["critic.0.bias"]))
    h = torch.tanh(F.linear(h, cri["critic.2.weight"], cri["critic.2.bias"]))
    return F.linear(h, cri["critic.4.weight"], cri["critic.4.bias"])


def get_action_logprob(pol, x, noise, act_low, act_high, clip_rescale=True):
    """ref: policy.py:61-73.  `noise` replaces the draw inside Normal.sample() (= loc + scale * randn)."""
    mean = policy_mean(pol, x)
    std = torch.exp(pol["policy_logstd"].expand_as(mean))
    probs = torch.distributions.Normal(mean, std)
    action = mean + std * noise
    if clip_rescale:
        clipped = torch.clip(action, -1, 1)
        env_action = act_low + (0.5 * (clipped + 1.0) * (act_high - act_low))
    else:
        env_action = action
    return action, env_action, probs.log_prob(action).sum(1)


def get_logprob_entropy(pol, x, action):
    """ref: policy.py:76-82."""
    mean = policy_mean(pol, x)
    std = torch.exp(pol["policy_logstd"].expand_as(mean))
    probs = torch.distributions.Normal(mean, std)
    return probs.log_prob(action).sum(1), probs.entropy().sum(1)


def get_deterministic_action(pol, x, act_low, act_high, clip_rescale=True):
    """ref: policy.py:85-93."""
    action = policy_mean(pol, x)
    if clip_rescale:
        clipped = torch.clip(action, -1, 1)
        return act_low + (0.5 * (clipped + 1.0) * (act_high - act_low))
    return action


# ---------------------------------------------------------------------------------------------------------- GAE
def gae(rewards, terminations, values, next_values, gamma, gae_lambda):
    """ref: ppo.py:110-118 (same expression order; gamma * gae_lambda multiplied as Python floats)."""
    delta = rewards + gamma * next_values * (1 - terminations) - values
    advantages = torch.zeros_like(rewards)
    lastgaelam = torch.zeros_like(rewards[0])
    for t in range(values.shape[0] - 1, -1, -1):
        lastgaelam = advantages[t] = delta[t] + gamma * gae_lambda * (1 - terminations[t]) * lastgaelam
    returns = advantages + values
    return advantages, returns


# ----------------------------------------------------------------------------------------------- bf16 autocast variant
def autocast_bf16(enabled):
    """The reference's mixed-precision mode (`bf16_mixed_precision_training`, ppo.py:123,155,208,253): torch autocast around the forward and
    loss computations, backward / clipping / Adam outside.  The reference hard-codes device_type="cuda" and refuses the mode elsewhere
    (ppo.py:66-67); here — and in tests/golden/make_golden_ppo.py's bf16 run — the CPU autocast context stands in for it: the cast rules
    that matter on this path are the same (Linear runs in bf16 with bf16 outputs, tanh stays in its input dtype, everything that mixes
    a bf16 and an fp32 tensor promotes to fp32)."""
    return torch.autocast("cpu", dtype=torch.bfloat16, enabled=enabled)


def gae_mixed_precision(rewards, terminations, values, next_values, gamma, gae_lambda):
    """ref: calculate_gae_advantages_and_returns_mixed_precision (ppo.py:98-107).  `next_values` arrives as a bf16 tensor, so
    `gamma * next_values` is rounded to bf16 before it meets the fp32 operands; the recurrence itself is fp32."""
    delta = rewards + gamma * next_values * (1 - terminations) - values
    advantages = torch.zeros_like(rewards)
    lastgaelam = torch.zeros_like(rewards[0])
    for t in range(values.shape[0] - 1, -1, -1):
        lastgaelam = advantages[t] = delta[t] + gamma * gae_lambda * (1 - terminations[t]) * lastgaelam
    return advantages, advantages + values


# ----------------------------------------------------------------------------------------------- losses / update
def policy_loss(pol, states, actions, log_probs, advantages, clip_range, entropy_coef):
    """ref: ppo.py:124-141. Returns (loss, pg_loss, entropy_loss, approx_kl, clip_fraction)."""
    new_log_prob, entropy = get_logprob_entropy(pol, states, actions)
    logratio = new_log_prob - log_probs
    ratio = logratio.exp()
    with torch.no_grad():
        approx_kl = torch.mean((torch.exp(logratio) - 1) - logratio)
        clip_fraction = torch.mean((torch.abs(ratio - 1) > clip_range).float())
    adv = (advantages - advantages.mean()) / (advantages.std() + 1e-8)
    pg_loss1 = -adv * ratio
    pg_loss2 = -adv * torch.clamp(ratio, 1 - clip_range, 1 + clip_range)
    pg_loss = torch.maximum(pg_loss1, pg_loss2).mean()
    entropy_loss = entropy.mean()
    loss = pg_loss - entropy_coef * entropy_loss
    return loss, pg_loss, entropy_loss, approx_kl, clip_fraction


def critic_loss(cri, states, returns, critic_coef):
    """ref: ppo.py:156-157."""
    new_value = critic_value(cri, states).reshape(-1)
    return critic_coef * (0.5 * (new_value - returns) ** 2).mean()


class Learner:
    """Policy + critic + two Adam optimisers, as PPO.__init__ builds them (ref: ppo.py:79-88)."""

    def __init__(self, pol, cri, lr=3e-4, clip_range=0.2, entropy_coef=0.0, critic_coef=0.5, max_grad_norm=0.5, bf16=False):
        self.bf16 = bf16  # bf16_mixed_precision_training: the two loss functions run under autocast (ppo.py:123,155)
        self.pol = {k: v.clone().requires_grad_(True) for k, v in pol.items()}
        self.cri = {k: v.clone().requires_grad_(True) for k, v in cri.items()}
        self.popt = torch.optim.Adam([self.pol[k] for k in POLICY_KEYS], lr=lr)
        self.copt = torch.optim.Adam([self.cri[k] for k in CRITIC_KEYS], lr=lr)
        self.clip_range, self.entropy_coef, self.critic_coef, self.max_grad_norm = clip_range, entropy_coef, critic_coef, max_grad_norm

    def set_lr(self, lr):
        for opt in (self.popt, self.copt):
            for g in opt.param_groups:
                g["lr"] = lr

    def grads(self, states, actions, log_probs, advantages, returns):
        """Gradients of one minibatch without stepping (for kernel-level parity checks)."""
        self.popt.zero_grad()
        self.copt.zero_grad()
        loss, pg, ent, kl, cf = policy_loss(self.pol, states, actions, log_probs, advantages, self.clip_range, self.entropy_coef)
        loss.backward()
        closs = critic_loss(self.cri, states, returns, self.critic_coef)
        closs.backward()
        gp = {k: self.pol[k].grad.clone() for k in POLICY_KEYS}
        gc = {k: self.cri[k].grad.clone() for k in CRITIC_KEYS}
        return gp, gc, dict(pg_loss=pg.item(), critic_loss=closs.item(), entropy_loss=ent.item(), approx_kl=kl.item(), clip_fraction=cf.item())

    def minibatch_step(self, states, actions, log_probs, advantages, returns):
        """ref: policy_loss_fn + critic_loss_fn (ppo.py:121-166): backward, clip_grad_norm_, Adam.step for each net."""
        self.popt.zero_grad()
        with autocast_bf16(self.bf16):
            loss, pg, ent, kl, cf = policy_loss(self.pol, states, actions, log_probs, advantages, self.clip_range, self.entropy_coef)
        loss.backward()
        pnorm = torch.nn.utils.clip_grad_norm_([self.pol[k] for k in POLICY_KEYS], self.max_grad_norm)
        self.popt.step()
        self.copt.zero_grad()
        with autocast_bf16(self.bf16):
            closs = critic_loss(self.cri, states, returns, self.critic_coef)
        closs.backward()
        cnorm = torch.nn.utils.clip_grad_norm_([self.cri[k] for k in CRITIC_KEYS], self.max_grad_norm)
        self.copt.step()
        return dict(pg_loss=pg.item(), critic_loss=closs.item(), entropy_loss=ent.item(), approx_kl=kl.item(),
                    clip_fraction=cf.item(), policy_grad_norm=pnorm.item(), critic_grad_norm=cnorm.item())

    def update(self, batch, perms, minibatch_size):
        """ref: ppo.py:265-294.  batch: dict of flattened (T*N, ...) tensors; perms: list of index arrays, one per epoch."""
        out = []
        B = batch["states"].shape[0]
        for perm in perms:
            for start in range(0, B, minibatch_size):
                idx = torch.as_tensor(perm[start:start + minibatch_size])
                out.append(self.minibatch_step(batch["states"][idx], batch["actions"][idx], batch["log_probs"][idx],
                                               batch["advantages"][idx], batch["returns"][idx]))
        return out


# ---------------------------------------------------------------------------------------------- numpy RNG restatement
_PCG_MULT = 0x2360ED051FC65DA44385DF649FCCF645
_M128 = (1 << 128) - 1


class Pcg64Py:
    """Pure-Python restatement of numpy's PCG64 + next_uint32 buffering; small cases only.
    State is taken from numpy's own seeding (np.random.default_rng(seed).bit_generator.state)."""

    def __init__(self, seed):
        st = np.random.default_rng(seed).bit_generator.state
        self.state, self.inc = st["state"]["state"], st["state"]["inc"]
        self.has_uint32, self.uinteger = st["has_uint32"], st["uinteger"]

    def next64(self):
        self.state = (self.state * _PCG_MULT + self.inc) & _M128
        hi, lo = self.state >> 64, self.state & ((1 << 64) - 1)
        x, rot = hi ^ lo, self.state >> 122
        return ((x >> rot) | (x << ((64 - rot) & 63))) & ((1 << 64) - 1)

    def next32(self):
        if self.has_uint32:
            self.has_uint32 = 0
            return self.uinteger
        v = self.next64()
        self.has_uint32, self.uinteger = 1, v >> 32
        return v & 0xFFFFFFFF


def pcg64_shuffle_py(rng, a):
    """Generator.shuffle on a 1-D array: Fisher-Yates from the top with masked rejection (ref call site: ppo.py:276)."""
    n = len(a)
    for i in range(n - 1, 0, -1):
        mask = i
        for s in (1, 2, 4, 8, 16, 32):
            mask |= mask >> s
        while True:
            j = rng.next32() & mask
            if j <= i:
                break
        a[i], a[j] = a[j], a[i]
    return a


# ------------------------------------------------------------------------------------------- whole-iteration loop
class SyntheticVecEnv:
    """Synthetic Box(obs)/Box(act) vector env with the TORCH data interface on CPU tensors: obs ~ N(0,1), reward ~ N(0,1),
    terminated ~ Bernoulli(p).  Test scaffolding shared by the oracle timing loop (SURVEY.md §8 d)."""

    def __init__(self, nr_envs, obs_dim, act_dim, seed=1, p_term=0.01):
        self.nr_envs, self.obs_dim, self.act_dim, self.p_term = nr_envs, obs_dim, act_dim, p_term
        self.gen = torch.Generator().manual_seed(seed)
        self.act_low = torch.full((act_dim,), -1.0)
        self.act_high = torch.full((act_dim,), 1.0)

    def reset(self):
        return torch.randn(self.nr_envs, self.obs_dim, generator=self.gen)

    def step(self, action):
        obs = torch.randn(self.nr_envs, self.obs_dim, generator=self.gen)
        rew = torch.randn(self.nr_envs, generator=self.gen)
        term = torch.rand(self.nr_envs, generator=self.gen) < self.p_term
        return obs, rew, term, torch.zeros(self.nr_envs, dtype=torch.bool)


def rollout(learner, env, state, nr_steps, gen):
    """ref: ppo.py:203-246 (acting) with the TORCH-interface branch (actual_next_state = next_state)."""
    N = env.nr_envs
    states = torch.zeros(nr_steps, N, env.obs_dim)
    next_states = torch.zeros(nr_steps, N, env.obs_dim)
    actions = torch.zeros(nr_steps, N, env.act_dim)
    rewards, values, terms, log_probs = (torch.zeros(nr_steps, N) for _ in range(4))
    with torch.no_grad():
        for step in range(nr_steps):
            noise = torch.randn(N, env.act_dim, generator=gen)
            action, env_action, logp = get_action_logprob(learner.pol, state, noise, env.act_low, env.act_high)
            value = critic_value(learner.cri, state)
            next_state, reward, terminated, truncated = env.step(env_action)
            states[step], next_states[step], actions[step] = state, next_state, action
            rewards[step], values[step], terms[step], log_probs[step] = reward, value.reshape(-1), terminated.float(), logp
            state = next_state
    return dict(states=states, next_states=next_states, actions=actions, rewards=rewards, values=values,
                terminations=terms, log_probs=log_probs), state


def advantages_and_returns(learner, batch, gamma, gae_lambda):
    """ref: ppo.py:253-258."""
    with torch.no_grad():
        next_values = critic_value(learner.cri, batch["next_states"]).squeeze(-1)
        return gae(batch["rewards"], batch["terminations"], batch["values"], next_values, gamma, gae_lambda)


def flatten(batch):
    """ref: ppo.py:265-270."""
    return {k: v.reshape((-1,) + v.shape[2:]) for k, v in batch.items()}
