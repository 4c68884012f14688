"""CPU oracle for the PPO+LSTM path — TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as oracle/ppo_oracle.py).

Restates rl_x/algorithms/ppo_lstm/flax/{policy.py,critic.py,ppo_lstm.py} (nico-bohlinger/RL-X @ 46d8e26) in plain PyTorch (CPU, fp32,
autograd).  The reference implementation is JAX/Flax; JAX, Flax and Optax are NOT in this image and not vendored under /root/reference
(pyproject.toml pins `jax[cpu]<=0.7.2`, `flax<=0.12.0`, `optax>=0.2.6`), so the reference cannot be executed here.

Parity status: **PARITY UNPINNED.**  There are no golden vectors for this path: the third-party pieces are restated from their
published definitions —
  * flax.linen.Dense            y = x @ kernel + bias, kernel [in, out]
  * flax.linen.LayerNorm        eps 1e-6, fast variance  var = max(0, E[x^2] - E[x]^2),  y = (x - mean) * (rsqrt(var + eps) * scale) + bias
  * flax.linen.OptimizedLSTMCell  same function as LSTMCell: i,f,o = sigmoid, g = tanh of  x @ W_i* + h @ W_h* + b_h*  (bias on the
                                hidden side only); c' = f*c + i*g; h' = o*tanh(c'); carry = (c, h), output = h'
  * optax.clip_by_global_norm   g if ||g|| < max_norm else g / ||g|| * max_norm          (no +1e-6 in the divisor, unlike torch)
  * optax.adam                  m_hat / (sqrt(v_hat) + 1e-8), both bias-corrected        (same function as torch.optim.Adam)
and anchored on the reference's own call sites.  What IS checked (tests/test_oracle_vs_reference.py::test_ppo_lstm_*): the LSTM cell
against torch.nn.LSTMCell, LayerNorm against torch.nn.functional.layer_norm, the optimiser against torch.optim.Adam, the episode
reset rule of forward_sequence (policy.py:127-146) by splitting sequences at the done flags, and the GAE scan (ppo_lstm.py:125-138)
against the PPO oracle's.
"""
import math

import torch

from oracle import ppo_oracle as O

LN_EPS = 1e-6
GATES = ("i", "f", "g", "o")


# ------------------------------------------------------------------------------------------------------ parameters
def _orthogonal(shape, gain, gen):
    w = torch.empty(shape[1], shape[0])  # flax kernels are [in, out]; torch's orthogonal_ works on [out, in]
    torch.nn.init.orthogonal_(w, gain=gain, generator=gen)
    return w.t().contiguous()


def init_params(obs_dim, act_dim, hidden=256, enc=128, lstm=64, std_dev=1.0, share_encoder=False, seed=0, combine="concat"):
    """Same shapes and initialiser families as Policy.setup / Critic (policy.py:47-70, critic.py:22-30).  The values are NOT the
    reference's: they come from torch's generator, not from jax.random (no JAX here)."""
    g = torch.Generator().manual_seed(seed)
    s2 = math.sqrt(2.0)

    def dense(i, o, gain):
        return {"kernel": _orthogonal((i, o), gain, g), "bias": torch.zeros(o)}

    def ln(n):
        return {"scale": torch.ones(n), "bias": torch.zeros(n)}

    pol = {"lstm_obs_encoder_dense": dense(obs_dim, enc, s2), "lstm_obs_encoder_ln": ln(enc)}
    if not share_encoder:
        pol["obs_encoder_dense"], pol["obs_encoder_ln"] = dense(obs_dim, enc, s2), ln(enc)
    cell = {}
    for k in GATES:
        cell["i" + k] = {"kernel": torch.randn(enc, lstm, generator=g) / math.sqrt(enc)}                      # lecun_normal
        cell["h" + k] = {"kernel": _orthogonal((lstm, lstm), 1.0, g), "bias": torch.zeros(lstm)}              # orthogonal, zeros
    pol["lstm"], pol["lstm_ln"] = cell, ln(lstm)
    if combine == "film":  # policy.py:57-59
        pol["lstm_film_gamma"], pol["lstm_film_beta"] = dense(lstm, enc, s2), dense(lstm, enc, s2)
    pol["torso_dense1"], pol["torso_dense2"] = dense(enc if combine == "film" else enc + lstm, hidden, s2), dense(hidden, hidden, s2)
    pol["mean_head"] = dense(hidden, act_dim, 0.01)
    pol["policy_logstd"] = torch.full((1, act_dim), math.log(std_dev))
    cri = {"Dense_0": dense(obs_dim, hidden, s2), "Dense_1": dense(hidden, hidden, s2), "Dense_2": dense(hidden, 1, 1.0)}
    return pol, cri


def tree_leaves(tree, prefix=""):
    """[(dotted name, tensor)] in a fixed order."""
    out = []
    for k in sorted(tree):
        v = tree[k]
        out.extend(tree_leaves(v, prefix + k + ".") if isinstance(v, dict) else [(prefix + k, v)])
    return out


def tree_map(fn, tree):
    return {k: (tree_map(fn, v) if isinstance(v, dict) else fn(v)) for k, v in tree.items()}


# ----------------------------------------------------------------------------------------------------------- layers
def dense(p, x):
    return x @ p["kernel"] + p["bias"]


def layer_norm(p, x):
    """flax.linen.LayerNorm with its defaults (epsilon=1e-6, use_fast_variance=True)."""
    mean = x.mean(-1, keepdim=True)
    var = torch.clamp((x * x).mean(-1, keepdim=True) - mean * mean, min=0.0)
    mul = torch.rsqrt(var + LN_EPS) * p["scale"]
    return (x - mean) * mul + p["bias"]


def lstm_cell(p, carry, x):
    """flax.linen.OptimizedLSTMCell.__call__(carry=(c, h), inputs) -> ((c', h'), h')."""
    c, h = carry
    z = {k: x @ p["i" + k]["kernel"] + h @ p["h" + k]["kernel"] + p["h" + k]["bias"] for k in GATES}
    i, f, g, o = torch.sigmoid(z["i"]), torch.sigmoid(z["f"]), torch.tanh(z["g"]), torch.sigmoid(z["o"])
    new_c = f * c + i * g
    new_h = o * torch.tanh(new_c)
    return (new_c, new_h), new_h


def encode(pol, obs, which):
    """policy.py:79-92."""
    return torch.tanh(layer_norm(pol[which + "_ln"], dense(pol[which + "_dense"], obs)))


def decode(pol, obs_latent, lstm_latent):
    """policy.py:95-112; lstm_obs_combine_method is "film" iff the tree holds the FiLM layers (policy.py:57-59), else "concat"."""
    lstm_latent = torch.tanh(layer_norm(pol["lstm_ln"], lstm_latent))
    if "lstm_film_gamma" in pol:
        torso_in = obs_latent * dense(pol["lstm_film_gamma"], lstm_latent) + dense(pol["lstm_film_beta"], lstm_latent)
    else:
        torso_in = torch.cat([obs_latent, lstm_latent], dim=-1)
    h = torch.tanh(dense(pol["torso_dense1"], torso_in))
    h = torch.tanh(dense(pol["torso_dense2"], h))
    return dense(pol["mean_head"], h), pol["policy_logstd"]


def apply_one_step(pol, obs, carry):
    """policy.py:115-125.  obs [N, obs]; carry ([N, L], [N, L])."""
    lstm_obs_latent = encode(pol, obs, "lstm_obs_encoder")
    carry, hidden = lstm_cell(pol["lstm"], carry, lstm_obs_latent)
    obs_latent = lstm_obs_latent if "obs_encoder_dense" not in pol else encode(pol, obs, "obs_encoder")
    mean, logstd = decode(pol, obs_latent, hidden)
    return mean, logstd, carry


def forward_sequence(pol, obs_seq, done_seq, init_carry):
    """policy.py:128-146, batched over envs instead of vmapped: obs_seq [T, N, obs], done_seq [T, N] (done AFTER step t),
    init_carry ([N, L], [N, L]).  The carry is zeroed before step t when the episode ended at t-1."""
    T = obs_seq.shape[0]
    done_prev = torch.cat([torch.zeros_like(done_seq[:1]), done_seq[:-1]], dim=0).to(obs_seq.dtype)
    carry, means = init_carry, []
    for t in range(T):
        keep = (1.0 - done_prev[t]).unsqueeze(-1)
        carry = (carry[0] * keep, carry[1] * keep)
        mean_t, _, carry = apply_one_step(pol, obs_seq[t], carry)
        means.append(mean_t)
    return torch.stack(means), pol["policy_logstd"]


def critic_value(cri, x):
    """critic.py:22-30."""
    h = torch.tanh(dense(cri["Dense_0"], x))
    h = torch.tanh(dense(cri["Dense_1"], h))
    return dense(cri["Dense_2"], h)


# ------------------------------------------------------------------------------------------------------- acting / GAE
def get_action_and_value(pol, cri, state, carry, noise, act_low, act_high, clip_rescale=True):
    """ppo_lstm.py:107-118 with the normal draw passed in (jax.random is not reproducible here)."""
    mean, logstd, next_carry = apply_one_step(pol, state, carry)
    std = torch.exp(logstd)
    action = mean + std * noise
    log_prob = (-0.5 * ((action - mean) / std) ** 2 - 0.5 * math.log(2.0 * math.pi) - logstd).sum(1)
    value = critic_value(cri, state).reshape(-1)
    processed = act_low + 0.5 * (torch.clamp(action, -1, 1) + 1.0) * (act_high - act_low) if clip_rescale else action
    return processed, action, value, log_prob, next_carry


def gae(rewards, terminations, values, next_values, gamma, gae_lambda):
    """ppo_lstm.py:125-138: the same recurrence as PPO's (ppo.py:110-118), written as a scan from delta[-1]."""
    return O.gae(rewards, terminations, values, next_values, gamma, gae_lambda)


# -------------------------------------------------------------------------------------------------------------- update
def loss_fn(pol, cri, states, actions, log_probs, returns, advantages_norm, dones, init_carry, clip_range, entropy_coef, critic_coef):
    """ppo_lstm.py:143-184 for one minibatch of envs; [T, n_env, ...] tensors; means over (T, n_env) as `mean_vmapped_loss_fn` takes.
    (The reference's vmapped per-env code broadcasts the (1, act) log-std against [T, act] and so carries a replicated [T, T] block
    whose mean equals the plain per-step mean restated here.)"""
    mean, logstd = forward_sequence(pol, states, dones, init_carry)
    std = torch.exp(logstd)
    new_log_prob = (-0.5 * ((actions - mean) / std) ** 2 - 0.5 * math.log(2.0 * math.pi) - logstd).sum(-1)
    entropy_loss = (logstd + 0.5 * math.log(2.0 * math.pi * math.e)).sum(-1)  # [1]
    logratio = new_log_prob - log_probs
    ratio = torch.exp(logratio)
    approx_kl = (ratio - 1) - logratio
    clip_fraction = (torch.abs(ratio - 1) > clip_range).float()
    pg_loss = torch.maximum(-advantages_norm * ratio, -advantages_norm * torch.clamp(ratio, 1 - clip_range, 1 + clip_range))
    new_value = critic_value(cri, states).squeeze(-1)
    critic_loss = 0.5 * (new_value - returns) ** 2
    loss = (pg_loss - entropy_coef * entropy_loss + critic_coef * critic_loss).mean()
    metrics = {"loss/policy_gradient_loss": pg_loss.mean().item(), "loss/critic_loss": critic_loss.mean().item(),
               "loss/entropy_loss": entropy_loss.mean().item(), "policy_ratio/approx_kl": approx_kl.mean().item(),
               "policy_ratio/clip_fraction": clip_fraction.mean().item()}
    return loss, metrics


class OptaxAdam:
    """optax.chain(clip_by_global_norm(max_norm), adam(lr)) on one parameter tree (ppo_lstm.py:88-103)."""

    def __init__(self, leaves, lr, max_norm, b1=0.9, b2=0.999, eps=1e-8):
        self.leaves, self.lr, self.max_norm, self.b1, self.b2, self.eps = leaves, lr, max_norm, b1, b2, eps
        self.mu = [torch.zeros_like(p) for p in leaves]
        self.nu = [torch.zeros_like(p) for p in leaves]
        self.count = 0

    def step(self, grads, lr=None):
        lr = self.lr if lr is None else lr
        g_norm = torch.sqrt(sum((g * g).sum() for g in grads))
        if not bool(g_norm < self.max_norm):
            grads = [g / g_norm * self.max_norm for g in grads]
        self.count += 1
        bc1, bc2 = 1 - self.b1 ** self.count, 1 - self.b2 ** self.count
        with torch.no_grad():
            for p, g, m, v in zip(self.leaves, grads, self.mu, self.nu):
                m.mul_(self.b1).add_(g, alpha=1 - self.b1)
                v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
                p.add_(-(lr * (m / bc1) / (torch.sqrt(v / bc2) + self.eps)))
        return float(g_norm)


class Learner:
    def __init__(self, pol, cri, lr=3e-4, clip_range=0.2, entropy_coef=0.0, critic_coef=0.5, max_grad_norm=0.5):
        self.pol = tree_map(lambda v: v.clone().requires_grad_(True), pol)
        self.cri = tree_map(lambda v: v.clone().requires_grad_(True), cri)
        self.pleaves = [v for _, v in tree_leaves(self.pol)]
        self.cleaves = [v for _, v in tree_leaves(self.cri)]
        self.popt, self.copt = OptaxAdam(self.pleaves, lr, max_grad_norm), OptaxAdam(self.cleaves, lr, max_grad_norm)
        self.clip_range, self.entropy_coef, self.critic_coef = clip_range, entropy_coef, critic_coef

    def grads(self, mb):
        """mb: dict of [T, n_env, ...] tensors + init_carry; advantages raw (normalised here, ppo_lstm.py:196-197: jnp.std has ddof 0)."""
        adv = mb["advantages"]
        adv = (adv - adv.mean()) / (adv.std(unbiased=False) + 1e-8)
        loss, metrics = loss_fn(self.pol, self.cri, mb["states"], mb["actions"], mb["log_probs"], mb["returns"], adv, mb["dones"],
                                mb["init_carry"], self.clip_range, self.entropy_coef, self.critic_coef)
        g = torch.autograd.grad(loss, self.pleaves + self.cleaves, allow_unused=True)
        g = [torch.zeros_like(p) if x is None else x for x, p in zip(g, self.pleaves + self.cleaves)]
        return g[:len(self.pleaves)], g[len(self.pleaves):], metrics

    def minibatch_step(self, mb, lr=None):
        gp, gc, metrics = self.grads(mb)
        metrics["gradients/policy_grad_norm"] = self.popt.step(gp, lr)
        metrics["gradients/critic_grad_norm"] = self.copt.step(gc, lr)
        return metrics

    def update(self, batch, env_index_rows, lr=None):
        """ppo_lstm.py:186-222.  batch: [T, N, ...] tensors + init_policy_carry ([N, L], [N, L]); env_index_rows: the rows of
        `batch_env_indices` (jax.random.permutation output in the reference — an input here)."""
        out = []
        for idx in env_index_rows:
            idx = torch.as_tensor(idx)
            mb = {k: batch[k][:, idx] for k in ["states", "actions", "log_probs", "returns", "advantages", "dones"]}
            mb["init_carry"] = tuple(c[idx] for c in batch["init_policy_carry"])
            out.append(self.minibatch_step(mb, lr))
        return out
