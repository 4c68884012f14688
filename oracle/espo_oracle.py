"""CPU oracle for the ESPO update — TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as oracle/ppo_oracle.py).

ESPO (rl_x/algorithms/espo/pytorch/espo.py @ 46d8e26) shares PPO's networks, rollout and GAE (oracle/ppo_oracle.py restates those);
what differs is the update (espo.py:123-166, 246-275), restated here with the same torch primitives:

  * every "epoch" is ONE minibatch drawn with `rng.choice(batch_size, size=minibatch_size, replace=False)` (espo.py:256),
  * the surrogate is unclipped, `pg_loss = mean(-A_norm * ratio)` (espo.py:138),
  * `ratio_delta = mean|ratio - 1|` (or median, espo.py:57-63,133) is logged and the loop stops as soon as it exceeds
    `max_ratio_delta` (espo.py:273-274) — after that step has been applied.

Parity status: PINNED against `tests/golden/espo_small.npz`, captured from the executed reference by
`tests/golden/make_golden_espo.py` (checked by tests/test_oracle_vs_reference.py::test_espo_update_matches_reference).
"""
import torch

from oracle import ppo_oracle as O


def policy_loss(pol, states, actions, log_probs, advantages, entropy_coef, delta_op=torch.mean):
    """ref: espo.py:125-141. Returns (loss, pg_loss, entropy_loss, approx_kl, ratio_delta)."""
    new_log_prob, entropy = O.get_logprob_entropy(pol, states, actions)
    logratio = new_log_prob - log_probs
    ratio = logratio.exp()
    with torch.no_grad():
        approx_kl = torch.mean((torch.exp(logratio) - 1) - logratio)
        ratio_delta = delta_op(torch.abs(ratio - 1))
    adv = (advantages - advantages.mean()) / (advantages.std() + 1e-8)
    pg_loss = (-adv * ratio).mean()
    entropy_loss = entropy.mean()
    loss = pg_loss - entropy_coef * entropy_loss
    return loss, pg_loss, entropy_loss, approx_kl, ratio_delta


class Learner(O.Learner):
    """Policy + critic + two Adam optimisers as ESPO.__init__ builds them (espo.py:84-92)."""

    def __init__(self, pol, cri, lr=3e-4, entropy_coef=0.0, critic_coef=0.5, max_grad_norm=0.5, max_ratio_delta=0.25, delta_op=torch.mean):
        super().__init__(pol, cri, lr=lr, clip_range=float("inf"), entropy_coef=entropy_coef, critic_coef=critic_coef, max_grad_norm=max_grad_norm)
        self.max_ratio_delta, self.delta_op = max_ratio_delta, delta_op

    def minibatch_step(self, states, actions, log_probs, advantages, returns):
        """ref: policy_loss_fn + critic_loss_fn (espo.py:123-166)."""
        self.popt.zero_grad()
        loss, pg, ent, kl, rd = policy_loss(self.pol, states, actions, log_probs, advantages, self.entropy_coef, self.delta_op)
        loss.backward()
        pnorm = torch.nn.utils.clip_grad_norm_([self.pol[k] for k in O.POLICY_KEYS], self.max_grad_norm)
        self.popt.step()
        self.copt.zero_grad()
        closs = O.critic_loss(self.cri, states, returns, self.critic_coef)
        closs.backward()
        cnorm = torch.nn.utils.clip_grad_norm_([self.cri[k] for k in O.CRITIC_KEYS], self.max_grad_norm)
        self.copt.step()
        return dict(pg_loss=pg.item(), critic_loss=closs.item(), entropy_loss=ent.item(), approx_kl=kl.item(), ratio_delta=rd.item(),
                    policy_grad_norm=pnorm.item(), critic_grad_norm=cnorm.item())

    def update(self, batch, draw, max_epochs):
        """ref: espo.py:254-274.  batch: flattened (T*N, ...) tensors; draw(): the next minibatch index array.  Returns the metric
        dicts of the steps that were made."""
        out = []
        for _ in range(max_epochs):
            idx = torch.as_tensor(draw())
            m = self.minibatch_step(batch["states"][idx], batch["actions"][idx], batch["log_probs"][idx], batch["advantages"][idx],
                                    batch["returns"][idx])
            out.append(m)
            if m["ratio_delta"] > self.max_ratio_delta:
                break
        return out
