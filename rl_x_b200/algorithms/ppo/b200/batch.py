class Batch:
    """Device-resident rollout buffer with the reference's field names (rl_x/algorithms/ppo/pytorch/batch.py:1-11).
    `states` has T+1 time slots: slot t+1 is both next_states[t] (TORCH-interface envs, ppo.py:224-232) and the input of step
    t+1, so the separate `next_states` tensor only exists for NUMPY-interface envs (final-observation patching)."""

    def __init__(self, states, next_states, actions, rewards, values, terminations, log_probs, advantages, returns):
        self.states = states
        self.next_states = next_states
        self.actions = actions
        self.rewards = rewards
        self.values = values
        self.terminations = terminations
        self.log_probs = log_probs
        self.advantages = advantages
        self.returns = returns
