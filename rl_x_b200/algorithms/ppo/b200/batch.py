class Batch:
    """Device-resident rollout buffer.  Field names are the reference's (rl_x/algorithms/ppo/pytorch/batch.py) because the training loop
    and the tests address them by name.  `states` has T+1 time slots: slot t+1 is both next_states[t] (TORCH-interface envs,
    ppo.py:224-232) and the input of step t+1, so a separate `next_states` tensor only exists for NUMPY-interface envs (final-observation
    patching) and is None otherwise."""

    __slots__ = ("states", "next_states", "actions", "rewards", "values", "terminations", "log_probs", "advantages", "returns")

    def __init__(self, *positional, **tensors):
        # the reference's signature: Batch(states, next_states, actions, rewards, values, terminations, log_probs, advantages, returns),
        # positionally or by keyword (ppo.py:171-181 uses keywords)
        if len(positional) > len(self.__slots__) or set(self.__slots__[:len(positional)]) & set(tensors):
            raise TypeError(f"Batch takes the fields {self.__slots__} once each")
        tensors.update(zip(self.__slots__, positional))
        missing = set(self.__slots__) - set(tensors)
        if missing or len(tensors) != len(self.__slots__):
            raise TypeError(f"Batch needs exactly the fields {self.__slots__}; missing {sorted(missing)}, got {sorted(tensors)}")
        for name in self.__slots__:
            setattr(self, name, tensors[name])

    def nbytes(self):
        return sum(getattr(self, n).numel() * getattr(self, n).element_size() for n in self.__slots__ if getattr(self, n) is not None)
