class RLXInfo:
    """The three helper methods the PPO loop expects from an env wrapper (reference: custom_mujoco/ant/warp_torch/wrappers.py:4-51,
    gym/mujoco/humanoid_v4/wrappers.py:4-32).  The synthetic stream has no episode statistics, so logging info is empty and
    no device->host sync happens per step."""

    def __init__(self, env):
        self.env = env

    def get_logging_info_dict(self, info):
        return {}

    def get_final_observation_at_index(self, info, index):
        # the synthetic next observation is i.i.d.: the "final" observation of a finished episode is the returned one
        return self.env._obs[self.env.t % self.env.ring_length][index].numpy()

    def get_final_observations_batch(self, info, indices):
        """Vectorised form of get_final_observation_at_index (optional extension used by rl_x_b200's PPO)."""
        return self.env._obs[self.env.t % self.env.ring_length].numpy()[indices]

    def get_final_info_value_at_index(self, info, key, index):
        return 0.0

    def get_final_info_values_batch(self, info, key, indices):
        """Vectorised form of get_final_info_value_at_index (optional extension used by rl_x_b200's PPO)."""
        return [0.0] * len(indices)

    def __getattr__(self, name):
        return getattr(self.env, name)
