"""Pendulum-v1 as a NumPy vector environment (BASELINE.json configs[0]).

gymnasium is absent from this image, so the published dynamics of gymnasium's classic_control Pendulum-v1 are restated here:
max_speed 8, max_torque 2, dt 0.05, g 10, m 1, l 1; u = clip(a, -2, 2); cost = angle_normalize(th)^2 + 0.1 thdot^2 + 0.001 u^2;
thdot' = clip(thdot + (3g/(2l) sin th + 3/(m l^2) u) dt, -8, 8); th' = th + thdot' dt; obs = (cos th, sin th, thdot) as float32;
reset: th ~ U(-pi, pi), thdot ~ U(-1, 1); TimeLimit of 200 steps (truncation).  The vector layer follows the gymnasium 0.29 autoreset
convention the reference's wrappers rely on (gym/mujoco/humanoid_v4/wrappers.py:9-32): when an episode ends, `step` returns the first
observation of the next episode and puts the last observation / info of the finished one into info["final_observation"] /
info["final_info"] (object arrays with "_"-prefixed masks); RecordEpisodeStatistics' episode_return / episode_length land in final_info.
Each sub-environment owns np.random.default_rng(seed + i), as gymnasium seeds its sub-environments."""
import numpy as np


class BoxSpace:
    def __init__(self, low, high, shape):
        self.shape = tuple(shape)
        self.dtype = np.float32
        self.low = np.broadcast_to(np.asarray(low, dtype=np.float32), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=np.float32), self.shape).copy()


def angle_normalize(x):
    return ((x + np.pi) % (2 * np.pi)) - np.pi


class PendulumVecEnv:
    max_speed, max_torque, dt, g, m, l = 8.0, 2.0, 0.05, 10.0, 1.0, 1.0
    max_episode_steps = 200

    def __init__(self, nr_envs, seed):
        self.nr_envs = int(nr_envs)
        self.single_observation_space = BoxSpace([-1.0, -1.0, -self.max_speed], [1.0, 1.0, self.max_speed], (3,))
        self.single_action_space = BoxSpace(-self.max_torque, self.max_torque, (1,))
        self.rngs = [np.random.default_rng(int(seed) + i) for i in range(self.nr_envs)]
        self.th = np.zeros(self.nr_envs)
        self.thdot = np.zeros(self.nr_envs)
        self.elapsed = np.zeros(self.nr_envs, dtype=np.int64)
        self.episode_return = np.zeros(self.nr_envs)
        self.episode_length = np.zeros(self.nr_envs)

    def _reset_one(self, i):
        self.th[i], self.thdot[i] = self.rngs[i].uniform(low=[-np.pi, -1.0], high=[np.pi, 1.0])
        self.elapsed[i] = 0
        self.episode_return[i] = 0.0
        self.episode_length[i] = 0.0

    def _obs(self):
        return np.stack([np.cos(self.th), np.sin(self.th), self.thdot], axis=1).astype(np.float32)

    def reset(self, seed=None):
        if seed is not None:
            self.rngs = [np.random.default_rng(int(seed) + i) for i in range(self.nr_envs)]
        for i in range(self.nr_envs):
            self._reset_one(i)
        return self._obs(), {}

    def step(self, action):
        u = np.clip(np.asarray(action, dtype=np.float64).reshape(self.nr_envs), -self.max_torque, self.max_torque)
        costs = angle_normalize(self.th) ** 2 + 0.1 * self.thdot ** 2 + 0.001 * u ** 2
        thdot = self.thdot + (3 * self.g / (2 * self.l) * np.sin(self.th) + 3.0 / (self.m * self.l ** 2) * u) * self.dt
        self.thdot = np.clip(thdot, -self.max_speed, self.max_speed)
        self.th = self.th + self.thdot * self.dt
        self.elapsed += 1
        reward = -costs
        self.episode_return += reward
        self.episode_length += 1
        terminated = np.zeros(self.nr_envs, dtype=bool)
        truncated = self.elapsed >= self.max_episode_steps
        obs = self._obs()
        info = {}
        done = terminated | truncated
        if done.any():
            final_obs = np.full(self.nr_envs, None, dtype=object)
            final_info = np.full(self.nr_envs, None, dtype=object)
            for i in np.nonzero(done)[0]:
                final_obs[i] = obs[i].copy()
                final_info[i] = {"episode_return": float(self.episode_return[i]), "episode_length": float(self.episode_length[i])}
                self._reset_one(i)
            info = {"final_observation": final_obs, "_final_observation": done.copy(), "final_info": final_info, "_final_info": done.copy()}
            obs = self._obs()
        return obs, reward.astype(np.float64), terminated, truncated, info

    def close(self):
        pass
