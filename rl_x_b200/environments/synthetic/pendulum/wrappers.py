class RLXInfo:
    """The helper methods the PPO / SAC loops expect from a gym-style vector env, with the behaviour of the reference's wrapper
    (gym/mujoco/humanoid_v4/wrappers.py:4-32) on gymnasium-0.29-style info dicts."""

    def __init__(self, env):
        self.env = env

    def get_logging_info_dict(self, info):
        keys_to_remove = ["final_observation", "final_info"]
        logging_info = {key: info[key][info["_" + key]].tolist() for key in list(info.keys())
                        if key not in keys_to_remove and not key.startswith("_") and len(info[key][info["_" + key]]) > 0}
        if "final_info" in info:
            for done, final_info in zip(info["_final_info"], info["final_info"]):
                if done:
                    for key, info_value in final_info.items():
                        if key not in keys_to_remove:
                            logging_info.setdefault(key, []).append(info_value)
        return logging_info

    def get_final_observation_at_index(self, info, index):
        return info["final_observation"][index]

    def get_final_info_value_at_index(self, info, key, index):
        return info["final_info"][index][key]

    def __getattr__(self, name):
        return getattr(self.env, name)
