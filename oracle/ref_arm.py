"""Reference arm of the benchmark: the UNMODIFIED reference classes (oracle/_ref, staged by oracle/make_ref.py) timed on the host cores.

TEST / BENCH INFRASTRUCTURE ONLY: imported by bench.py's `--impl reference` / `cpu_baseline` legs and by tests/, never by rl_x_b200/.

What runs is `rl_x.algorithms.ppo.pytorch.ppo.PPO(config, env, env, run_path, None).train()` exactly as the reference's runner would
call it (rl_x/runner/runner.py:334-341), with
  * a stub `ml_collections.config_dict.ConfigDict` (the real package is absent from the image; every default_config.py imports it),
  * `algorithm.device = "cpu"`, `bf16_mixed_precision_training = False` (mandatory off-CUDA, ppo.py:69-70), `compile_mode = "default"`
    through Inductor's C++ backend with CXX=/usr/bin/g++, falling back to eager (TORCHDYNAMO_DISABLE) if Inductor
    cannot compile on this host — the fallback is reported, never silent,
  * the synthetic Box(obs)/Box(act) vector env of BASELINE.json configs[1] with the TORCH data interface on CPU tensors
    (obs ~ N(0,1), reward ~ N(0,1), terminated ~ Bernoulli(0.01)) — scaffolding, not reference code.
Iteration times are wall-clock stamps taken in the model's own `start_logging` hook (called once per iteration after the save step,
ppo.py:366), i.e. the interval the reference itself reports as `time/sps` (ppo.py:359-363).
"""
import os
import sys
import time
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


class _ConfigDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    __setattr__ = dict.__setitem__


def _install_stub():
    if "ml_collections" in sys.modules and hasattr(sys.modules["ml_collections"], "config_dict"):
        return
    mc = types.ModuleType("ml_collections")
    cd = types.ModuleType("ml_collections.config_dict")
    cd.ConfigDict = _ConfigDict
    mc.config_dict = cd
    sys.modules["ml_collections"] = mc
    sys.modules["ml_collections.config_dict"] = cd


def import_reference():
    """Put oracle/_ref first on sys.path and import the reference PPO / SAC modules from it."""
    from oracle import make_ref
    if not make_ref.available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/make_ref.py` in the build container (needs /root/reference)")
    make_ref.verify()
    _install_stub()
    os.environ.setdefault("WANDB_MODE", "disabled")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import rl_x.algorithms.ppo.pytorch.ppo as refppo
    if not os.path.abspath(refppo.__file__).startswith(REF_DIR):
        raise RuntimeError(f"rl_x resolved to {refppo.__file__}, not to the staged copy")
    return refppo


class _Space:
    def __init__(self, shape, low=None, high=None):
        self.shape, self.low, self.high = shape, low, high


class SyntheticTorchEnv:
    """Box(obs)/Box(act in [-1, 1]) vector env, TORCH data interface on CPU tensors (SURVEY.md §8 d)."""

    def __init__(self, nr_envs, obs_dim, act_dim, seed=1, p_term=0.01):
        from rl_x.environments.action_space_type import ActionSpaceType
        from rl_x.environments.data_interface_type import DataInterfaceType
        from rl_x.environments.observation_space_type import ObservationSpaceType

        class Props:
            observation_space_type = ObservationSpaceType.FLAT_VALUES
            action_space_type = ActionSpaceType.CONTINUOUS
            data_interface_type = DataInterfaceType.TORCH

        self.general_properties = Props
        self.nr_envs, self.obs_dim, self.act_dim, self.p_term = nr_envs, obs_dim, act_dim, p_term
        self.single_observation_space = _Space((obs_dim,))
        self.single_action_space = _Space((act_dim,), np.full(act_dim, -1.0, np.float32), np.full(act_dim, 1.0, np.float32))
        self.gen = torch.Generator().manual_seed(seed)

    def reset(self):
        return torch.randn(self.nr_envs, self.obs_dim, generator=self.gen), {}

    def step(self, action):
        obs = torch.randn(self.nr_envs, self.obs_dim, generator=self.gen)
        rew = torch.randn(self.nr_envs, generator=self.gen)
        term = torch.rand(self.nr_envs, generator=self.gen) < self.p_term
        return obs, rew, term, torch.zeros(self.nr_envs, dtype=torch.bool), {}

    def get_logging_info_dict(self, info):
        return {}

    def close(self):
        pass


class _Done(Exception):
    pass


def run_ppo(nr_envs, nr_steps, obs_dim, act_dim, hidden, minibatch, epochs, warmup, steps, threads, compile_mode="default",
            max_seconds=None, min_steps=3):
    """Train the reference PPO for `warmup` + up to `steps` REAL iterations on CPU; returns a dict with the per-iteration seconds of the
    timed iterations.  `max_seconds`: stop early (after at least `min_steps` timed iterations) once the timed part exceeded it."""
    os.environ["CXX"] = "/usr/bin/g++"  # the image default /opt/gcc/bin/g++ cannot find libgomp.spec
    os.environ["CC"] = "/usr/bin/gcc"
    torch.set_num_threads(threads)
    refppo = import_reference()
    from rl_x.algorithms.ppo.pytorch.default_config import get_config
    a = get_config("ppo.pytorch")
    a.device, a.bf16_mixed_precision_training, a.compile_mode = "cpu", False, compile_mode
    a.nr_steps, a.nr_epochs, a.minibatch_size, a.nr_hidden_units = nr_steps, epochs, minibatch, hidden
    B = nr_envs * nr_steps
    a.total_timesteps = float(B * (warmup + steps + 1))
    cfg = _ConfigDict(algorithm=a, environment=_ConfigDict(seed=1, nr_envs=nr_envs),
                      runner=_ConfigDict(save_model=False, track_console=False, track_tb=False, track_wandb=False))
    env = SyntheticTorchEnv(nr_envs, obs_dim, act_dim, seed=1)
    model = refppo.PPO(cfg, env, env, "/tmp/rlx_ref_arm", None)
    stamps, logged = [], {}

    def start_logging(step):
        stamps.append(time.perf_counter())
        n_timed = len(stamps) - 1 - warmup
        if n_timed >= steps or (max_seconds is not None and n_timed >= min_steps and stamps[-1] - stamps[warmup] > max_seconds):
            raise _Done()

    def log(name, value, step):
        logged.setdefault(name, []).append(float(value))

    model.start_logging, model.log, model.end_logging = start_logging, log, (lambda *a_, **k_: None)
    t_begin = time.perf_counter()
    stamps.append(t_begin)
    try:
        model.train()
    except _Done:
        pass
    iters = np.diff(np.asarray(stamps))  # iters[0] contains compilation
    timed = iters[warmup:]
    return {"seconds": [float(x) for x in timed], "first_iteration_s": float(iters[0]), "warmup_done": int(min(warmup, len(iters))),
            "phases": {k: [float(x) for x in v[warmup:]] for k, v in logged.items() if k.startswith("time/") and "sps" not in k},
            "threads": threads, "compile_mode": compile_mode, "torch_compile": os.environ.get("TORCHDYNAMO_DISABLE", "0") != "1",
            "class": f"{refppo.PPO.__module__}.{refppo.PPO.__name__}", "file": os.path.relpath(refppo.__file__, os.path.dirname(HERE))}


def main(argv=None):
    """python -m oracle.ref_arm --envs N --nr-steps T ... : one JSON object on stdout (bench.py runs this in a child process so that
    CXX / TORCHDYNAMO_DISABLE apply before torch is imported)."""
    import argparse
    import json
    ap = argparse.ArgumentParser()
    for name, default in (("envs", 4096), ("nr-steps", 128), ("obs", 376), ("act", 17), ("hidden", 256), ("minibatch", 32768), ("epochs", 10),
                          ("warmup", 1), ("steps", 3), ("threads", 8), ("min-steps", 3)):
        ap.add_argument(f"--{name}", type=int, default=default)
    ap.add_argument("--max-seconds", type=float, default=None)
    ap.add_argument("--compile-mode", default="default")
    a = ap.parse_args(argv)
    r = run_ppo(a.envs, a.nr_steps, a.obs, a.act, a.hidden, a.minibatch, a.epochs, a.warmup, a.steps, a.threads, a.compile_mode,
                a.max_seconds, a.min_steps)
    print("REF_ARM_JSON " + json.dumps(r))


if __name__ == "__main__":
    main()
