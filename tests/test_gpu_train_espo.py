"""ESPO plugin (rl_x_b200/algorithms/espo/b200) against the executed reference: the golden environment stream and action noise are
replayed through ESPO.train(); the number of update steps before the ratio_delta stop, the weights after every iteration and the
logged means must match (tests/golden/espo_small.npz, made by tests/golden/make_golden_espo.py)."""
import numpy as np
import pytest
import torch

from test_gpu_train import ReplayEnv, _reference_noise

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["simt", "tcgen05"])
def gemm_engine(request):
    from rl_x_b200 import _native as nt
    lib = nt.load()
    lib.rlx_set_gemm_engine(1 if request.param == "tcgen05" else 0)
    yield request.param
    lib.rlx_set_gemm_engine(0)


def _config(g, engine, **algo):
    from rl_x_b200.config_dict import ConfigDict
    from rl_x_b200.algorithms.espo.b200.default_config import get_config
    a = get_config("espo.b200")
    a.nr_steps, a.minibatch_size, a.max_epochs, a.nr_hidden_units = g.T, g.mb, g.max_epochs, g.hidden
    a.total_timesteps = g.N * g.T * g.iterations
    a.std_dev, a.entropy_coef, a.anneal_learning_rate = g.std_dev, g.entropy_coef, g.anneal
    a.learning_rate, a.max_ratio_delta, a.critic_coef, a.max_grad_norm = g.lr, g.max_ratio_delta, g.critic_coef, g.max_grad_norm
    a.gamma, a.gae_lambda, a.gemm_engine = g.gamma, g.gae_lambda, engine
    for k, v in algo.items():
        a[k] = v
    return ConfigDict(algorithm=a, environment=ConfigDict(seed=g.seed, nr_envs=g.N),
                      runner=ConfigDict(save_model=False, track_console=False, track_tb=False, track_wandb=False, load_model=""))


@pytest.mark.parametrize("interface", ["NUMPY", "TORCH"])
def test_espo_train_reproduces_reference_run(golden_espo, interface, gemm_engine):
    from rl_x_b200.algorithms.espo.b200.espo import ESPO
    g = golden_espo
    env = ReplayEnv(g, interface)
    model = ESPO(_config(g, gemm_engine), env, env, "/tmp/rlx_test_run", None)
    eps = _reference_noise(g)
    calls = {"n": 0}

    def draw(step):
        i = calls["n"]
        calls["n"] += 1
        return eps[i]

    model._draw_noise = draw
    logged, snaps, steps = [], [], []
    model.log = lambda name, value, step: logged.append((name, float(value), int(step)))
    orig = model.start_logging

    def start_logging(step):
        b = model.batch
        snaps.append(dict(adv=b.advantages.cpu().numpy().copy(), ret=b.returns.cpu().numpy().copy(), sd=model.params.state_dicts()))
        steps.append(model._espo_steps)
        orig(step)

    model.start_logging = start_logging
    model.train()
    assert len(snaps) == g.iterations
    # the stop rule fired after the same number of update steps as in the reference (3, 4, 8)
    assert steps == [int(g[f"iter{it}/nr_epochs"]) for it in range(g.iterations)]
    np.testing.assert_allclose(torch.stack(env.actions).numpy(), g["env_actions"], rtol=1e-4, atol=5e-6)
    for it, s in enumerate(snaps):
        np.testing.assert_allclose(s["adv"], g[f"iter{it}/advantages"], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(s["ret"], g[f"iter{it}/returns"], rtol=1e-4, atol=2e-5)
        pol_ref, cri_ref = g.params(f"iter{it}")
        pol, cri = s["sd"]
        for name, v in {**pol_ref, **cri_ref}.items():
            ours = (pol if name in pol else cri)[name].numpy()
            rel = float(np.linalg.norm(ours - v) / np.linalg.norm(v))
            assert rel <= 2e-5, (it, name, rel)
    for n in ["loss/critic_loss", "loss/entropy_loss", "policy_ratio/ratio_delta", "gradients/policy_grad_norm", "gradients/critic_grad_norm",
              "lr/learning_rate", "v_value/explained_variance", "policy/std_dev", "optim/nr_epochs", "steps/nr_env_steps", "steps/nr_updates",
              "steps/nr_episodes"]:
        ours = [v for m, v, _ in logged if m == n]
        np.testing.assert_allclose(ours, g[f"metric/{n}"], rtol=2e-4, atol=1e-6, err_msg=n)
    for n, atol in [("loss/policy_gradient_loss", 2e-5), ("policy_ratio/approx_kl", 1e-6)]:
        ours = [v for m, v, _ in logged if m == n]
        np.testing.assert_allclose(ours, g[f"metric/{n}"], rtol=1e-3, atol=atol, err_msg=n)


@pytest.mark.parametrize("operator", ["mean", "median"])
def test_espo_update_steps_match_oracle_on_fresh_data(gemm_engine, operator):
    """One ESPO iteration on random rollout data of Humanoid-like width (376 -> 256 -> 256 -> 17) against oracle/espo_oracle.py: same stop
    step, same weights, for both delta_calc_operators (espo.py:57-63; median = torch.median = the lower median)."""
    from oracle import espo_oracle as E
    from oracle import ppo_oracle as O
    from rl_x_b200.algorithms.espo.b200.espo import ESPO
    from rl_x_b200.config_dict import ConfigDict
    from rl_x_b200.algorithms.espo.b200.default_config import get_config
    from test_gpu_train import _Space, _props
    N, T, obs, act, hid, mb = 8, 32, 376, 17, 256, 64
    a = get_config("espo.b200")
    a.nr_steps, a.minibatch_size, a.max_epochs, a.nr_hidden_units, a.total_timesteps = T, mb, 10, hid, N * T
    a.learning_rate, a.max_ratio_delta, a.gemm_engine, a.entropy_coef = 1e-3, (0.03 if operator == "mean" else 0.025), gemm_engine, 0.005
    a.delta_calc_operator = operator
    cfg = ConfigDict(algorithm=a, environment=ConfigDict(seed=11, nr_envs=N),
                     runner=ConfigDict(save_model=False, track_console=False, track_tb=False, track_wandb=False, load_model=""))

    class Env:
        general_properties = _props("TORCH")
        single_observation_space = _Space((obs,))
        single_action_space = _Space((act,), np.full(act, -1.0, np.float32), np.full(act, 1.0, np.float32))
        gen = torch.Generator(device="cuda").manual_seed(3)

        def reset(self):
            return torch.randn(N, obs, device="cuda", generator=self.gen), {}

        def step(self, action):
            return (torch.randn(N, obs, device="cuda", generator=self.gen), torch.randn(N, device="cuda", generator=self.gen),
                    torch.rand(N, device="cuda", generator=self.gen) < 0.05, torch.zeros(N, dtype=torch.bool, device="cuda"), {})

        def get_logging_info_dict(self, info):
            return {}

        def close(self):
            pass

    env = Env()
    model = ESPO(cfg, env, env, "/tmp/rlx_test_run", None)
    pol0, cri0 = model.params.state_dicts()
    model.log = lambda *a_: None
    model.train()
    b = model.batch
    batch = O.flatten({"states": b.states[:T].cpu(), "actions": b.actions.cpu(), "log_probs": b.log_probs.cpu(), "advantages": b.advantages.cpu(),
                       "returns": b.returns.cpu()})
    L = E.Learner(pol0, cri0, lr=1e-3, entropy_coef=0.005, critic_coef=a.critic_coef, max_grad_norm=a.max_grad_norm, max_ratio_delta=a.max_ratio_delta,
                  delta_op=torch.mean if operator == "mean" else torch.median)
    rng = np.random.default_rng(11)
    torch.set_num_threads(1)
    metrics = L.update(batch, lambda: rng.choice(N * T, size=mb, replace=False), 10)
    assert len(metrics) == model._espo_steps
    ours_rd = model.metrics_host.numpy()[:model._espo_steps, 4]
    # the first step starts from ratio == 1: this build's rollout and update log-probs are the same kernel arithmetic (exactly 0), torch's
    # eager forward differs from its own rollout by fp32 rounding (~1e-6)
    np.testing.assert_allclose(ours_rd, [m["ratio_delta"] for m in metrics], rtol=2e-4, atol=5e-6)
    pol, cri = model.params.state_dicts()
    for name in O.POLICY_KEYS + O.CRITIC_KEYS:
        ours, ref = (pol if name in pol else cri)[name].numpy(), (L.pol if name in L.pol else L.cri)[name].detach().numpy()
        rel = float(np.linalg.norm(ours - ref) / np.linalg.norm(ref))
        # zero-initialised biases are, after <= 10 Adam steps, sums of per-step moves of size ~lr = 1e-3 whose direction g/sqrt(v) amplifies
        # fp32 rounding of near-zero gradient entries: elementwise bound 0.05 * lr instead of the relative norm
        assert rel <= 2e-5 or float(np.abs(ours - ref).max()) <= 0.05 * 1e-3, (name, rel, float(np.abs(ours - ref).max()))
