"""GPU parity tests of the SAC path: fused update / acting / replay sampling (through the C-ABI) against the executed reference's
golden vectors (tests/golden/sac_small.npz) and against the CPU oracle (oracle/sac_oracle.py) at BASELINE config-4 sizes."""
import os

import numpy as np
import pytest
import torch

from oracle import sac_oracle as S

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True, params=["simt", "tcgen05"])
def gemm_engine(request):
    """Every test runs once per GEMM engine (the engine is a process-wide switch of the library)."""
    from rl_x_b200 import _native as nt
    yield request.param
    nt.load().rlx_set_gemm_engine(0)


class _Space:
    def __init__(self, shape, low=None, high=None):
        self.shape, self.low, self.high = shape, low, high
        self._rng = np.random.default_rng(0)

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(np.float32)


class StubEnv:
    def __init__(self, n, obs, act, low, high, interface="NUMPY"):
        from rl_x_b200.environments.types import ActionSpaceType, ObservationSpaceType, DataInterfaceType

        class P:
            observation_space_type = ObservationSpaceType.FLAT_VALUES
            action_space_type = ActionSpaceType.CONTINUOUS
            data_interface_type = getattr(DataInterfaceType, interface)
        self.general_properties = P
        self.n, self.obs, self.torch = n, obs, interface == "TORCH"
        self.single_observation_space = _Space((obs,))
        self.single_action_space = _Space((act,), np.full(act, low, np.float32), np.full(act, high, np.float32))
        self.rng = np.random.default_rng(5)
        self.t = 0

    def _o(self, x):
        return torch.from_numpy(x).to(DEV) if self.torch else x

    def reset(self):
        return self._o(self.rng.standard_normal((self.n, self.obs)).astype(np.float32)), {}

    def step(self, action):
        self.t += 1
        self.final = self.rng.standard_normal((self.n, self.obs)).astype(np.float32)
        return (self._o(self.rng.standard_normal((self.n, self.obs)).astype(np.float32)), self._o(self.rng.standard_normal(self.n).astype(np.float32)),
                self._o(self.rng.random(self.n) < 0.1), self._o(np.full(self.n, self.t % 5 == 0)), {})

    def get_logging_info_dict(self, info):
        return {}

    def get_final_observation_at_index(self, info, i):
        return self.final[i]

    def get_final_info_value_at_index(self, info, key, i):
        return 0.0

    def close(self):
        pass


def _model(N, obs, act, hidden, batch, low, high, seed=2, interface="NUMPY", **algo):
    import inspect
    from rl_x_b200.config_dict import ConfigDict
    # pick up the engine of the running parametrisation
    for fr in inspect.stack():
        if "gemm_engine" in fr.frame.f_locals and isinstance(fr.frame.f_locals["gemm_engine"], str):
            algo.setdefault("gemm_engine", fr.frame.f_locals["gemm_engine"])
            break
    from rl_x_b200.algorithms.sac.b200.default_config import get_config
    from rl_x_b200.algorithms.sac.b200.sac import SAC
    a = get_config("sac.b200")
    a.nr_hidden_units, a.batch_size = hidden, batch
    for k, v in algo.items():
        a[k] = v
    cfg = ConfigDict(algorithm=a, environment=ConfigDict(seed=seed, nr_envs=N),
                     runner=ConfigDict(save_model=False, track_console=False, track_tb=False, track_wandb=False, load_model=""))
    env = StubEnv(N, obs, act, low, high, interface)
    return SAC(cfg, env, env, "/tmp/rlx_sac_test", None), env


def _named(z, prefix, net):
    return {k.split("/", 2)[2]: torch.from_numpy(z[k]) for k in z.files if k.startswith(f"{prefix}/{net}/")}


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def test_sac_updates_reproduce_reference_golden(gemm_engine):
    from conftest import GOLDEN_DIR
    from rl_x_b200 import _native as nt
    z = np.load(os.path.join(GOLDEN_DIR, "sac_small.npz"))
    N, obs, act, hidden, batch, nupd, seed = (int(x) for x in z["meta"])
    gamma, tau, lr, ls_min, ls_max, tgt, low, high = (float(x) for x in z["meta_f"])
    model, _ = _model(N, obs, act, hidden, batch, low, high, seed=seed)
    # identical initialisation for the same seed (same RNG stream / construction order as the reference)
    pol0, qs0 = model.state_dicts()
    for k, v in _named(z, "init", "policy").items():
        np.testing.assert_allclose(pol0[k].numpy(), v.numpy(), rtol=1e-6, atol=1e-7, err_msg=k)
    for net in ("q1", "q2", "q1_target", "q2_target"):
        for k, v in _named(z, "init", net).items():
            np.testing.assert_allclose(qs0[net][k].numpy(), v.numpy(), rtol=1e-6, atol=1e-7, err_msg=f"{net}/{k}")
    model.load_named(_named(z, "init", "policy"), _named(z, "init", "q1"), _named(z, "init", "q2"), _named(z, "init", "q1_target"), _named(z, "init", "q2_target"))
    for u in range(nupd):
        b = {k: torch.from_numpy(z[f"batch{u}/{k}"]).to(DEV).contiguous() for k in ["states", "next_states", "actions", "rewards", "terminations"]}
        eps = [torch.from_numpy(z[f"batch{u}/eps_next"]).to(DEV), torch.from_numpy(z[f"batch{u}/eps_cur"]).to(DEV)]
        model._draw_eps = lambda n, e=eps: e.pop(0)
        model.metric_sums.zero_()
        model.update((b["states"], b["next_states"], b["actions"], b["rewards"], b["terminations"]))
        m = model.metrics.cpu().numpy()
        for i, name in enumerate(nt.SAC_METRIC_NAMES):
            ref = float(z[f"metric/{name}"][u])
            assert abs(float(m[i]) - ref) <= 2e-5 * max(1.0, abs(ref)), (u, name, float(m[i]), ref)
    pol, qs = model.state_dicts()
    for k, v in _named(z, "final", "policy").items():
        assert _rel(pol[k].numpy(), v.numpy()) < 2e-5, k
    for net in ("q1", "q2", "q1_target", "q2_target"):
        for k, v in _named(z, "final", net).items():
            assert _rel(qs[net][k].numpy(), v.numpy()) < 2e-5, (net, k)
    np.testing.assert_allclose(model.log_alpha.cpu().numpy(), z["final/log_alpha"], rtol=1e-5, atol=1e-8)
    assert model.steps.cpu().tolist() == [nupd, nupd, nupd]


def test_sac_update_vs_oracle_at_config4_sizes(gemm_engine):
    """Box(17) / Box(6), hidden 256, batch 4096 (BASELINE.json configs[3])."""
    from rl_x_b200 import _native as nt
    obs, act, hidden, B = 17, 6, 256, 4096
    model, _ = _model(1, obs, act, hidden, B, -1.0, 1.0)
    pol, q1, q2 = S.init_params(obs, act, hidden, seed=3)
    L = S.Learner(pol, q1, q2, torch.full((act,), -1.0), torch.full((act,), 1.0), log_alpha=-0.3)
    model.load_named(pol, q1, q2, q1, q2)
    model.log_alpha.fill_(-0.3)
    g = torch.Generator().manual_seed(0)
    for u in range(2):
        s, ns = torch.randn(B, obs, generator=g), torch.randn(B, obs, generator=g)
        ac = torch.tanh(torch.randn(B, act, generator=g))
        r, d = torch.randn(B, generator=g), (torch.rand(B, generator=g) < 0.05).float()
        e1, e2 = torch.randn(B, act, generator=g), torch.randn(B, act, generator=g)
        ref = L.update(s, ns, ac, r, d, e1, e2)
        eps = [e1.to(DEV), e2.to(DEV)]
        model._draw_eps = lambda n, e=eps: e.pop(0)
        model.update(tuple(t.to(DEV).contiguous() for t in (s, ns, ac, r, d)))
        m = model.metrics.cpu().numpy()
        for i, name in enumerate(nt.SAC_METRIC_NAMES):
            assert abs(float(m[i]) - ref[name]) <= 2e-5 * max(1.0, abs(ref[name])), (u, name, float(m[i]), ref[name])
    polg, qs = model.state_dicts()
    for k in S.POLICY_KEYS:
        assert _rel(polg[k].numpy(), L.pol[k].detach().numpy()) < 1e-5, k
    for net, d in (("q1", L.q1), ("q2", L.q2), ("q1_target", L.q1t), ("q2_target", L.q2t)):
        for k in S.Q_KEYS:
            assert _rel(qs[net][k].numpy(), d[k].detach().numpy()) < 1e-5, (net, k)


def test_sac_act_vs_oracle(gemm_engine):
    obs, act, hidden, n = 17, 6, 256, 333
    model, _ = _model(n, obs, act, hidden, 64, -2.0, 0.5)
    pol, q1, q2 = S.init_params(obs, act, hidden, seed=4)
    model.load_named(pol, q1, q2, q1, q2)
    g = torch.Generator().manual_seed(1)
    x, eps = torch.randn(n, obs, generator=g), torch.randn(n, act, generator=g)
    low, high = torch.full((act,), -2.0), torch.full((act,), 0.5)
    with torch.no_grad():
        a_ref, s_ref, lp_ref = S.policy_get_action(pol, x, eps, low, high)
        d_ref = S.policy_deterministic(pol, x, low, high)
    a_t, a_env, lp = torch.empty(n, act, device=DEV), torch.empty(n, act, device=DEV), torch.empty(n, device=DEV)
    model.k.act(model.policy, x.to(DEV), eps.to(DEV), model.d_low, model.d_high, model.ws, action_tanh=a_t, env_action=a_env, logp=lp)
    np.testing.assert_allclose(a_t.cpu().numpy(), a_ref.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(a_env.cpu().numpy(), s_ref.numpy(), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(lp.cpu().numpy(), lp_ref.numpy().reshape(-1), rtol=1e-5, atol=1e-4)
    d = torch.empty(n, act, device=DEV)
    model.k.act(model.policy, x.to(DEV), None, model.d_low, model.d_high, model.ws, deterministic=True, env_action=d)
    np.testing.assert_allclose(d.cpu().numpy(), d_ref.numpy(), rtol=1e-5, atol=2e-6)


def test_replay_sampling_is_numpy_exact():
    from rl_x_b200 import _native as nt
    from rl_x_b200.algorithms.sac.b200.replay_buffer import ReplayBuffer
    N, obs, act = 4, 5, 2
    rb = ReplayBuffer(64, N, (obs,), (act,), nt.Pcg64Generator(7), torch.device(DEV))
    ref_rng = np.random.default_rng(7)
    g = np.random.default_rng(0)
    host = dict(s=np.zeros((16, N, obs), np.float32), ns=np.zeros((16, N, obs), np.float32), a=np.zeros((16, N, act), np.float32),
                r=np.zeros((16, N), np.float32), t=np.zeros((16, N), np.float32))
    for step in range(21):  # wraps around the 16-slot ring
        s, ns, a = g.standard_normal((N, obs)).astype(np.float32), g.standard_normal((N, obs)).astype(np.float32), g.standard_normal((N, act)).astype(np.float32)
        r, t = g.standard_normal(N).astype(np.float32), g.random(N) < 0.3
        pos = step % 16
        host["s"][pos], host["ns"][pos], host["a"][pos], host["r"][pos], host["t"][pos] = s, ns, a, r, t
        rb.add(s, ns, a, r, t)
        size = min(step + 1, 16)
        out = rb.sample(37)
        i1, i2 = ref_rng.integers(size, size=37), ref_rng.integers(N, size=37)
        for got, key in zip(out, ["s", "ns", "a", "r", "t"]):
            assert np.array_equal(got.cpu().numpy(), host[key][i1, i2]), (step, key)


@pytest.mark.parametrize("interface", ["NUMPY", "TORCH"])
def test_sac_train_loop_runs(interface, gemm_engine):
    model, env = _model(4, 17, 6, 64, 32, -2.0, 1.0, interface=interface, learning_starts=24, total_timesteps=120, logging_frequency=8, buffer_size=512)
    logged = []
    model.log = lambda name, value, step: logged.append((name, float(value), step))
    model.train()
    names = {n for n, _, _ in logged}
    for n in ["loss/q_loss", "loss/policy_loss", "loss/entropy_loss", "entropy/alpha", "q_value/q_value", "gradients/critic_grad_norm", "steps/nr_updates"]:
        assert n in names, n
    assert all(np.isfinite(v) for _, v, _ in logged)
    assert model.nr_updates == (120 - 24) // 4  # updates run once global_step > learning_starts (sac.py:212)
    assert int(model.steps[0].item()) == model.nr_updates


def test_sac_checkpoint_has_the_reference_optimizer_layout(tmp_path):
    """save() writes the reference's keys (sac.py:381-396) incl. the three torch.optim.Adam state dicts in parameters() order
    (policy: torso, mean, log_std; q: q1 then q2), and load() restores moments and step counters from them."""
    from rl_x_b200.algorithms.sac.b200.sac import SAC, POLICY_PARAM_ORDER
    N, obs, act, hidden = 2, 9, 3, 32
    model, env = _model(N, obs, act, hidden, 16, np.full(act, -1.0, np.float32), np.full(act, 1.0, np.float32))
    model.save_path = str(tmp_path)
    g = torch.Generator(device="cuda").manual_seed(0)
    for t in (model.m_policy, model.m_q, model.m_la, model.policy, model.q):
        t.normal_(generator=g)
    for t in (model.v_policy, model.v_q, model.v_la):
        t.uniform_(generator=g)
    model.steps.copy_(torch.tensor([7, 9, 8]))
    model.log_alpha.fill_(-0.3)
    model.save()
    ck = torch.load(str(tmp_path / "best.model"), weights_only=False)
    assert set(ck) == {"config_algorithm", "policy_state_dict", "q1_state_dict", "q2_state_dict", "q1_target_state_dict", "q2_target_state_dict",
                       "log_alpha", "policy_optimizer_state_dict", "q_optimizer_state_dict", "entropy_optimizer_state_dict"}
    assert list(ck["policy_state_dict"]) == list(POLICY_PARAM_ORDER)
    po, qo, eo = ck["policy_optimizer_state_dict"], ck["q_optimizer_state_dict"], ck["entropy_optimizer_state_dict"]
    assert len(po["state"]) == 8 and len(qo["state"]) == 12 and len(eo["state"]) == 1
    assert tuple(po["state"][4]["exp_avg"].shape) == (act, hidden) and tuple(po["state"][5]["exp_avg"].shape) == (act,)  # mean.weight, mean.bias
    assert tuple(qo["state"][6]["exp_avg"].shape) == (hidden, obs + act) and float(qo["state"][0]["step"]) == 9.0
    # the dicts drive real torch optimizers over modules shaped like the reference's
    pol = torch.nn.ParameterList([torch.nn.Parameter(v.clone()) for v in ck["policy_state_dict"].values()])
    opt = torch.optim.Adam(pol.parameters(), lr=1e-3)
    opt.load_state_dict(po)
    assert torch.equal(opt.state[pol[6]]["exp_avg"], po["state"][6]["exp_avg"])
    # round trip
    cfg = model.config
    cfg.runner.load_model = str(tmp_path / "best.model")
    m2 = SAC.load(cfg, env, env, "/tmp/rlx_sac_test2", None, [])
    for name in ("policy", "log_alpha", "m_policy", "v_policy", "m_la", "v_la", "steps"):
        assert torch.equal(getattr(m2, name), getattr(model, name)), name
    # q-net tensors by name (the flat buffers pad every net to a 256-byte stride; the padding is not part of a checkpoint)
    (_, qs1), (_, qs2) = model.state_dicts(), m2.state_dicts()
    for net in qs1:
        for key in qs1[net]:
            assert torch.equal(qs1[net][key], qs2[net][key]), (net, key)
    ov1, ov2 = model._optimizer_views(), m2._optimizer_views()
    for a, b in zip(ov1["q"][0] + ov1["q"][1], ov2["q"][0] + ov2["q"][1]):
        assert torch.equal(a, b)
