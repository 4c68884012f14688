// ThreadSanitizer driver for the block-cooperative LSTM recurrence kernels (csrc/lstm.cu: lstm_seq_fwd_kernel / lstm_seq_bwd_kernel).
// Test scaffolding, built by tests/test_lstm_emulation.py::test_persistent_recurrence_is_race_free_under_tsan:
//     g++ -std=c++17 -O1 -g -fsanitize=thread -pthread -DRLX_EMU -x c++ rl_x_b200/csrc/lstm.cu tests/emu_tsan_lstm.cpp -o emu_tsan_lstm
// The emulation runs every CUDA thread of a block as an OS thread and __syncthreads as a real barrier (csrc/dual_build.cuh), so a
// barrier that is missing between a shared-memory write and another thread's read is a data race TSan reports (exit code 66).
// One fwd+bwd call with the one-launch-per-direction recurrence on; the result is compared with the per-step path bit for bit.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "../include/rlx_b200.h"

static int run_case(rlx_lstm_dims d, long long T, long long n, unsigned long long expect_launches) {
  const long long R = T * n;
  int64_t poff[RLX_LSTM_POLICY_NSEG + 1], coff[RLX_LSTM_CRITIC_NSEG + 1];
  if (rlx_lstm_param_layout(&d, poff, coff)) return 2;
  std::mt19937 gen(5);
  std::normal_distribution<float> nd(0.f, 1.f);
  auto randn = [&](size_t cnt, float s) { std::vector<float> v(cnt); for (auto& x : v) x = s * nd(gen); return v; };
  std::vector<float> P = randn(poff[RLX_LSTM_POLICY_NSEG], 0.3f), Cc = randn(coff[RLX_LSTM_CRITIC_NSEG], 0.3f);
  std::vector<float> states = randn(R * d.obs_dim, 1.f), actions = randn(R * d.act_dim, 1.f), logp = randn(R, 0.1f), adv = randn(R, 1.f), ret = randn(R, 1.f);
  std::vector<float> dones(R), ic = randn(n * d.lstm_dim, 0.5f), ih = randn(n * d.lstm_dim, 0.5f);
  for (auto& x : dones) x = (gen() % 5 == 0) ? 1.f : 0.f;
  for (auto& x : logp) x -= 2.5f;
  float stats[2] = {0.f, 1.f};
  const size_t nbytes = rlx_lstm_minibatch_workspace_bytes(&d, T, n);
  std::vector<float> out[2][3];
  for (int persistent = 0; persistent < 2; ++persistent) {
    std::vector<float> ws(nbytes / 4, NAN), gP(P.size(), NAN), gC(Cc.size(), NAN), metrics(8, 0.f);
    rlx_lstm_minibatch_args a;
    memset(&a, 0, sizeof(a));
    a.dims = d; a.T = T; a.n_env = n;
    a.states = states.data(); a.actions = actions.data(); a.log_probs = logp.data(); a.advantages = adv.data(); a.returns = ret.data();
    a.dones = dones.data(); a.init_c = ic.data(); a.init_h = ih.data(); a.adv_stats = stats;
    a.policy_params = P.data(); a.critic_params = Cc.data(); a.policy_grads = gP.data(); a.critic_grads = gC.data();
    a.clip_range = 0.2f; a.entropy_coef = 0.01f; a.critic_coef = 0.5f;
    a.metrics = metrics.data(); a.workspace = ws.data(); a.workspace_bytes = nbytes;
    rlx_set_lstm_persistent(persistent);
    if (rlx_lstm_ppo_minibatch_fwdbwd_f32(&a, nullptr)) return 3;
    out[persistent][0] = gP; out[persistent][1] = gC; out[persistent][2] = metrics;
  }
  if (rlx_lstm_persistent_launch_count() != expect_launches) { printf("the persistent path did not run\n"); return 4; }
  for (int k = 0; k < 3; ++k) {
    for (float x : out[1][k]) if (!std::isfinite(x)) { printf("non-finite output\n"); return 5; }
    if (memcmp(out[0][k].data(), out[1][k].data(), out[0][k].size() * sizeof(float))) { printf("persistent and per-step paths differ\n"); return 6; }
  }
  return 0;
}

int main() {
  // lstm_dim 4: 32 envs per block, 7 envs -> one block, most of its env slots inactive;  lstm_dim 64 (the reference width): 2 envs per
  // block of 128 threads, the recurrent kernel takes 64 KB of (emulated) shared memory, 3 envs -> the second block has an inactive slot
  int rc = run_case(rlx_lstm_dims{6, 2, 12, 8, 4, 0}, 9, 7, 2);
  if (!rc) rc = run_case(rlx_lstm_dims{6, 2, 16, 16, 64, 0}, 5, 3, 4);
  if (!rc) printf("ok\n");
  return rc;
}
