// FastSAC's n-step replay sampling (rl_x/algorithms/fastsac/pytorch/replay_buffer.py:34-96) — see rlx_replay_sample_nstep_f32 in
// include/rlx_b200.h.  Two flat launches: (1) one thread per sample walks the <= 32 steps of its window (n-step reward, effective
// length, first done / truncation -> the ring row the bootstrap state is taken from), (2) one thread per output float copies the
// observation / next-observation / action rows (consecutive threads = consecutive floats of a row: coalesced).
// Dual build (dual_build.cuh): the host emulation of this very file is checked against the golden vectors of the executed reference
// in tests/test_lstm_emulation.py::test_emulated_nstep_replay_matches_reference_golden.
#include "dual_build.cuh"

namespace rlx {
namespace nstep {

struct P {
  const long long *idx_t, *idx_e;
  long long n, capacity, nr_envs, size, pos;
  int obs, act, n_steps;
  const float *discounts, *states, *next_states, *actions, *rewards, *dones, *truncs;
  float *o_states, *o_next, *o_actions, *o_rewards, *o_dones, *o_truncs, *o_eff;
  long long* final_row;  // [n] scratch: ring row (t * nr_envs + e) of the bootstrap state
};

__device__ __forceinline__ long long gtid() { return (long long)blockIdx.x * blockDim.x + threadIdx.x; }

__global__ void scalars_kernel(const P p) {
  const long long i = gtid();
  if (i >= p.n) return;
  const long long t0 = p.idx_t[i], e = p.idx_e[i];
  const long long src0 = t0 * p.nr_envs + e;
  if (p.n_steps == 1) {  // replay_buffer.py:36-47
    p.o_rewards[i] = p.rewards[src0];
    p.o_dones[i] = p.dones[src0];
    p.o_truncs[i] = p.truncs[src0];
    p.o_eff[i] = 1.f;
    p.final_row[i] = src0;
    return;
  }
  const bool full = p.size >= p.capacity;
  const long long last_idx = ((p.pos - 1) % p.capacity + p.capacity) % p.capacity;
  float mask = 1.f, acc = 0.f, eff = 0.f;
  int first_done = p.n_steps - 1, first_trunc = p.n_steps - 1;
  bool seen_done = false, seen_trunc = false;
  for (int j = 0; j < p.n_steps; ++j) {
    const long long t = (t0 + j) % p.capacity;
    const long long src = t * p.nr_envs + e;
    const float d = p.dones[src];
    float tr = p.truncs[src];
    if (full && t == last_idx) tr = (d > 0.f) ? tr : 1.f;  // :50-57 newest row of a full ring: the episode continues outside the ring
    acc = rn_add(acc, rn_mul(rn_mul(p.rewards[src], mask), p.discounts[j]));  // (r * mask) * discount, summed in step order
    eff = rn_add(eff, mask);
    if (!seen_done && d > 0.f) { first_done = j; seen_done = true; }
    if (!seen_trunc && tr > 0.f) { first_trunc = j; seen_trunc = true; }
    mask = rn_mul(mask, rn_sub(1.f, d));                    // cumprod(1 - dones shifted by one step)
  }
  const int off = first_done < first_trunc ? first_done : first_trunc;
  const long long tf = (t0 + off) % p.capacity;
  const long long fin = tf * p.nr_envs + e;
  float trf = p.truncs[fin];
  if (full && tf == last_idx) trf = (p.dones[fin] > 0.f) ? trf : 1.f;
  p.o_rewards[i] = acc;
  p.o_dones[i] = p.dones[fin];
  p.o_truncs[i] = trf;
  p.o_eff[i] = eff;
  p.final_row[i] = fin;
}

// thread = one output float of [states | next_states | actions] of one sample
__global__ void rows_kernel(const P p) {
  const long long id = gtid();
  const long long per = 2LL * p.obs + p.act;
  if (id >= p.n * per) return;
  const long long i = id / per, k = id % per;
  const long long src0 = p.idx_t[i] * p.nr_envs + p.idx_e[i];
  if (k < p.obs) {
    p.o_states[i * p.obs + k] = p.states[src0 * p.obs + k];
  } else if (k < 2LL * p.obs) {
    const long long c = k - p.obs;
    p.o_next[i * p.obs + c] = p.next_states[p.final_row[i] * p.obs + c];
  } else {
    const long long c = k - 2LL * p.obs;
    p.o_actions[i * p.act + c] = p.actions[src0 * p.act + c];
  }
}

}  // namespace nstep
}  // namespace rlx

using namespace rlx;

extern "C" int rlx_replay_sample_nstep_f32(const int64_t* idx_t, const int64_t* idx_e, int64_t n, int64_t capacity, int64_t nr_envs,
                                           int64_t obs_dim, int64_t act_dim, int32_t n_steps, const float* discounts, int64_t size,
                                           int64_t pos, const float* states, const float* next_states, const float* actions,
                                           const float* rewards, const float* dones, const float* truncations, float* out_states,
                                           float* out_next_states, float* out_actions, float* out_rewards, float* out_dones,
                                           float* out_truncations, float* out_effective_n_steps, int64_t* workspace, void* stream) {
  RLX_CHECK_ARG(n >= 0 && capacity > 0 && nr_envs > 0 && obs_dim > 0 && act_dim > 0, "bad sizes");
  RLX_CHECK_ARG(n_steps >= 1 && n_steps <= 32, "n_steps must be in [1, 32]");
  RLX_CHECK_ARG(size >= 0 && size <= capacity && pos >= 0 && pos < capacity, "bad ring state");
  if (n == 0) return RLX_OK;
  RLX_CHECK_ARG(idx_t && idx_e && states && next_states && actions && rewards && dones && truncations, "null input");
  RLX_CHECK_ARG(n_steps == 1 || discounts != nullptr, "discounts is required for n_steps > 1");
  RLX_CHECK_ARG(out_states && out_next_states && out_actions && out_rewards && out_dones && out_truncations && out_effective_n_steps && workspace,
                "null output / workspace");
  nstep::P p{(const long long*)idx_t, (const long long*)idx_e, n, capacity, nr_envs, size, pos, (int)obs_dim, (int)act_dim, n_steps, discounts,
             states, next_states, actions, rewards, dones, truncations, out_states, out_next_states, out_actions, out_rewards, out_dones,
             out_truncations, out_effective_n_steps, (long long*)workspace};
  RLX_FLAT_LAUNCH(nstep::scalars_kernel, (long long)n, stream, p);
  RLX_FLAT_LAUNCH(nstep::rows_kernel, (long long)n * (2 * obs_dim + act_dim), stream, p);
  return RLX_OK;
}
