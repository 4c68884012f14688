// HBM-bound data-movement kernels of the hot path: rollout row stores, minibatch gather by permutation,
// per-minibatch advantage statistics, SAC replay sample-gather and Polyak averaging.
#include "common.cuh"

namespace rlx {

// 128-bit streaming accessors (rows are touched once: do not pollute L1)
__device__ __forceinline__ float4 ld_stream4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream4(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}

// ------------------------------------------------------------------------------------------------ rollout store
__global__ void __launch_bounds__(256) rollout_store_kernel(const float* __restrict__ reward, const uint8_t* __restrict__ terminated,
                                                            const uint8_t* __restrict__ truncated, const float* __restrict__ next_obs,
                                                            long long n, long long obs_dim, float* __restrict__ rewards_row,
                                                            float* __restrict__ term_row, float* __restrict__ next_dst,
                                                            long long* done_count, int vec, float* __restrict__ ep_return,
                                                            float* __restrict__ ep_length, float* __restrict__ done_return_row,
                                                            float* __restrict__ done_length_row) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (next_obs != nullptr && next_dst != nullptr) {
    const long long tot = n * obs_dim;
    if (vec) {
      const float4* s = reinterpret_cast<const float4*>(next_obs);
      float4* d = reinterpret_cast<float4*>(next_dst);
      for (long long i = tid; i < tot / 4; i += stride) st_stream4(d + i, ld_stream4(s + i));
    } else {
      for (long long i = tid; i < tot; i += stride) next_dst[i] = next_obs[i];
    }
  }
  int dones = 0;
  if (terminated != nullptr) {
    for (long long i = tid; i < n; i += stride) {
      const bool te = terminated[i] != 0;
      const bool tr = truncated != nullptr && truncated[i] != 0;
      if (rewards_row && reward) rewards_row[i] = reward[i];
      if (term_row) term_row[i] = te ? 1.f : 0.f;
      dones += (te || tr) ? 1 : 0;
      if (ep_return != nullptr) {
        // device-side episode statistics with the semantics of the reference's torch-interface env
        // (custom_mujoco/ant/warp_torch/environment.py:159-178): step count and return accumulate, a finished episode publishes both
        // for this step and restarts from zero.  A length of 0 in the row means "no episode ended here".
        const float ret = ep_return[i] + reward[i];
        const float len = ep_length[i] + 1.f;
        const bool done = te || tr;
        done_return_row[i] = done ? ret : 0.f;
        done_length_row[i] = done ? len : 0.f;
        ep_return[i] = done ? 0.f : ret;
        ep_length[i] = done ? 0.f : len;
      }
    }
  }
  if (done_count != nullptr) {
    // warp-aggregated count, one atomic per warp that saw a done
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dones += __shfl_xor_sync(0xffffffffu, dones, o);
    if ((threadIdx.x & 31) == 0 && dones > 0) atomicAdd(reinterpret_cast<unsigned long long*>(done_count), (unsigned long long)dones);
  }
}

// -------------------------------------------------------------------------------------------- minibatch gather
// One warp per gathered row: the 8-byte index is read once and broadcast, the state row (obs_dim*4 bytes, 1504 B for
// obs 376) moves as 128-bit streaming loads/stores, the action row and the three scalars ride along.
struct GatherP {
  const long long* idx;
  long long count;
  int obs, act;
  const float *states, *actions, *logp, *adv, *ret;
  float *o_states, *o_actions, *o_logp, *o_adv, *o_ret;
  int vec_obs;
  long long o_ld;  // row pitch of o_states
};

__global__ void __launch_bounds__(256) gather_minibatch_kernel(const GatherP p) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < p.count; r += nwarps) {
    const long long src = p.idx[r];
    if (p.o_states) {
      if (p.vec_obs) {
        const float4* s = reinterpret_cast<const float4*>(p.states + src * p.obs);
        float4* d = reinterpret_cast<float4*>(p.o_states + r * p.o_ld);
        const int nv = p.obs >> 2;
        for (int i = lane; i < nv; i += 32) st_stream4(d + i, ld_stream4(s + i));
      } else {
        const float* s = p.states + src * p.obs;
        float* d = p.o_states + r * p.o_ld;
        for (int i = lane; i < p.obs; i += 32) d[i] = s[i];
      }
      if (p.o_ld > p.obs && lane < (int)(p.o_ld - p.obs)) p.o_states[r * p.o_ld + p.obs + lane] = (lane == 0) ? 1.f : 0.f;
    }
    if (p.o_actions) {
      const float* s = p.actions + src * p.act;
      float* d = p.o_actions + r * p.act;
      for (int i = lane; i < p.act; i += 32) d[i] = s[i];
    }
    if (lane == 0) {
      if (p.o_logp) p.o_logp[r] = p.logp[src];
      if (p.o_adv) p.o_adv[r] = p.adv[src];
      if (p.o_ret) p.o_ret[r] = p.ret[src];
    }
  }
}

// ------------------------------------------------------------------------------------- advantage mean / std
// One CTA per minibatch, two passes (mean, then centred sum of squares) for fp32 accuracy.  ref: ppo.py:133-134
__global__ void __launch_bounds__(512) advantage_stats_kernel(const float* __restrict__ adv, long long count, long long mb,
                                                              float* __restrict__ stats) {
  __shared__ float sh[34];
  const long long b0 = (long long)blockIdx.x * mb;
  const long long n = min(mb, count - b0);
  const float* __restrict__ a = adv + b0;
  float s = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) s += a[i];
  const float mean = block_sum(s, sh) / (float)n;
  float q = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = a[i] - mean;
    q = fmaf(d, d, q);
  }
  q = block_sum(q, sh);
  if (threadIdx.x == 0) {
    stats[2 * blockIdx.x] = mean;
    stats[2 * blockIdx.x + 1] = sqrtf(q / (float)(n - 1));  // unbiased (ddof = 1); n == 1 gives NaN like torch
  }
}

// Sharded form of the statistics above: a rank holds only its rows of every minibatch, so the mean / centred sum of squares are
// built from per-rank partial sums with an all-reduce in between.  One CTA per segment (minibatch); centre == nullptr: plain sums.
__global__ void __launch_bounds__(512) segment_moments_kernel(const float* __restrict__ x, const long long* __restrict__ offsets,
                                                              const float* __restrict__ gsum, const float* __restrict__ gcount,
                                                              float* __restrict__ out) {
  __shared__ float sh[34];
  const long long b0 = offsets[blockIdx.x], n = offsets[blockIdx.x + 1] - b0;
  const float* __restrict__ a = x + b0;
  const float centre = gsum ? gsum[blockIdx.x] / gcount[blockIdx.x] : 0.f;
  float s = 0.f;
  if (gsum) {
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      const float d = a[i] - centre;
      s = fmaf(d, d, s);
    }
  } else {
    for (long long i = threadIdx.x; i < n; i += blockDim.x) s += a[i];
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) out[blockIdx.x] = s;
}

// ----------------------------------------------------------------------------------------- SAC replay gather
struct ReplayGatherP {
  const long long *idx_t, *idx_e;
  long long n, nr_envs;
  int obs, act;
  const float *states, *next_states, *actions, *rewards, *terms;
  float *o_states, *o_next, *o_actions, *o_rewards, *o_terms;
};
__global__ void __launch_bounds__(256) replay_gather_kernel(const ReplayGatherP p) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < p.n; r += nwarps) {
    const long long src = p.idx_t[r] * p.nr_envs + p.idx_e[r];
    for (int i = lane; i < p.obs; i += 32) {
      p.o_states[r * p.obs + i] = p.states[src * p.obs + i];
      p.o_next[r * p.obs + i] = p.next_states[src * p.obs + i];
    }
    for (int i = lane; i < p.act; i += 32) p.o_actions[r * p.act + i] = p.actions[src * p.act + i];
    if (lane == 0) {
      p.o_rewards[r] = p.rewards[src];
      p.o_terms[r] = p.terms[src];
    }
  }
}

__global__ void __launch_bounds__(256) polyak_kernel(float* __restrict__ target, const float* __restrict__ online, long long n, float tau) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    // sac.py:240-241: target.mul_(1 - tau); target.add_(online * tau)  -- two roundings, kept as written
    target[i] = __fadd_rn(__fmul_rn(target[i], 1.f - tau), __fmul_rn(online[i], tau));
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }


}  // namespace rlx

using namespace rlx;

extern "C" int rlx_rollout_store_f32(const float* reward, const uint8_t* terminated, const uint8_t* truncated, const float* next_obs,
                                     int64_t n, int64_t obs_dim, float* rewards_row, float* terminations_row, float* next_obs_dst,
                                     int64_t* done_count, void* stream) {
  return rlx_rollout_store_stats_f32(reward, terminated, truncated, next_obs, n, obs_dim, rewards_row, terminations_row, next_obs_dst, done_count,
                                     nullptr, nullptr, nullptr, nullptr, stream);
}

extern "C" int rlx_rollout_store_stats_f32(const float* reward, const uint8_t* terminated, const uint8_t* truncated, const float* next_obs,
                                           int64_t n, int64_t obs_dim, float* rewards_row, float* terminations_row, float* next_obs_dst,
                                           int64_t* done_count, float* episode_return, float* episode_length, float* done_return_row,
                                           float* done_length_row, void* stream) {
  RLX_CHECK_ARG(n >= 0 && obs_dim >= 0, "negative size");
  if (n == 0) return RLX_OK;
  RLX_CHECK_ARG((reward && terminated) || (next_obs && next_obs_dst), "nothing to store");
  RLX_CHECK_ARG(!rewards_row || reward, "rewards_row given without reward");
  RLX_CHECK_ARG(!episode_return || (reward && terminated && episode_length && done_return_row && done_length_row),
                "episode statistics need reward, terminated and all four statistics arrays");
  const int vec = (next_obs && next_obs_dst && aligned16(next_obs) && aligned16(next_obs_dst) && ((n * obs_dim) % 4 == 0)) ? 1 : 0;
  const long long work = (next_obs && next_obs_dst) ? (vec ? n * obs_dim / 4 : n * obs_dim) : n;
  const unsigned grid = (unsigned)std::min<long long>(ceil_div(std::max<long long>(work, n), 256), (long long)sm_count() * 8);
  RLX_LAUNCH_C(KC_STORE, 0, ((next_obs && next_obs_dst) ? 8.0 * n * obs_dim : 0.0) + 10.0 * n, rollout_store_kernel, grid, 256, 0, stream, reward, terminated, truncated, next_obs, (long long)n, (long long)obs_dim,
             rewards_row, terminations_row, next_obs_dst, (long long*)done_count, vec, episode_return, episode_length, done_return_row, done_length_row);
  return RLX_OK;
}

extern "C" int rlx_gather_minibatch_f32(const int64_t* idx, int64_t count, int64_t obs_dim, int64_t act_dim, const float* states,
                                        const float* actions, const float* log_probs, const float* advantages, const float* returns,
                                        float* out_states, float* out_actions, float* out_log_probs, float* out_advantages,
                                        float* out_returns, int64_t out_states_ld, void* stream) {
  RLX_CHECK_ARG(count >= 0 && obs_dim > 0 && act_dim > 0, "bad sizes");
  if (count == 0) return RLX_OK;
  RLX_CHECK_ARG(idx != nullptr, "idx is null");
  RLX_CHECK_ARG(!out_states || states, "states is null");
  RLX_CHECK_ARG(!out_actions || actions, "actions is null");
  RLX_CHECK_ARG(!out_log_probs || log_probs, "log_probs is null");
  RLX_CHECK_ARG(!out_advantages || advantages, "advantages is null");
  RLX_CHECK_ARG(!out_returns || returns, "returns is null");
  GatherP p{(const long long*)idx, count, (int)obs_dim, (int)act_dim, states, actions, log_probs, advantages, returns,
            out_states, out_actions, out_log_probs, out_advantages, out_returns, 0, 0};
  p.o_ld = out_states_ld > 0 ? out_states_ld : obs_dim;
  RLX_CHECK_ARG(p.o_ld >= obs_dim && p.o_ld - obs_dim <= 32, "out_states_ld must be in [obs_dim, obs_dim + 32]");
  p.vec_obs = (obs_dim % 4 == 0 && p.o_ld % 4 == 0 && aligned16(states) && aligned16(out_states)) ? 1 : 0;
  // 8 warps per CTA, grid sized to a multiple of the SM count (persistent-style grid-stride loop over rows)
  const long long want = ceil_div(count, 8);
  const unsigned grid = (unsigned)std::min<long long>(want, (long long)sm_count() * 16);
  RLX_LAUNCH_C(KC_GATHER, 0, (double)count * (8.0 + 8.0 * (obs_dim + act_dim + 3)), gather_minibatch_kernel, grid, 256, 0, stream, p);
  return RLX_OK;
}

extern "C" int rlx_advantage_stats_f32(const float* adv, int64_t count, int64_t mb, float* stats, void* stream) {
  RLX_CHECK_ARG(count >= 0 && mb > 0, "bad sizes");
  if (count == 0) return RLX_OK;
  RLX_CHECK_ARG(adv && stats, "null pointer");
  const unsigned grid = (unsigned)ceil_div(count, mb);
  RLX_LAUNCH_C(KC_ADV_STATS, 0, 4.0 * count, advantage_stats_kernel, grid, 512, 0, stream, adv, (long long)count, (long long)mb, stats);
  return RLX_OK;
}

extern "C" int rlx_segment_moments_f32(const float* x, const int64_t* offsets, int64_t nseg, const float* gsum, const float* gcount,
                                       float* out, void* stream) {
  RLX_CHECK_ARG(nseg >= 0, "negative segment count");
  if (nseg == 0) return RLX_OK;
  RLX_CHECK_ARG(x && offsets && out, "null pointer");
  RLX_CHECK_ARG((gsum == nullptr) == (gcount == nullptr), "gsum and gcount go together");
  RLX_LAUNCH_C(KC_ADV_STATS, 0, 0, segment_moments_kernel, (unsigned)nseg, 512, 0, stream, x, (const long long*)offsets, gsum, gcount, out);
  return RLX_OK;
}

extern "C" int rlx_replay_sample_gather_f32(const int64_t* idx_t, const int64_t* idx_e, int64_t n, int64_t nr_envs, int64_t obs_dim,
                                            int64_t act_dim, const float* states, const float* next_states, const float* actions,
                                            const float* rewards, const float* terminations, float* out_states,
                                            float* out_next_states, float* out_actions, float* out_rewards, float* out_terminations,
                                            void* stream) {
  RLX_CHECK_ARG(n >= 0 && nr_envs > 0 && obs_dim > 0 && act_dim > 0, "bad sizes");
  if (n == 0) return RLX_OK;
  RLX_CHECK_ARG(idx_t && idx_e && states && next_states && actions && rewards && terminations, "null input");
  RLX_CHECK_ARG(out_states && out_next_states && out_actions && out_rewards && out_terminations, "null output");
  ReplayGatherP p{(const long long*)idx_t, (const long long*)idx_e, n, nr_envs, (int)obs_dim, (int)act_dim, states, next_states,
                  actions, rewards, terminations, out_states, out_next_states, out_actions, out_rewards, out_terminations};
  const unsigned grid = (unsigned)std::min<long long>(ceil_div(n, 8), (long long)sm_count() * 16);
  RLX_LAUNCH(replay_gather_kernel, grid, 256, 0, stream, p);
  return RLX_OK;
}

extern "C" int rlx_polyak_f32(float* target, const float* online, int64_t n, float tau, void* stream) {
  RLX_CHECK_ARG(n >= 0, "negative size");
  if (n == 0) return RLX_OK;
  RLX_CHECK_ARG(target && online, "null pointer");
  const unsigned grid = (unsigned)std::min<long long>(ceil_div(n, 256), (long long)sm_count() * 8);
  RLX_LAUNCH(polyak_kernel, grid, 256, 0, stream, target, online, (long long)n, tau);
  return RLX_OK;
}
