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
                                                            long long* done_count, int vec) {
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


// ----------------------------------------------------------------------------------- FastSAC n-step replay sample
// One warp per sample.  Lane j < n_steps loads the scalars of step j; every lane then walks the (<= 32) steps through shuffles, so the
// whole warp knows the final row and copies the observation / action rows with coalesced 128-byte accesses.
struct NStepP {
  const long long *idx_t, *idx_e;
  long long n, capacity, nr_envs, size, pos;
  int obs, act, n_steps;
  const float *discounts, *states, *next_states, *actions, *rewards, *dones, *truncs;
  float *o_states, *o_next, *o_actions, *o_rewards, *o_dones, *o_truncs, *o_eff;
};
__global__ void __launch_bounds__(256) replay_nstep_kernel(const NStepP p) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const bool full = p.size >= p.capacity;
  const long long last_idx = ((p.pos - 1) % p.capacity + p.capacity) % p.capacity;
  for (long long i = warp; i < p.n; i += nwarps) {
    const long long t0 = p.idx_t[i], e = p.idx_e[i];
    const long long src0 = t0 * p.nr_envs + e;
    long long fin = src0;  // ring row the next state / done / truncation are read from
    float reward, done, trunc, eff;
    if (p.n_steps == 1) {
      reward = p.rewards[src0];
      done = p.dones[src0];
      trunc = p.truncs[src0];
      eff = 1.f;
    } else {
      float r = 0.f, d = 0.f, tr = 0.f, disc = 0.f;
      if (lane < p.n_steps) {
        const long long t = (t0 + lane) % p.capacity;
        const long long src = t * p.nr_envs + e;
        r = p.rewards[src];
        d = p.dones[src];
        tr = p.truncs[src];
        if (full && t == last_idx) tr = (d > 0.f) ? tr : 1.f;  // newest row: the episode continues outside the ring
        disc = p.discounts[lane];
      }
      float mask = 1.f, acc = 0.f;
      eff = 0.f;
      int first_done = p.n_steps - 1, first_trunc = p.n_steps - 1;
      bool seen_done = false, seen_trunc = false;
      for (int j = 0; j < p.n_steps; ++j) {
        const float rj = __shfl_sync(0xffffffffu, r, j), dj = __shfl_sync(0xffffffffu, d, j), tj = __shfl_sync(0xffffffffu, tr, j);
        const float cj = __shfl_sync(0xffffffffu, disc, j);
        acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(rj, mask), cj));  // (r * mask) * discount, summed in step order
        eff += mask;
        if (!seen_done && dj > 0.f) { first_done = j; seen_done = true; }
        if (!seen_trunc && tj > 0.f) { first_trunc = j; seen_trunc = true; }
        mask = __fmul_rn(mask, __fsub_rn(1.f, dj));  // cumprod(1 - dones shifted by one step)
      }
      const int off = min(first_done, first_trunc);
      reward = acc;
      done = __shfl_sync(0xffffffffu, d, off);
      trunc = __shfl_sync(0xffffffffu, tr, off);
      fin = ((t0 + off) % p.capacity) * p.nr_envs + e;
    }
    for (int k = lane; k < p.obs; k += 32) {
      p.o_states[i * p.obs + k] = p.states[src0 * p.obs + k];
      p.o_next[i * p.obs + k] = p.next_states[fin * p.obs + k];
    }
    for (int k = lane; k < p.act; k += 32) p.o_actions[i * p.act + k] = p.actions[src0 * p.act + k];
    if (lane == 0) {
      p.o_rewards[i] = reward;
      p.o_dones[i] = done;
      p.o_truncs[i] = trunc;
      p.o_eff[i] = eff;
    }
  }
}

}  // namespace rlx

using namespace rlx;

extern "C" int rlx_rollout_store_f32(const float* reward, const uint8_t* terminated, const uint8_t* truncated, const float* next_obs,
                                     int64_t n, int64_t obs_dim, float* rewards_row, float* terminations_row, float* next_obs_dst,
                                     int64_t* done_count, void* stream) {
  RLX_CHECK_ARG(n >= 0 && obs_dim >= 0, "negative size");
  if (n == 0) return RLX_OK;
  RLX_CHECK_ARG((reward && terminated) || (next_obs && next_obs_dst), "nothing to store");
  RLX_CHECK_ARG(!rewards_row || reward, "rewards_row given without reward");
  const int vec = (next_obs && next_obs_dst && aligned16(next_obs) && aligned16(next_obs_dst) && ((n * obs_dim) % 4 == 0)) ? 1 : 0;
  const long long work = (next_obs && next_obs_dst) ? (vec ? n * obs_dim / 4 : n * obs_dim) : n;
  const unsigned grid = (unsigned)std::min<long long>(ceil_div(std::max<long long>(work, n), 256), (long long)sm_count() * 8);
  RLX_LAUNCH_C(KC_STORE, 0, ((next_obs && next_obs_dst) ? 8.0 * n * obs_dim : 0.0) + 10.0 * n, rollout_store_kernel, grid, 256, 0, stream, reward, terminated, truncated, next_obs, (long long)n, (long long)obs_dim,
             rewards_row, terminations_row, next_obs_dst, (long long*)done_count, vec);
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

extern "C" int rlx_replay_sample_nstep_f32(const int64_t* idx_t, const int64_t* idx_e, int64_t n, int64_t capacity, int64_t nr_envs,
                                           int64_t obs_dim, int64_t act_dim, int32_t n_steps, const float* discounts, int64_t size,
                                           int64_t pos, const float* states, const float* next_states, const float* actions,
                                           const float* rewards, const float* dones, const float* truncations, float* out_states,
                                           float* out_next_states, float* out_actions, float* out_rewards, float* out_dones,
                                           float* out_truncations, float* out_effective_n_steps, void* stream) {
  RLX_CHECK_ARG(n >= 0 && capacity > 0 && nr_envs > 0 && obs_dim > 0 && act_dim > 0, "bad sizes");
  RLX_CHECK_ARG(n_steps >= 1 && n_steps <= 32, "n_steps must be in [1, 32]");
  RLX_CHECK_ARG(size >= 0 && size <= capacity && pos >= 0 && pos < capacity, "bad ring state");
  if (n == 0) return RLX_OK;
  RLX_CHECK_ARG(idx_t && idx_e && states && next_states && actions && rewards && dones && truncations, "null input");
  RLX_CHECK_ARG(n_steps == 1 || discounts != nullptr, "discounts is required for n_steps > 1");
  RLX_CHECK_ARG(out_states && out_next_states && out_actions && out_rewards && out_dones && out_truncations && out_effective_n_steps,
                "null output");
  NStepP p{(const long long*)idx_t, (const long long*)idx_e, n, capacity, nr_envs, size, pos, (int)obs_dim, (int)act_dim, n_steps,
           discounts, states, next_states, actions, rewards, dones, truncations, out_states, out_next_states, out_actions, out_rewards,
           out_dones, out_truncations, out_effective_n_steps};
  const unsigned grid = (unsigned)std::min<long long>(ceil_div(n, 8), (long long)sm_count() * 16);
  // algorithmic bytes: two index words, the n_steps scalar triples, two observation rows and one action row read, the same rows written
  const double bytes = (double)n * (16.0 + 12.0 * n_steps + 4.0 * (4.0 * obs_dim + 2.0 * act_dim + 4.0));
  RLX_LAUNCH_C(KC_GATHER, 0, bytes, replay_nstep_kernel, grid, 256, 0, stream, p);
  return RLX_OK;
}
