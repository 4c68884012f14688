// Output-layer ("head") kernels of the PPO policy/critic MLP pair.
//
// The last layers are skinny (hidden -> act and hidden -> 1), i.e. HBM-bound row reductions, so they are not run as
// GEMMs: one warp owns one row of the second hidden activation H2 = [H2p | H2c] ([M, 2H]), computes the act+1 dot
// products with warp-shuffle reductions and then everything that hangs off them in registers:
//   rollout head : action sampling, log-prob, action clip/rescale, value         (ref: policy.py:61-73, critic.py:44-46)
//   train head   : log-prob, ratio, clipped surrogate, value loss, their gradients wrt mean / logstd / value, the
//                  back-propagated dZ2 = dH2 * (1 - H2^2) and the metric sums   (ref: ppo.py:121-141,153-157)
#pragma once
#include "common.cuh"

namespace rlx {

constexpr float kLogSqrt2Pi = 0.91893853320467274178f;  // log(sqrt(2*pi)), torch/distributions/normal.py:101

// ---------------------------------------------------------------------------- counter-based normal generator
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
// Standard normal number `j` of row `row` for call `offset` under `seed` (Box-Muller on Philox output).
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t offset, uint32_t row, uint32_t j) {
  uint32_t c[4] = {row, j >> 2, (uint32_t)offset, (uint32_t)(offset >> 32)};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const uint32_t a = (j & 2) ? c[2] : c[0], b = (j & 2) ? c[3] : c[1];
  const float u1 = ((float)a + 0.5f) * 2.3283064365386963e-10f;  // (0, 1]
  const float u2 = ((float)b + 0.5f) * 2.3283064365386963e-10f;
  const float r = sqrtf(-2.f * logf(fminf(u1, 1.f)));
  float s, co;
  sincospif(2.f * u2, &s, &co);
  return (j & 1) ? r * s : r * co;
}

struct HeadP {
  int M, H, act;
  const float* H2;      // [M, 2H]
  const float* W3p;     // [act, H]
  const float* W3c;     // [H]
  const float* b3p;     // [act]
  const float* b3c;     // [1]
  const float* logstd;  // [act]
  // rollout
  const float* noise;   // [M, act] or null
  unsigned long long seed, offset;
  const float* act_low;
  const float* act_high;
  int clip_rescale, deterministic;
  float* action;        // [M, act]
  float* env_action;    // [M, act]
  float* logp_out;      // [M]
  float* value_out;     // [M]
  // train
  const float* actions; // [M, act]
  const float* logp_old;
  const float* adv;
  const float* ret;
  const float* adv_stats;  // [2]
  float inv_mg;            // 1 / m_global
  float clip_range, critic_coef;
  float* dZ2;              // [M, 2H]
  float* dhead;            // [M, act+1]   (dmean | dv)
  float* block_partials;   // [gridDim.x, 2*act+5+2H]  (db3p | db3c | dlogstd | pg | vloss | kl | clipfrac | db2p[H] | db2c[H])
};

// Loads the row's H2 slices and returns the act means (lane a holds mean a; a+32 in mean_hi) and the value (all lanes).
template <int NCH>
__device__ __forceinline__ void head_row_forward(const HeadP& p, const float* __restrict__ sW3p, const float* __restrict__ sW3c,
                                                 long long row, int lane, float (&hp)[NCH], float (&hc)[NCH], float& mean_lo,
                                                 float& mean_hi, float& value) {
  const float* __restrict__ h = p.H2 + row * (2LL * p.H);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int j = lane + 32 * c;
    hp[c] = (j < p.H) ? h[j] : 0.f;
    hc[c] = (j < p.H) ? h[p.H + j] : 0.f;
  }
  mean_lo = 0.f;
  mean_hi = 0.f;
  for (int a = 0; a < p.act; ++a) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int j = lane + 32 * c;
      if (j < p.H) s = fmaf(hp[c], sW3p[a * p.H + j], s);
    }
    s = warp_sum(s) + p.b3p[a];
    if (lane == (a & 31)) {
      if (a < 32) mean_lo = s; else mean_hi = s;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int j = lane + 32 * c;
    if (j < p.H) s = fmaf(hc[c], sW3c[j], s);
  }
  value = warp_sum(s) + p.b3c[0];
}

__device__ __forceinline__ void head_stage_weights(const HeadP& p, float* sW3p, float* sW3c) {
  for (int i = threadIdx.x; i < p.act * p.H; i += blockDim.x) sW3p[i] = p.W3p[i];
  for (int i = threadIdx.x; i < p.H; i += blockDim.x) sW3c[i] = p.W3c[i];
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ rollout head
template <int NCH>
__global__ void __launch_bounds__(256) ppo_head_rollout_kernel(const HeadP p) {
  extern __shared__ float smem[];
  float* sW3p = smem;
  float* sW3c = smem + p.act * p.H;
  head_stage_weights(p, sW3p, sW3c);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (long long row = (long long)blockIdx.x * nw + wib; row < p.M; row += (long long)gridDim.x * nw) {
    float hp[NCH], hc[NCH], mean_lo, mean_hi, value;
    head_row_forward<NCH>(p, sW3p, sW3c, row, lane, hp, hc, mean_lo, mean_hi, value);
    if (p.value_out && lane == 0) p.value_out[row] = value;
    float lp_sum = 0.f;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int a = lane + 32 * half;
      if (a < p.act) {
        const float mean = half ? mean_hi : mean_lo;
        float x = mean;
        if (!p.deterministic) {
          const float ls = p.logstd[a];
          const float sd = expf(ls);
          const float eps = p.noise ? p.noise[row * p.act + a] : philox_normal(p.seed, p.offset, (uint32_t)row, (uint32_t)a);
          x = __fadd_rn(mean, __fmul_rn(sd, eps));  // Normal.sample(): loc + scale * eps  (no FMA contraction)
          // Normal.log_prob (torch/distributions/normal.py:87-103)
          const float var = sd * sd;
          const float d = x - mean;
          lp_sum += -(d * d) / (2.f * var) - logf(sd) - kLogSqrt2Pi;
        }
        if (p.action) p.action[row * p.act + a] = x;
        if (p.env_action) {
          float e = x;
          if (p.clip_rescale) {  // policy.py:68-70
            const float c = fminf(fmaxf(x, -1.f), 1.f);
            const float lo = p.act_low[a], hi = p.act_high[a];
            e = lo + (0.5f * (c + 1.f)) * (hi - lo);
          }
          p.env_action[row * p.act + a] = e;
        }
      }
    }
    if (!p.deterministic && p.logp_out) {
      lp_sum = warp_sum(lp_sum);
      if (lane == 0) p.logp_out[row] = lp_sum;
    }
  }
}

// -------------------------------------------------------------------------------------------------- train head
template <int NCH>
__global__ void __launch_bounds__(256) ppo_head_train_kernel(const HeadP p) {
  extern __shared__ float smem[];
  float* sW3p = smem;
  float* sW3c = smem + p.act * p.H;
  float* sred = sW3c + p.H;  // [nw][2*act+5+2H]
  head_stage_weights(p, sW3p, sW3c);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int npart = 2 * p.act + 5 + 2 * p.H;
  float acc_db2p[NCH], acc_db2c[NCH];  // column sums of dZ2 = layer-2 bias gradients (lane owns columns lane + 32c)
#pragma unroll
  for (int c = 0; c < NCH; ++c) acc_db2p[c] = acc_db2c[c] = 0.f;

  const float adv_mean = p.adv_stats[0];
  const float adv_den = p.adv_stats[1] + 1e-8f;  // ppo.py:134
  const float clip_lo = 1.f - p.clip_range, clip_hi = 1.f + p.clip_range;

  // per-lane accumulators: lane a holds component a (and a+32)
  float acc_db3p[2] = {0.f, 0.f}, acc_dls[2] = {0.f, 0.f};
  float acc_db3c = 0.f, acc_pg = 0.f, acc_vl = 0.f, acc_kl = 0.f, acc_cf = 0.f;

  for (long long row = (long long)blockIdx.x * nw + wib; row < p.M; row += (long long)gridDim.x * nw) {
    float hp[NCH], hc[NCH], mean_lo, mean_hi, value;
    head_row_forward<NCH>(p, sW3p, sW3c, row, lane, hp, hc, mean_lo, mean_hi, value);

    // new log-prob (policy.py:76-82)
    float dmu[2] = {0.f, 0.f}, zz[2] = {0.f, 0.f};  // (x-mean)/var and (x-mean)^2/var per owned component
    float lp = 0.f;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int a = lane + 32 * half;
      if (a < p.act) {
        const float mean = half ? mean_hi : mean_lo;
        const float sd = expf(p.logstd[a]);
        const float var = sd * sd;
        const float d = p.actions[row * p.act + a] - mean;
        lp += -(d * d) / (2.f * var) - logf(sd) - kLogSqrt2Pi;
        dmu[half] = d / var;
        zz[half] = d * d / var;
      }
    }
    const float logp_new = warp_sum(lp);
    const float logratio = logp_new - p.logp_old[row];
    const float ratio = expf(logratio);
    const float A = (p.adv[row] - adv_mean) / adv_den;
    const float pg1 = -A * ratio;
    const float pg2 = -A * fminf(fmaxf(ratio, clip_lo), clip_hi);
    const float pg = fmaxf(pg1, pg2);
    // d pg / d ratio: torch.maximum splits ties evenly, clamp passes gradient on the closed interval.
    const float w1 = (pg1 > pg2) ? 1.f : ((pg1 == pg2) ? 0.5f : 0.f);
    const float inr = (ratio >= clip_lo && ratio <= clip_hi) ? 1.f : 0.f;
    const float dratio = -A * (w1 + (1.f - w1) * inr);
    const float dlogp = dratio * ratio * p.inv_mg;
    const float verr = value - p.ret[row];
    const float dv = p.critic_coef * verr * p.inv_mg;

    acc_pg += pg;
    acc_vl += 0.5f * verr * verr;
    acc_kl += (ratio - 1.f) - logratio;
    acc_cf += (fabsf(ratio - 1.f) > p.clip_range) ? 1.f : 0.f;
    acc_db3c += dv;

    float dmean[2];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      dmean[half] = dlogp * dmu[half];
      acc_db3p[half] += dmean[half];
      acc_dls[half] += dlogp * (zz[half] - 1.f);  // d logp / d logstd = (x-mean)^2/var - 1
      const int a = lane + 32 * half;
      if (a < p.act) p.dhead[row * (p.act + 1) + a] = dmean[half];
    }
    if (lane == 0) p.dhead[row * (p.act + 1) + p.act] = dv;

    // dZ2 = (dhead @ W3) * (1 - H2^2)
    float dz[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) dz[c] = 0.f;
    for (int a = 0; a < p.act; ++a) {
      const float g = __shfl_sync(0xffffffffu, (a < 32) ? dmean[0] : dmean[1], a & 31);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int j = lane + 32 * c;
        if (j < p.H) dz[c] = fmaf(g, sW3p[a * p.H + j], dz[c]);
      }
    }
    float* __restrict__ out = p.dZ2 + row * (2LL * p.H);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int j = lane + 32 * c;
      if (j < p.H) {
        const float zp = dz[c] * (1.f - hp[c] * hp[c]);
        const float zc = dv * sW3c[j] * (1.f - hc[c] * hc[c]);
        out[j] = zp;
        out[p.H + j] = zc;
        acc_db2p[c] += zp;
        acc_db2c[c] += zc;
      }
    }
  }

  // ---- block reduction of the per-warp accumulators (fixed order => deterministic)
  float* my = sred + wib * npart;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int a = lane + 32 * half;
    if (a < p.act) {
      my[a] = acc_db3p[half];
      my[p.act + 1 + a] = acc_dls[half];
    }
  }
  if (lane == 0) {
    my[p.act] = acc_db3c;
    my[2 * p.act + 1] = acc_pg;
    my[2 * p.act + 2] = acc_vl;
    my[2 * p.act + 3] = acc_kl;
    my[2 * p.act + 4] = acc_cf;
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int j = lane + 32 * c;
    if (j < p.H) {
      my[2 * p.act + 5 + j] = acc_db2p[c];
      my[2 * p.act + 5 + p.H + j] = acc_db2c[c];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < npart; i += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < nw; ++w) s += sred[w * npart + i];
    p.block_partials[(long long)blockIdx.x * npart + i] = s;
  }
}

// ------------------------------------------------------------------ weight gradient of the head (thread per column)
// part[chunk][(act+1)*H]: rows 0..act-1 = dW3p, row act = dW3c, summed over the chunk's rows.
struct HeadWgradP {
  int M, H, act, rows_per_chunk;
  const float* H2;     // [M, 2H]
  const float* dhead;  // [M, act+1]
  float* part;         // [nchunk, (act+1)*H]
};

template <int ACTMAX>
__global__ void __launch_bounds__(256) ppo_head_wgrad_kernel(const HeadWgradP p) {
  extern __shared__ float sd[];  // [rows_per_chunk][act+1]
  const int chunk = blockIdx.x;
  const long long r0 = (long long)chunk * p.rows_per_chunk;
  const int nrows = (int)min((long long)p.rows_per_chunk, (long long)p.M - r0);
  const int w = p.act + 1;
  for (int i = threadIdx.x; i < nrows * w; i += blockDim.x) sd[i] = p.dhead[r0 * w + i];
  __syncthreads();
  const int jj = blockIdx.y * blockDim.x + threadIdx.x;  // column of H2 in [0, 2H)
  if (jj >= 2 * p.H) return;
  const bool is_pol = jj < p.H;
  float acc[ACTMAX];
#pragma unroll
  for (int a = 0; a < ACTMAX; ++a) acc[a] = 0.f;
  const float* __restrict__ col = p.H2 + r0 * (2LL * p.H) + jj;
  if (is_pol) {
#pragma unroll 4
    for (int r = 0; r < nrows; ++r) {
      const float h = col[(long long)r * 2 * p.H];
#pragma unroll
      for (int a = 0; a < ACTMAX; ++a)
        if (a < p.act) acc[a] = fmaf(sd[r * w + a], h, acc[a]);
    }
    float* out = p.part + (long long)chunk * w * p.H + jj;
#pragma unroll
    for (int a = 0; a < ACTMAX; ++a)
      if (a < p.act) out[(long long)a * p.H] = acc[a];
  } else {
    float s = 0.f;
#pragma unroll 4
    for (int r = 0; r < nrows; ++r) s = fmaf(sd[r * w + p.act], col[(long long)r * 2 * p.H], s);
    p.part[(long long)chunk * w * p.H + (long long)p.act * p.H + (jj - p.H)] = s;
  }
}

}  // namespace rlx
