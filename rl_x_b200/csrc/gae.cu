// GAE backward scan (ref: calculate_gae_advantages_and_returns, rl_x/algorithms/ppo/pytorch/ppo.py:110-118).
//
// HBM-bound: 20-24 algorithmic bytes per (t, env).  The recurrence is serial in t but independent per env, and the
// loads do not depend on the recurrence, so a CTA owns a strip of 32 envs and
//   phase 1: all 256 threads stream the [TT x 32] tiles of rewards / terminations / values (/ next_values) into
//            shared memory (fully coalesced 128-byte rows, every load independent => deep memory-level parallelism),
//   phase 2: one warp (lane = env) runs the bit-exact sequential recurrence out of shared memory,
//   phase 3: all threads stream advantages / returns back out, coalesced.
// Tiles of TT time steps are processed from the end of the rollout backwards; the carry (last advantage, next value)
// stays in the scanning warp's registers.  Arithmetic follows the reference's fp32 evaluation order with explicit
// __fmul_rn/__fadd_rn so that no FMA contraction changes a bit.
#include "common.cuh"

namespace rlx {

constexpr int GAE_ENVS = 32;   // envs per CTA (one 128-byte row segment)

struct GaeP {
  const float* r;
  const float* term;
  const float* v;
  const float* nv;      // [T, N] or null
  const float* last_v;  // [N] (used when nv == null)
  float* adv;
  float* ret;
  long long T, N;
  float gamma_f;        // (float)gamma
  float gl_f;           // (float)(gamma * lambda), product taken in double as TorchScript does
  int bf16;             // bf16-autocast mode (ppo.py:98-107): next_values is a bf16 tensor there, so `gamma * next_values` is a bf16 product
};

// GAE_TT: time steps per shared-memory tile.  128 = the whole config-2 rollout in one DRAM round trip (few CTAs, latency matters);
// 64 = half the registers and shared memory per CTA, so more CTAs overlap their load / scan / store phases (many envs, HBM matters).
template <int GAE_TT>
__global__ void __launch_bounds__(256) gae_kernel(const GaeP p) {
  extern __shared__ __align__(16) float gae_smem[];
  float (*s_r)[GAE_ENVS] = reinterpret_cast<float (*)[GAE_ENVS]>(gae_smem);
  float (*s_t)[GAE_ENVS] = s_r + GAE_TT;
  float (*s_v)[GAE_ENVS] = s_t + GAE_TT;
  float (*s_x)[GAE_ENVS] = s_v + GAE_TT;  // next_values in, advantages out
  const long long n0 = (long long)blockIdx.x * GAE_ENVS;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const long long n = n0 + lane;
  const bool valid = n < p.N;
  const bool has_nv = p.nv != nullptr;

  float last = 0.f;                                         // lastgaelam
  float vnext = (valid && !has_nv) ? p.last_v[n] : 0.f;     // value of the state after step t (shortcut mode)

  for (long long t_hi = p.T; t_hi > 0; t_hi -= GAE_TT) {
    const long long t_lo = (t_hi > GAE_TT) ? t_hi - GAE_TT : 0;
    const int nt = (int)(t_hi - t_lo);
    // phase 1: coalesced tile loads (row = time step, 32 consecutive envs).  All of a thread's loads are issued before the first
    // shared-memory store, so the tile costs one DRAM round trip instead of one per row.
    {
      constexpr int ROWS = GAE_TT / 8;  // 8 warps (the launch uses 256 threads)
      float rr[ROWS], tm[ROWS], vv[ROWS], xx[ROWS];
#pragma unroll
      for (int j = 0; j < ROWS; ++j) {
        const int tt = wrp + j * 8;
        const bool ok = valid && tt < nt;
        const long long g = (t_lo + tt) * p.N + n;
        rr[j] = ok ? p.r[g] : 0.f;
        tm[j] = ok ? p.term[g] : 0.f;
        vv[j] = ok ? p.v[g] : 0.f;
        xx[j] = (ok && has_nv) ? p.nv[g] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < ROWS; ++j) {
        const int tt = wrp + j * 8;
        if (tt < nt) {
          s_r[tt][lane] = rr[j];
          s_t[tt][lane] = tm[j];
          s_v[tt][lane] = vv[j];
          if (has_nv) s_x[tt][lane] = xx[j];
        }
      }
    }
    __syncthreads();
    // phase 2a, all warps: everything that does not depend on the running lastgaelam.  s_r <- delta, s_t <- gamma*lambda*(1-term).
    // (same fp32 operations, in the same order, as the reference's expression: only their scheduling changes)
    {
      const float v_after_tile = vnext;  // shortcut mode: value of the state after the tile's last step
      for (int tt = wrp; tt < nt; tt += nw) {
        const float r = s_r[tt][lane], tm = s_t[tt][lane], v = s_v[tt][lane];
        const float nv = has_nv ? s_x[tt][lane] : (tt + 1 < nt ? s_v[tt + 1][lane] : v_after_tile);
        const float nonterm = __fsub_rn(1.f, tm);
        // delta = rewards + gamma * next_values * (1 - terminations) - values
        const float gnv = p.bf16 ? bf16r(__fmul_rn(p.gamma_f, nv)) : __fmul_rn(p.gamma_f, nv);
        const float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(gnv, nonterm)), v);
        s_r[tt][lane] = delta;
        s_t[tt][lane] = __fmul_rn(p.gl_f, nonterm);
      }
      vnext = s_v[0][lane];  // the tile below (earlier steps) continues from this tile's first value
    }
    __syncthreads();
    // phase 2b: the sequential recurrence, one warp, lane = env: lastgaelam = delta + (gamma*lambda*(1-term)) * lastgaelam
    if (wrp == 0) {
      int tt = nt - 1;
      for (; tt >= 7; tt -= 8) {
        float d[8], k[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { d[j] = s_r[tt - j][lane]; k[j] = s_t[tt - j][lane]; }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          last = __fadd_rn(d[j], __fmul_rn(k[j], last));
          s_x[tt - j][lane] = last;
        }
      }
      for (; tt >= 0; --tt) {
        last = __fadd_rn(s_r[tt][lane], __fmul_rn(s_t[tt][lane], last));
        s_x[tt][lane] = last;
      }
    }
    __syncthreads();
    // phase 3: coalesced stores; returns = advantages + values
    for (int tt = wrp; tt < nt; tt += nw) {
      if (valid) {
        const long long g = (t_lo + tt) * p.N + n;
        const float a = s_x[tt][lane];
        p.adv[g] = a;
        p.ret[g] = __fadd_rn(a, s_v[tt][lane]);
      }
    }
    __syncthreads();
  }
}

}  // namespace rlx

extern "C" int rlx_gae_f32(const float* rewards, const float* terminations, const float* values, const float* next_values,
                           const float* last_value, int64_t T, int64_t N, double gamma, double gae_lambda, float* advantages,
                           float* returns, void* stream) {
  using namespace rlx;
  RLX_CHECK_ARG(T >= 0 && N >= 0, "T, N must be non-negative");
  if (T == 0 || N == 0) return RLX_OK;
  RLX_CHECK_ARG(rewards && terminations && values && advantages && returns, "null tensor");
  RLX_CHECK_ARG(next_values || last_value, "either next_values or last_value is required");
  GaeP p{rewards, terminations, values, next_values, last_value, advantages, returns, T, N, (float)gamma,
         (float)(gamma * gae_lambda), g_autocast_bf16};
  const unsigned grid = (unsigned)ceil_div(N, GAE_ENVS);
  const double bytes = (next_values ? 24.0 : 20.0) * (double)T * (double)N;
  static bool attr_set = false;
  if (!attr_set) {
    RLX_CHECK_CUDA(cudaFuncSetAttribute(gae_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 128 * GAE_ENVS * (int)sizeof(float)));
    RLX_CHECK_CUDA(cudaFuncSetAttribute(gae_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 64 * GAE_ENVS * (int)sizeof(float)));
    attr_set = true;
  }
  if ((long long)grid > 2LL * sm_count() && T > 64) {
    RLX_LAUNCH_C(KC_GAE, 0, bytes, gae_kernel<64>, grid, 256, 4ull * 64 * GAE_ENVS * sizeof(float), stream, p);
  } else {
    RLX_LAUNCH_C(KC_GAE, 0, bytes, gae_kernel<128>, grid, 256, 4ull * 128 * GAE_ENVS * sizeof(float), stream, p);
  }
  return RLX_OK;
}
