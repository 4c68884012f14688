// GAE backward scan (ref: calculate_gae_advantages_and_returns, rl_x/algorithms/ppo/pytorch/ppo.py:110-118).
//
// HBM-bound: 20-24 algorithmic bytes per (t, env).  The recurrence is serial in t but independent per env, and the
// loads do not depend on the recurrence, so a CTA owns a strip of 32 envs and
//   phase 1: the [TT x 32] tiles of rewards / terminations / values (/ next_values) are staged into shared memory by TMA
//            (cp.async.bulk.tensor.2d: one box per array, completion on an mbarrier; one elected thread issues them) when the
//            arrays are 16-byte aligned with N % 4 == 0 - otherwise by all 256 threads (coalesced 128-byte rows, every load
//            independent => deep memory-level parallelism),
//   phase 2: one warp (lane = env) runs the bit-exact sequential recurrence out of shared memory,
//   phase 3: all threads stream advantages / returns back out, coalesced.
// Tiles of TT time steps are processed from the end of the rollout backwards; the carry (last advantage, next value)
// stays in the scanning warp's registers.  Arithmetic follows the reference's fp32 evaluation order with explicit
// __fmul_rn/__fadd_rn so that no FMA contraction changes a bit.
#include "gemm_tc_common.cuh"  // mbarrier / TMA wrappers and the tensor-map encoder entry point

namespace rlx {

constexpr int GAE_ENVS = 32;   // envs per CTA (one 128-byte row segment)

struct GaeMaps {
  CUtensorMap r, term, v, nv;  // 2-D fp32 tensors [T, N], box [GAE_TT, 32], no swizzle
};

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
template <int GAE_TT, bool TMA>
__global__ void __launch_bounds__(256) gae_kernel(const GaeP p, const __grid_constant__ GaeMaps maps) {
  extern __shared__ __align__(128) float gae_smem[];
  __shared__ __align__(8) uint64_t tma_bar;
  uint32_t tma_phase = 0;
  if (TMA) {
    if (threadIdx.x == 0) {
      tc::mbar_init(&tma_bar, 1);
      tc::fence_barrier_init();
    }
    __syncthreads();
  }
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
    // phase 1 (TMA): one box per array; rows beyond the rollout / envs beyond N are zero-filled by the hardware, rows of the tile that
    // belong to the tile above (short lowest tile) are loaded and ignored.  The previous tile's generic-proxy accesses to the same
    // shared memory are ordered before these async-proxy writes by the __syncthreads that ended it plus the proxy fence.
    if (TMA) {
      if (threadIdx.x == 0) {
        tc::fence_proxy_async();
        const int c0 = (int)n0, c1 = (int)t_lo;
        tc::mbar_arrive_expect_tx(&tma_bar, (has_nv ? 4u : 3u) * GAE_TT * GAE_ENVS * (uint32_t)sizeof(float));
        tc::tma_load_2d(&maps.r, &tma_bar, s_r, c0, c1);
        tc::tma_load_2d(&maps.term, &tma_bar, s_t, c0, c1);
        tc::tma_load_2d(&maps.v, &tma_bar, s_v, c0, c1);
        if (has_nv) tc::tma_load_2d(&maps.nv, &tma_bar, s_x, c0, c1);
      }
      tc::mbar_wait(&tma_bar, tma_phase);
      tma_phase ^= 1;
    } else
    // phase 1 (fallback): coalesced tile loads (row = time step, 32 consecutive envs).  All of a thread's loads are issued before the
    // first shared-memory store, so the tile costs one DRAM round trip instead of one per row.
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

int g_gae_tma = 1;  // rlx_set_gae_tma: 0 = always stage the tiles with ordinary loads

}  // namespace rlx

extern "C" int rlx_set_gae_tma(int on) {
  rlx::g_gae_tma = on ? 1 : 0;
  return rlx::g_gae_tma;
}

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
    RLX_CHECK_CUDA(cudaFuncSetAttribute(gae_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 128 * GAE_ENVS * (int)sizeof(float)));
    RLX_CHECK_CUDA(cudaFuncSetAttribute(gae_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 64 * GAE_ENVS * (int)sizeof(float)));
    RLX_CHECK_CUDA(cudaFuncSetAttribute(gae_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 128 * GAE_ENVS * (int)sizeof(float)));
    RLX_CHECK_CUDA(cudaFuncSetAttribute(gae_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 64 * GAE_ENVS * (int)sizeof(float)));
    attr_set = true;
  }
  const bool tt64 = (long long)grid > 2LL * sm_count() && T > 64;
  // TMA staging needs 16-byte aligned bases and a 16-byte row pitch (N % 4 == 0), and tensor-map extents below 2^31
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  GaeMaps maps{};
  bool tma = g_gae_tma && (N % 4 == 0) && N < (1LL << 31) && T < (1LL << 31) && al16(rewards) && al16(terminations) && al16(values) &&
             (!next_values || al16(next_values)) && tc::get_encode_fn() != nullptr;
  if (tma) {
    const int tt = tt64 ? 64 : 128;
    auto mk = [&](CUtensorMap* m, const float* base) {
      cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)T};
      cuuint64_t gstride[1] = {(cuuint64_t)N * 4};
      cuuint32_t box[2] = {(cuuint32_t)GAE_ENVS, (cuuint32_t)tt};
      cuuint32_t estr[2] = {1, 1};
      return tc::get_encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    tma = mk(&maps.r, rewards) && mk(&maps.term, terminations) && mk(&maps.v, values) && mk(&maps.nv, next_values ? next_values : values);
  }
  if (tt64) {
    if (tma) RLX_LAUNCH_C(KC_GAE, 0, bytes, (gae_kernel<64, true>), grid, 256, 4ull * 64 * GAE_ENVS * sizeof(float), stream, p, maps);
    else RLX_LAUNCH_C(KC_GAE, 0, bytes, (gae_kernel<64, false>), grid, 256, 4ull * 64 * GAE_ENVS * sizeof(float), stream, p, maps);
  } else {
    if (tma) RLX_LAUNCH_C(KC_GAE, 0, bytes, (gae_kernel<128, true>), grid, 256, 4ull * 128 * GAE_ENVS * sizeof(float), stream, p, maps);
    else RLX_LAUNCH_C(KC_GAE, 0, bytes, (gae_kernel<128, false>), grid, 256, 4ull * 128 * GAE_ENVS * sizeof(float), stream, p, maps);
  }
  return RLX_OK;
}
