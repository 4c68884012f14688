// tcgen05 GEMM engine on CTA PAIRS (cta_group::2): the same 3xTF32 fp32-equivalent GEMM as gemm_tc.cu, with one 256 x BN output tile per
// pair of CTAs on the two SMs of a TPC.
//
// Why: gemm_tc.cu is bound by shared-memory bandwidth, not by the tensor pipe (per 128x128x32 k-block an SM moves 192 KB
// through shared memory in 768 MMA cycles = 250 B/clk against 128 B/clk).  With cta_group::2 each CTA stages its own 128 rows of A and
// only HALF of the B tile; the tensor cores of the pair read the other half from the peer's shared memory.  Per CTA and k-block of a
// 256x256 pair tile: 32 KB TMA writes + 32 KB splitter reads + 32 KB splitter writes + 3 x (16 + 16) KB operand reads = 192 KB in
// 1536 MMA cycles = 125 B/clk.
//
// Structure (cluster of 2 CTAs, 448 threads each, persistent over a static schedule of pair tiles).  Every CTA keeps its OWN TMA ring and
// splitter (unchanged from gemm_tc.cu: local full barriers, local shared memory); only the hand-offs around the MMA cross the pair:
//   split_bar[s]      lives in the LEADER (cluster rank 0); the 4 splitter warps of BOTH CTAs arrive on it (remote arrive for the peer)
//   MMA               one thread of the leader issues tcgen05.mma.cta_group::2.kind::tf32 (M = 256: rows 0-127 accumulate in the leader's
//                     TMEM, rows 128-255 in the peer's, same column addresses)
//   empty_bar[s]      local in each CTA; tcgen05.commit.cta_group::2 ... multicast::cluster releases the ring slot in BOTH CTAs
//   tmem_full_bar[a]  local in each CTA, multicast commit; each CTA's epilogue drains its own 128 TMEM lanes
//   tmem_empty_bar[a] lives in the leader; the epilogue warps of both CTAs arrive on it
// TMEM is allocated with tcgen05.alloc.cta_group::2 by one warp of each CTA, with cluster barriers around allocation and teardown: a CTA
// must not exit while its peer can still arrive on its barriers or the pair's MMAs can still read its shared memory.
#include "gemm_tc_common.cuh"

namespace rlx {
namespace tc {

constexpr int PM = 256;  // UMMA M of a CTA pair

// ------------------------------------------------------------------------------------------------ cluster / 2-CTA PTX wrappers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory object in the CTA with cluster rank `rank`
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
// Remote arrive WITHOUT a cluster-scope release: ptxas turns `.release.cluster` into MEMBAR.ALL.GPU + error barriers in front of the arrive,
// i.e. the arriving thread drains every outstanding global access first (measured: the pair kernel ran 1.5x slower than the single-CTA
// one with it, gpurun_out/r2_bench4_*).  Nothing here needs it: what the arrival publishes is shared memory already fenced towards the
// async proxy (splitter) or tensor-memory reads ordered by tcgen05.fence::before_thread_sync (epilogue) - the form CUTLASS' ClusterBarrier
// uses for the same hand-offs.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Waits of the pair kernel carry a watchdog: a protocol error between the two CTAs must end in a trap (reported as a launch failure),
// never in a hung device.  try_wait suspends the thread for a hardware-defined time slice, so the loop is not a hot spin.
__device__ __forceinline__ unsigned long long global_ns2() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
template <bool CLUSTER>
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  unsigned long long t0 = 0;
  for (uint32_t it = 0;; ++it) {
    uint32_t ok;
    if (false && CLUSTER) {
      // (an explicit cluster-scope acquire makes ptxas invalidate L1 after every successful wait; the default form is what CUTLASS uses)
      asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                   : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } else {
      asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                   : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    }
    if (ok) return;
    if ((it & 0xFFu) == 0xFFu) {
      const unsigned long long now = global_ns2();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) {
        printf("tc_gemm2_kernel: block %d thread %d waited 4 s on barrier %u (parity %u)\n", (int)blockIdx.x, (int)threadIdx.x, addr, parity);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) { mbar_wait_wd<true>(bar, parity); }
__device__ __forceinline__ void mbar_wait_local(uint64_t* bar, uint32_t parity) { mbar_wait_wd<false>(bar, parity); }
// generic-proxy writes to this CTA's shared memory -> visible to async-proxy reads (the tensor cores of both CTAs read it through the async
// proxy); the unqualified `fence.proxy.async` also compiles to MEMBAR.ALL.GPU
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// arrive::one on the barrier at this shared-memory offset in every CTA of `mask` once all MMAs issued so far have retired
__device__ __forceinline__ void umma_commit2(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Per CTA: raw A (128 x 32 fp32) + raw half-B (BN/2 x 32 fp32) + their lo tiles per ring stage.
template <int BN>
struct Cfg2 {
  static constexpr int BH = BN / 2;                             // B rows staged by one CTA
  static constexpr int A_BYTES = BM * BK * 4;                   // 16 KB
  static constexpr int B_BYTES = BH * BK * 4;                   // 16 / 8 KB
  static constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);   // raw + lo: 64 / 48 KB
  static constexpr int STAGES = (BN == 128) ? 4 : 3;
  static constexpr int ACC_STAGES = (BN == 128) ? 2 : 1;   // BN = 192 / 256: one accumulator pair fills (most of) the 512 columns
  static constexpr int TMEM_COLS = 512;
  static constexpr int AUX_BYTES = 1024 + EPI_WARPS * 32 * 33 * 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + AUX_BYTES + 1024;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory of an sm_100 CTA");
  static_assert(ACC_STAGES * 2 * BN <= 512, "TMEM columns");
};

// BF16: the bf16-autocast variant (single-pass MMAs on bf16-valued operands, bf16 roundings in the epilogue), a compile-time switch so that
// the fp32-equivalent kernels carry none of it
template <int BN, bool A_KMAJ, bool B_KMAJ, int EPI, bool BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
    tc_gemm2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const TcParams p) {
  using C_ = Cfg2<BN>;
  constexpr int STAGES = C_::STAGES;
  constexpr int ACC_STAGES = C_::ACC_STAGES;
  constexpr int BH = C_::BH;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* aux = smem + STAGES * C_::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);            // [STAGES]      local: own TMA landed
  uint64_t* split_bar = full_bar + STAGES;                          // [STAGES]      leader's copy is used: lo tiles of BOTH CTAs written
  uint64_t* empty_bar = split_bar + STAGES;                         // [STAGES]      local: the pair's MMAs of this slot retired
  uint64_t* tmem_full_bar = empty_bar + STAGES;                     // [ACC_STAGES]  local
  uint64_t* tmem_empty_bar = tmem_full_bar + ACC_STAGES;            // [ACC_STAGES]  leader's copy is used
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + ACC_STAGES);
  float* stage_smem = reinterpret_cast<float*>(aux + 1024);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&split_bar[s], 8);   // 4 splitter warps x 2 CTAs
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < ACC_STAGES; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 2 * EPI_WARPS);
    }
    fence_barrier_init();
  }
  cluster_sync_all();  // barriers of both CTAs are initialised before anybody arrives remotely
  if (warp == 1) tmem_alloc2(tmem_ptr_smem, C_::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int tiles_per_z = p.tiles_m * p.tiles_n;  // tiles_m counts PAIR tiles (256 rows)
  const int num_tiles = tiles_per_z * p.batch * p.splits;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  auto tile_coords = [&](int tile, int& zb, int& zs, int& m0, int& n0, int& kbeg, int& nkb) {
    const int z = tile / tiles_per_z, r = tile % tiles_per_z;
    zb = z / p.splits;
    zs = z % p.splits;
    m0 = (r / p.tiles_n) * PM + (int)rank * BM;   // this CTA's 128 rows of A / of the accumulator
    n0 = (r % p.tiles_n) * BN;                    // first column of the PAIR tile
    kbeg = zs * p.kchunk;
    const int kend = min(p.K, kbeg + p.kchunk);
    nkb = (kend - kbeg + BK - 1) / BK;
  };

  if (warp == 0) {
    // ===================================================== TMA producer (both CTAs: own A rows, own half of B)
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        int zb, zs, m0, n0, kbeg, nkb;
        tile_coords(tile, zb, zs, m0, n0, kbeg, nkb);
        const int nb0 = n0 + (int)rank * BH;  // this CTA's B rows
        auto issue = [&](int kb, bool prefetch_only, uint8_t* sa, uint8_t* sb, uint64_t* bar) {
          const int k0 = kbeg + kb * BK;
          if (A_KMAJ) {
            const int c0 = p.a_k_off * zb + k0, c1 = p.a_mn_off * zb + m0;
            if (prefetch_only) tma_prefetch_2d(&tmap_a, c0, c1); else tma_load_2d(&tmap_a, bar, sa, c0, c1);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 32; ++j) {
              const int c0 = p.a_mn_off * zb + m0 + 32 * j, c1 = p.a_k_off * zb + k0;
              if (prefetch_only) tma_prefetch_2d(&tmap_a, c0, c1); else tma_load_2d(&tmap_a, bar, sa + j * (BK * 128), c0, c1);
            }
          }
          if (B_KMAJ) {
            const int c0 = p.b_k_off * zb + k0, c1 = p.b_mn_off * zb + nb0;
            if (prefetch_only) tma_prefetch_2d(&tmap_b, c0, c1); else tma_load_2d(&tmap_b, bar, sb, c0, c1);
          } else {
#pragma unroll
            for (int j = 0; j < BH / 32; ++j) {
              const int c0 = p.b_mn_off * zb + nb0 + 32 * j, c1 = p.b_k_off * zb + k0;
              if (prefetch_only) tma_prefetch_2d(&tmap_b, c0, c1); else tma_load_2d(&tmap_b, bar, sb + j * (BK * 128), c0, c1);
            }
          }
        };
        constexpr int PF = STAGES + 3;
        for (int kb = 0; kb < min(PF, nkb); ++kb) issue(kb, true, nullptr, nullptr, nullptr);
        for (int kb = 0; kb < nkb; ++kb) {
          if (kb + PF < nkb) issue(kb + PF, true, nullptr, nullptr, nullptr);
          mbar_wait_local(&empty_bar[s], ph ^ 1);
          uint8_t* st = smem + s * C_::STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[s], C_::A_BYTES + C_::B_BYTES);
          issue(kb, false, st, st + C_::A_BYTES, &full_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (one thread of the LEADER CTA)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN, !A_KMAJ, !B_KMAJ, PM);
      constexpr uint32_t A_LBO = A_KMAJ ? 16u : (uint32_t)(BK * 128), B_LBO = B_KMAJ ? 16u : (uint32_t)(BK * 128);
      constexpr uint32_t A_SBO = A_KMAJ ? 1024u : 512u, B_SBO = B_KMAJ ? 1024u : 512u;
      constexpr uint32_t A_LT = A_KMAJ ? 2u : 1u, B_LT = B_KMAJ ? 2u : 1u;
      constexpr uint32_t A_KSTEP = A_KMAJ ? (UMMA_K * 4) : (UMMA_K * 128), B_KSTEP = B_KMAJ ? (UMMA_K * 4) : (UMMA_K * 128);
      // descriptors hold CTA-relative shared-memory addresses: the same offsets are valid in the peer, whose tiles sit at the same places
      const uint32_t s0 = smem_u32(smem);
      const uint64_t da0 = make_smem_desc(s0, A_LBO, A_SBO, A_LT);
      const uint64_t db0 = make_smem_desc(s0 + C_::A_BYTES, B_LBO, B_SBO, B_LT);
      constexpr uint64_t LO_OFF = (uint64_t)((C_::A_BYTES + C_::B_BYTES) >> 4);
      int s = 0, acc = 0;
      uint32_t ph = 0, acc_ph = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        int zb, zs, m0, n0, kbeg, nkb;
        tile_coords(tile, zb, zs, m0, n0, kbeg, nkb);
        mbar_wait_cluster(&tmem_empty_bar[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_main = tmem_base + (uint32_t)(acc * 2 * BN);
        const uint32_t d_corr = d_main + (uint32_t)BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait_cluster(&split_bar[s], ph);  // both CTAs: TMA data landed and lo tiles written
          tc_fence_after();
          const uint64_t soff = (uint64_t)((uint32_t)(s * C_::STAGE_BYTES) >> 4);
#pragma unroll
          for (int kk = 0; kk < BK / UMMA_K; ++kk) {
            const uint64_t da = da0 + soff + (uint64_t)((kk * A_KSTEP) >> 4);
            const uint64_t db = db0 + soff + (uint64_t)((kk * B_KSTEP) >> 4);
            const uint64_t da_lo = da + LO_OFF;
            const uint64_t db_lo = db + LO_OFF;
            const uint32_t accum = (kb | kk) != 0 ? 1u : 0u;
            if (!BF16) {
              umma2_tf32(d_corr, da_lo, db, idesc, accum);  // lo * hi   } small terms, own accumulator
              umma2_tf32(d_corr, da, db_lo, idesc, 1u);     // hi * lo   }
            }
            umma2_tf32(d_main, da, db, idesc, accum);       // hi * hi  (bf16-valued operands: the whole product)
          }
          umma_commit2(&empty_bar[s], 3);  // ring slot free in both CTAs
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit2(&tmem_full_bar[acc], 3);  // accumulators complete: both epilogues may drain
        if (++acc == ACC_STAGES) { acc = 0; acc_ph ^= 1; }
      }
    }
  } else if (warp >= SPLIT_WARP0 && warp < EPI_WARP0) {
    // ===================================================== splitter (both CTAs, own tiles): lo = x - tf32_trunc(x)
    const int t = threadIdx.x - SPLIT_WARP0 * 32;
    constexpr int NV = (C_::A_BYTES + C_::B_BYTES) / 16;
    int s = 0;
    uint32_t ph = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      int zb, zs, m0, n0, kbeg, nkb;
      tile_coords(tile, zb, zs, m0, n0, kbeg, nkb);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait_local(&full_bar[s], ph);
        const uint4* raw = reinterpret_cast<const uint4*>(smem + s * C_::STAGE_BYTES);
        float4* lo = reinterpret_cast<float4*>(smem + s * C_::STAGE_BYTES + C_::A_BYTES + C_::B_BYTES);
#pragma unroll 8
        for (int i = t; i < (BF16 ? 0 : NV); i += 128) {
          const uint4 v = raw[i];
          float4 o;
          o.x = __uint_as_float(v.x) - __uint_as_float(v.x & 0xFFFFE000u);
          o.y = __uint_as_float(v.y) - __uint_as_float(v.y & 0xFFFFE000u);
          o.z = __uint_as_float(v.z) - __uint_as_float(v.z & 0xFFFFE000u);
          o.w = __uint_as_float(v.w) - __uint_as_float(v.w & 0xFFFFE000u);
          lo[i] = o;
        }
        fence_proxy_async_all();  // generic-proxy writes -> visible to the async-proxy reads of BOTH tensor cores
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(map_to_rank(smem_u32(&split_bar[s]), 0));
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===================================================== epilogue (both CTAs, own 128 TMEM lanes = own 128 rows)
    const int q = warp & 3;
    float* tile = stage_smem + (warp - EPI_WARP0) * (32 * 33);
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int tile_i = cluster_id; tile_i < num_tiles; tile_i += num_clusters) {
      int zb, zs, m0, n0, kbeg, nkb;
      tile_coords(tile_i, zb, zs, m0, n0, kbeg, nkb);
      mbar_wait_local(&tmem_full_bar[acc], acc_ph);
      tc_fence_after();
      const int row0 = m0 + q * 32;
      float* cbase = p.C + p.c_batch_off * zb + p.c_split_off * zs;
      constexpr bool HAS_AUX = (EPI == TC_EPI_DTANH || EPI == TC_EPI_DRELU);
      constexpr bool HAS_BIAS = (EPI == TC_EPI_BIAS_TANH || EPI == TC_EPI_BIAS_RELU || EPI == TC_EPI_BIAS);
      const float* abase = HAS_AUX ? (p.aux + p.aux_batch_off * zb) : nullptr;
      const float* bias = HAS_BIAS ? (p.bias + p.bias_batch_off * zb) : nullptr;
      const int rows_valid = min(32, p.M - row0);
      const int eg = (warp - EPI_WARP0) >> 2;
#pragma unroll 1
      for (int c = eg; c < BN / 32; c += EPI_WARPS / 4) {
        uint32_t r[32], rc[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * BN + c * 32);
        tmem_ld_32x32b_x32(taddr, r);
        tmem_ld_32x32b_x32(taddr + BN, rc);
        tmem_ld_wait();
        const int nb = n0 + c * 32;
        if (p.transpose_out) {
          const int m = row0 + lane;
          if (m < p.M) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = nb + j;
              const float v = BF16 ? __uint_as_float(r[j]) : __uint_as_float(r[j]) + __uint_as_float(rc[j]);
              if (n < p.N) {
                if (m < p.m_main) cbase[(long long)n * p.ldc + m] = v;
                else if (m == p.m_main && p.extra_col != nullptr) p.extra_col[p.extra_batch_off * zb + p.extra_split_off * zs + n] = v;
              }
            }
          }
          continue;
        }
        if (rows_valid > 0 && nb < p.N) {
#pragma unroll
          for (int j = 0; j < 32; ++j)  // main + correction (fp32 RN); single-pass mode never wrote the correction accumulator
            tile[lane * 33 + j] = BF16 ? __uint_as_float(r[j]) : __uint_as_float(r[j]) + __uint_as_float(rc[j]);
          __syncwarp();
          const int n = nb + lane;
          const bool n_ok = n < p.N;
          float bv = 0.f;
          if (HAS_BIAS && n_ok) bv = bias[n];
          float x[32];
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) x[rr] = tile[rr * 33 + lane];
          if (HAS_AUX) {
            float hv[32];
#pragma unroll
            for (int rr = 0; rr < 32; ++rr) hv[rr] = (rr < rows_valid && n_ok) ? abase[(long long)(row0 + rr) * p.ldaux + n] : 0.f;
#pragma unroll
            for (int rr = 0; rr < 32; ++rr)  // bf16 mode: linear-backward output rounded to bf16, then tanh_backward rounded to bf16
              x[rr] = (EPI == TC_EPI_DTANH) ? bf16r_if(bf16r_if(x[rr], BF16) * (1.f - hv[rr] * hv[rr]), BF16) : ((hv[rr] > 0.f) ? x[rr] : 0.f);
          }
          if (HAS_BIAS) {
#pragma unroll
            for (int rr = 0; rr < 32; ++rr) {
              const float z = (EPI == TC_EPI_BIAS_RELU) ? x[rr] + bv : bf16r_if(x[rr] + bv, BF16);  // bf16 mode: Linear output is a bf16 tensor
              x[rr] = (EPI == TC_EPI_BIAS_TANH) ? bf16r_if(tanh_fast(z), BF16) : ((EPI == TC_EPI_BIAS_RELU) ? fmaxf(z, 0.f) : z);
            }
          }
          if (n < p.n_main) {
            float* cp = cbase + (long long)row0 * p.ldc + n;
#pragma unroll
            for (int rr = 0; rr < 32; ++rr)
              if (rr < rows_valid) cp[(long long)rr * p.ldc] = x[rr];
          } else if (n == p.n_main && n_ok && p.extra_col != nullptr) {
            float* ep = p.extra_col + p.extra_batch_off * zb + p.extra_split_off * zs + row0;
#pragma unroll
            for (int rr = 0; rr < 32; ++rr)
              if (rr < rows_valid) ep[rr] = x[rr];
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(map_to_rank(smem_u32(&tmem_empty_bar[acc]), 0));
      if (++acc == ACC_STAGES) { acc = 0; acc_ph ^= 1; }
    }
  }

  // ---- teardown: nobody leaves before the peer is done with this CTA's barriers, shared memory and tensor memory
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, C_::TMEM_COLS);
  }
}

template <int BN, bool A_KMAJ, bool B_KMAJ, int EPI, bool BF16 = false>
static int launch_cfg2(const TcOperand& A, const TcOperand& B, TcParams p, int kclass, cudaStream_t stream) {
  CUtensorMap ta, tb;
  int rc = make_tmap(&ta, A.base, A.rows, A.cols, A.ld, 32, A_KMAJ ? BM : BK, A_KMAJ);
  if (rc) return rc;
  rc = make_tmap(&tb, B.base, B.rows, B.cols, B.ld, 32, B_KMAJ ? BN / 2 : BK, B_KMAJ);
  if (rc) return rc;
  p.tiles_m = (int)ceil_div(p.M, PM);
  p.tiles_n = (int)ceil_div(p.N, BN);
  const long long tiles = (long long)p.tiles_m * p.tiles_n * p.batch * p.splits;
  if (tiles <= 0) return RLX_OK;
  auto kern = tc_gemm2_kernel<BN, A_KMAJ, B_KMAJ, EPI, BF16>;
  static bool attr_done = false;
  if (!attr_done) {
    RLX_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<BN>::SMEM_BYTES));
    attr_done = true;
  }
  const unsigned grid = 2u * (unsigned)std::min<long long>(tiles, sm_count() / 2);
  const double flops = 2.0 * p.M * p.N * (double)p.K * p.batch;
  const double bytes = 4.0 * p.batch * ((double)p.M * p.K + (double)p.N * p.K + (double)p.M * p.N * p.splits);
  RLX_LAUNCH_C(kclass, flops, bytes, kern, grid, NUM_THREADS, Cfg2<BN>::SMEM_BYTES, stream, ta, tb, p);
  return RLX_OK;
}

// Dispatch for gemm_tc.cu::tc_gemm_t: the same (layout, epilogue) combinations as the single-CTA engine.
int tc2_launch(int bn, bool a_kmaj, bool b_kmaj, int epi, const TcOperand& A, const TcOperand& B, const TcParams& p, int kclass, cudaStream_t stream) {
#define RLX_TC2_DISPATCH(BN_)                                                                                                    \
  do {                                                                                                                          \
    if (p.bf16 && a_kmaj && b_kmaj && epi == TC_EPI_BIAS_TANH) return launch_cfg2<BN_, true, true, TC_EPI_BIAS_TANH, true>(A, B, p, kclass, stream); \
    if (p.bf16 && a_kmaj && !b_kmaj && epi == TC_EPI_DTANH) return launch_cfg2<BN_, true, false, TC_EPI_DTANH, true>(A, B, p, kclass, stream);     \
    if (p.bf16 && !a_kmaj && !b_kmaj && epi == TC_EPI_NONE) return launch_cfg2<BN_, false, false, TC_EPI_NONE, true>(A, B, p, kclass, stream);     \
    if (p.bf16) return RLX_ERR_UNSUPPORTED;                                                                                     \
    if (a_kmaj && b_kmaj && epi == TC_EPI_BIAS_TANH) return launch_cfg2<BN_, true, true, TC_EPI_BIAS_TANH>(A, B, p, kclass, stream); \
    if (a_kmaj && b_kmaj && epi == TC_EPI_NONE) return launch_cfg2<BN_, true, true, TC_EPI_NONE>(A, B, p, kclass, stream);       \
    if (a_kmaj && b_kmaj && epi == TC_EPI_BIAS_RELU) return launch_cfg2<BN_, true, true, TC_EPI_BIAS_RELU>(A, B, p, kclass, stream); \
    if (a_kmaj && b_kmaj && epi == TC_EPI_BIAS) return launch_cfg2<BN_, true, true, TC_EPI_BIAS>(A, B, p, kclass, stream);       \
    if (a_kmaj && !b_kmaj && epi == TC_EPI_DRELU) return launch_cfg2<BN_, true, false, TC_EPI_DRELU>(A, B, p, kclass, stream);   \
    if (a_kmaj && !b_kmaj && epi == TC_EPI_DTANH) return launch_cfg2<BN_, true, false, TC_EPI_DTANH>(A, B, p, kclass, stream);   \
    if (a_kmaj && !b_kmaj && epi == TC_EPI_NONE) return launch_cfg2<BN_, true, false, TC_EPI_NONE>(A, B, p, kclass, stream);     \
    if (!a_kmaj && !b_kmaj && epi == TC_EPI_NONE) return launch_cfg2<BN_, false, false, TC_EPI_NONE>(A, B, p, kclass, stream);   \
  } while (0)
  if (bn == 192) {  // N = 377 (obs + 1) = 2 x 192: the layer-1 weight gradient, both operands MN-major
    if (!a_kmaj && !b_kmaj && epi == TC_EPI_NONE && p.bf16) return launch_cfg2<192, false, false, TC_EPI_NONE, true>(A, B, p, kclass, stream);
    if (!a_kmaj && !b_kmaj && epi == TC_EPI_NONE) return launch_cfg2<192, false, false, TC_EPI_NONE>(A, B, p, kclass, stream);
    return RLX_ERR_UNSUPPORTED;
  }
  if (bn == 256) {
    RLX_TC2_DISPATCH(256);
  } else {
    RLX_TC2_DISPATCH(128);
  }
#undef RLX_TC2_DISPATCH
  return RLX_ERR_UNSUPPORTED;
}

}  // namespace tc
}  // namespace rlx
