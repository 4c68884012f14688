// Shared pieces of the tcgen05 GEMM engines (gemm_tc.cu: one CTA per tile, gemm_tc2.cu: CTA pairs with cta_group::2): tile constants,
// kernel parameters, PTX wrappers, UMMA descriptors, TMA tensor maps.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gemm_simt.cuh"

namespace rlx {
namespace tc {

constexpr int BM = 128;        // UMMA M (cta_group::1)
constexpr int BK = 32;         // fp32 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 8;      // tf32: 32 bytes of K per instruction
constexpr int NUM_THREADS = 448;
constexpr int SPLIT_WARP0 = 2, EPI_WARP0 = 6, EPI_WARPS = 8;  // two epilogue warp-groups: group g takes column chunks g, g+2, ...

enum TcEpi { TC_EPI_NONE = 0, TC_EPI_BIAS_TANH = 1, TC_EPI_DTANH = 2, TC_EPI_BIAS_RELU = 3, TC_EPI_DRELU = 4, TC_EPI_BIAS = 5 };

struct TcParams {
  int M, N, K;               // per-z output is [M, N]; K = full reduction extent
  int batch, splits, kchunk; // z = batch * splits; kchunk multiple of BK
  int tiles_m, tiles_n;
  // TMA coordinate offsets per batch index (elements)
  int a_mn_off, a_k_off, b_mn_off, b_k_off;
  float* C;
  long long ldc, c_batch_off, c_split_off;
  int n_main;                // columns >= n_main are not stored to C; column == n_main goes to extra_col (bias-gradient trick)
  int transpose_out;         // store C^T: element (m, n) at C[n * ldc + m]; rows >= m_main are not stored, row == m_main goes to extra_col[n]
  int m_main;
  float* extra_col;          // [z][M] or null
  long long extra_batch_off, extra_split_off;
  const float* bias;         // TC_EPI_BIAS_TANH
  long long bias_batch_off;
  const float* aux;          // TC_EPI_DTANH: activation values, same indexing as C
  long long ldaux, aux_batch_off;
  int single;                // bf16-autocast mode: operands are bf16 values = exact TF32 operands -> ONE MMA per k-slice, no lo tiles
  int bf16;                  // round Linear outputs / activations / activation gradients to bf16 in the epilogue (see GemmP::bf16)
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem)),
               "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
// pull a box into L2 ahead of time (no shared memory, no barrier): hides DRAM latency that the 2-3 stage smem ring cannot
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"((uint64_t)map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], TF32 inputs, FP32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor, SWIZZLE_128B (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4   [32,46) stride byte offset >> 4   [46,48) version = 1
//   [61,64) layout type = 2 (SWIZZLE_128B)
//   K-major operands use SWIZZLE_128B (type 2: 16-byte chunks XOR row%8, 8-row groups of 1024 B).  MN-major 32-bit operands
//   must use SWIZZLE_128B_BASE32B (type 1: 32-byte chunks XOR row%4, 4-row groups of 512 B) — the only MN-major layout the
//   tensor core accepts for tf32 (cutlass sm100_common.inl:92); TMA writes it with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 @ [4,6); a/b format TF32 = 2 @ [7,10) / [10,13);
// a_major @ 15, b_major @ 16 (0 = K-major, 1 = MN-major); N >> 3 @ [17,23); M >> 4 @ [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn, bool b_mn, int m = BM) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}


// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D fp32 row-major tensor [rows, cols] with row pitch ld (elements); box = [box_cols (<= 32), box_rows], SWIZZLE_128B.
static inline int make_tmap(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld, int box_cols, int box_rows, bool kmaj) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("tcgen05 engine: cuTensorMapEncodeTiled is unavailable");
    return RLX_ERR_UNSUPPORTED;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  kmaj ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("tcgen05 engine: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box=%dx%d", (int)r, rows, cols, ld, box_cols, box_rows);
    return RLX_ERR_CUDA;
  }
  return RLX_OK;
}

struct TcOperand {
  const float* base;
  long long rows, cols, ld;  // global tensor as allocated: [rows, cols], pitch ld
};


}  // namespace tc
}  // namespace rlx
