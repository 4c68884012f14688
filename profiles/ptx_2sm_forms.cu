// Round-2 preparation: the PTX forms the cta_group::2 version of gemm_tc.cu needs, assembled here with nvcc 12.9 for sm_100a
// (nvcc -gencode arch=compute_100a,code=sm_100a -c profiles/ptx_2sm_forms.cu: UTCATOMSWS.2CTA, UTMALDG.2D.2CTA, UTCMMA...).
// NOT executed yet and not part of the library build (rl_x_b200/build.py compiles rl_x_b200/csrc only).
#include <cuda.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap tmap, uint32_t idesc) {
  __shared__ __align__(1024) uint8_t tile[16384];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t taddr;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[0])), "r"(2));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&taddr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // 2-SM TMA load: both CTAs issue it, the transaction bytes land on CTA 0's barrier (peer bit cleared)
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(tile)),
                 "l"((uint64_t)&tmap), "r"(smem_u32(&bar[0]) & 0xFEFFFFFFu), "r"(0), "r"(0)
                 : "memory");
    // peer -> leader arrive
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(&bar[1]) & 0xFEFFFFFFu) : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[0]) & 0xFEFFFFFFu), "r"(1024) : "memory");
    if (rank == 0) {
      uint64_t da = 0, db = 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(taddr),
          "l"(da), "l"(db), "r"(idesc), "r"(1)
          : "memory");
      // A operand from tensor memory
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(taddr),
          "r"(taddr + 256), "l"(db), "r"(idesc), "r"(1)
          : "memory");
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar[1])),
                   "h"((uint16_t)3)
                   : "memory");
    }
  }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(512) : "memory");
}
