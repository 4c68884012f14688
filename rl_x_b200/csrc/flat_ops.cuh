// Generic "flat" kernels shared by the dual-build sources (lstm.cu, fastsac.cu): column sums, split reductions, sum of squares.
// One thread = one (row chunk, column) or one element; no shared memory, warp primitives or barriers (see dual_build.cuh).
#pragma once
#include "dual_build.cuh"

namespace rlx {
namespace flat {

constexpr int kColChunk = 256;  // rows per partial of the column-sum reductions

static __device__ __forceinline__ long long gtid() { return (long long)blockIdx.x * blockDim.x + threadIdx.x; }

// partial column sums of X [R, W] (bias gradients): thread = (row chunk, column)
static __global__ void colsum_partial_kernel(const float* __restrict__ X, int ldx, long long R, int W, float* __restrict__ part) {
  const long long id = gtid();
  const long long nchunk = (R + kColChunk - 1) / kColChunk;
  if (id >= nchunk * W) return;
  const long long ch = id / W;
  const int j = (int)(id % W);
  const long long r1 = ch * kColChunk + kColChunk < R ? ch * kColChunk + kColChunk : R;
  float s = 0.f;
  for (long long r = ch * kColChunk; r < r1; ++r) s += X[r * ldx + j];
  part[id] = s;
}

// out[i] = scale * sum_s part[s * len + i] (+ add)      thread = element
static __global__ void reduce_parts_kernel(const float* __restrict__ part, long long nparts, long long len, float scale, float add,
                                    float* __restrict__ out) {
  const long long i = gtid();
  if (i >= len) return;
  float s = 0.f;
  for (long long k = 0; k < nparts; ++k) s += part[k * len + i];
  out[i] = s * scale + add;
}

// sum of squares per 1024-element chunk; thread = chunk
static __global__ void sumsq_partial_kernel(const float* __restrict__ g, long long n, float* __restrict__ part) {
  const long long c = gtid();
  const long long nchunk = (n + 1023) / 1024;
  if (c >= nchunk) return;
  const long long i1 = c * 1024 + 1024 < n ? c * 1024 + 1024 : n;
  float s = 0.f;
  for (long long i = c * 1024; i < i1; ++i) s += g[i] * g[i];
  part[c] = s;
}
static __global__ void sumsq_final_kernel(const float* __restrict__ part, long long nchunk, float* __restrict__ norm_out, long long* __restrict__ step) {
  if (gtid() != 0) return;
  float s = 0.f;
  for (long long c = 0; c < nchunk; ++c) s += part[c];
  norm_out[0] = sqrtf(s);
  step[0] += 1;
}
// bias gradient: db[o] = sum_r dY[r, o]
static int colsum(const float* X, int ldx, long long R, int W, float* col_ws, float scale, float add, float* out, cudaStream_t st) {
  const long long nchunk = ceil_div(R, kColChunk);
  RLX_FLAT_LAUNCH(colsum_partial_kernel, nchunk * W, st, X, ldx, R, W, col_ws);
  RLX_FLAT_LAUNCH(reduce_parts_kernel, (long long)W, st, col_ws, nchunk, (long long)W, scale, add, out);
  return RLX_OK;
}

}  // namespace flat
}  // namespace rlx
