// tcgen05 (5th-gen tensor core) 3xTF32 GEMM engine — placeholder until the engine lands; the SIMT engine is used.
#include "common.cuh"

namespace rlx {
int tc_supported(const rlx_ppo_dims&) { return 0; }
int tc_mlp_hidden_forward(const rlx_ppo_dims&, const float*, const float*, long long, float*, float*, void*, size_t, cudaStream_t) {
  return RLX_ERR_UNSUPPORTED;
}
size_t tc_workspace_bytes(const rlx_ppo_dims&, long long, bool) { return 0; }
}  // namespace rlx
