// Library-level state: error strings, launch counter, engine selection, device properties.
#include "common.cuh"

namespace rlx {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launch_count{0};
int g_gemm_engine = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;  // B200
  }
  return n;
}

}  // namespace rlx

extern "C" int rlx_version(void) { return 1; }
extern "C" const char* rlx_last_error_string(void) { return rlx::g_err; }
extern "C" uint64_t rlx_launch_count(void) { return rlx::g_launch_count.load(); }
extern "C" void rlx_reset_launch_count(void) { rlx::g_launch_count.store(0); }
extern "C" int rlx_get_gemm_engine(void) { return rlx::g_gemm_engine; }

extern "C" int64_t rlx_ppo_param_count(const rlx_ppo_dims* d) {
  if (d == nullptr || !rlx::dims_ok(*d)) {
    rlx::set_error("rlx_ppo_param_count: unsupported dims");
    return RLX_ERR_INVALID_ARG;
  }
  return rlx::make_layout(*d).total();
}

extern "C" int rlx_ppo_param_layout(const rlx_ppo_dims* d, int64_t* offsets, int32_t* is_critic) {
  RLX_CHECK_ARG(d != nullptr && rlx::dims_ok(*d), "unsupported dims");
  RLX_CHECK_ARG(offsets != nullptr, "offsets is null");
  const rlx::PpoLayout L = rlx::make_layout(*d);
  for (int i = 0; i <= RLX_PPO_NSEG; ++i) offsets[i] = L.off[i];
  if (is_critic)
    for (int i = 0; i < RLX_PPO_NSEG; ++i) is_critic[i] = rlx::seg_is_critic(i) ? 1 : 0;
  return RLX_OK;
}
