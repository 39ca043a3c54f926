// Error plumbing and ABI bookkeeping shared by every entry point of libfvb200.
#include "common.cuh"

#include <string.h>

namespace fvb {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    (void)cudaGetLastError();  // clear the non-sticky launch error so the next call starts clean
    return FVB_E_CUDA;
  }
  return FVB_OK;
}

}  // namespace fvb

extern "C" int fvb_abi_version(void) { return FVB_ABI_VERSION; }
extern "C" const char* fvb_last_error(void) { return fvb::g_err; }
extern "C" uint64_t fvb_launch_count(void) { return fvb::g_launches.load(std::memory_order_relaxed); }
