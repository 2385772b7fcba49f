// Error plumbing, init, launch accounting for libc2d.
#include <stdarg.h>
#include <stdlib.h>
#include <atomic>

#include "common.cuh"

namespace c2d {

static thread_local char g_err[512] = "";
static thread_local const char* g_last_kernel = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  g_last_kernel = what;
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return C2D_ERR_CUDA;
  }
  return C2D_OK;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("C2D_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int init_tc(int device);   // gemm_tc.cu

}  // namespace c2d

extern "C" {

int c2d_abi_version(void) { return C2D_ABI_VERSION; }

const char* c2d_last_error(void) { return c2d::g_err; }

unsigned long long c2d_launch_count(void) { return c2d::g_launches.load(); }

const char* c2d_last_kernel(void) { return c2d::g_last_kernel; }

int c2d_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    c2d::set_error("c2d_init: no CUDA device (%s); libc2d has no CPU fallback",
                   e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return C2D_ERR_CUDA;
  }
  C2D_REQUIRE(device >= 0 && device < n, "c2d_init: bad device %d (have %d)", device, n);
  C2D_CUDA(cudaSetDevice(device));
  cudaDeviceProp p;
  C2D_CUDA(cudaGetDeviceProperties(&p, device));
  if (p.major != 10) {
    c2d::set_error("c2d_init: device %d is sm_%d%d; libc2d is built for sm_100a only", device, p.major, p.minor);
    return C2D_ERR_UNSUPPORTED;
  }
  return c2d::init_tc(device);
}

}  // extern "C"
