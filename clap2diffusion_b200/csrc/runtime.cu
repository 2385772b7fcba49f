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
long long splitk_workspace_bytes();

// ---- per-device context: everything the library remembers between calls -------------------------------------------
struct DevCtx {
  bool inited = false;
  int sms = 0;
  void* ws = nullptr;          // caller-owned split-K workspace (c2d_set_workspace)
  size_t ws_bytes = 0;
};
static DevCtx g_ctx[C2D_MAX_DEVICES];

int cur_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < C2D_MAX_DEVICES) ? dev : 0;
}

int num_sms() {
  DevCtx& c = g_ctx[cur_device()];
  if (!c.sms) {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    c.sms = n > 0 ? n : 148;
  }
  return c.sms;
}

float* splitk_workspace(size_t* bytes) {
  const DevCtx& c = g_ctx[cur_device()];
  *bytes = c.ws_bytes;
  return reinterpret_cast<float*>(c.ws);
}

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
  C2D_REQUIRE(device >= 0 && device < n && device < c2d::C2D_MAX_DEVICES, "c2d_init: bad device %d (have %d)", device, n);
  int major = 0, minor = 0;
  C2D_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  C2D_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  if (major != 10 || minor != 0) {     // sm_100a cubins do not run on sm_103 / sm_120
    c2d::set_error("c2d_init: device %d is sm_%d%d; libc2d is built for sm_100a (B200) only", device, major, minor);
    return C2D_ERR_UNSUPPORTED;
  }
  // the process's current device is left as the caller set it (PyTorch owns it); nothing here needs a context switch
  int rc = c2d::init_tc(device);
  if (rc == C2D_OK) c2d::g_ctx[device].inited = true;
  return rc;
}

long long c2d_splitk_workspace_bytes(void) { return c2d::splitk_workspace_bytes(); }

int c2d_set_workspace(int device, void* workspace, long long bytes) {
  C2D_REQUIRE(device >= 0 && device < c2d::C2D_MAX_DEVICES, "c2d_set_workspace: bad device %d", device);
  C2D_REQUIRE(bytes >= 0 && (workspace != nullptr || bytes == 0), "c2d_set_workspace: null workspace with %lld bytes", bytes);
  C2D_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "c2d_set_workspace: workspace must be 256-byte aligned");
  c2d::g_ctx[device].ws = workspace;
  c2d::g_ctx[device].ws_bytes = workspace ? (size_t)bytes : 0;
  return C2D_OK;
}

int c2d_destroy(int device) {
  C2D_REQUIRE(device >= 0 && device < c2d::C2D_MAX_DEVICES, "c2d_destroy: bad device %d", device);
  // nothing device-side is owned by the library (tensor maps are passed by value per launch, workspaces belong to
  // the caller): destroying the context forgets the caller's workspace and the cached device facts.
  c2d::g_ctx[device] = c2d::DevCtx();
  return C2D_OK;
}

}  // extern "C"
