// Shared helpers for the c2d kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/c2d.h"

namespace c2d {

// ---- error plumbing: no exceptions cross the C ABI; message via c2d_last_error() -------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);   // returns C2D_OK or C2D_ERR_CUDA (peeks cudaGetLastError)

#define C2D_REQUIRE(cond, ...)                         \
  do {                                                 \
    if (!(cond)) {                                     \
      c2d::set_error(__VA_ARGS__);                     \
      return C2D_ERR_ARG;                              \
    }                                                  \
  } while (0)

#define C2D_CUDA(call)                                                         \
  do {                                                                         \
    cudaError_t e_ = (call);                                                   \
    if (e_ != cudaSuccess) {                                                   \
      c2d::set_error("%s failed: %s", #call, cudaGetErrorString(e_));          \
      return C2D_ERR_CUDA;                                                     \
    }                                                                          \
  } while (0)

// ---- dtype helpers ---------------------------------------------------------------------------
typedef __nv_bfloat16 bf16;

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 8-element vector load/store (32 B for float, 16 B for bf16); pointer must be aligned accordingly.
template <typename T> struct Vec8;
template <> struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&o)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&o)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(o[4], o[5], o[6], o[7]);
  }
};
template <> struct Vec8<bf16> {
  static __device__ __forceinline__ void load(const bf16* p, float (&o)[8]) {
    uint4 r = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      o[2 * i] = f.x; o[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store(bf16* p, const float (&o)[8]) {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = r;
  }
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
// accurate SiLU for fp32 parity mode
__device__ __forceinline__ float silu_acc(float x) { return x / (1.0f + expf(-x)); }

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == C2D_ACT_GELU) return gelu_erf(v);
  if (act == C2D_ACT_SILU) return silu_acc(v);
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// attention problem descriptor shared by the SIMT and tcgen05 attention kernels (see c2d_attention)
struct AttnParams {
  const void *q, *k, *v;
  void* o;
  int Nq, Nkv, d, heads;
  long long ldq, ldk, ldv, ldo, bsq, bsk, bsv, bso;   // row / batch strides in elements
  float scale;
  const uint8_t* mask;   // [B][Nkv] (1 = keep) or null
};

// optional inputs / outputs of the tcgen05 GEMM (see c2d_linear_ex)
struct GemmExtras {
  const void* x2;        // second A source: A = [x | x2] along K (first K1 columns from x)
  int K1, ldx2;
  long long* stats;      // per-channel fixed-point statistics of y: [M / stats_rows][N][2]
  int stats_rows;
};

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace c2d
