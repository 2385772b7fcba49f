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

// GELU for the bf16 tensor-core epilogues: erf by Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7, far below the bf16
// rounding of the result) -- 2 MUFU + ~12 FMA-pipe instructions instead of erff's ~40 with branches.  The GEGLU
// projection's epilogue evaluates one GELU per output element and is instruction-bound.
__device__ __forceinline__ float erf_fast(float x) {
  const float ax = fabsf(x);
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * ax * -1.4426950408889634f));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  return copysignf(fmaf(-poly * t, e, 1.0f), x);
}
__device__ __forceinline__ float gelu_fast(float x) { return 0.5f * x * (1.0f + erf_fast(x * 0.70710678118654752440f)); }
// (v0, v1) *= gelu(g0, g1): the same A&S 7.1.26 erf as gelu_fast, evaluated for two values at once with the packed
// f32x2 FMA / MUL instructions of sm_100 (FFMA2 / FMUL2): ~10.5 issue slots per element instead of ~17 -- the GEGLU
// epilogue spends ~60 % of its instructions here and is issue-bound at K = 320.  erf = copysign(q e - 1, z) takes the
// magnitude 1 - q e from the negated value, which saves the packed negation.
__device__ __forceinline__ void gelu_mul2(float& v0, float& v1, float g0, float g1) {
  asm("{\n\t"
      ".reg .b64 x, z, az, u, t, w, e, p, k, q, r, h, v;\n\t"
      ".reg .b32 z0, z1, a0, a1, u0, u1, t0, t1, w0, w1, e0, e1, r0, r1;\n\t"
      "mov.b64 x, {%2, %3};\n\t"
      "mov.b64 k, 0x3F3504F33F3504F3;\n\t"        // 0.70710678
      "mul.rn.f32x2 z, x, k;\n\t"
      "mov.b64 {z0, z1}, z;\n\t"
      "and.b32 a0, z0, 0x7FFFFFFF;\n\t"
      "and.b32 a1, z1, 0x7FFFFFFF;\n\t"
      "mov.b64 az, {a0, a1};\n\t"
      "mov.b64 k, 0x3EA7BA053EA7BA05;\n\t"        // 0.3275911
      "mov.b64 u, 0x3F8000003F800000;\n\t"        // 1.0
      "fma.rn.f32x2 u, az, k, u;\n\t"
      "mov.b64 {u0, u1}, u;\n\t"
      "rcp.approx.ftz.f32 t0, u0;\n\t"
      "rcp.approx.ftz.f32 t1, u1;\n\t"
      "mov.b64 t, {t0, t1};\n\t"
      "mul.rn.f32x2 w, az, az;\n\t"
      "mov.b64 k, 0xBFB8AA3BBFB8AA3B;\n\t"        // -log2(e)
      "mul.rn.f32x2 w, w, k;\n\t"
      "mov.b64 {w0, w1}, w;\n\t"
      "ex2.approx.ftz.f32 e0, w0;\n\t"
      "ex2.approx.ftz.f32 e1, w1;\n\t"
      "mov.b64 e, {e0, e1};\n\t"
      "mov.b64 k, 0x3F87DC223F87DC22;\n\t"        // 1.061405429
      "mov.b64 p, 0xBFBA00E3BFBA00E3;\n\t"        // -1.453152027
      "fma.rn.f32x2 p, t, k, p;\n\t"
      "mov.b64 k, 0x3FB5F0E33FB5F0E3;\n\t"        // 1.421413741
      "fma.rn.f32x2 p, p, t, k;\n\t"
      "mov.b64 k, 0xBE91A98EBE91A98E;\n\t"        // -0.284496736
      "fma.rn.f32x2 p, p, t, k;\n\t"
      "mov.b64 k, 0x3E8279063E827906;\n\t"        // 0.254829592
      "fma.rn.f32x2 p, p, t, k;\n\t"
      "mul.rn.f32x2 q, p, t;\n\t"
      "mov.b64 k, 0xBF800000BF800000;\n\t"        // -1.0
      "fma.rn.f32x2 r, q, e, k;\n\t"              // q e - 1 = -(1 - q e)
      "mov.b64 {r0, r1}, r;\n\t"
      "lop3.b32 r0, r0, 0x7FFFFFFF, z0, 0xE2;\n\t" // copysign: (r & 0x7fffffff) | (z & 0x80000000)
      "lop3.b32 r1, r1, 0x7FFFFFFF, z1, 0xE2;\n\t"
      "mov.b64 r, {r0, r1};\n\t"
      "mov.b64 k, 0x3F0000003F000000;\n\t"        // 0.5
      "mul.rn.f32x2 h, x, k;\n\t"
      "fma.rn.f32x2 h, h, r, h;\n\t"              // 0.5 x (1 + erf)
      "mov.b64 v, {%0, %1};\n\t"
      "mul.rn.f32x2 v, v, h;\n\t"
      "mov.b64 {%0, %1}, v;\n\t"
      "}"
      : "+f"(v0), "+f"(v1)
      : "f"(g0), "f"(g1));
}
__device__ __forceinline__ float silu_fast(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
__device__ __forceinline__ float apply_act_fast(float v, int act) {
  if (act == C2D_ACT_GELU) return gelu_fast(v);
  if (act == C2D_ACT_SILU) return silu_fast(v);
  if (act == C2D_ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == C2D_ACT_GELU) return gelu_erf(v);
  if (act == C2D_ACT_SILU) return silu_acc(v);
  if (act == C2D_ACT_RELU) return fmaxf(v, 0.f);
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

// ---- programmatic dependent launch (PDL).  A kernel launched through launch_pdl() may start -- barrier / TMEM set-up,
// tensor-map prefetch, index arithmetic -- while the tail of its predecessor on the stream is still running; it must
// call pdl_wait() before its FIRST global-memory access of any kind (inputs written by the predecessor, and outputs
// whose buffer the predecessor may still be reading), and calls pdl_trigger() right after so that its own successor
// may be scheduled as soon as all of this grid's CTAs are resident.  Without the launch attribute both are no-ops.
// C2D_PDL=0 launches everything fully serialised (A/B runs).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// attention problem descriptor shared by the SIMT and tcgen05 attention kernels (see c2d_attention)
struct AttnParams {
  const void *q, *k, *v;
  void* o;
  int Nq, Nkv, d, heads;
  long long ldq, ldk, ldv, ldo, bsq, bsk, bsv, bso;   // row / batch strides in elements
  float scale;
  const uint8_t* mask;   // [B][Nkv] (1 = keep) or null
  // optional (training forward): per-row log2-domain log-sum-exp of the scaled scores, [B][heads][Nq]; the kernel that
  // produces it sets *lse_written (host) -- kernels without the output leave both untouched
  float* lse;
  int* lse_written;
};

// optional inputs / outputs of the tcgen05 GEMM (see c2d_linear_ex)
struct GemmExtras {
  const void* x2;        // second A source: A = [x | x2] along K (first K1 columns from x)
  int K1, ldx2;
  long long* stats;      // per-channel fixed-point statistics of y: [M / stats_rows][N][2]
  int stats_rows;
  long long* rowstats_out;        // per-row fixed-point (sum, sumsq) of y: [M][2]
  const long long* ln_stats;      // folded LayerNorm: row statistics of x ...
  const float* ln_colsum;         // ... and column sums of the gamma-scaled weight
  float ln_eps;
};

// ---- per-device state.  The library keeps NO process-global mutable compute state: what it remembers (largest
// dynamic-smem size set per kernel instantiation, the caller-provided split-K workspace, the SM count) is per device.
constexpr int C2D_MAX_DEVICES = 16;
int cur_device();                      // runtime.cu: cudaGetDevice clamped to [0, C2D_MAX_DEVICES)
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: `set_bytes` (one static array per
// kernel instantiation) remembers the largest size already granted on each device.
template <typename K>
static inline int ensure_dyn_smem(K kern, int bytes, int (&set_bytes)[C2D_MAX_DEVICES], const char* what) {
  const int dev = cur_device();
  if (bytes <= set_bytes[dev]) return C2D_OK;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("%s: cudaFuncSetAttribute(%d B dynamic smem) failed on device %d: %s", what, bytes, dev, cudaGetErrorString(e));
    return C2D_ERR_CUDA;
  }
  set_bytes[dev] = bytes;
  return C2D_OK;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
int num_sms();                         // runtime.cu: SM count of the CURRENT device (cached per device)

}  // namespace c2d
