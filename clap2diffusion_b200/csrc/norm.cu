// GroupNorm(+SiLU) and LayerNorm over channels-last activations.  Bandwidth-bound: 128-bit loads,
// warp-shuffle / smem reductions, fp32 partials promoted to double for the cross-CTA combine so the
// fp32 parity mode (rel-L2 <= 1e-4 over 50 steps) is not limited by E[x^2]-E[x]^2 cancellation.
#include "common.cuh"

namespace c2d {

constexpr int GN_THREADS = 256;
constexpr int GN_MAX_ITERS = 2;        // C <= 8 * 256 * 2 = 4096
constexpr int GN_MAX_GROUPS = 64;
constexpr int GN_UNROLL = 4;

template <typename T>
__device__ __forceinline__ void gn_load(const T* __restrict__ x, const T* __restrict__ x2, long long row, int c,
                                        int C1, int C2, float (&f)[8]) {
  if (c < C1) Vec8<T>::load(x + row * C1 + c, f);
  else Vec8<T>::load(x2 + row * C2 + (c - C1), f);
}

// SiLU: accurate in the fp32 parity mode; MUFU ex2 + rcp in bf16 mode (the result is rounded to bf16 anyway)
template <typename T> __device__ __forceinline__ float gn_silu(float v);
template <> __device__ __forceinline__ float gn_silu<float>(float v) { return silu_acc(v); }
// x * sigmoid(x) = h + h * tanh(h), h = x / 2: one MUFU.TANH (rel. error ~2^-11, below the bf16 rounding of the result)
template <> __device__ __forceinline__ float gn_silu<bf16>(float v) {
  const float h = 0.5f * v;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// grid: (slabs, B).  Each CTA reduces `rows_per_cta` rows of one image into stats[b][g][{sum,sumsq}].
template <typename T>
__global__ void __launch_bounds__(GN_THREADS, 3)
gn_stats_kernel(const T* __restrict__ x, const T* __restrict__ x2, double* __restrict__ stats, int HW, int C1, int C2,
                int groups, int rows_per_cta) {
  const int C = C1 + C2, nvec = C >> 3, cpg = C / groups;
  const int nvec_eff = nvec < GN_THREADS ? nvec : GN_THREADS;
  const int rows_in_flight = GN_THREADS / nvec_eff;
  const int my_vec = threadIdx.x % nvec_eff, my_rl = threadIdx.x / nvec_eff;
  const int b = blockIdx.y;
  float acc[GN_MAX_ITERS][8], acq[GN_MAX_ITERS][8];
#pragma unroll
  for (int it = 0; it < GN_MAX_ITERS; ++it)
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[it][j] = 0.f; acq[it][j] = 0.f; }
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(HW, r0 + rows_per_cta);
  if (my_rl < rows_in_flight) {
    const bool has2 = my_vec + nvec_eff < nvec;            // second channel vector (only when C > 2048)
    int r = r0 + my_rl;
    // GN_UNROLL independent 128-bit loads in flight per thread, issued unconditionally (memory-level parallelism)
    for (; r + (GN_UNROLL - 1) * rows_in_flight < r1; r += GN_UNROLL * rows_in_flight) {
      float f[GN_UNROLL][8];
#pragma unroll
      for (int u = 0; u < GN_UNROLL; ++u) gn_load<T>(x, x2, (long long)b * HW + r + u * rows_in_flight, my_vec * 8, C1, C2, f[u]);
#pragma unroll
      for (int u = 0; u < GN_UNROLL; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[0][j] += f[u][j]; acq[0][j] += f[u][j] * f[u][j]; }
      if (has2) {
#pragma unroll
        for (int u = 0; u < GN_UNROLL; ++u) gn_load<T>(x, x2, (long long)b * HW + r + u * rows_in_flight, (my_vec + nvec_eff) * 8, C1, C2, f[u]);
#pragma unroll
        for (int u = 0; u < GN_UNROLL; ++u)
#pragma unroll
          for (int j = 0; j < 8; ++j) { acc[1][j] += f[u][j]; acq[1][j] += f[u][j] * f[u][j]; }
      }
    }
    for (; r < r1; r += rows_in_flight) {
      float f[8];
      gn_load<T>(x, x2, (long long)b * HW + r, my_vec * 8, C1, C2, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[0][j] += f[j]; acq[0][j] += f[j] * f[j]; }
      if (has2) {
        gn_load<T>(x, x2, (long long)b * HW + r, (my_vec + nvec_eff) * 8, C1, C2, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[1][j] += f[j]; acq[1][j] += f[j] * f[j]; }
      }
    }
  }
  // per-(row lane, channel) fp32 partials -> smem, then one thread per group folds them in double
  // (rows_in_flight * C == 8 * GN_THREADS floats per array whenever C <= 8 * GN_THREADS)
  extern __shared__ float s_part[];
  float* s_ps = s_part;
  float* s_pq = s_part + rows_in_flight * C;
  if (my_rl < rows_in_flight) {
#pragma unroll
    for (int it = 0; it < GN_MAX_ITERS; ++it) {
      const int vec = my_vec + it * nvec_eff;
      if (vec < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s_ps[my_rl * C + vec * 8 + j] = acc[it][j];
          s_pq[my_rl * C + vec * 8 + j] = acq[it][j];
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < groups) {
    double a = 0.0, q = 0.0;
    for (int rl = 0; rl < rows_in_flight; ++rl)
      for (int c = threadIdx.x * cpg; c < (threadIdx.x + 1) * cpg; ++c) {
        a += (double)s_ps[rl * C + c];
        q += (double)s_pq[rl * C + c];
      }
    atomicAdd(&stats[((long long)b * groups + threadIdx.x) * 2 + 0], a);
    atomicAdd(&stats[((long long)b * groups + threadIdx.x) * 2 + 1], q);
  }
}

template <typename T>
__global__ void __launch_bounds__(GN_THREADS, 3)
gn_apply_kernel(const T* __restrict__ x, const T* __restrict__ x2, const float* __restrict__ gamma,
                const float* __restrict__ beta, const double* __restrict__ stats, T* __restrict__ y,
                T* __restrict__ raw, int HW, int C1, int C2, int groups, float eps, int silu, int rows_per_cta) {
  extern __shared__ float sm[];      // scale[C], shift[C]
  const int C = C1 + C2, nvec = C >> 3, cpg = C / groups;
  float* s_a = sm;
  float* s_b = sm + C;
  const int b = blockIdx.y;
  __shared__ float s_mean[GN_MAX_GROUPS], s_rstd[GN_MAX_GROUPS];
  if (threadIdx.x < groups) {
    const double inv_n = 1.0 / ((double)HW * (double)cpg);
    double mean = stats[((long long)b * groups + threadIdx.x) * 2 + 0] * inv_n;
    double var = stats[((long long)b * groups + threadIdx.x) * 2 + 1] * inv_n - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[threadIdx.x] = (float)mean;
    s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += GN_THREADS) {
    const int g = c / cpg;
    const float a = s_rstd[g] * gamma[c];
    s_a[c] = a;
    s_b[c] = beta[c] - s_mean[g] * a;
  }
  __syncthreads();
  const int nvec_eff = nvec < GN_THREADS ? nvec : GN_THREADS;
  const int rows_in_flight = GN_THREADS / nvec_eff;
  const int my_vec = threadIdx.x % nvec_eff, my_rl = threadIdx.x / nvec_eff;
  if (my_rl >= rows_in_flight) return;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(HW, r0 + rows_per_cta);
  for (int vec = my_vec; vec < nvec; vec += nvec_eff) {
    const int c = vec * 8;
    float a[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { a[j] = s_a[c + j]; sh[j] = s_b[c + j]; }
    int r = r0 + my_rl;
    for (; r + (GN_UNROLL - 1) * rows_in_flight < r1; r += GN_UNROLL * rows_in_flight) {
      float f[GN_UNROLL][8];
#pragma unroll
      for (int u = 0; u < GN_UNROLL; ++u) gn_load<T>(x, x2, (long long)b * HW + r + u * rows_in_flight, c, C1, C2, f[u]);
#pragma unroll
      for (int u = 0; u < GN_UNROLL; ++u) {
        const long long row = (long long)b * HW + r + u * rows_in_flight;
        if (raw) Vec8<T>::store(raw + row * C + c, f[u]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v = f[u][j] * a[j] + sh[j];
          f[u][j] = silu ? gn_silu<T>(v) : v;
        }
        Vec8<T>::store(y + row * C + c, f[u]);
      }
    }
    for (; r < r1; r += rows_in_flight) {
      const long long row = (long long)b * HW + r;
      float f[8];
      gn_load<T>(x, x2, row, c, C1, C2, f);
      if (raw) Vec8<T>::store(raw + row * C + c, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = f[j] * a[j] + sh[j];
        f[j] = silu ? gn_silu<T>(v) : v;
      }
      Vec8<T>::store(y + row * C + c, f);
    }
  }
}

// =====================================================================================================
// GroupNorm from producer-side channel statistics (bf16 product path).
//   chan stats: long long [B][C][2] = 2^20 fixed-point (sum, sum of squares) per channel, accumulated with integer
//   atomics by the producing GEMM / conv epilogue (gemm_tc.cu: stats_commit) or by chan_stats_kernel below --
//   integer adds commute, so the statistics (and therefore the whole denoising trajectory) are bit-reproducible.
//   The apply kernel is ONE pass: read x (and the skip tensor x2), write y.  Algorithmic traffic 4 B / element.
// =====================================================================================================
constexpr int GN2_MAX_THREADS = 512;
constexpr int GN2_UNROLL = 4;
constexpr double STATS_INV_SCALE = 1.0 / 1048576.0;

template <typename T>
__global__ void __launch_bounds__(GN2_MAX_THREADS, 2)
gn_apply2_kernel(const T* __restrict__ x, const T* __restrict__ x2, const long long* __restrict__ st1,
                 const long long* __restrict__ st2, const float* __restrict__ gamma, const float* __restrict__ beta,
                 T* __restrict__ y, int HW, int C1, int C2, int groups, float eps, int silu, int rows_per_cta) {
  pdl_wait();            // x and its statistics come from the kernel(s) just before this one
  pdl_trigger();
  const int C = C1 + C2, nvec = C >> 3, cpg = C / groups;
  const int b = blockIdx.y;
  const int rif = blockDim.x / nvec;                    // rows in flight (blockDim is a multiple of nvec)
  const int my_vec = threadIdx.x % nvec, my_rl = threadIdx.x / nvec;
  const int c0 = my_vec * 8;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(HW, r0 + rows_per_cta);
  const T* src;
  long long rowstride;
  if (c0 < C1) { src = x + (long long)b * HW * C1 + c0; rowstride = C1; }
  else { src = x2 + (long long)b * HW * C2 + (c0 - C1); rowstride = C2; }
  T* dst = y + (long long)b * HW * C + c0;

  // first batch of loads goes out before the statistics are touched
  int r = r0 + my_rl;
  uint4 pre[GN2_UNROLL];
  static_assert(sizeof(T) == 2, "gn_apply2 is the bf16 path");
#pragma unroll
  for (int u = 0; u < GN2_UNROLL; ++u) {
    const int rr = r + u * rif;
    pre[u] = make_uint4(0u, 0u, 0u, 0u);
    if (rr < r1) pre[u] = *reinterpret_cast<const uint4*>(src + (long long)rr * rowstride);
  }

  __shared__ float s_mean[GN_MAX_GROUPS], s_rstd[GN_MAX_GROUPS];
  {
    // one warp per group (round robin): lanes stride over the group's channels, fixed-order butterfly in double
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int g = warp; g < groups; g += nwarps) {
      long long a = 0, q = 0;
      for (int c = g * cpg + lane; c < (g + 1) * cpg; c += 32) {
        const long long* sp = c < C1 ? st1 + ((long long)b * C1 + c) * 2 : st2 + ((long long)b * C2 + (c - C1)) * 2;
        a += sp[0];
        q += sp[1];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      if (lane == 0) {
        const double inv_n = 1.0 / ((double)HW * (double)cpg);
        const double mean = (double)a * STATS_INV_SCALE * inv_n;
        double var = (double)q * STATS_INV_SCALE * inv_n - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[g] = (float)mean;
        s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
      }
    }
  }
  __syncthreads();
  float a[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j, g = c / cpg;
    a[j] = s_rstd[g] * __ldg(gamma + c);
    sh[j] = __ldg(beta + c) - s_mean[g] * a[j];
  }
  auto emit = [&](const uint4& raw, int rr) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
    float f[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = fmaf(f[j], a[j], sh[j]);
      f[j] = silu ? gn_silu<T>(v) : v;
    }
    Vec8<T>::store(dst + (long long)rr * C, f);
  };
  for (;;) {
#pragma unroll
    for (int u = 0; u < GN2_UNROLL; ++u) {
      const int rr = r + u * rif;
      if (rr < r1) emit(pre[u], rr);
    }
    r += GN2_UNROLL * rif;
    if (r >= r1) break;
#pragma unroll
    for (int u = 0; u < GN2_UNROLL; ++u) {
      const int rr = r + u * rif;
      if (rr < r1) pre[u] = *reinterpret_cast<const uint4*>(src + (long long)rr * rowstride);
    }
  }
}

// Stand-alone producer of the same statistics for tensors that do not come out of a tcgen05 epilogue (conv_in).
template <typename T>
__global__ void __launch_bounds__(GN2_MAX_THREADS, 2)
chan_stats_kernel(const T* __restrict__ x, unsigned long long* __restrict__ stats, int HW, int C, int rows_per_cta) {
  const int nvec = C >> 3;
  const int b = blockIdx.y;
  const int rif = blockDim.x / nvec;
  const int my_vec = threadIdx.x % nvec, my_rl = threadIdx.x / nvec;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(HW, r0 + rows_per_cta);
  const T* src = x + (long long)b * HW * C + my_vec * 8;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
  for (int r = r0 + my_rl; r < r1; r += GN2_UNROLL * rif) {
    float f[GN2_UNROLL][8];
#pragma unroll
    for (int u = 0; u < GN2_UNROLL; ++u) {
      const int rr = r + u * rif;
      if (rr < r1) Vec8<T>::load(src + (long long)rr * C, f[u]);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[u][j] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < GN2_UNROLL; ++u)
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += f[u][j]; q[j] = fmaf(f[u][j], f[u][j], q[j]); }
  }
  // fixed-order fold over the row lanes through smem, then one integer atomic per (channel, moment)
  extern __shared__ float s_part[];            // [rif][C][2]
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s_part[((size_t)my_rl * C + my_vec * 8 + j) * 2 + 0] = s[j];
    s_part[((size_t)my_rl * C + my_vec * 8 + j) * 2 + 1] = q[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float t = 0.f;
    for (int rl = 0; rl < rif; ++rl) t += s_part[(size_t)rl * C * 2 + i];
    atomicAdd(stats + (size_t)b * C * 2 + i, (unsigned long long)__float2ll_rn(t * 1048576.0f));
  }
}

static int gn2_threads(int C) {
  const int nvec = C / 8;
  return (GN2_MAX_THREADS / nvec) * nvec;
}
static void gn2_grid(int B, int HW, int rif, dim3* grid, int* rows_per_cta) {
  int slabs = (2 * num_sms()) / B;               // at most one full wave of 2 CTAs per SM
  if (slabs < 1) slabs = 1;
  int rpc = ceil_div(HW, slabs);
  rpc = ceil_div(rpc, rif) * rif;               // whole row batches per CTA
  if (rpc < rif) rpc = rif;
  *rows_per_cta = rpc;
  *grid = dim3(ceil_div(HW, rpc), B);
}

// ---- GroupNorm(+SiLU) backward, bf16 fast path (training step): three coalesced passes with the channel-vector
// mapping of gn_apply2 instead of one CTA per (sample, group) walking a strided slab.
//   pass 1  chan_stats_kernel: per-channel (sum x, sum x^2) -> group mean / rstd
//   pass 2  gn_bwd_reduce_kernel: g = dy * act'(z) * gamma; per-channel (sum g, sum g * xhat) -> double atomics
//   pass 3  gn_bwd_apply_kernel: dx = rstd * (g - mean_group(g) - xhat * mean_group(g * xhat)) (+ add)
// Group statistics of both kinds are rebuilt per CTA from the per-channel arrays (one warp per group).
__device__ __forceinline__ void gnb_group_stats(const long long* __restrict__ st, const double* __restrict__ gst, int b, int C,
                                                int groups, int HW, float eps, float* s_mean, float* s_rstd, float* s_m1,
                                                float* s_m2) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5, cpg = C / groups;
  for (int g = warp; g < groups; g += nwarps) {
    long long a = 0, q = 0;
    double g1 = 0.0, g2 = 0.0;
    for (int c = g * cpg + lane; c < (g + 1) * cpg; c += 32) {
      const long long* sp = st + ((long long)b * C + c) * 2;
      a += sp[0]; q += sp[1];
      if (gst) { g1 += gst[((long long)b * C + c) * 2]; g2 += gst[((long long)b * C + c) * 2 + 1]; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
      g1 += __shfl_xor_sync(0xffffffffu, g1, o);
      g2 += __shfl_xor_sync(0xffffffffu, g2, o);
    }
    if (lane == 0) {
      const double inv_n = 1.0 / ((double)HW * (double)cpg);
      const double mean = (double)a * STATS_INV_SCALE * inv_n;
      double var = (double)q * STATS_INV_SCALE * inv_n - mean * mean;
      if (var < 0.0) var = 0.0;
      s_mean[g] = (float)mean;
      s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
      if (gst) { s_m1[g] = (float)(g1 * inv_n); s_m2[g] = (float)(g2 * inv_n); }
    }
  }
}

template <bool APPLY>
__global__ void __launch_bounds__(GN2_MAX_THREADS, 2)
gn_bwd_pass_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, const long long* __restrict__ st, double* __restrict__ gst,
                   const float* __restrict__ gamma, const float* __restrict__ beta, const bf16* __restrict__ add, bf16* __restrict__ dx,
                   int HW, int C, int groups, float eps, int silu, int rows_per_cta) {
  const int nvec = C >> 3, cpg = C / groups;
  const int b = blockIdx.y;
  const int rif = blockDim.x / nvec;
  const int my_vec = threadIdx.x % nvec, my_rl = threadIdx.x / nvec;
  const int c0 = my_vec * 8;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(HW, r0 + rows_per_cta);
  __shared__ float s_mean[GN_MAX_GROUPS], s_rstd[GN_MAX_GROUPS], s_m1[GN_MAX_GROUPS], s_m2[GN_MAX_GROUPS];
  gnb_group_stats(st, APPLY ? gst : nullptr, b, C, groups, HW, eps, s_mean, s_rstd, s_m1, s_m2);
  __syncthreads();
  float mean[8], rstd[8], ga[8], be[8], m1[8], m2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j, g = c / cpg;
    mean[j] = s_mean[g]; rstd[j] = s_rstd[g];
    ga[j] = __ldg(gamma + c); be[j] = __ldg(beta + c);
    m1[j] = APPLY ? s_m1[g] : 0.f; m2[j] = APPLY ? s_m2[g] : 0.f;
  }
  const long long base = (long long)b * HW * C + c0;
  float a1[8], a2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a1[j] = 0.f; a2[j] = 0.f; }
  for (int r = r0 + my_rl; r < r1; r += rif) {
    const long long o = base + (long long)r * C;
    float fx[8], fd[8];
    Vec8<bf16>::load(x + o, fx);
    Vec8<bf16>::load(dy + o, fd);
    float out[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (fx[j] - mean[j]) * rstd[j];
      float d = fd[j];
      if (silu) {
        const float z = fmaf(ga[j], xh, be[j]);
        const float sg = 1.f / (1.f + __expf(-z));
        d *= sg * (1.f + z * (1.f - sg));
      }
      const float gg = d * ga[j];
      if (APPLY) out[j] = rstd[j] * (gg - m1[j] - xh * m2[j]);
      else { a1[j] += gg; a2[j] = fmaf(gg, xh, a2[j]); }
    }
    if (APPLY) {
      if (add) {
        float fa[8];
        Vec8<bf16>::load(add + o, fa);
#pragma unroll
        for (int j = 0; j < 8; ++j) out[j] += fa[j];
      }
      Vec8<bf16>::store(dx + o, out);
    }
  }
  if (!APPLY) {
    extern __shared__ float s_part[];            // [rif][C][2]
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s_part[((size_t)my_rl * C + c0 + j) * 2 + 0] = a1[j];
      s_part[((size_t)my_rl * C + c0 + j) * 2 + 1] = a2[j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
      float t = 0.f;
      for (int rl = 0; rl < rif; ++rl) t += s_part[(size_t)rl * C * 2 + i];
      atomicAdd(gst + (size_t)b * C * 2 + i, (double)t);
    }
  }
}

// ws: caller-zeroed, B * C * 2 int64 (channel statistics of x) followed by B * C * 2 doubles (adjoint sums); x_stats
// (optional): the channel statistics the forward pass already produced (c2d_channel_stats / a GEMM epilogue) -- pass 1 is
// then skipped
int group_norm_bwd_fast(const void* x, const void* dy, const float* gamma, const float* beta, const void* add, void* dx, void* ws,
                        const long long* x_stats, int B, int HW, int C, int groups, float eps, int silu, cudaStream_t s) {
  const int threads = gn2_threads(C), rif = threads / (C / 8);
  dim3 grid;
  int rpc;
  gn2_grid(B, HW, rif, &grid, &rpc);
  const size_t smem = (size_t)rif * C * 2 * sizeof(float);
  if (smem > 48 * 1024) return -1;
  long long* st_ws = reinterpret_cast<long long*>(ws);
  double* gst = reinterpret_cast<double*>(st_ws + (size_t)B * C * 2);
  const long long* st = x_stats ? x_stats : st_ws;
  if (!x_stats) {
    chan_stats_kernel<bf16><<<grid, threads, smem, s>>>((const bf16*)x, reinterpret_cast<unsigned long long*>(st_ws), HW, C, rpc);
    if (int rc = check_launch("group_norm_bwd")) return rc;
  }
  gn_bwd_pass_kernel<false><<<grid, threads, smem, s>>>((const bf16*)x, (const bf16*)dy, st, gst, gamma, beta, nullptr, nullptr, HW, C,
                                                        groups, eps, silu, rpc);
  if (int rc = check_launch("group_norm_bwd")) return rc;
  gn_bwd_pass_kernel<true><<<grid, threads, 0, s>>>((const bf16*)x, (const bf16*)dy, st, gst, gamma, beta, (const bf16*)add, (bf16*)dx, HW,
                                                    C, groups, eps, silu, rpc);
  return check_launch("group_norm_bwd");
}

// ---- LayerNorm for narrow rows (C <= 8 * SUB, SUB = 8 | 16 lanes per row): 32 / SUB rows side by side in a warp, ROWS
// such passes per lane in flight, sub-warp butterfly reductions.  The CLAP tower's first stage (C = 96: 12 vectors) ran the
// warp-per-row kernel with 20 of 32 lanes idle, at 18 % of the HBM roof.
template <typename T, int SUB, int ROWS>
__global__ void ln_sub_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                              T* __restrict__ y, int M, int C, float eps) {
  constexpr int RPP = 32 / SUB;                       // rows per pass
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int sub = lane / SUB, vec = lane % SUB;
  const int row0 = warp * (RPP * ROWS) + sub;
  const int nvec = C >> 3;
  const bool act = vec < nvec;
  float v[ROWS][8];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int row = row0 + r * RPP;
    if (act && row < M) Vec8<T>::load(x + (long long)row * C + vec * 8, v[r]);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[r][j] = 0.f;
    }
  }
  float g[8], bt[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { g[j] = act ? gamma[vec * 8 + j] : 0.f; bt[j] = act ? beta[vec * 8 + j] : 0.f; }
  const float inv_c = 1.f / (float)C;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[r][j];
#pragma unroll
    for (int o = SUB / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * inv_c;
    float q = 0.f;
    if (act) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[r][j] - mean; q = fmaf(d, d, q); }
    }
#pragma unroll
    for (int o = SUB / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * inv_c + eps);
    const int row = row0 + r * RPP;
    if (act && row < M) {
      float o8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o8[j] = (v[r][j] - mean) * rstd * g[j] + bt[j];
      Vec8<T>::store(y + (long long)row * C + vec * 8, o8);
    }
  }
}

// ---- LayerNorm: one warp handles ROWS rows at once (ROWS x ITERS independent 128-bit loads in flight per lane),
// rows kept in registers, two-pass statistics in registers.  C % 8 == 0 and C <= 256 * ITERS.
template <typename T, int ITERS, int ROWS>
__global__ void ln_vec_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                              T* __restrict__ y, int M, int C, float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int row0 = warp * ROWS;
  if (row0 >= M) return;
  const int nvec = C >> 3;
  float v[ROWS][ITERS][8];
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int vec = lane + it * 32;
      if (row0 + r < M && vec < nvec) Vec8<T>::load(x + (long long)(row0 + r) * C + vec * 8, v[r][it]);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[r][it][j] = 0.f;
      }
    }
  float mean[ROWS], rstd[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    float s = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[r][it][j];
    mean[r] = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int vec = lane + it * 32;
      if (vec < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { float d = v[r][it][j] - mean[r]; q += d * d; }
      }
    }
    rstd[r] = rsqrtf(warp_sum(q) / (float)C + eps);
  }
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int vec = lane + it * 32;
    if (vec < nvec) {
      float g[8], bt[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { g[j] = gamma[vec * 8 + j]; bt[j] = beta[vec * 8 + j]; }
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        if (row0 + r < M) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = (v[r][it][j] - mean[r]) * rstd[r] * g[j] + bt[j];
          Vec8<T>::store(y + (long long)(row0 + r) * C + vec * 8, o);
        }
      }
    }
  }
}

template <typename T>
__global__ void ln_generic_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, T* __restrict__ y, int M, int C, float eps) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const T* xr = x + (long long)warp * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += to_f<T>(xr[c]);
  float mean = warp_sum(s) / (float)C;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) { float d = to_f<T>(xr[c]) - mean; q += d * d; }
  float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  T* yr = y + (long long)warp * C;
  for (int c = lane; c < C; c += 32) yr[c] = from_f<T>((to_f<T>(xr[c]) - mean) * rstd * gamma[c] + beta[c]);
}

}  // namespace c2d

using namespace c2d;

extern "C" {

int c2d_group_norm(const void* x, const void* x2, const float* gamma, const float* beta, void* y, void* raw_cat,
                   double* stats_ws, int B, int HW, int C1, int C2, int groups, float eps, int silu, int dtype,
                   void* stream) {
  C2D_REQUIRE(x && gamma && beta && y && stats_ws, "group_norm: null pointer");
  C2D_REQUIRE(B > 0 && HW > 0 && C1 > 0 && C2 >= 0 && (C2 == 0 || x2), "group_norm: bad dims / missing x2");
  int C = C1 + C2;
  C2D_REQUIRE(C1 % 8 == 0 && C2 % 8 == 0, "group_norm: C1=%d, C2=%d must be multiples of 8", C1, C2);
  C2D_REQUIRE(groups > 0 && groups <= GN_MAX_GROUPS && C % groups == 0, "group_norm: bad groups %d for C=%d", groups, C);
  C2D_REQUIRE(C <= 8 * GN_THREADS * GN_MAX_ITERS, "group_norm: C=%d too large", C);
  C2D_REQUIRE(!raw_cat || C2 > 0, "group_norm: raw_cat only meaningful with a second source");
  cudaStream_t s = (cudaStream_t)stream;
  C2D_CUDA(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * B * groups, s));
  // slabs: aim for ~4 CTAs per SM over the whole batch
  int target = num_sms() * 4;
  int slabs = ceil_div(target, B);
  if (slabs > HW) slabs = HW;
  if (slabs < 1) slabs = 1;
  int rows_per_cta = ceil_div(HW, slabs);
  slabs = ceil_div(HW, rows_per_cta);
  dim3 grid(slabs, B);
  size_t smem = sizeof(float) * 2 * C;
  const int nvec_eff_h = (C / 8) < GN_THREADS ? (C / 8) : GN_THREADS;
  size_t smem_stats = sizeof(float) * 2 * (size_t)(GN_THREADS / nvec_eff_h) * C;
  if (dtype == C2D_F32) {
    gn_stats_kernel<float><<<grid, GN_THREADS, smem_stats, s>>>((const float*)x, (const float*)x2, stats_ws, HW, C1, C2, groups, rows_per_cta);
    int rc = check_launch("gn_stats");
    if (rc) return rc;
    gn_apply_kernel<float><<<grid, GN_THREADS, smem, s>>>((const float*)x, (const float*)x2, gamma, beta, stats_ws,
                                                           (float*)y, (float*)raw_cat, HW, C1, C2, groups, eps, silu, rows_per_cta);
  } else if (dtype == C2D_BF16) {
    gn_stats_kernel<bf16><<<grid, GN_THREADS, smem_stats, s>>>((const bf16*)x, (const bf16*)x2, stats_ws, HW, C1, C2, groups, rows_per_cta);
    int rc = check_launch("gn_stats");
    if (rc) return rc;
    gn_apply_kernel<bf16><<<grid, GN_THREADS, smem, s>>>((const bf16*)x, (const bf16*)x2, gamma, beta, stats_ws,
                                                          (bf16*)y, (bf16*)raw_cat, HW, C1, C2, groups, eps, silu, rows_per_cta);
  } else {
    set_error("group_norm: bad dtype %d", dtype);
    return C2D_ERR_ARG;
  }
  return check_launch("gn_apply");
}

int c2d_channel_stats(const void* x, long long* chan_stats, int B, int HW, int C, int dtype, void* stream) {
  C2D_REQUIRE(x && chan_stats && B > 0 && HW > 0 && C > 0, "channel_stats: bad args");
  C2D_REQUIRE(dtype == C2D_BF16, "channel_stats: bf16 product path only (fp32 mode uses c2d_group_norm)");
  C2D_REQUIRE(C % 8 == 0 && C / 8 <= GN2_MAX_THREADS, "channel_stats: C=%d must be a multiple of 8 and <= %d", C, 8 * GN2_MAX_THREADS);
  const int threads = gn2_threads(C), rif = threads / (C / 8);
  dim3 grid;
  int rpc;
  gn2_grid(B, HW, rif, &grid, &rpc);
  const size_t smem = sizeof(float) * 2 * (size_t)rif * C;
  C2D_REQUIRE(smem <= 48 * 1024, "channel_stats: C=%d needs %zu B of shared memory", C, smem);
  chan_stats_kernel<bf16><<<grid, threads, smem, (cudaStream_t)stream>>>((const bf16*)x, (unsigned long long*)chan_stats, HW, C, rpc);
  return check_launch("chan_stats");
}

int c2d_group_norm_apply(const void* x, const void* x2, const long long* stats1, const long long* stats2, const float* gamma,
                         const float* beta, void* y, int B, int HW, int C1, int C2, int groups, float eps, int silu,
                         int dtype, void* stream) {
  C2D_REQUIRE(x && stats1 && gamma && beta && y, "group_norm_apply: null pointer");
  C2D_REQUIRE(B > 0 && HW > 0 && C1 > 0 && C2 >= 0 && (C2 == 0 || (x2 && stats2)), "group_norm_apply: bad dims / missing second source");
  C2D_REQUIRE(dtype == C2D_BF16, "group_norm_apply: bf16 product path only (fp32 mode uses c2d_group_norm)");
  const int C = C1 + C2;
  C2D_REQUIRE(C1 % 8 == 0 && C2 % 8 == 0 && C / 8 <= GN2_MAX_THREADS, "group_norm_apply: C1=%d C2=%d", C1, C2);
  C2D_REQUIRE(groups > 0 && groups <= GN_MAX_GROUPS && C % groups == 0, "group_norm_apply: bad groups %d for C=%d", groups, C);
  const int threads = gn2_threads(C), rif = threads / (C / 8);
  dim3 grid;
  int rpc;
  gn2_grid(B, HW, rif, &grid, &rpc);
  launch_pdl(gn_apply2_kernel<bf16>, grid, dim3(threads), 0, (cudaStream_t)stream, (const bf16*)x, (const bf16*)x2, stats1, stats2, gamma,
             beta, (bf16*)y, HW, C1, C2, groups, eps, silu, rpc);
  return check_launch("gn_apply");
}

int c2d_layer_norm(const void* x, const float* gamma, const float* beta, void* y, int M, int C, float eps, int dtype,
                   void* stream) {
  C2D_REQUIRE(x && gamma && beta && y && M > 0 && C > 0, "layer_norm: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  const int threads = 128, wpc = threads / 32;
  C2D_REQUIRE(dtype == C2D_F32 || dtype == C2D_BF16, "layer_norm: bad dtype %d", dtype);
#define LN_GO(T, ITERS, ROWS)                                                                              \
  ln_vec_kernel<T, ITERS, ROWS><<<ceil_div(ceil_div(M, ROWS), wpc), threads, 0, s>>>((const T*)x, gamma, beta, (T*)y, M, C, eps)
#define LN_SUB(T, SUB, ROWS)                                                                               \
  ln_sub_kernel<T, SUB, ROWS><<<ceil_div(ceil_div(M, (32 / SUB) * ROWS), wpc), threads, 0, s>>>((const T*)x, gamma, beta, (T*)y, M, C, eps)
#define LN_DISPATCH(T)                                                                                     \
  if (C % 8 == 0 && C <= 64) LN_SUB(T, 8, 4);                                                              \
  else if (C % 8 == 0 && C <= 128) LN_SUB(T, 16, 4);                                                       \
  else if (C % 8 == 0 && C <= 256) LN_GO(T, 1, 8);                                                         \
  else if (C % 8 == 0 && C <= 512) LN_GO(T, 2, 4);                                                         \
  else if (C % 8 == 0 && C <= 768) LN_GO(T, 3, 2);                                                         \
  else if (C % 8 == 0 && C <= 1280) LN_GO(T, 5, 1);                                                        \
  else ln_generic_kernel<T><<<ceil_div(M, wpc), threads, 0, s>>>((const T*)x, gamma, beta, (T*)y, M, C, eps)
  if (dtype == C2D_F32) { LN_DISPATCH(float); }
  else { LN_DISPATCH(bf16); }
#undef LN_DISPATCH
#undef LN_SUB
#undef LN_GO
  return check_launch("layer_norm");
}

}  // extern "C"
