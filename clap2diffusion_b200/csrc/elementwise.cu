// Bandwidth-bound elementwise / layout kernels: vectorised (128-bit where the shape allows), coalesced,
// grid-stride with grids sized in multiples of the SM count.
#include "common.cuh"

namespace c2d {

static inline int ew_grid(long long work_items, int threads) {
  long long blocks = (work_items + threads - 1) / threads;
  long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ------------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void unary_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long n, int act) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  long long n8 = n >> 3;
  for (long long v = i; v < n8; v += stride) {
    float f[8];
    Vec8<TI>::load(x + v * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = apply_act(f[j], act);
    Vec8<TO>::store(y + v * 8, f);
  }
  for (long long e = n8 * 8 + i; e < n; e += stride) y[e] = from_f<TO>(apply_act(to_f<TI>(x[e]), act));
}

template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  long long n8 = n >> 3;
  for (long long v = i; v < n8; v += stride) {
    float fa[8], fb[8];
    Vec8<T>::load(a + v * 8, fa);
    Vec8<T>::load(b + v * 8, fb);
#pragma unroll
    for (int j = 0; j < 8; ++j) fa[j] += fb[j];
    Vec8<T>::store(y + v * 8, fa);
  }
  for (long long e = n8 * 8 + i; e < n; e += stride) y[e] = from_f<T>(to_f<T>(a[e]) + to_f<T>(b[e]));
}

// y[m, f] = x[m, f] * gelu(x[m, F + f]);  F % 8 == 0
template <typename T>
__global__ void geglu_kernel(const T* __restrict__ x, T* __restrict__ y, long long M, int F) {
  int f8 = F >> 3;
  long long total = M * f8;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += stride) {
    long long m = v / f8;
    int c = (int)(v - m * f8) * 8;
    float a[8], g[8];
    Vec8<T>::load(x + m * 2 * F + c, a);
    Vec8<T>::load(x + m * 2 * F + F + c, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] *= gelu_erf(g[j]);
    Vec8<T>::store(y + m * F + c, a);
  }
}

// nearest 2x upsample, NHWC, C % 8 == 0
template <typename T>
__global__ void upsample2x_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C) {
  int c8 = C >> 3;
  long long total = (long long)B * 2 * H * 2 * W * c8;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += stride) {
    int c = (int)(v % c8) * 8;
    long long p = v / c8;
    int ox = (int)(p % (2 * W));
    long long q = p / (2 * W);
    int oy = (int)(q % (2 * H));
    int b = (int)(q / (2 * H));
    const T* src = x + (((long long)b * H + (oy >> 1)) * W + (ox >> 1)) * C + c;
    float f[8];
    Vec8<T>::load(src, f);
    Vec8<T>::store(y + p * C + c, f);
  }
}

// channel concat, C1 % 8 == 0 and C2 % 8 == 0
template <typename T>
__global__ void concat_kernel(const T* __restrict__ x1, const T* __restrict__ x2, T* __restrict__ y, long long rows,
                              int C1, int C2) {
  int C = C1 + C2, c8 = C >> 3;
  long long total = rows * c8;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += stride) {
    long long r = v / c8;
    int c = (int)(v - r * c8) * 8;
    float f[8];
    if (c < C1) Vec8<T>::load(x1 + r * C1 + c, f);
    else Vec8<T>::load(x2 + r * C2 + (c - C1), f);
    Vec8<T>::store(y + r * C + c, f);
  }
}

// NCHW fp32 [B][C][HW] <-> NHWC T [B][HW][C] through a 32x32 smem tile (coalesced both sides)
template <typename T, bool TO_NHWC>
__global__ void layout_kernel(const void* __restrict__ xin, void* __restrict__ yout, int C, int HW) {
  __shared__ float tile[32][33];
  int b = blockIdx.z;
  int hw0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
  if (TO_NHWC) {
    const float* x = reinterpret_cast<const float*>(xin) + (long long)b * C * HW;
    T* y = reinterpret_cast<T*>(yout) + (long long)b * C * HW;
    for (int j = ty; j < 32; j += 8) {
      int c = c0 + j, hw = hw0 + tx;
      tile[j][tx] = (c < C && hw < HW) ? x[(long long)c * HW + hw] : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
      int hw = hw0 + j, c = c0 + tx;
      if (c < C && hw < HW) y[(long long)hw * C + c] = from_f<T>(tile[tx][j]);
    }
  } else {
    const T* x = reinterpret_cast<const T*>(xin) + (long long)b * C * HW;
    float* y = reinterpret_cast<float*>(yout) + (long long)b * C * HW;
    for (int j = ty; j < 32; j += 8) {
      int hw = hw0 + j, c = c0 + tx;
      tile[j][tx] = (c < C && hw < HW) ? to_f<T>(x[(long long)hw * C + c]) : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
      int c = c0 + j, hw = hw0 + tx;
      if (c < C && hw < HW) y[(long long)c * HW + hw] = tile[tx][j];
    }
  }
}

__global__ void timestep_embedding_kernel(const float* __restrict__ t, float* __restrict__ out, int B, int dim) {
  int half = dim >> 1;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  int b = i / half, j = i - b * half;
  // exp(-ln(10000) * j / half); computed in fp32 like torch, evaluated with accurate expf/sincosf
  float freq = expf(-9.210340371976184f * (float)j / (float)half);
  float ang = t[b] * freq;
  float s, c;
  sincosf(ang, &s, &c);
  out[(long long)b * dim + j] = c;          // flip_sin_to_cos: cos first
  out[(long long)b * dim + half + j] = s;
}

// CFG combine + scheduler update + next-input duplication, one thread per latent element.
template <typename T>
__global__ void cfg_sched_kernel(const T* __restrict__ eps2, float* __restrict__ x, T* __restrict__ xin2,
                                 float* __restrict__ trace, int B, int HW, float g,
                                 const float* __restrict__ coef, int cpitch) {
  const float ca = coef[0], cb = coef[1], in_scale = coef[2];
  long long n = (long long)B * 4 * HW;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // i indexes NCHW fp32 latent: b, c, p
  int p = (int)(i % HW);
  int c = (int)((i / HW) % 4);
  int b = (int)(i / (4LL * HW));
  long long nh = ((long long)b * HW + p) * 4 + c;            // NHWC index inside one CFG half
  long long half = (long long)B * HW * 4;
  float eu = to_f<T>(eps2[nh]);
  float ec = to_f<T>(eps2[half + nh]);
  float e = eu + g * (ec - eu);
  float xn = ca * x[i] + cb * e;
  x[i] = xn;
  if (trace) trace[i] = xn;
  T xi = from_f<T>(xn * in_scale);
  // the next UNet input may carry padded channels (cpitch = 8: conv_in then runs on the tensor cores); only the
  // four real channels are written, the caller zero-initialises the padding once
  const long long no = ((long long)b * HW + p) * cpitch + c;
  xin2[no] = xi;
  xin2[(long long)B * HW * cpitch + no] = xi;
}

// [Cout][Cin][3][3] fp32 -> [Cout][3][3][Cin] T
template <typename T>
__global__ void pack_conv_kernel(const float* __restrict__ w, T* __restrict__ out, int Cout, int Cin) {
  long long n = (long long)Cout * Cin * 9;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int ci = (int)(i % Cin);
    int tap = (int)((i / Cin) % 9);
    int co = (int)(i / (9LL * Cin));
    out[i] = from_f<T>(w[((long long)co * Cin + ci) * 9 + tap]);
  }
}

// interleave rows of [2F][K]: packed row r -> source row:  blk = r / 128, j = r % 128;
//   j < 64 -> a-row blk*64 + j ; else g-row F + blk*64 + (j-64).   F % 64 == 0.
template <typename T>
__global__ void pack_geglu_kernel(const float* __restrict__ w, const float* __restrict__ bias, T* __restrict__ wo,
                                  float* __restrict__ bo, int F, int K) {
  long long n = 2LL * F * K;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int k = (int)(i % K);
    int r = (int)(i / K);
    int blk = r >> 7, j = r & 127;
    int src = (j < 64) ? (blk * 64 + j) : (F + blk * 64 + (j - 64));
    wo[i] = from_f<T>(w[(long long)src * K + k]);
    if (k == 0 && bias) bo[r] = bias[src];
  }
}

// LayerNorm folded into the following linear layer (one warp per output row n):
//   w'[n][k] = w[n][k] * gamma[k]     colsum[n] = sum_k round_bf16(w'[n][k])     bias'[n] = bias[n] + sum_k beta[k] w[n][k]
// so that  LN(x) W^T + b  =  rstd (x W'^T - mean colsum) + bias'.  colsum is taken over the ROUNDED weights the tensor
// core will multiply, so the mean term cancels exactly what the GEMM accumulates.
template <typename T>
__global__ void pack_lnfold_kernel(const float* __restrict__ w, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const float* __restrict__ bias, T* __restrict__ wo, float* __restrict__ colsum,
                                   float* __restrict__ bo, int N, int K) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (n >= N) return;
  float cs = 0.f, bs = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float wv = w[(long long)n * K + k];
    const float wg = wv * gamma[k];
    wo[(long long)n * K + k] = from_f<T>(wg);
    cs += __bfloat162float(__float2bfloat16_rn(wg));
    bs = fmaf(beta[k], wv, bs);
  }
  cs = warp_sum(cs);
  bs = warp_sum(bs);
  if (lane == 0) {
    colsum[n] = cs;
    bo[n] = bs + (bias ? bias[n] : 0.f);
  }
}

// ------------------------------------------------------------------------------------------------
// AudioAttnProcessor context step (one CTA per batch element).  See c2d.h.
//   g[k][j]   = gelu(W1[j,:] . a[b,k,:] + b1[j])                       k < K, j < Hb
//   gb[p][j]  = mean_{k in pool p} g[k][j]                               (linear: pooling commutes with W2)
//   v[p][d]   = W2[d,:] . gb[p] + b2[d]
//   ADD:    out[b,t,d] = ehs[b,t,d] + sigmoid(alpha) * v[0][d]
//   CONCAT: out[b,0:T] = ehs[b];  out[b,T+p,d] = v[p][d]
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void audio_context_kernel(const T* __restrict__ ehs, const T* __restrict__ audio, const T* __restrict__ w1,
                                     const float* __restrict__ b1, const T* __restrict__ w2,
                                     const float* __restrict__ b2, const float* __restrict__ alpha,
                                     T* __restrict__ out, int Tn, int D, int K, int Da, int Hb, int mode) {
  extern __shared__ float sm[];
  int P = (mode == C2D_AUDIO_ADD) ? 1 : (K > 4 ? 4 : K);
  float* g = sm;                 // [K][Hb]
  float* gb = g + K * Hb;        // [P][Hb]
  float* v = gb + P * Hb;        // [P][D]
  int b = blockIdx.x;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const T* a = audio + (long long)b * K * Da;
  const bool vec = (Da % 8 == 0) && (D % 8 == 0) && ((reinterpret_cast<uintptr_t>(w1) | reinterpret_cast<uintptr_t>(audio) |
                                                     reinterpret_cast<uintptr_t>(ehs) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  for (int o = warp; o < K * Hb; o += nwarp) {
    int k = o / Hb, j = o - k * Hb;
    float acc = 0.f;
    if (vec) {          // 16-byte loads, all of a lane's chunks in flight (the scalar loop was one load latency per element)
      const T* wr = w1 + (long long)j * Da;
      const T* ar = a + (long long)k * Da;
#pragma unroll 4
      for (int i = lane * 8; i < Da; i += 256) {
        float fw[8], fa[8];
        Vec8<T>::load(wr + i, fw);
        Vec8<T>::load(ar + i, fa);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc = fmaf(fw[u], fa[u], acc);
      }
    } else {
      for (int i = lane; i < Da; i += 32) acc += to_f<T>(w1[(long long)j * Da + i]) * to_f<T>(a[(long long)k * Da + i]);
    }
    acc = warp_sum(acc);
    if (lane == 0) g[o] = gelu_erf(acc + b1[j]);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < P * Hb; o += blockDim.x) {
    int p = o / Hb, j = o - p * Hb;
    int k0, k1;
    if (mode == C2D_AUDIO_ADD) { k0 = 0; k1 = K; }
    else if (K > 4) { k0 = (p * K) / 4; k1 = ((p + 1) * K + 3) / 4; }     // adaptive_avg_pool1d bins
    else { k0 = p; k1 = p + 1; }
    float s = 0.f;
    for (int k = k0; k < k1; ++k) s += g[k * Hb + j];
    gb[o] = s / (float)(k1 - k0);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < P * D; o += blockDim.x) {
    int p = o / D, d = o - p * D;
    float acc = b2[d];
    for (int j = 0; j < Hb; ++j) acc += to_f<T>(w2[(long long)d * Hb + j]) * gb[p * Hb + j];
    v[o] = acc;
  }
  __syncthreads();
  const T* e = ehs + (long long)b * Tn * D;
  if (mode == C2D_AUDIO_ADD) {
    float gate = 1.f / (1.f + expf(-alpha[0]));
    T* o = out + (long long)b * Tn * D;
    if (vec) {
      for (int i = threadIdx.x * 8; i < Tn * D; i += blockDim.x * 8) {
        float f[8];
        Vec8<T>::load(e + i, f);
        const int d0 = i % D;
#pragma unroll
        for (int u = 0; u < 8; ++u) f[u] = fmaf(gate, v[d0 + u], f[u]);
        Vec8<T>::store(o + i, f);
      }
    } else {
      for (int i = threadIdx.x; i < Tn * D; i += blockDim.x) o[i] = from_f<T>(to_f<T>(e[i]) + gate * v[i % D]);
    }
  } else {
    T* o = out + (long long)b * (Tn + P) * D;
    for (int i = threadIdx.x; i < Tn * D; i += blockDim.x) o[i] = e[i];
    for (int i = threadIdx.x; i < P * D; i += blockDim.x) o[Tn * D + i] = from_f<T>(v[i]);
  }
}

}  // namespace c2d

using namespace c2d;

#define DISPATCH_T(dtype, ...)                                   \
  if ((dtype) == C2D_F32) { typedef float T; __VA_ARGS__ }       \
  else if ((dtype) == C2D_BF16) { typedef bf16 T; __VA_ARGS__ }  \
  else { set_error("bad dtype %d", (int)(dtype)); return C2D_ERR_ARG; }

extern "C" {

int c2d_timestep_embedding(const float* t, float* out, int B, int dim, void* stream) {
  C2D_REQUIRE(t && out && B > 0 && dim > 0 && dim % 2 == 0, "timestep_embedding: bad args");
  int n = B * (dim / 2);
  timestep_embedding_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(t, out, B, dim);
  return check_launch("timestep_embedding");
}

int c2d_unary(const void* x, void* y, long long n, int act, int dtype_in, int dtype_out, void* stream) {
  C2D_REQUIRE(x && y && n > 0, "unary: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  int g = ew_grid((n + 7) / 8, 256);
  if (dtype_in == C2D_F32 && dtype_out == C2D_F32) unary_kernel<float, float><<<g, 256, 0, s>>>((const float*)x, (float*)y, n, act);
  else if (dtype_in == C2D_F32 && dtype_out == C2D_BF16) unary_kernel<float, bf16><<<g, 256, 0, s>>>((const float*)x, (bf16*)y, n, act);
  else if (dtype_in == C2D_BF16 && dtype_out == C2D_F32) unary_kernel<bf16, float><<<g, 256, 0, s>>>((const bf16*)x, (float*)y, n, act);
  else if (dtype_in == C2D_BF16 && dtype_out == C2D_BF16) unary_kernel<bf16, bf16><<<g, 256, 0, s>>>((const bf16*)x, (bf16*)y, n, act);
  else { set_error("unary: bad dtype"); return C2D_ERR_ARG; }
  return check_launch("unary");
}

int c2d_add(const void* a, const void* b, void* y, long long n, int dtype, void* stream) {
  C2D_REQUIRE(a && b && y && n > 0, "add: bad args");
  int g = ew_grid((n + 7) / 8, 256);
  DISPATCH_T(dtype, add_kernel<T><<<g, 256, 0, (cudaStream_t)stream>>>((const T*)a, (const T*)b, (T*)y, n);)
  return check_launch("add");
}

int c2d_geglu(const void* x, void* y, int M, int F, int dtype, void* stream) {
  C2D_REQUIRE(x && y && M > 0 && F > 0 && F % 8 == 0, "geglu: bad args (F %% 8)");
  int g = ew_grid((long long)M * (F / 8), 256);
  DISPATCH_T(dtype, geglu_kernel<T><<<g, 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, M, F);)
  return check_launch("geglu");
}

int c2d_upsample2x(const void* x, void* y, int B, int H, int W, int C, int dtype, void* stream) {
  C2D_REQUIRE(x && y && B > 0 && H > 0 && W > 0 && C % 8 == 0, "upsample2x: bad args (C %% 8)");
  int g = ew_grid((long long)B * 4 * H * W * (C / 8), 256);
  DISPATCH_T(dtype, upsample2x_kernel<T><<<g, 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, B, H, W, C);)
  return check_launch("upsample2x");
}

int c2d_concat(const void* x1, const void* x2, void* y, long long rows, int C1, int C2, int dtype, void* stream) {
  C2D_REQUIRE(x1 && x2 && y && rows > 0 && C1 % 8 == 0 && C2 % 8 == 0, "concat: bad args (C %% 8)");
  int g = ew_grid(rows * ((C1 + C2) / 8), 256);
  DISPATCH_T(dtype, concat_kernel<T><<<g, 256, 0, (cudaStream_t)stream>>>((const T*)x1, (const T*)x2, (T*)y, rows, C1, C2);)
  return check_launch("concat");
}

int c2d_nchw_to_nhwc(const float* x, void* y, int B, int C, int HW, int dtype, void* stream) {
  C2D_REQUIRE(x && y && B > 0 && C > 0 && HW > 0, "nchw_to_nhwc: bad args");
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), B), blk(32, 8);
  DISPATCH_T(dtype, layout_kernel<T, true><<<grid, blk, 0, (cudaStream_t)stream>>>(x, y, C, HW);)
  return check_launch("nchw_to_nhwc");
}

int c2d_nhwc_to_nchw(const void* x, float* y, int B, int C, int HW, int dtype, void* stream) {
  C2D_REQUIRE(x && y && B > 0 && C > 0 && HW > 0, "nhwc_to_nchw: bad args");
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), B), blk(32, 8);
  DISPATCH_T(dtype, layout_kernel<T, false><<<grid, blk, 0, (cudaStream_t)stream>>>(x, y, C, HW);)
  return check_launch("nhwc_to_nchw");
}

int c2d_cfg_sched_step(const void* eps2, float* x, void* xin2, float* trace, int B, int HW, float guidance,
                       const float* coef, int xin_cpitch, int dtype, void* stream) {
  C2D_REQUIRE(eps2 && x && xin2 && coef && B > 0 && HW > 0, "cfg_sched_step: bad args");
  C2D_REQUIRE(xin_cpitch >= 4, "cfg_sched_step: xin2 channel pitch %d < 4", xin_cpitch);
  long long n = (long long)B * 4 * HW;
  DISPATCH_T(dtype, cfg_sched_kernel<T><<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
                        (const T*)eps2, x, (T*)xin2, trace, B, HW, guidance, coef, xin_cpitch);)
  return check_launch("cfg_sched_step");
}

int c2d_pack_conv3x3(const float* w, void* out, int Cout, int Cin, int dtype, void* stream) {
  C2D_REQUIRE(w && out && Cout > 0 && Cin > 0, "pack_conv3x3: bad args");
  int g = ew_grid((long long)Cout * Cin * 9, 256);
  DISPATCH_T(dtype, pack_conv_kernel<T><<<g, 256, 0, (cudaStream_t)stream>>>(w, (T*)out, Cout, Cin);)
  return check_launch("pack_conv3x3");
}

int c2d_pack_geglu(const float* w, const float* bias, void* w_out, float* bias_out, int F, int K, int dtype,
                   void* stream) {
  C2D_REQUIRE(w && w_out && F > 0 && F % 64 == 0 && K > 0, "pack_geglu: bad args (F %% 64)");
  C2D_REQUIRE(!bias || bias_out, "pack_geglu: bias_out missing");
  int g = ew_grid(2LL * F * K, 256);
  DISPATCH_T(dtype, pack_geglu_kernel<T><<<g, 256, 0, (cudaStream_t)stream>>>(w, bias, (T*)w_out, bias_out, F, K);)
  return check_launch("pack_geglu");
}

int c2d_pack_lnfold(const float* w, const float* gamma, const float* beta, const float* bias, void* w_out, float* colsum,
                    float* bias_out, int N, int K, int dtype, void* stream) {
  C2D_REQUIRE(w && gamma && beta && w_out && colsum && bias_out && N > 0 && K > 0, "pack_lnfold: bad args");
  const int threads = 256, rows_per_cta = threads / 32;
  DISPATCH_T(dtype, pack_lnfold_kernel<T><<<ceil_div(N, rows_per_cta), threads, 0, (cudaStream_t)stream>>>(
                        w, gamma, beta, bias, (T*)w_out, colsum, bias_out, N, K);)
  return check_launch("pack_lnfold");
}

int c2d_audio_context(const void* ehs, const void* audio, const void* w1, const float* b1, const void* w2,
                      const float* b2, const float* alpha, void* ehs_out, int B, int T_, int D, int K, int Da, int Hb,
                      int mode, int dtype, void* stream) {
  C2D_REQUIRE(ehs && audio && w1 && b1 && w2 && b2 && ehs_out, "audio_context: null pointer");
  C2D_REQUIRE(mode == C2D_AUDIO_ADD || mode == C2D_AUDIO_CONCAT, "audio_context: bad mode %d", mode);
  C2D_REQUIRE(mode != C2D_AUDIO_ADD || alpha, "audio_context: alpha required for ADD");
  C2D_REQUIRE(B > 0 && T_ > 0 && D > 0 && K > 0 && Da > 0 && Hb > 0, "audio_context: bad dims");
  int P = (mode == C2D_AUDIO_ADD) ? 1 : (K > 4 ? 4 : K);
  size_t smem = sizeof(float) * ((size_t)K * Hb + (size_t)P * Hb + (size_t)P * D);
  C2D_REQUIRE(smem <= 48 * 1024, "audio_context: K*Hb + P*(Hb+D) too large for smem (%zu B)", smem);
  DISPATCH_T(dtype, audio_context_kernel<T><<<B, 256, smem, (cudaStream_t)stream>>>(
                        (const T*)ehs, (const T*)audio, (const T*)w1, b1, (const T*)w2, b2, alpha, (T*)ehs_out, T_, D,
                        K, Da, Hb, mode);)
  return check_launch("audio_context");
}

}  // extern "C"
