// FFMA (CUDA-core) GEMM with fp32 accumulation: the fp32 parity-mode path for every dense contraction
// (rel-L2 <= 1e-4 rules out TF32/bf16 tensor-core math, SURVEY §7 "Hard parts") and the path for shapes
// the tcgen05 kernels do not take (tiny M, K % 8 != 0, Cin = 4).  Two A-operand loaders:
//   plain   : x[M][K] row-major
//   im2col  : 3x3 / pad 1 / stride {1,2} / optional nearest-2x upsample gather from NHWC activations,
//             k = tap * Cin + c  (weights packed [Cout][3][3][Cin])
// Tile 128x128x16, 256 threads, 8x8 register micro-tile, register-staged global prefetch.
#include "common.cuh"

namespace c2d {

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 16, SG_THREADS = 256, SG_LD = SG_BM + 4;

struct SimtGemmParams {
  const void* x;      // A source (plain or NHWC image)
  const void* w;      // [N][K]
  const float* bias;  // [N] or null
  const float* rowvec;  // [M / rows_per_vec][N] or null
  const void* residual;
  void* y;
  int M, N, K;
  long long ldx, ldy, ldr;
  int rows_per_vec;
  int act;
  int vec_a, vec_b;   // 8-wide vector loads legal for A / B
  // im2col geometry
  int H, W, Cin, Ho, Wo, stride, up, pad;
};

template <typename T>
__device__ __forceinline__ void load8_guard(const T* base, long long off, int valid, bool vec, float (&f)[8]) {
  if (vec && valid >= 8) {
    Vec8<T>::load(base + off, f);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (j < valid) ? to_f<T>(base[off + j]) : 0.f;
  }
}

template <typename T, bool CONV>
__device__ __forceinline__ void load_a(const SimtGemmParams& p, int m, int k0, float (&f)[8]) {
  if (m >= p.M || k0 >= p.K) {
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = 0.f;
    return;
  }
  const T* x = reinterpret_cast<const T*>(p.x);
  if (!CONV) {
    load8_guard<T>(x, (long long)m * p.ldx + k0, p.K - k0, p.vec_a, f);
  } else {
    int ox = m % p.Wo;
    int t = m / p.Wo;
    int oy = t % p.Ho;
    int b = t / p.Ho;
    if (p.vec_a) {   // Cin % 8 == 0: the 8-chunk lies inside one tap
      int tap = k0 / p.Cin, c = k0 - tap * p.Cin;
      int ky = tap / 3, kx = tap - ky * 3;
      int iy = oy * p.stride + ky - p.pad, ix = ox * p.stride + kx - p.pad;
      int Hin = p.up ? 2 * p.H : p.H, Win = p.up ? 2 * p.W : p.W;
      if (iy < 0 || iy >= Hin || ix < 0 || ix >= Win) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
      } else {
        if (p.up) { iy >>= 1; ix >>= 1; }
        Vec8<T>::load(x + (((long long)b * p.H + iy) * p.W + ix) * p.Cin + c, f);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int k = k0 + j;
        float v = 0.f;
        if (k < p.K) {
          int tap = k / p.Cin, c = k - tap * p.Cin;
          int ky = tap / 3, kx = tap - ky * 3;
          int iy = oy * p.stride + ky - p.pad, ix = ox * p.stride + kx - p.pad;
          int Hin = p.up ? 2 * p.H : p.H, Win = p.up ? 2 * p.W : p.W;
          if (iy >= 0 && iy < Hin && ix >= 0 && ix < Win) {
            if (p.up) { iy >>= 1; ix >>= 1; }
            v = to_f<T>(x[(((long long)b * p.H + iy) * p.W + ix) * p.Cin + c]);
          }
        }
        f[j] = v;
      }
    }
  }
}

template <typename T>
__device__ __forceinline__ void load_b(const SimtGemmParams& p, int n, int k0, float (&f)[8]) {
  if (n >= p.N || k0 >= p.K) {
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = 0.f;
    return;
  }
  load8_guard<T>(reinterpret_cast<const T*>(p.w), (long long)n * p.K + k0, p.K - k0, p.vec_b, f);
}

template <typename T, bool CONV>
__global__ void __launch_bounds__(SG_THREADS)
gemm_simt_kernel(const SimtGemmParams p) {
  __shared__ __align__(16) float As[SG_BK][SG_LD];
  __shared__ __align__(16) float Bs[SG_BK][SG_LD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const int lrow = tid >> 1, lk = (tid & 1) * 8;
  const int ty = tid >> 4, tx = tid & 15;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[8], rb[8];
  load_a<T, CONV>(p, m0 + lrow, lk, ra);
  load_b<T>(p, n0 + lrow, lk, rb);
  const int ktiles = (p.K + SG_BK - 1) / SG_BK;
  for (int kt = 0; kt < ktiles; ++kt) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { As[lk + j][lrow] = ra[j]; Bs[lk + j][lrow] = rb[j]; }
    __syncthreads();
    if (kt + 1 < ktiles) {
      load_a<T, CONV>(p, m0 + lrow, (kt + 1) * SG_BK + lk, ra);
      load_b<T>(p, n0 + lrow, (kt + 1) * SG_BK + lk, rb);
    }
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  T* y = reinterpret_cast<T*>(p.y);
  const T* res = reinterpret_cast<const T*>(p.residual);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= p.M) continue;
    const float* rv = p.rowvec ? p.rowvec + (long long)(m / p.rows_per_vec) * p.N : nullptr;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[n];
      if (rv) v += rv[n];
      v = apply_act(v, p.act);
      if (res) v += to_f<T>(res[(long long)m * p.ldr + n]);
      y[(long long)m * p.ldy + n] = from_f<T>(v);
    }
  }
}

static inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

int launch_gemm_simt(SimtGemmParams& p, bool conv, int dtype, cudaStream_t s) {
  dim3 grid(ceil_div(p.N, SG_BN), ceil_div(p.M, SG_BM));
  size_t es = dtype == C2D_F32 ? 4 : 2;
  p.vec_b = (p.K % 8 == 0) && aligned(p.w, 8 * es);
  if (conv) p.vec_a = (p.Cin % 8 == 0) && aligned(p.x, 8 * es);
  else p.vec_a = (p.K % 8 == 0) && (p.ldx % 8 == 0) && aligned(p.x, 8 * es);
  if (dtype == C2D_F32) {
    if (conv) gemm_simt_kernel<float, true><<<grid, SG_THREADS, 0, s>>>(p);
    else gemm_simt_kernel<float, false><<<grid, SG_THREADS, 0, s>>>(p);
  } else if (dtype == C2D_BF16) {
    if (conv) gemm_simt_kernel<bf16, true><<<grid, SG_THREADS, 0, s>>>(p);
    else gemm_simt_kernel<bf16, false><<<grid, SG_THREADS, 0, s>>>(p);
  } else {
    set_error("gemm_simt: bad dtype %d", dtype);
    return C2D_ERR_ARG;
  }
  return check_launch(conv ? "conv3x3_simt" : "linear_simt");
}

int linear_simt(const void* x, const void* w, const float* bias, const float* rowvec, int rows_per_vec,
                const void* residual, void* y, int M, int N, int K, int ldx, int ldy, int ldr, int act, int dtype,
                cudaStream_t s) {
  SimtGemmParams p = {};
  p.x = x; p.w = w; p.bias = bias; p.rowvec = rowvec; p.residual = residual; p.y = y;
  p.M = M; p.N = N; p.K = K; p.ldx = ldx; p.ldy = ldy; p.ldr = ldr;
  p.rows_per_vec = rows_per_vec > 0 ? rows_per_vec : 1;
  p.act = act;
  return launch_gemm_simt(p, false, dtype, s);
}

int conv3x3_simt(const void* x, const void* w, const float* bias, const float* rowvec, const void* residual, void* y,
                 int B, int H, int W, int Cin, int Cout, int stride, int up, int dtype, cudaStream_t s, int pad) {
  SimtGemmParams p = {};
  int Hin = up ? 2 * H : H, Win = up ? 2 * W : W;
  p.Ho = (Hin - 1) / stride + 1;   // pad 1, k 3: floor((Hin + 2 - 3)/stride) + 1
  p.Wo = (Win - 1) / stride + 1;   // pad 0 + (0,1,0,1) padding, stride 2, even H / W: the same H/2 x W/2
  p.pad = pad;
  p.x = x; p.w = w; p.bias = bias; p.rowvec = rowvec; p.residual = residual; p.y = y;
  p.M = B * p.Ho * p.Wo; p.N = Cout; p.K = 9 * Cin;
  p.ldx = 0; p.ldy = Cout; p.ldr = Cout;
  p.rows_per_vec = p.Ho * p.Wo;
  p.act = C2D_ACT_NONE;
  p.H = H; p.W = W; p.Cin = Cin; p.stride = stride; p.up = up;
  return launch_gemm_simt(p, true, dtype, s);
}

}  // namespace c2d
