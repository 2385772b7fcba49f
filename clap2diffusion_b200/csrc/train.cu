// Backward kernels of the stage-3 fine-tune step (BASELINE config 5: audio attention processors
// trainable, SD-1.5 UNet frozen).  With the UNet frozen the backward pass is ACTIVATION gradients only: every dense
// layer's dgrad is the forward tcgen05 / FFMA kernel again on a transposed (linear) or flipped + transposed (3x3
// convolution) copy of the frozen weight, so the kernels here are only the ones that have no forward twin:
//   GroupNorm(+SiLU) backward, LayerNorm backward, GEGLU backward, flash-attention backward (recompute form: a
//   log-sum-exp / delta pre-pass, a dQ pass and a dK/dV pass; fp32 accumulation on the FFMA pipe), the adjoints of the
//   stride-2 convolution gather (zero insertion) and of the nearest 2x upsample (2x2 sum), channel slicing (adjoint of
//   the skip concat), the loss (weighted MSE + its gradient), column sums / gate / GELU adjoints of the processor's
//   audio branch (reference models/audio_attention_processor.py:85-99), and the optimiser (global-norm clip + AdamW,
//   reference scripts/train_stage3.py:33-38, :182-188).
// All take fp32 or bf16 activations (same dtype switch as the forward kernels) and accumulate in fp32.
#include <float.h>

#include "common.cuh"

namespace c2d {

static inline int tr_grid(long long work_items, int threads) {
  long long blocks = (work_items + threads - 1) / threads;
  long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

__device__ __forceinline__ float block_sum(float v, float* sh) {   // sh: >= 32 floats; all threads get the total
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
  if (w == 0) {
    t = warp_sum(t);
    if (lane == 0) sh[0] = t;
  }
  __syncthreads();
  return sh[0];
}

// ------------------------------------------------------------------------------------------------ GroupNorm (+SiLU) backward
// One CTA per (group, sample): statistics of x are recomputed (pass 1), then the two group means of the adjoint
// (pass 2), then dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) with g = dz * gamma, dz = dy * silu'(z) (pass 3).
template <typename T>
__global__ void __launch_bounds__(512)
gn_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, const float* __restrict__ gamma,
              const float* __restrict__ beta, const T* __restrict__ add, T* __restrict__ dx, int HW, int C, int groups,
              float eps, int silu) {
  __shared__ float sh[32];
  const int g = blockIdx.x, b = blockIdx.y;
  const int cpg = C / groups, c0 = g * cpg;
  const long long base = (long long)b * HW * C + c0;
  const int n = HW * cpg;
  float s = 0.f, ss = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int r = i / cpg, c = i - r * cpg;
    const float v = to_f<T>(x[base + (long long)r * C + c]);
    s += v; ss += v * v;
  }
  const float inv_n = 1.f / (float)n;
  const float mean = block_sum(s, sh) * inv_n;
  const float var = fmaxf(block_sum(ss, sh) * inv_n - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int r = i / cpg, c = i - r * cpg;
    const long long o = base + (long long)r * C + c;
    const float xh = (to_f<T>(x[o]) - mean) * rstd;
    const float ga = gamma[c0 + c];
    float d = to_f<T>(dy[o]);
    if (silu) {
      const float z = ga * xh + beta[c0 + c];
      const float sg = 1.f / (1.f + expf(-z));
      d *= sg * (1.f + z * (1.f - sg));
    }
    const float gg = d * ga;
    s1 += gg; s2 += gg * xh;
  }
  const float m1 = block_sum(s1, sh) * inv_n;
  const float m2 = block_sum(s2, sh) * inv_n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int r = i / cpg, c = i - r * cpg;
    const long long o = base + (long long)r * C + c;
    const float xh = (to_f<T>(x[o]) - mean) * rstd;
    const float ga = gamma[c0 + c];
    float d = to_f<T>(dy[o]);
    if (silu) {
      const float z = ga * xh + beta[c0 + c];
      const float sg = 1.f / (1.f + expf(-z));
      d *= sg * (1.f + z * (1.f - sg));
    }
    float v = rstd * (d * ga - m1 - xh * m2);
    if (add) v += to_f<T>(add[o]);
    dx[o] = from_f<T>(v);
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm backward
// One warp per row: dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) (+ add), g = dy * gamma.
template <typename T>
__global__ void ln_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, const float* __restrict__ gamma,
                              const T* __restrict__ add, T* __restrict__ dx, int M, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const T* xr = x + (long long)row * C;
  const T* dr = dy + (long long)row * C;
  float s = 0.f, ss = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float v = to_f<T>(xr[c]);
    s += v; ss += v * v;
  }
  const float inv = 1.f / (float)C;
  const float mean = warp_sum(s) * inv;
  const float var = fmaxf(warp_sum(ss) * inv - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float xh = (to_f<T>(xr[c]) - mean) * rstd;
    const float gg = to_f<T>(dr[c]) * gamma[c];
    s1 += gg; s2 += gg * xh;
  }
  const float m1 = warp_sum(s1) * inv, m2 = warp_sum(s2) * inv;
  const T* ar = add ? add + (long long)row * C : nullptr;
  T* o = dx + (long long)row * C;
  for (int c = lane; c < C; c += 32) {
    const float xh = (to_f<T>(xr[c]) - mean) * rstd;
    float v = rstd * (to_f<T>(dr[c]) * gamma[c] - m1 - xh * m2);
    if (ar) v += to_f<T>(ar[c]);
    o[c] = from_f<T>(v);
  }
}

// ------------------------------------------------------------------------------------------------ GEGLU backward
// y = a * gelu(g), ag = [a | g] rows of 2F:  da = dy * gelu(g), dg = dy * a * (Phi(g) + g phi(g))
template <typename T>
__global__ void geglu_bwd_kernel(const T* __restrict__ ag, const T* __restrict__ dy, T* __restrict__ dag, long long M, int F) {
  const long long total = M * F;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long m = i / F;
    const int f = (int)(i - m * F);
    const float a = to_f<T>(ag[m * 2 * F + f]), g = to_f<T>(ag[m * 2 * F + F + f]), d = to_f<T>(dy[i]);
    const float cdf = 0.5f * (1.f + erff(g * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * g * g);
    dag[m * 2 * F + f] = from_f<T>(d * g * cdf);
    dag[m * 2 * F + F + f] = from_f<T>(d * a * (cdf + g * pdf));
  }
}

// elementwise: dz = dh[row / K] * gelu'(z) / K   (adjoint of mean over K tokens of gelu(z); fp32, audio branch)
__global__ void gelu_bwd_bcast_kernel(const float* __restrict__ z, const float* __restrict__ dh, float* __restrict__ dz,
                                      int rows, int H, int K) {
  const int total = rows * H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / H, j = i - r * H;
    const float g = z[i];
    const float cdf = 0.5f * (1.f + erff(g * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * g * g);
    dz[i] = dh[(r / K) * H + j] * (cdf + g * pdf) / (float)K;
  }
}

// ------------------------------------------------------------------------------------------------ layout adjoints
// z[b][2y][2x][c] = x[b][y][x][c], zeros elsewhere: adjoint of the stride-2 gather of a pad-1 3x3 convolution
template <typename T>
__global__ void zero_insert2x_kernel(const T* __restrict__ x, T* __restrict__ z, int B, int H, int W, int C) {
  const long long total = (long long)B * 2 * H * 2 * W * C;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % C);
    long long p = i / C;
    const int xx = (int)(p % (2 * W)); p /= 2 * W;
    const int yy = (int)(p % (2 * H));
    const int b = (int)(p / (2 * H));
    T v = from_f<T>(0.f);
    if (!(xx & 1) && !(yy & 1)) v = x[(((long long)b * H + (yy >> 1)) * W + (xx >> 1)) * C + c];
    z[i] = v;
  }
}
// y[b][y][x][c] = sum of the 2x2 block of x[b][2y..2y+1][2x..2x+1][c]: adjoint of the nearest 2x upsample
template <typename T>
__global__ void sumpool2x2_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C) {
  const long long total = (long long)B * H * W * C;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % C);
    long long p = i / C;
    const int xx = (int)(p % W); p /= W;
    const int yy = (int)(p % H);
    const int b = (int)(p / H);
    const long long r0 = (((long long)b * 2 * H + 2 * yy) * 2 * W + 2 * xx) * C + c;
    const long long r1 = r0 + (long long)2 * W * C;
    y[i] = from_f<T>(to_f<T>(x[r0]) + to_f<T>(x[r0 + C]) + to_f<T>(x[r1]) + to_f<T>(x[r1 + C]));
  }
}
// y[r][0:Cs] = x[r][c0 : c0+Cs] (+ add[r][0:Cs]): adjoint of the channel concat, with the fan-in sum folded in
template <typename T>
__global__ void slice_channels_kernel(const T* __restrict__ x, const T* __restrict__ add, T* __restrict__ y, long long rows,
                                      int C, int c0, int Cs) {
  const long long total = rows * Cs;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long r = i / Cs;
    const int c = (int)(i - r * Cs);
    float v = to_f<T>(x[r * C + c0 + c]);
    if (add) v += to_f<T>(add[i]);
    y[i] = from_f<T>(v);
  }
}

// ------------------------------------------------------------------------------------------------ loss
// pred: NHWC [B][HW][C] (dtype T), target: fp32 NCHW [B][C][HW].  grad = weight * 2 (pred - target) / n (NHWC, T);
// loss (double, accumulated) += weight * sum((pred - target)^2) / n
template <typename T>
__global__ void mse_loss_grad_kernel(const T* __restrict__ pred, const float* __restrict__ target, T* __restrict__ grad,
                                     double* __restrict__ loss, int B, int HW, int C, float weight) {
  __shared__ float sh[32];
  const long long n = (long long)B * HW * C;
  const float k = 2.f * weight / (float)n;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long p = i / C;
    const int hw = (int)(p % HW);
    const int b = (int)(p / HW);
    const float d = to_f<T>(pred[i]) - target[((long long)b * C + c) * HW + hw];
    acc += d * d;
    grad[i] = from_f<T>(k * d);
  }
  const float tot = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(loss, (double)tot * (double)weight / (double)n);
}

// ------------------------------------------------------------------------------------------------ small reductions
// out[b][c] (fp32) = sum_r x[b][r][c]   (x dtype T; grid: (ceil(C / 128), B))
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, float* __restrict__ out, int R, int C, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (c >= C) return;
  const T* p = x + (long long)b * R * C + c;
  float s = 0.f;
  for (int r = 0; r < R; ++r) s += to_f<T>(p[(long long)r * C]);
  float* o = out + (long long)b * C + c;
  *o = accumulate ? *o + s : s;
}

// gate adjoint of  ehs' = ehs + sigmoid(alpha) * af:   dalpha += gate (1 - gate) <s, af>,  daf = gate * s   (fp32, n = B*D)
__global__ void gate_bwd_kernel(const float* __restrict__ s, const float* __restrict__ af, const float* __restrict__ alpha,
                                float* __restrict__ daf, float* __restrict__ dalpha, int n) {
  __shared__ float sh[32];
  const float gate = 1.f / (1.f + expf(-alpha[0]));
  float acc = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    acc += s[i] * af[i];
    daf[i] = gate * s[i];
  }
  const float tot = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(dalpha, tot * gate * (1.f - gate));
}

// ------------------------------------------------------------------------------------------------ optimiser
__global__ void sumsq_kernel(const float* __restrict__ x, long long n, double* __restrict__ out) {
  __shared__ float sh[32];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) acc += x[i] * x[i];
  const float tot = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(out, (double)tot);
}
// torch.nn.utils.clip_grad_norm_: scale = min(1, max_norm / (norm + 1e-6))
__global__ void clip_scale_kernel(const double* __restrict__ sumsq, float max_norm, float* __restrict__ scale, float* __restrict__ norm_out) {
  const float norm = (float)sqrt(*sumsq);
  *scale = fminf(1.f, max_norm / (norm + 1e-6f));
  if (norm_out) *norm_out = norm;
}
// torch.optim.AdamW (decoupled weight decay, bias correction), gradient pre-multiplied by *grad_scale
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             long long n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2,
                             const float* __restrict__ grad_scale) {
  const float gs = grad_scale ? *grad_scale : 1.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    pi -= lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
    p[i] = pi;
  }
}

// The same update with the schedule on the device: sched[t] = (learning rate, 1 - beta1^(t+1), 1 - beta2^(t+1)) for optimiser step
// t, *step_dev = number of steps taken so far.  Nothing about the step is a kernel argument, so the launch can sit in a CUDA
// graph that is replayed for every step (the caller increments *step_dev after the launch, inside the same graph).
__global__ void adamw_sched_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                   long long n, const float* __restrict__ sched, int sched_len, const int* __restrict__ step_dev,
                                   float b1, float b2, float eps, float wd, const float* __restrict__ grad_scale) {
  int t = *step_dev;
  t = t < 0 ? 0 : (t >= sched_len ? sched_len - 1 : t);
  const float lr = sched[3 * t], bc1 = sched[3 * t + 1], bc2 = sched[3 * t + 2];
  const float gs = grad_scale ? *grad_scale : 1.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    pi -= lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
    p[i] = pi;
  }
}

// ------------------------------------------------------------------------------------------------ attention backward
// Recompute-form flash attention backward on the FFMA pipe (fp32 accumulation; fp32 or bf16 storage).
//   S = scale Q K^T, P = softmax(S), O = P V;  D_i = <dO_i, O_i>;  dV = P^T dO;  dP = dO V^T;  dS = P o (dP - D);
//   dQ = scale dS K;  dK = scale dS^T Q.
// 64 x 64 score tiles, 256 threads (16 x 16, 4 x 4 scores each), operands staged in shared memory as fp32.
constexpr int AB_T = 64, AB_THREADS = 256, AB_MAXD = 160;

struct AttnBwdParams {
  const void *q, *k, *v, *o, *dout;
  void *dq, *dk, *dv;
  float *lse, *delta;                   // [B][heads][Nq] each
  int Nq, Nkv, d, heads;
  long long ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;      // row strides (elements)
  long long bsq, bsk, bsv, bso, bsdo, bsdq, bsdk, bsdv;      // batch strides
  float scale;
};

template <typename T>
__device__ __forceinline__ void ab_load_tile(float* dst, int dp, const T* src, long long ld, int row0, int nrows_valid, int d) {
  // dst[64][dp] <- src rows [row0, row0 + 64) x d columns (zeros past nrows_valid)
  for (int i = threadIdx.x; i < AB_T * d; i += AB_THREADS) {
    const int r = i / d, c = i - r * d;
    dst[r * dp + c] = (row0 + r < nrows_valid) ? to_f<T>(src[(long long)(row0 + r) * ld + c]) : 0.f;
  }
}

// pass 0: lse2[i] = log2(sum_j 2^(s2_ij)) with s2 = S * log2(e), and delta[i] = <dO_i, O_i>
template <typename T>
__global__ void __launch_bounds__(AB_THREADS)
attn_bwd_prep_kernel(const AttnBwdParams p) {
  extern __shared__ float smem[];
  const int dp = p.d + 4;
  float* sQ = smem;
  float* sK = sQ + AB_T * dp;
  float* sS = sK + AB_T * dp;              // [64][65]
  __shared__ float s_m[AB_T], s_l[AB_T];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AB_T;
  const T* q = reinterpret_cast<const T*>(p.q) + b * p.bsq + (long long)h * p.d;
  const T* k = reinterpret_cast<const T*>(p.k) + b * p.bsk + (long long)h * p.d;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float sc2 = p.scale * 1.4426950408889634f;
  ab_load_tile<T>(sQ, dp, q, p.ldq, q0, p.Nq, p.d);
  if (threadIdx.x < AB_T) { s_m[threadIdx.x] = -FLT_MAX; s_l[threadIdx.x] = 0.f; }
  for (int k0 = 0; k0 < p.Nkv; k0 += AB_T) {
    __syncthreads();
    ab_load_tile<T>(sK, dp, k, p.ldk, k0, p.Nkv, p.d);
    __syncthreads();
    float s[4][4] = {};
    for (int c = 0; c < p.d; ++c) {
      float qa[4], kb[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) { qa[a] = sQ[(ty * 4 + a) * dp + c]; kb[a] = sK[(tx * 4 + a) * dp + c]; }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int e = 0; e < 4; ++e) s[a][e] = fmaf(qa[a], kb[e], s[a][e]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        sS[(ty * 4 + a) * 65 + tx * 4 + e] = (k0 + tx * 4 + e < p.Nkv) ? s[a][e] * sc2 : -FLT_MAX;
    __syncthreads();
    if (threadIdx.x < AB_T) {
      const int r = threadIdx.x;
      float m = s_m[r];
      float mx = m;
      for (int j = 0; j < AB_T; ++j) mx = fmaxf(mx, sS[r * 65 + j]);
      float l = s_l[r] * exp2f(m - mx);
      for (int j = 0; j < AB_T; ++j) l += exp2f(sS[r * 65 + j] - mx);
      s_m[r] = mx; s_l[r] = l;
    }
  }
  __syncthreads();
  if (threadIdx.x < AB_T && q0 + threadIdx.x < p.Nq) {
    const int r = threadIdx.x, i = q0 + r;
    const long long idx = ((long long)b * p.heads + h) * p.Nq + i;
    p.lse[idx] = s_m[r] + log2f(s_l[r]);
    const T* orow = reinterpret_cast<const T*>(p.o) + b * p.bso + (long long)i * p.ldo + (long long)h * p.d;
    const T* drow = reinterpret_cast<const T*>(p.dout) + b * p.bsdo + (long long)i * p.lddo + (long long)h * p.d;
    float acc = 0.f;
    for (int c = 0; c < p.d; ++c) acc += to_f<T>(orow[c]) * to_f<T>(drow[c]);
    p.delta[idx] = acc;
  }
}

// shared by the dQ and dK/dV passes: scores and dP of one 64 x 64 tile -> P and dS (scaled) in shared memory
__device__ __forceinline__ void ab_scores(const float* sQ, const float* sK, const float* sV, const float* sdO, int dp, int d,
                                          const float* s_lse, const float* s_dl, float sc2, float scale, int q0, int k0,
                                          int Nq, int Nkv, float* sP, float* sdS) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float s[4][4] = {}, dpv[4][4] = {};
  for (int c = 0; c < d; ++c) {
    float qa[4], oa[4], kb[4], vb[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      qa[a] = sQ[(ty * 4 + a) * dp + c]; oa[a] = sdO[(ty * 4 + a) * dp + c];
      kb[a] = sK[(tx * 4 + a) * dp + c]; vb[a] = sV[(tx * 4 + a) * dp + c];
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        s[a][e] = fmaf(qa[a], kb[e], s[a][e]);
        dpv[a][e] = fmaf(oa[a], vb[e], dpv[a][e]);
      }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int r = ty * 4 + a;
    const bool rok = q0 + r < Nq;
    const float lse = s_lse[r], dl = s_dl[r];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int cidx = tx * 4 + e;
      const bool ok = rok && (k0 + cidx < Nkv);
      const float pv = ok ? exp2f(s[a][e] * sc2 - lse) : 0.f;
      sP[r * 65 + cidx] = pv;
      sdS[r * 65 + cidx] = pv * (dpv[a][e] - dl) * scale;
    }
  }
}

// NCOL = ceil(d / 16): output columns per thread (compile-time so that the accumulators stay in registers)
template <typename T, int NCOL>
__global__ void __launch_bounds__(AB_THREADS)
attn_bwd_dq_kernel(const AttnBwdParams p) {
  extern __shared__ float smem[];
  const int dp = p.d + 4;
  float* sQ = smem;
  float* sdO = sQ + AB_T * dp;
  float* sK = sdO + AB_T * dp;
  float* sV = sK + AB_T * dp;
  float* sP = sV + AB_T * dp;               // [64][65]
  float* sdS = sP + AB_T * 65;
  __shared__ float s_lse[AB_T], s_dl[AB_T];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AB_T;
  const long long ho = (long long)h * p.d;
  const T* q = reinterpret_cast<const T*>(p.q) + b * p.bsq + ho;
  const T* k = reinterpret_cast<const T*>(p.k) + b * p.bsk + ho;
  const T* v = reinterpret_cast<const T*>(p.v) + b * p.bsv + ho;
  const T* dO = reinterpret_cast<const T*>(p.dout) + b * p.bsdo + ho;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float sc2 = p.scale * 1.4426950408889634f;
  ab_load_tile<T>(sQ, dp, q, p.ldq, q0, p.Nq, p.d);
  ab_load_tile<T>(sdO, dp, dO, p.lddo, q0, p.Nq, p.d);
  if (threadIdx.x < AB_T) {
    const int i = q0 + threadIdx.x;
    const long long idx = ((long long)b * p.heads + h) * p.Nq + i;
    s_lse[threadIdx.x] = i < p.Nq ? p.lse[idx] : 0.f;
    s_dl[threadIdx.x] = i < p.Nq ? p.delta[idx] : 0.f;
  }
  float acc[4][NCOL] = {};
  for (int k0 = 0; k0 < p.Nkv; k0 += AB_T) {
    __syncthreads();
    ab_load_tile<T>(sK, dp, k, p.ldk, k0, p.Nkv, p.d);
    ab_load_tile<T>(sV, dp, v, p.ldv, k0, p.Nkv, p.d);
    __syncthreads();
    ab_scores(sQ, sK, sV, sdO, dp, p.d, s_lse, s_dl, sc2, p.scale, q0, k0, p.Nq, p.Nkv, sP, sdS);
    __syncthreads();
    for (int j = 0; j < AB_T; ++j) {
      float ds[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) ds[a] = sdS[(ty * 4 + a) * 65 + j];
#pragma unroll
      for (int cj = 0; cj < NCOL; ++cj) {
        const int c = tx + 16 * cj;
        const float kv = c < p.d ? sK[j * dp + c] : 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) acc[a][cj] = fmaf(ds[a], kv, acc[a][cj]);
      }
    }
  }
  T* dq = reinterpret_cast<T*>(p.dq) + b * p.bsdq + ho;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = q0 + ty * 4 + a;
    if (i >= p.Nq) continue;
#pragma unroll
    for (int cj = 0; cj < NCOL; ++cj) {
      const int c = tx + 16 * cj;
      if (c < p.d) dq[(long long)i * p.lddq + c] = from_f<T>(acc[a][cj]);
    }
  }
}

template <typename T, int NCOL>
__global__ void __launch_bounds__(AB_THREADS)
attn_bwd_dkv_kernel(const AttnBwdParams p) {
  extern __shared__ float smem[];
  const int dp = p.d + 4;
  float* sQ = smem;
  float* sdO = sQ + AB_T * dp;
  float* sK = sdO + AB_T * dp;
  float* sV = sK + AB_T * dp;
  float* sP = sV + AB_T * dp;
  float* sdS = sP + AB_T * 65;
  __shared__ float s_lse[AB_T], s_dl[AB_T];
  const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * AB_T;
  const long long ho = (long long)h * p.d;
  const T* q = reinterpret_cast<const T*>(p.q) + b * p.bsq + ho;
  const T* k = reinterpret_cast<const T*>(p.k) + b * p.bsk + ho;
  const T* v = reinterpret_cast<const T*>(p.v) + b * p.bsv + ho;
  const T* dO = reinterpret_cast<const T*>(p.dout) + b * p.bsdo + ho;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float sc2 = p.scale * 1.4426950408889634f;
  ab_load_tile<T>(sK, dp, k, p.ldk, k0, p.Nkv, p.d);
  ab_load_tile<T>(sV, dp, v, p.ldv, k0, p.Nkv, p.d);
  float acck[4][NCOL] = {}, accv[4][NCOL] = {};
  for (int q0 = 0; q0 < p.Nq; q0 += AB_T) {
    __syncthreads();
    ab_load_tile<T>(sQ, dp, q, p.ldq, q0, p.Nq, p.d);
    ab_load_tile<T>(sdO, dp, dO, p.lddo, q0, p.Nq, p.d);
    if (threadIdx.x < AB_T) {
      const int i = q0 + threadIdx.x;
      const long long idx = ((long long)b * p.heads + h) * p.Nq + i;
      s_lse[threadIdx.x] = i < p.Nq ? p.lse[idx] : 0.f;
      s_dl[threadIdx.x] = i < p.Nq ? p.delta[idx] : 0.f;
    }
    __syncthreads();
    ab_scores(sQ, sK, sV, sdO, dp, p.d, s_lse, s_dl, sc2, p.scale, q0, k0, p.Nq, p.Nkv, sP, sdS);
    __syncthreads();
    // key rows ty*4 + a of this tile, columns tx + 16 cj:  dV += P^T dO,  dK += dS^T Q
    for (int i = 0; i < AB_T; ++i) {
      float pv[4], ds[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) { pv[a] = sP[i * 65 + ty * 4 + a]; ds[a] = sdS[i * 65 + ty * 4 + a]; }
#pragma unroll
      for (int cj = 0; cj < NCOL; ++cj) {
        const int c = tx + 16 * cj;
        const float ov = c < p.d ? sdO[i * dp + c] : 0.f;
        const float qv = c < p.d ? sQ[i * dp + c] : 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          accv[a][cj] = fmaf(pv[a], ov, accv[a][cj]);
          acck[a][cj] = fmaf(ds[a], qv, acck[a][cj]);
        }
      }
    }
  }
  T* dk = reinterpret_cast<T*>(p.dk) + b * p.bsdk + ho;
  T* dv = reinterpret_cast<T*>(p.dv) + b * p.bsdv + ho;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int j = k0 + ty * 4 + a;
    if (j >= p.Nkv) continue;
#pragma unroll
    for (int cj = 0; cj < NCOL; ++cj) {
      const int c = tx + 16 * cj;
      if (c < p.d) {
        dk[(long long)j * p.lddk + c] = from_f<T>(acck[a][cj]);
        dv[(long long)j * p.lddv + c] = from_f<T>(accv[a][cj]);
      }
    }
  }
}

// the pre-pass kernel is shared by every NCOL instantiation: ONE record of the dynamic-smem size granted to it
template <typename T>
static int (&ab_prep_smem_set())[C2D_MAX_DEVICES] {
  static int set[C2D_MAX_DEVICES] = {};
  return set;
}

template <typename T, int NCOL>
static int attention_bwd_n(const AttnBwdParams& p, int B, cudaStream_t s) {
  const int dp = p.d + 4;
  const int smem_prep = (2 * AB_T * dp + AB_T * 65) * (int)sizeof(float);
  const int smem_main = (4 * AB_T * dp + 2 * AB_T * 65) * (int)sizeof(float);
  static int set_dq[C2D_MAX_DEVICES] = {}, set_dkv[C2D_MAX_DEVICES] = {};
  if (int rc = ensure_dyn_smem(attn_bwd_prep_kernel<T>, smem_prep, ab_prep_smem_set<T>(), "attention_bwd")) return rc;
  if (int rc = ensure_dyn_smem(attn_bwd_dq_kernel<T, NCOL>, smem_main, set_dq, "attention_bwd")) return rc;
  if (int rc = ensure_dyn_smem(attn_bwd_dkv_kernel<T, NCOL>, smem_main, set_dkv, "attention_bwd")) return rc;
  dim3 gq(ceil_div(p.Nq, AB_T), p.heads, B), gk(ceil_div(p.Nkv, AB_T), p.heads, B);
  attn_bwd_prep_kernel<T><<<gq, AB_THREADS, smem_prep, s>>>(p);
  if (int rc = check_launch("attention_bwd")) return rc;
  attn_bwd_dq_kernel<T, NCOL><<<gq, AB_THREADS, smem_main, s>>>(p);
  if (int rc = check_launch("attention_bwd")) return rc;
  attn_bwd_dkv_kernel<T, NCOL><<<gk, AB_THREADS, smem_main, s>>>(p);
  return check_launch("attention_bwd");
}

template <typename T>
static int attention_bwd_t(const AttnBwdParams& p, int B, cudaStream_t s) {
  const int ncol = (p.d + 15) / 16;
  if (ncol <= 3) return attention_bwd_n<T, 3>(p, B, s);      // d <= 48  (SD-1.5 level 0: 40)
  if (ncol <= 5) return attention_bwd_n<T, 5>(p, B, s);      // d <= 80
  return attention_bwd_n<T, AB_MAXD / 16>(p, B, s);          // d <= 160
}

// norm.cu: three coalesced passes (bf16, C % 8 == 0); returns -1 when the shape does not fit
int group_norm_bwd_fast(const void* x, const void* dy, const float* gamma, const float* beta, const void* add, void* dx, void* ws,
                        const long long* x_stats, int B, int HW, int C, int groups, float eps, int silu, cudaStream_t s);

// attn_bwd_tc.cu: tcgen05 kernels (bf16, head dims up to 192)
struct AttnBwdArgs {
  const void *q, *k, *v, *o, *dout;
  void *dq, *dk, *dv;
  float *lse, *delta;
  int B, heads, Nq, Nkv, d;
  long long ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv, bsq, bsk, bsv, bso, bsdo, bsdq, bsdk, bsdv;
  float scale;
  int have_lse;
};
bool attention_bwd_tc_supported(const AttnBwdArgs& a);
int attention_bwd_tc(const AttnBwdArgs& a, cudaStream_t s);

}  // namespace c2d

using namespace c2d;

#define TR_DISPATCH(dtype, CALL_F32, CALL_BF16)   \
  do {                                            \
    if ((dtype) == C2D_F32) { CALL_F32; }         \
    else if ((dtype) == C2D_BF16) { CALL_BF16; }  \
    else { set_error("bad dtype %d", (dtype)); return C2D_ERR_ARG; } \
  } while (0)

extern "C" {

int c2d_group_norm_bwd(const void* x, const void* dy, const float* gamma, const float* beta, const void* add, void* dx, void* ws,
                       const long long* x_stats, int B, int HW, int C, int groups, float eps, int silu, int dtype, void* stream) {
  C2D_REQUIRE(x && dy && gamma && beta && dx, "group_norm_bwd: null pointer");
  C2D_REQUIRE(B > 0 && HW > 0 && C > 0 && groups > 0 && C % groups == 0, "group_norm_bwd: bad dims");
  cudaStream_t s = (cudaStream_t)stream;
  if (ws && dtype == C2D_BF16 && C % 8 == 0 && C / 8 <= 512 && groups <= 32) {
    const int rc = group_norm_bwd_fast(x, dy, gamma, beta, add, dx, ws, x_stats, B, HW, C, groups, eps, silu, s);
    if (rc >= 0) return rc;
  }
  dim3 grid(groups, B);
  TR_DISPATCH(dtype,
              (gn_bwd_kernel<float><<<grid, 512, 0, s>>>((const float*)x, (const float*)dy, gamma, beta, (const float*)add, (float*)dx, HW, C, groups, eps, silu)),
              (gn_bwd_kernel<bf16><<<grid, 512, 0, s>>>((const bf16*)x, (const bf16*)dy, gamma, beta, (const bf16*)add, (bf16*)dx, HW, C, groups, eps, silu)));
  return check_launch("group_norm_bwd");
}

int c2d_layer_norm_bwd(const void* x, const void* dy, const float* gamma, const void* add, void* dx, int M, int C, float eps,
                       int dtype, void* stream) {
  C2D_REQUIRE(x && dy && gamma && dx && M > 0 && C > 0, "layer_norm_bwd: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  const int wpb = 8;
  TR_DISPATCH(dtype,
              (ln_bwd_kernel<float><<<ceil_div(M, wpb), wpb * 32, 0, s>>>((const float*)x, (const float*)dy, gamma, (const float*)add, (float*)dx, M, C, eps)),
              (ln_bwd_kernel<bf16><<<ceil_div(M, wpb), wpb * 32, 0, s>>>((const bf16*)x, (const bf16*)dy, gamma, (const bf16*)add, (bf16*)dx, M, C, eps)));
  return check_launch("layer_norm_bwd");
}

int c2d_geglu_bwd(const void* ag, const void* dy, void* dag, int M, int F, int dtype, void* stream) {
  C2D_REQUIRE(ag && dy && dag && M > 0 && F > 0, "geglu_bwd: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  const int g = tr_grid((long long)M * F, 256);
  TR_DISPATCH(dtype, (geglu_bwd_kernel<float><<<g, 256, 0, s>>>((const float*)ag, (const float*)dy, (float*)dag, M, F)),
              (geglu_bwd_kernel<bf16><<<g, 256, 0, s>>>((const bf16*)ag, (const bf16*)dy, (bf16*)dag, M, F)));
  return check_launch("geglu_bwd");
}

int c2d_attention_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, void* dq, void* dk, void* dv,
                      float* lse_ws, float* delta_ws, int B, int heads, int Nq, int Nkv, int d, long long ldq, long long ldk,
                      long long ldv, long long ldo, long long lddo, long long lddq, long long lddk, long long lddv,
                      long long bsq, long long bsk, long long bsv, long long bso, long long bsdo, long long bsdq,
                      long long bsdk, long long bsdv, float scale, int have_lse, int dtype, void* stream) {
  C2D_REQUIRE(q && k && v && o && dout && dq && dk && dv && lse_ws && delta_ws, "attention_bwd: null pointer");
  C2D_REQUIRE(B > 0 && heads > 0 && Nq > 0 && Nkv > 0 && d > 0 && d <= AB_MAXD, "attention_bwd: bad dims (head dim <= %d)", AB_MAXD);
  AttnBwdParams p = {q, k, v, o, dout, dq, dk, dv, lse_ws, delta_ws, Nq, Nkv, d, heads,
                     ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv, bsq, bsk, bsv, bso, bsdo, bsdq, bsdk, bsdv, scale};
  if (dtype == C2D_F32) return attention_bwd_t<float>(p, B, (cudaStream_t)stream);
  if (dtype == C2D_BF16) {
    const AttnBwdArgs a = {q, k, v, o, dout, dq, dk, dv, lse_ws, delta_ws, B, heads, Nq, Nkv, d, ldq, ldk, ldv, ldo, lddo, lddq, lddk,
                           lddv, bsq, bsk, bsv, bso, bsdo, bsdq, bsdk, bsdv, scale, have_lse};
    if (attention_bwd_tc_supported(a)) return attention_bwd_tc(a, (cudaStream_t)stream);      // tensor cores
    return attention_bwd_t<bf16>(p, B, (cudaStream_t)stream);                                 // head dims > 128: FFMA kernels
  }
  set_error("attention_bwd: bad dtype %d", dtype);
  return C2D_ERR_ARG;
}

int c2d_zero_insert2x(const void* x, void* z, int B, int H, int W, int C, int dtype, void* stream) {
  C2D_REQUIRE(x && z && B > 0 && H > 0 && W > 0 && C > 0, "zero_insert2x: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  const int g = tr_grid((long long)B * 4 * H * W * C, 256);
  TR_DISPATCH(dtype, (zero_insert2x_kernel<float><<<g, 256, 0, s>>>((const float*)x, (float*)z, B, H, W, C)),
              (zero_insert2x_kernel<bf16><<<g, 256, 0, s>>>((const bf16*)x, (bf16*)z, B, H, W, C)));
  return check_launch("zero_insert2x");
}

int c2d_sumpool2x2(const void* x, void* y, int B, int H, int W, int C, int dtype, void* stream) {
  C2D_REQUIRE(x && y && B > 0 && H > 0 && W > 0 && C > 0, "sumpool2x2: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  const int g = tr_grid((long long)B * H * W * C, 256);
  TR_DISPATCH(dtype, (sumpool2x2_kernel<float><<<g, 256, 0, s>>>((const float*)x, (float*)y, B, H, W, C)),
              (sumpool2x2_kernel<bf16><<<g, 256, 0, s>>>((const bf16*)x, (bf16*)y, B, H, W, C)));
  return check_launch("sumpool2x2");
}

int c2d_slice_channels(const void* x, const void* add, void* y, long long rows, int C, int c0, int Cs, int dtype, void* stream) {
  C2D_REQUIRE(x && y && rows > 0 && C > 0 && c0 >= 0 && Cs > 0 && c0 + Cs <= C, "slice_channels: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  const int g = tr_grid(rows * Cs, 256);
  TR_DISPATCH(dtype, (slice_channels_kernel<float><<<g, 256, 0, s>>>((const float*)x, (const float*)add, (float*)y, rows, C, c0, Cs)),
              (slice_channels_kernel<bf16><<<g, 256, 0, s>>>((const bf16*)x, (const bf16*)add, (bf16*)y, rows, C, c0, Cs)));
  return check_launch("slice_channels");
}

int c2d_mse_loss_grad(const void* pred, const float* target, void* grad, double* loss, int B, int HW, int C, float weight,
                      int dtype, void* stream) {
  C2D_REQUIRE(pred && target && grad && loss && B > 0 && HW > 0 && C > 0, "mse_loss_grad: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  const int g = tr_grid((long long)B * HW * C, 256);
  TR_DISPATCH(dtype, (mse_loss_grad_kernel<float><<<g, 256, 0, s>>>((const float*)pred, target, (float*)grad, loss, B, HW, C, weight)),
              (mse_loss_grad_kernel<bf16><<<g, 256, 0, s>>>((const bf16*)pred, target, (bf16*)grad, loss, B, HW, C, weight)));
  return check_launch("mse_loss_grad");
}

int c2d_colsum(const void* x, float* out, int B, int R, int C, int accumulate, int dtype, void* stream) {
  C2D_REQUIRE(x && out && B > 0 && R > 0 && C > 0, "colsum: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid(ceil_div(C, 128), B);
  TR_DISPATCH(dtype, (colsum_kernel<float><<<grid, 128, 0, s>>>((const float*)x, out, R, C, accumulate)),
              (colsum_kernel<bf16><<<grid, 128, 0, s>>>((const bf16*)x, out, R, C, accumulate)));
  return check_launch("colsum");
}

int c2d_gate_bwd(const float* s_in, const float* af, const float* alpha, float* daf, float* dalpha, int n, void* stream) {
  C2D_REQUIRE(s_in && af && alpha && daf && dalpha && n > 0, "gate_bwd: bad args");
  gate_bwd_kernel<<<tr_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(s_in, af, alpha, daf, dalpha, n);
  return check_launch("gate_bwd");
}

int c2d_gelu_bwd_bcast(const float* z, const float* dh, float* dz, int rows, int H, int K, void* stream) {
  C2D_REQUIRE(z && dh && dz && rows > 0 && H > 0 && K > 0 && rows % K == 0, "gelu_bwd_bcast: bad args");
  gelu_bwd_bcast_kernel<<<tr_grid((long long)rows * H, 256), 256, 0, (cudaStream_t)stream>>>(z, dh, dz, rows, H, K);
  return check_launch("gelu_bwd_bcast");
}

int c2d_sumsq(const float* x, long long n, double* out, void* stream) {
  C2D_REQUIRE(x && out && n > 0, "sumsq: bad args");
  sumsq_kernel<<<tr_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, out);
  return check_launch("sumsq");
}

int c2d_clip_scale(const double* sumsq, float max_norm, float* scale, float* norm_out, void* stream) {
  C2D_REQUIRE(sumsq && scale && max_norm > 0.f, "clip_scale: bad args");
  clip_scale_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sumsq, max_norm, scale, norm_out);
  return check_launch("clip_scale");
}

int c2d_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int step, const float* grad_scale, void* stream) {
  C2D_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "adamw_step: bad args");
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  adamw_kernel<<<tr_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                   weight_decay, bc1, bc2, grad_scale);
  return check_launch("adamw_step");
}

int c2d_adamw_step_sched(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, const float* sched,
                         int sched_len, const int* step_dev, float beta1, float beta2, float eps, float weight_decay,
                         const float* grad_scale, void* stream) {
  C2D_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0 && sched && sched_len > 0 && step_dev, "adamw_step_sched: bad args");
  adamw_sched_kernel<<<tr_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, sched, sched_len, step_dev,
                                                                         beta1, beta2, eps, weight_decay, grad_scale);
  return check_launch("adamw_step");
}

}  // extern "C"
