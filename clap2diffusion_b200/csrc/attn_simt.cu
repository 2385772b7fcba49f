// FFMA flash-attention forward (online softmax, fp32 everywhere on chip): the fp32 parity-mode path
// for both the spatial self-attention and the text/audio cross-attention core, and the general path
// for the small audio-side attentions (head_dim 32..96, <= 77 tokens).
//   o[b, n, h*d + j] = sum_m softmax_m(q[b,n,h,:] . k[b,m,h,:] * scale + mask) v[b,m,h,j]
// One CTA = 64 queries of one (batch, head); keys streamed in tiles of 64; K and V share one smem tile.
// Never materialises the [B*heads, N, Nkv] probabilities the reference builds
// (models/audio_attention_processor.py:129-130).
#include <float.h>

#include "common.cuh"

namespace c2d {

constexpr int FA_BQ = 64, FA_BK = 64, FA_THREADS = 256;

template <typename T>
__device__ __forceinline__ void load_tile(float* dst, int DP, const T* src, long long ld, int row0, int nrows_valid,
                                          int d) {
  // 64 rows x d cols, Vec8 loads, zero-fill rows past the end
  const int vpr = d >> 3;
  for (int i = threadIdx.x; i < 64 * vpr; i += FA_THREADS) {
    int r = i / vpr, c = (i - r * vpr) * 8;
    float f[8];
    if (r < nrows_valid) Vec8<T>::load(src + (long long)(row0 + r) * ld + c, f);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = 0.f;
    }
    float* p = dst + r * DP + c;
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
}

template <typename T, int DJ>   // DJ = ceil(d / 16)
__global__ void __launch_bounds__(FA_THREADS)
attn_simt_kernel(const AttnParams p) {
  extern __shared__ __align__(16) float sm[];
  const int d = p.d, DP = d + 4;
  float* Qs = sm;                    // [64][DP]
  float* KVs = Qs + 64 * DP;         // [64][DP]
  float* Ps = KVs + 64 * DP;         // [64][68]
  constexpr int PLD = 68;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * FA_BQ;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const T* qb = reinterpret_cast<const T*>(p.q) + (long long)b * p.bsq + (long long)h * d;
  const T* kb = reinterpret_cast<const T*>(p.k) + (long long)b * p.bsk + (long long)h * d;
  const T* vb = reinterpret_cast<const T*>(p.v) + (long long)b * p.bsv + (long long)h * d;
  const uint8_t* mk = p.mask ? p.mask + (long long)b * p.Nkv : nullptr;

  load_tile<T>(Qs, DP, qb, p.ldq, q0, min(FA_BQ, p.Nq - q0), d);

  float o[4][DJ];
  float m_run[4], l_run[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -INFINITY;
    l_run[i] = 0.f;
#pragma unroll
    for (int j = 0; j < DJ; ++j) o[i][j] = 0.f;
  }

  for (int k0 = 0; k0 < p.Nkv; k0 += FA_BK) {
    const int kvalid = min(FA_BK, p.Nkv - k0);
    __syncthreads();                                   // previous PV done with KVs / Ps
    load_tile<T>(KVs, DP, kb, p.ldk, k0, kvalid, d);
    __syncthreads();
    // ---- S = Q K^T: rows ty*4+i, keys tx + 16*c
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) s[i][c] = 0.f;
    for (int kk = 0; kk < d; kk += 4) {
      float4 qv[4], kv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4*>(&Qs[(ty * 4 + i) * DP + kk]);
#pragma unroll
      for (int c = 0; c < 4; ++c) kv[c] = *reinterpret_cast<const float4*>(&KVs[(tx + 16 * c) * DP + kk]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          s[i][c] = fmaf(qv[i].x, kv[c].x, s[i][c]);
          s[i][c] = fmaf(qv[i].y, kv[c].y, s[i][c]);
          s[i][c] = fmaf(qv[i].z, kv[c].z, s[i][c]);
          s[i][c] = fmaf(qv[i].w, kv[c].w, s[i][c]);
        }
    }
    // ---- scale, mask, online softmax (row state replicated over the 16 tx lanes of a row group)
    float alpha[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        int key = tx + 16 * c;
        float v = s[i][c] * p.scale;
        if (key >= kvalid) v = -INFINITY;
        else if (mk && !mk[k0 + key]) v = -FLT_MAX;
        s[i][c] = v;
        mx = fmaxf(mx, v);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      float m_new = fmaxf(m_run[i], mx);
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float e = expf(s[i][c] - m_new);
        s[i][c] = e;
        sum += e;
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
      alpha[i] = expf(m_run[i] - m_new);
      l_run[i] = l_run[i] * alpha[i] + sum;
      m_run[i] = m_new;
#pragma unroll
      for (int c = 0; c < 4; ++c) Ps[(ty * 4 + i) * PLD + tx + 16 * c] = s[i][c];
    }
    __syncthreads();                                   // all warps done reading K; P visible
    load_tile<T>(KVs, DP, vb, p.ldv, k0, kvalid, d);
    __syncthreads();
    // ---- O = alpha * O + P V : rows ty*4+i, cols tx + 16*j
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < DJ; ++j) o[i][j] *= alpha[i];
    for (int key = 0; key < FA_BK; key += 4) {
      float4 pv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pv[i] = *reinterpret_cast<const float4*>(&Ps[(ty * 4 + i) * PLD + key]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float vv[DJ];
#pragma unroll
        for (int j = 0; j < DJ; ++j) {
          int col = tx + 16 * j;
          vv[j] = (col < d) ? KVs[(key + u) * DP + col] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float pw = u == 0 ? pv[i].x : (u == 1 ? pv[i].y : (u == 2 ? pv[i].z : pv[i].w));
#pragma unroll
          for (int j = 0; j < DJ; ++j) o[i][j] = fmaf(pw, vv[j], o[i][j]);
        }
      }
    }
  }
  // ---- normalise and store
  T* ob = reinterpret_cast<T*>(p.o) + (long long)b * p.bso + (long long)h * d;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int row = q0 + ty * 4 + i;
    if (row >= p.Nq) continue;
    float inv = 1.f / l_run[i];
#pragma unroll
    for (int j = 0; j < DJ; ++j) {
      int col = tx + 16 * j;
      if (col < d) ob[(long long)row * p.ldo + col] = from_f<T>(o[i][j] * inv);
    }
  }
}

template <typename T, int DJ>
static int launch_attn(const AttnParams& p, int B, cudaStream_t s) {
  size_t smem = sizeof(float) * (2 * 64 * (p.d + 4) + 64 * 68);
  static int smem_set[C2D_MAX_DEVICES] = {};   // per template instantiation and device
  if (int rc = ensure_dyn_smem(attn_simt_kernel<T, DJ>, 120 * 1024, smem_set, "attn_simt")) return rc;
  dim3 grid(ceil_div(p.Nq, FA_BQ), p.heads, B);
  attn_simt_kernel<T, DJ><<<grid, FA_THREADS, smem, s>>>(p);
  return check_launch("attn_simt");
}

int attention_simt(const AttnParams& p, int B, int dtype, cudaStream_t s) {
  int dj = (p.d + 15) / 16;
#define GO(T)                                                         \
  if (dj <= 2) return launch_attn<T, 2>(p, B, s);                     \
  if (dj <= 3) return launch_attn<T, 3>(p, B, s);                     \
  if (dj <= 4) return launch_attn<T, 4>(p, B, s);                     \
  if (dj <= 5) return launch_attn<T, 5>(p, B, s);                     \
  if (dj <= 6) return launch_attn<T, 6>(p, B, s);                     \
  if (dj <= 10) return launch_attn<T, 10>(p, B, s);
  if (dtype == C2D_F32) { GO(float) }
  else if (dtype == C2D_BF16) { GO(bf16) }
#undef GO
  set_error("attention_simt: unsupported head_dim %d / dtype %d (SIMT path takes d <= 160)", p.d, dtype);
  return C2D_ERR_UNSUPPORTED;
}

// ---- row softmax (for the materialised large-head-dim route: VAE mid attention, d = 512) ----------
template <typename T>
__global__ void softmax_rows_kernel(const T* __restrict__ x, T* __restrict__ y, int M, int N, float scale) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const T* xr = x + (long long)warp * N;
  T* yr = y + (long long)warp * N;
  float mx = -INFINITY;
  for (int c = lane; c < N; c += 32) mx = fmaxf(mx, to_f<T>(xr[c]) * scale);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int c = lane; c < N; c += 32) sum += expf(to_f<T>(xr[c]) * scale - mx);
  sum = warp_sum(sum);
  float inv = 1.f / sum;
  for (int c = lane; c < N; c += 32) yr[c] = from_f<T>(expf(to_f<T>(xr[c]) * scale - mx) * inv);
}

// y[C][R] = x[R][C]^T   (batched over blockIdx.z)
template <typename T>
__global__ void transpose_kernel(const T* __restrict__ x, T* __restrict__ y, int R, int C) {
  __shared__ float tile[32][33];
  const T* xb = x + (long long)blockIdx.z * R * C;
  T* yb = y + (long long)blockIdx.z * R * C;
  int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < R && c < C) ? to_f<T>(xb[(long long)r * C + c]) : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    int c = c0 + j, r = r0 + threadIdx.x;
    if (r < R && c < C) yb[(long long)c * R + r] = from_f<T>(tile[threadIdx.x][j]);
  }
}

}  // namespace c2d

using namespace c2d;

extern "C" {

int c2d_softmax_rows(const void* x, void* y, int M, int N, float scale, int dtype, void* stream) {
  C2D_REQUIRE(x && y && M > 0 && N > 0, "softmax_rows: bad args");
  int grid = ceil_div(M, 4);
  if (dtype == C2D_F32) softmax_rows_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>((const float*)x, (float*)y, M, N, scale);
  else if (dtype == C2D_BF16) softmax_rows_kernel<bf16><<<grid, 128, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, M, N, scale);
  else { set_error("softmax_rows: bad dtype"); return C2D_ERR_ARG; }
  return check_launch("softmax_rows");
}

int c2d_transpose(const void* x, void* y, int batch, int R, int C, int dtype, void* stream) {
  C2D_REQUIRE(x && y && batch > 0 && R > 0 && C > 0, "transpose: bad args");
  dim3 grid(ceil_div(C, 32), ceil_div(R, 32), batch), blk(32, 8);
  if (dtype == C2D_F32) transpose_kernel<float><<<grid, blk, 0, (cudaStream_t)stream>>>((const float*)x, (float*)y, R, C);
  else if (dtype == C2D_BF16) transpose_kernel<bf16><<<grid, blk, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, R, C);
  else { set_error("transpose: bad dtype"); return C2D_ERR_ARG; }
  return check_launch("transpose");
}

}  // extern "C"
