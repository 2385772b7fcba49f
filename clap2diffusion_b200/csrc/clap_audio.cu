// CLAP HTSAT audio tower: the kernels that are not plain GEMM / LayerNorm (those go through c2d_linear / c2d_layer_norm).
//   log-mel front end   framing + Hann window (reflect padding), |DFT|^2, dB + folded BatchNorm   (the DFT and the
//                       mel projection themselves are fp32 GEMMs against constant matrices)
//   mel -> patches      bicubic 1001 -> 1024 frame resampling, 4-chunk fold to 256 x 256, 4 x 4 patch gather
//   Swin window attention with relative-position bias, cyclic shift and shift mask folded into the addressing
//   2 x 2 patch-merge gather, token mean, L2 normalisation
// Reference behaviour: transformers/models/clap/modeling_clap.py (cited per kernel) as called from
// /root/reference/models/audio_encoder.py:164-174.
#include <stdlib.h>

#include "common.cuh"

namespace c2d {

// frames[(b * n_frames + f)][k] = wave[b][reflect(f * hop + k - n_fft / 2)] * window[k]   (audio_utils.spectrogram, center=True)
__global__ void stft_frames_kernel(const float* __restrict__ wave, const float* __restrict__ window, float* __restrict__ frames,
                                   int T, int n_fft, int hop, int n_frames, long long total) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int k = (int)(i % n_fft);
    const long long row = i / n_fft;
    const int f = (int)(row % n_frames);
    const long long b = row / n_frames;
    int t = f * hop + k - n_fft / 2;
    if (t < 0) t = -t;
    if (t >= T) t = 2 * (T - 1) - t;
    frames[i] = wave[b * T + t] * window[k];
  }
}

// Split-bf16 framing for the tensor-core DFT (bf16 product mode): x = hi + lo with hi = bf16(x), lo = bf16(x - hi); one row
// of frames3 is [hi | lo | hi] (3 * n_fft columns) and meets the constant matrix rows [HI | HI | LO], so a single bf16 GEMM
// with fp32 accumulation yields hi HI + lo HI + hi LO -- the fp32 product up to the dropped lo LO term (2^-18 relative).
// Plain bf16 operands would put a noise floor 54 dB under each frame's strongest bin, which the log-mel would show.
__global__ void stft_frames_split_kernel(const float* __restrict__ wave, const float* __restrict__ window, bf16* __restrict__ frames3,
                                         int T, int n_fft, int hop, int n_frames, long long total8) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int per_row = n_fft >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += stride) {
    const int k0 = (int)(i % per_row) * 8;
    const long long row = i / per_row;
    const int f = (int)(row % n_frames);
    const long long b = row / n_frames;
    float hi[8], lo[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      int t = f * hop + k0 + u - n_fft / 2;
      if (t < 0) t = -t;
      if (t >= T) t = 2 * (T - 1) - t;
      const float x = wave[b * T + t] * window[k0 + u];
      hi[u] = __bfloat162float(__float2bfloat16_rn(x));
      lo[u] = x - hi[u];
    }
    bf16* dst = frames3 + row * 3 * n_fft + k0;
    Vec8<bf16>::store(dst, hi);
    Vec8<bf16>::store(dst + n_fft, lo);
    Vec8<bf16>::store(dst + 2 * n_fft, hi);
  }
}

// dft row m = [re(0..nb) at column 0 | im(0..nb) at column im_off], row pitch ld  ->  out[m][j] = re^2 + im^2
template <typename T>
__global__ void power_spectrum_kernel(const T* __restrict__ dft, float* __restrict__ out, int nb, int ld, int im_off, long long total) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int j = (int)(i % nb);
    const long long m = i / nb;
    const float re = to_f<T>(dft[m * ld + j]), im = to_f<T>(dft[m * ld + im_off + j]);
    out[i] = fmaf(re, re, im * im);
  }
}

// y[m][f] = (10 log10(max(x, floor))) * a[f] + b[f]     (power_to_db + ClapAudioEncoder.batch_norm in eval mode)
__global__ void log_mel_affine_kernel(const float* __restrict__ x, const float* __restrict__ a, const float* __restrict__ b,
                                      float* __restrict__ y, int F, float floor_v, long long total) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int f = (int)(i % F);
    const float db = 10.0f * log10f(fmaxf(x[i], floor_v));
    y[i] = fmaf(db, a[f], b[f]);
  }
}

__device__ __forceinline__ float cubic1(float t, float A) { return ((A + 2.f) * t - (A + 3.f)) * t * t + 1.f; }
__device__ __forceinline__ float cubic2(float t, float A) { return ((A * t - 5.f * A) * t + 8.f * A) * t - 4.f * A; }

// reshape_mel2img (modeling_clap.py:777-811) + the im2col of ClapAudioPatchEmbed's 4 x 4 / stride-4 convolution:
// token p = ph * 64 + pw, element r * 4 + s  <-  image[4 ph + r][4 pw + s],  image[c * 64 + f][t'] = mel(t = c * 256 + t', f),
// mel(t, f) bicubically resampled (align_corners, A = -0.75) from the n_in input frames.
template <typename T>
__global__ void clap_patches_kernel(const float* __restrict__ mel, T* __restrict__ patches, int n_in, int F, long long total) {
  const float A = -0.75f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float ratio = (float)(n_in - 1) / (float)(4 * 256 - 1);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int e = (int)(i & 15), r = e >> 2, s = e & 3;
    const long long tok = i >> 4;
    const int p = (int)(tok & 4095);
    const long long b = tok >> 12;
    const int ph = p >> 6, pw = p & 63;
    const int h = 4 * ph + r, w = 4 * pw + s;
    const int c = h >> 6, f = h & 63;
    const int t = c * 256 + w;
    const float src = ratio * (float)t;
    const int x0 = (int)floorf(src);
    const float fr = src - (float)x0;
    const float w0 = cubic2(fr + 1.f, A), w1 = cubic1(fr, A), w2 = cubic1(1.f - fr, A), w3 = cubic2(2.f - fr, A);
    const float* col = mel + b * (long long)n_in * F + f;
    auto at = [&](int x) { x = x < 0 ? 0 : (x > n_in - 1 ? n_in - 1 : x); return col[(long long)x * F]; };
    const float v = w0 * at(x0 - 1) + w1 * at(x0) + w2 * at(x0 + 1) + w3 * at(x0 + 2);
    patches[i] = from_f<T>(v);
  }
}

// Swin window attention (ClapAudioSelfAttention :376-423 inside ClapAudioLayer :584-607).  One CTA = one (window, head),
// thread = query token.  The cyclic shift (torch.roll by -shift), the window partition and their inverses are pure
// index arithmetic here: window token (iy, ix) of window (wy, wx) lives at image position ((wy*8+iy+shift) % H, ...).
// The shift mask (-100 between tokens of different wrapped regions, :525-551) is recomputed from the coordinates.
template <typename T>
__global__ void __launch_bounds__(64)
window_attention_kernel(const T* __restrict__ qkv, const float* __restrict__ bias, T* __restrict__ out, int H, int W, int C,
                        int heads, int shift, float scale) {
  constexpr int WS = 8, NT = 64, DMAX = 32;
  const int d = C / heads;
  const int nw = W / WS;
  const int win = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int wy = win / nw, wx = win % nw;
  const int i = threadIdx.x, iy = i >> 3, ix = i & 7;
  const int ys = wy * WS + iy, xs = wx * WS + ix;                       // coordinates in the shifted image
  const int y = (ys + shift) % H, x = (xs + shift) % W;                 // source token
  const long long tok = (long long)b * H * W + (long long)y * W + x;
  __shared__ float sk[NT][DMAX + 1], sv[NT][DMAX + 1];
  __shared__ int sreg[NT];
  const T* row = qkv + tok * 3 * C + head * d;
  float q[DMAX];
#pragma unroll
  for (int j = 0; j < DMAX; ++j) {
    q[j] = j < d ? to_f<T>(row[j]) * scale : 0.f;
    sk[i][j] = j < d ? to_f<T>(row[C + j]) : 0.f;
    sv[i][j] = j < d ? to_f<T>(row[2 * C + j]) : 0.f;
  }
  int reg = 0;
  if (shift > 0) {
    const int rh = ys < H - WS ? 0 : (ys < H - shift ? 1 : 2);
    const int rw = xs < W - WS ? 0 : (xs < W - shift ? 1 : 2);
    reg = rh * 3 + rw;
  }
  sreg[i] = reg;
  __syncthreads();
  const float* brow = bias + ((long long)head * NT + i) * NT;
  float sc[NT];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < DMAX; ++k) a = fmaf(q[k], sk[j][k], a);
    a += brow[j];
    if (sreg[j] != reg) a += -100.0f;
    sc[j] = a;
    mx = fmaxf(mx, a);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < NT; ++j) { sc[j] = __expf(sc[j] - mx); sum += sc[j]; }
  const float inv = 1.f / sum;
  float o[DMAX];
#pragma unroll
  for (int k = 0; k < DMAX; ++k) o[k] = 0.f;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const float pj = sc[j] * inv;
#pragma unroll
    for (int k = 0; k < DMAX; ++k) o[k] = fmaf(pj, sv[j][k], o[k]);
  }
  T* orow = out + tok * C + head * d;
#pragma unroll
  for (int k = 0; k < DMAX; ++k)
    if (k < d) orow[k] = from_f<T>(o[k]);
}

// Same operation for head_dim D = 24 (every HTSAT stage: 96/4, 192/8, 384/16, 768/32) or 32, organised for the FMA pipe: the
// first kernel reads one shared-memory scalar per FMA (LDS-bound at ~1/4 of the FFMA rate).  Here K and V rows sit in shared
// memory as 16-byte chunks, chunk c of row j at slot (c + j) & 7 (conflict-free row-per-thread fill), and are read back as
// float4 broadcasts (one LDS.128 per four FMAs).  WA_HPB heads of one window per CTA (thread = query token).
constexpr int WA_HPB = 2;
template <typename T, int D>
__device__ __forceinline__ void wa_load(const T* p, float* f) {
  if constexpr (sizeof(T) == 4) {
#pragma unroll
    for (int c = 0; c < D / 4; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(p + 4 * c);
      f[4 * c] = v.x; f[4 * c + 1] = v.y; f[4 * c + 2] = v.z; f[4 * c + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int c = 0; c < D / 8; ++c) Vec8<T>::load(p + 8 * c, *reinterpret_cast<float(*)[8]>(f + 8 * c));
  }
}

template <typename T, int D>
__global__ void __launch_bounds__(64 * WA_HPB)
window_attention_d_kernel(const T* __restrict__ qkv, const float* __restrict__ bias, T* __restrict__ out, int H, int W, int C,
                          int heads, int shift, float scale) {
  constexpr int WS = 8, NT = 64, NC = D / 4;
  static_assert(D % 8 == 0 && D <= 32, "window attention: head dim");
  const int nw = W / WS;
  const int win = blockIdx.x, hl = threadIdx.x >> 6, head = blockIdx.y * WA_HPB + hl, b = blockIdx.z;
  const int wy = win / nw, wx = win % nw;
  const int i = threadIdx.x & 63, iy = i >> 3, ix = i & 7;
  const int ys = wy * WS + iy, xs = wx * WS + ix;                       // coordinates in the shifted image
  const int y = (ys + shift) % H, x = (xs + shift) % W;                 // source token
  const long long tok = (long long)b * H * W + (long long)y * W + x;
  __shared__ float4 sk[WA_HPB][NT][8], sv[WA_HPB][NT][8];
  __shared__ int sreg[NT];
  const T* row = qkv + tok * 3 * C + head * D;
  float q[D];
  {
    float t[D];
    wa_load<T, D>(row, q);
#pragma unroll
    for (int j = 0; j < D; ++j) q[j] *= scale;
    wa_load<T, D>(row + C, t);
#pragma unroll
    for (int c = 0; c < NC; ++c) sk[hl][i][(c + i) & 7] = make_float4(t[4 * c], t[4 * c + 1], t[4 * c + 2], t[4 * c + 3]);
    wa_load<T, D>(row + 2 * C, t);
#pragma unroll
    for (int c = 0; c < NC; ++c) sv[hl][i][(c + i) & 7] = make_float4(t[4 * c], t[4 * c + 1], t[4 * c + 2], t[4 * c + 3]);
  }
  int reg = 0;
  if (shift > 0) {
    const int rh = ys < H - WS ? 0 : (ys < H - shift ? 1 : 2);
    const int rw = xs < W - WS ? 0 : (xs < W - shift ? 1 : 2);
    reg = rh * 3 + rw;
  }
  if (hl == 0) sreg[i] = reg;
  __syncthreads();
  const float4* brow = reinterpret_cast<const float4*>(bias + ((long long)head * NT + i) * NT);
  float sc[NT];
  float mx = -INFINITY;
#pragma unroll
  for (int j4 = 0; j4 < NT / 4; ++j4) {
    const float4 bv = __ldg(brow + j4);
    const float bj[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j4 * 4 + u;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int c = 0; c < NC; c += 2) {
        const float4 k0 = sk[hl][j][(c + j) & 7], k1 = sk[hl][j][(c + 1 + j) & 7];
        a0 = fmaf(q[4 * c], k0.x, a0); a0 = fmaf(q[4 * c + 1], k0.y, a0); a0 = fmaf(q[4 * c + 2], k0.z, a0); a0 = fmaf(q[4 * c + 3], k0.w, a0);
        a1 = fmaf(q[4 * c + 4], k1.x, a1); a1 = fmaf(q[4 * c + 5], k1.y, a1); a1 = fmaf(q[4 * c + 6], k1.z, a1); a1 = fmaf(q[4 * c + 7], k1.w, a1);
      }
      float a = (a0 + a1) + bj[u];
      if (sreg[j] != reg) a += -100.0f;
      sc[j] = a;
      mx = fmaxf(mx, a);
    }
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < NT; ++j) { sc[j] = __expf(sc[j] - mx); sum += sc[j]; }
  const float inv = 1.f / sum;
  float o[D];
#pragma unroll
  for (int k = 0; k < D; ++k) o[k] = 0.f;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const float pj = sc[j];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float4 v = sv[hl][j][(c + j) & 7];
      o[4 * c] = fmaf(pj, v.x, o[4 * c]); o[4 * c + 1] = fmaf(pj, v.y, o[4 * c + 1]);
      o[4 * c + 2] = fmaf(pj, v.z, o[4 * c + 2]); o[4 * c + 3] = fmaf(pj, v.w, o[4 * c + 3]);
    }
  }
#pragma unroll
  for (int k = 0; k < D; ++k) o[k] *= inv;
  T* orow = out + tok * C + head * D;
#pragma unroll
  for (int c = 0; c < D / 8; ++c) Vec8<T>::store(orow + 8 * c, *reinterpret_cast<const float(*)[8]>(o + 8 * c));
}

// ClapAudioPatchMerging gather (:712-727): out[b][y2 * W/2 + x2] = cat(x[2y2][2x2], x[2y2+1][2x2], x[2y2][2x2+1], x[2y2+1][2x2+1])
template <typename T>
__global__ void patch_merge_kernel(const T* __restrict__ x, T* __restrict__ out, int H, int W, int C, long long total) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int cv = C >> 3;                                                 // 8-element vectors per channel row
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int v = (int)(i % cv);
    long long r = i / cv;
    const int part = (int)(r & 3);
    r >>= 2;
    const int x2 = (int)(r % (W / 2));
    r /= (W / 2);
    const int y2 = (int)(r % (H / 2));
    const long long b = r / (H / 2);
    const int yy = 2 * y2 + (part & 1), xx = 2 * x2 + (part >> 1);
    float f[8];
    Vec8<T>::load(x + ((b * H + yy) * W + xx) * C + v * 8, f);
    Vec8<T>::store(out + ((b * (H / 2) + y2) * (W / 2) + x2) * 4LL * C + (long long)part * C + v * 8, f);
  }
}

// out[b][c] = mean_n x[b][n][c]   (the pooling of ClapAudioEncoder.forward :896-909 is a global mean over the 64 tokens)
template <typename T>
__global__ void token_mean_kernel(const T* __restrict__ x, float* __restrict__ out, int N, int C) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int n = 0; n < N; ++n) s += to_f<T>(x[((long long)b * N + n) * C + c]);
  out[(long long)b * C + c] = s / (float)N;
}

// F.normalize(x, dim=-1): x / max(||x||_2, eps); one warp per row
__global__ void l2_normalize_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int D, float eps) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= B) return;
  float s = 0.f;
  for (int j = lane; j < D; j += 32) { const float v = x[(long long)row * D + j]; s = fmaf(v, v, s); }
  s = warp_sum(s);
  const float inv = 1.f / fmaxf(sqrtf(s), eps);
  for (int j = lane; j < D; j += 32) y[(long long)row * D + j] = x[(long long)row * D + j] * inv;
}

static inline int ew_grid2(long long n, int block) {
  long long g = (n + block - 1) / block;
  const long long cap = (long long)num_sms() * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace c2d

using namespace c2d;

#define CLAP_DISPATCH_T(dtype, ...)                              \
  if ((dtype) == C2D_F32) { using T = float; __VA_ARGS__ }       \
  else if ((dtype) == C2D_BF16) { using T = bf16; __VA_ARGS__ }  \
  else { set_error("bad dtype %d", (int)(dtype)); return C2D_ERR_ARG; }

namespace c2d {
// clap_attn_tc.cu: tcgen05 kernel (bf16, head dim 24 | 32)
bool window_attention_tc_supported(const void* qkv, const float* bias, const void* out, int C, int heads);
int window_attention_tc(const void* qkv, const float* bias, void* out, int B, int H, int W, int C, int heads, int shift, float scale,
                        cudaStream_t s);
}  // namespace c2d

extern "C" {

int c2d_stft_frames(const float* wave, const float* window, float* frames, int B, int T, int n_fft, int hop, int n_frames,
                    void* stream) {
  C2D_REQUIRE(wave && window && frames && B > 0 && T > n_fft / 2 && n_fft > 0 && hop > 0 && n_frames > 0, "stft_frames: bad args");
  C2D_REQUIRE((n_frames - 1) * hop + n_fft / 2 < 2 * T - 1, "stft_frames: frames run past the reflect padding");
  const long long total = (long long)B * n_frames * n_fft;
  stft_frames_kernel<<<ew_grid2(total, 256), 256, 0, (cudaStream_t)stream>>>(wave, window, frames, T, n_fft, hop, n_frames, total);
  return check_launch("stft_frames");
}

int c2d_stft_frames_split(const float* wave, const float* window, void* frames3, int B, int T, int n_fft, int hop, int n_frames,
                          void* stream) {
  C2D_REQUIRE(wave && window && frames3 && B > 0 && T > n_fft / 2 && n_fft > 0 && n_fft % 8 == 0 && hop > 0 && n_frames > 0,
              "stft_frames_split: bad args");
  C2D_REQUIRE((n_frames - 1) * hop + n_fft / 2 < 2 * T - 1, "stft_frames_split: frames run past the reflect padding");
  C2D_REQUIRE((reinterpret_cast<uintptr_t>(frames3) & 15) == 0, "stft_frames_split: output must be 16-byte aligned");
  const long long total8 = (long long)B * n_frames * (n_fft / 8);
  stft_frames_split_kernel<<<ew_grid2(total8, 256), 256, 0, (cudaStream_t)stream>>>(wave, window, (bf16*)frames3, T, n_fft, hop,
                                                                                   n_frames, total8);
  return check_launch("stft_frames_split");
}

int c2d_power_spectrum(const void* dft, float* out, long long M, int nb, int ld, int im_off, int dtype, void* stream) {
  C2D_REQUIRE(dft && out && M > 0 && nb > 0 && im_off >= nb && ld >= im_off + nb, "power_spectrum: bad args");
  const long long total = M * nb;
  CLAP_DISPATCH_T(dtype, power_spectrum_kernel<T><<<ew_grid2(total, 256), 256, 0, (cudaStream_t)stream>>>((const T*)dft, out, nb, ld,
                                                                                                        im_off, total);)
  return check_launch("power_spectrum");
}

int c2d_log_mel_affine(const float* x, const float* a, const float* b, float* y, long long M, int F, float floor_value,
                       void* stream) {
  C2D_REQUIRE(x && a && b && y && M > 0 && F > 0 && floor_value > 0.f, "log_mel_affine: bad args");
  const long long total = M * F;
  log_mel_affine_kernel<<<ew_grid2(total, 256), 256, 0, (cudaStream_t)stream>>>(x, a, b, y, F, floor_value, total);
  return check_launch("log_mel_affine");
}

int c2d_clap_patches(const float* mel, void* patches, int B, int n_frames, int n_mel, int dtype, void* stream) {
  C2D_REQUIRE(mel && patches && B > 0 && n_frames >= 4 && n_mel == 64, "clap_patches: bad args (64 mel bins, >= 4 frames)");
  const long long total = (long long)B * 4096 * 16;
  CLAP_DISPATCH_T(dtype, clap_patches_kernel<T><<<ew_grid2(total, 256), 256, 0, (cudaStream_t)stream>>>(mel, (T*)patches, n_frames,
                                                                                                      n_mel, total);)
  return check_launch("clap_patches");
}

int c2d_window_attention(const void* qkv, const float* bias, void* out, int B, int H, int W, int C, int heads, int shift,
                         float scale, int dtype, void* stream) {
  C2D_REQUIRE(qkv && bias && out && B > 0 && H > 0 && W > 0 && heads > 0, "window_attention: bad args");
  C2D_REQUIRE(H % 8 == 0 && W % 8 == 0, "window_attention: H=%d, W=%d must be multiples of the 8 x 8 window", H, W);
  C2D_REQUIRE(C % heads == 0 && C / heads <= 32, "window_attention: head_dim %d > 32", C / heads);
  C2D_REQUIRE(shift >= 0 && shift < 8, "window_attention: bad shift %d", shift);
  static int fast = -1;                 // C2D_WINATTN=0: first-generation kernel, 1: float4-broadcast FFMA kernel (A/B runs)
  if (fast < 0) {
    const char* e = getenv("C2D_WINATTN");
    fast = !e ? 2 : (e[0] == '0' ? 0 : (e[0] == '1' ? 1 : 2));
  }
  if (fast == 2 && dtype == C2D_BF16 && window_attention_tc_supported(qkv, bias, out, C, heads))
    return window_attention_tc(qkv, bias, out, B, H, W, C, heads, shift, scale, (cudaStream_t)stream);
  const int d = C / heads;
  if (fast && (d == 24 || d == 32) && heads % WA_HPB == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0) {
    dim3 gridd((H / 8) * (W / 8), heads / WA_HPB, B);
    if (d == 24) {
      CLAP_DISPATCH_T(dtype, window_attention_d_kernel<T, 24><<<gridd, 64 * WA_HPB, 0, (cudaStream_t)stream>>>(
                                 (const T*)qkv, bias, (T*)out, H, W, C, heads, shift, scale);)
    } else {
      CLAP_DISPATCH_T(dtype, window_attention_d_kernel<T, 32><<<gridd, 64 * WA_HPB, 0, (cudaStream_t)stream>>>(
                                 (const T*)qkv, bias, (T*)out, H, W, C, heads, shift, scale);)
    }
    return check_launch("window_attention");
  }
  dim3 grid((H / 8) * (W / 8), heads, B);
  CLAP_DISPATCH_T(dtype, window_attention_kernel<T><<<grid, 64, 0, (cudaStream_t)stream>>>((const T*)qkv, bias, (T*)out, H, W, C,
                                                                                       heads, shift, scale);)
  return check_launch("window_attention");
}

int c2d_patch_merge(const void* x, void* out, int B, int H, int W, int C, int dtype, void* stream) {
  C2D_REQUIRE(x && out && B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "patch_merge: bad args");
  const long long total = (long long)B * (H / 2) * (W / 2) * 4 * (C / 8);
  CLAP_DISPATCH_T(dtype, patch_merge_kernel<T><<<ew_grid2(total, 256), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)out, H, W, C,
                                                                                                     total);)
  return check_launch("patch_merge");
}

int c2d_token_mean(const void* x, float* out, int B, int N, int C, int dtype, void* stream) {
  C2D_REQUIRE(x && out && B > 0 && N > 0 && C > 0, "token_mean: bad args");
  dim3 grid(ceil_div(C, 128), B);
  CLAP_DISPATCH_T(dtype, token_mean_kernel<T><<<grid, 128, 0, (cudaStream_t)stream>>>((const T*)x, out, N, C);)
  return check_launch("token_mean");
}

int c2d_l2_normalize(const float* x, float* y, int B, int D, float eps, void* stream) {
  C2D_REQUIRE(x && y && B > 0 && D > 0, "l2_normalize: bad args");
  l2_normalize_kernel<<<ceil_div(B, 4), 128, 0, (cudaStream_t)stream>>>(x, y, B, D, eps);
  return check_launch("l2_normalize");
}

}  // extern "C"
