// Small latency-bound kernels of the audio-conditioning side (once per image, <= 77 tokens):
// tiny-attention, broadcast add, soft level assignment, level->UNet routing, Norm-60, legacy 5-3-2 combine.
// Reference semantics: models/hierarchical_audio_v4.py:154-182 (compute_assignments), :325-369 (router),
// :849-864 (legacy weights + concat), scripts/inference.py:92-99 (Norm-60).
#include "common.cuh"

namespace c2d {

// ---- tiny attention: one warp per (b, h, query); any head_dim, Nkv <= 128 --------------------------
constexpr int SA_MAXKV = 128;

template <typename T>
__global__ void attn_small_kernel(const AttnParams p, int B) {
  int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int total = B * p.heads * p.Nq;
  if (gw >= total) return;
  int qi = gw % p.Nq;
  int h = (gw / p.Nq) % p.heads;
  int b = gw / (p.Nq * p.heads);
  const int d = p.d;
  const T* q = reinterpret_cast<const T*>(p.q) + (long long)b * p.bsq + (long long)qi * p.ldq + (long long)h * d;
  const T* kb = reinterpret_cast<const T*>(p.k) + (long long)b * p.bsk + (long long)h * d;
  const T* vb = reinterpret_cast<const T*>(p.v) + (long long)b * p.bsv + (long long)h * d;
  const uint8_t* mk = p.mask ? p.mask + (long long)b * p.Nkv : nullptr;
  float s[SA_MAXKV / 32];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < SA_MAXKV / 32; ++i) {
    int key = lane + 32 * i;
    float acc = -INFINITY;
    if (key < p.Nkv) {
      const T* kr = kb + (long long)key * p.ldk;
      float dot = 0.f;
      for (int j = 0; j < d; ++j) dot = fmaf(to_f<T>(q[j]), to_f<T>(kr[j]), dot);
      acc = dot * p.scale;
      if (mk && !mk[key]) acc = -3.4028234664e38f;
    }
    s[i] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < SA_MAXKV / 32; ++i) {
    float e = (lane + 32 * i < p.Nkv) ? expf(s[i] - mx) : 0.f;
    s[i] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  float inv = 1.f / sum;
  T* o = reinterpret_cast<T*>(p.o) + (long long)b * p.bso + (long long)qi * p.ldo + (long long)h * d;
  for (int j0 = 0; j0 < d; j0 += 32) {
    int j = j0 + lane;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < SA_MAXKV / 32; ++i) {
      if (32 * i >= p.Nkv) break;
      for (int l = 0; l < 32; ++l) {
        int key = l + 32 * i;
        float pw = __shfl_sync(0xffffffffu, s[i], l);
        if (key < p.Nkv && j < d) acc = fmaf(pw, to_f<T>(vb[(long long)key * p.ldv + j]), acc);
      }
    }
    if (j < d) o[j] = from_f<T>(acc * inv);
  }
}

int attention_small(const AttnParams& p, int B, int dtype, cudaStream_t s) {
  if (p.Nkv > SA_MAXKV) {
    set_error("attention_small: Nkv=%d > %d", p.Nkv, SA_MAXKV);
    return C2D_ERR_UNSUPPORTED;
  }
  long long warps = (long long)B * p.heads * p.Nq;
  int threads = 128;
  int grid = (int)((warps * 32 + threads - 1) / threads);
  if (dtype == C2D_F32) attn_small_kernel<float><<<grid, threads, 0, s>>>(p, B);
  else attn_small_kernel<bf16><<<grid, threads, 0, s>>>(p, B);
  return check_launch("attn_small");
}

// ---- y[b,k,:] = a[...] + b[...] with per-operand broadcast mode (0 full, 1 over tokens, 2 over batch) ----
template <typename T>
__global__ void bcast_add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ y, int B, int K, int D,
                                 int am, int bm) {
  long long n = (long long)B * K * D;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int dd = (int)(i % D);
    int k = (int)((i / D) % K);
    int bb = (int)(i / ((long long)D * K));
    long long ia = am == 0 ? i : (am == 1 ? (long long)bb * D + dd : (long long)k * D + dd);
    long long ib = bm == 0 ? i : (bm == 1 ? (long long)bb * D + dd : (long long)k * D + dd);
    y[i] = from_f<T>(to_f<T>(a[ia]) + to_f<T>(b[ib]));
  }
}

// ---- soft level assignment: one warp per token ---------------------------------------------------
//  assign[b,k,:] = softmax((10 * cos(tok, anchor_l) + W2 gelu(W1 tok + b1) + b2) / T)
template <typename T>
__global__ void hier_assign_kernel(const T* __restrict__ tok, const T* __restrict__ anchors, const T* __restrict__ w1,
                                   const float* __restrict__ b1, const T* __restrict__ w2, const float* __restrict__ b2,
                                   const float* __restrict__ temperature, float* __restrict__ assign, int rows, int D,
                                   int L, int Hg) {
  int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= rows) return;
  const T* t = tok + (long long)gw * D;
  float tn = 0.f;
  for (int j = lane; j < D; j += 32) { float v = to_f<T>(t[j]); tn += v * v; }
  tn = fmaxf(sqrtf(warp_sum(tn)), 1e-12f);                 // F.normalize eps
  float hid[16];                                            // Hg <= 16
  for (int g = 0; g < Hg; ++g) {
    float acc = 0.f;
    for (int j = lane; j < D; j += 32) acc += to_f<T>(w1[(long long)g * D + j]) * to_f<T>(t[j]);
    hid[g] = gelu_erf(warp_sum(acc) + b1[g]);
  }
  float logit[8];                                           // L <= 8
  float mx = -INFINITY;
  const float T_ = temperature[0];
  for (int l = 0; l < L; ++l) {
    const T* an = anchors + (long long)l * D;
    float dot = 0.f, nn = 0.f;
    for (int j = lane; j < D; j += 32) { float a = to_f<T>(an[j]); dot += a * to_f<T>(t[j]); nn += a * a; }
    dot = warp_sum(dot);
    nn = fmaxf(sqrtf(warp_sum(nn)), 1e-12f);
    float gate = b2[l];
    for (int g = 0; g < Hg; ++g) gate += to_f<T>(w2[l * Hg + g]) * hid[g];
    logit[l] = (10.f * dot / (tn * nn) + gate) / T_;
    mx = fmaxf(mx, logit[l]);
  }
  float sum = 0.f;
  for (int l = 0; l < L; ++l) { logit[l] = expf(logit[l] - mx); sum += logit[l]; }
  if (lane == 0)
    for (int l = 0; l < L; ++l) assign[(long long)gw * L + l] = logit[l] / sum;
}

// ---- level -> UNet routing (L = 3 levels, 3 scales) ----------------------------------------------
//  a = assign * hw[b]; a /= (sum_l a + 1e-8); r = a @ softmax(routing, dim=1);
//  routed[s][b,k,:] = tok10[b,k,:] * r[s] * sigmoid(gate_s)
template <typename T>
__global__ void hier_route_kernel(const T* __restrict__ tok10, const float* __restrict__ assign,
                                  const float* __restrict__ hw, const float* __restrict__ routing,
                                  const float* __restrict__ gates, T* __restrict__ r0, T* __restrict__ r1,
                                  T* __restrict__ r2, int K, int D) {
  int row = blockIdx.x;                 // b*K + k
  int b = row / K;
  __shared__ float r[3];
  if (threadIdx.x == 0) {
    float a[3], s = 0.f;
    for (int l = 0; l < 3; ++l) { a[l] = assign[(long long)row * 3 + l] * (hw ? hw[b * 3 + l] : 1.f); s += a[l]; }
    if (hw) for (int l = 0; l < 3; ++l) a[l] /= (s + 1e-8f);
    float rs[3] = {0.f, 0.f, 0.f};
    for (int l = 0; l < 3; ++l) {
      float m = fmaxf(routing[l * 3], fmaxf(routing[l * 3 + 1], routing[l * 3 + 2]));
      float e0 = expf(routing[l * 3] - m), e1 = expf(routing[l * 3 + 1] - m), e2 = expf(routing[l * 3 + 2] - m);
      float inv = 1.f / (e0 + e1 + e2);
      rs[0] += a[l] * e0 * inv; rs[1] += a[l] * e1 * inv; rs[2] += a[l] * e2 * inv;
    }
    for (int s_ = 0; s_ < 3; ++s_) r[s_] = rs[s_] * (1.f / (1.f + expf(-gates[s_])));
  }
  __syncthreads();
  const T* t = tok10 + (long long)row * D;
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    float v = to_f<T>(t[j]);
    r0[(long long)row * D + j] = from_f<T>(v * r[0]);
    r1[(long long)row * D + j] = from_f<T>(v * r[1]);
    r2[(long long)row * D + j] = from_f<T>(v * r[2]);
  }
}

// ---- Norm-60: y = x * target / mean(||x||_2 over tokens [and batch]) -----------------------------
template <typename T>
__global__ void norm_scale_kernel(const T* __restrict__ x, T* __restrict__ y, int rows_per_group, int D, float target) {
  __shared__ float s_part[32];
  __shared__ float s_scale;
  long long base = (long long)blockIdx.x * rows_per_group * D;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  float acc = 0.f;
  for (int r = warp; r < rows_per_group; r += nwarp) {
    float q = 0.f;
    for (int j = lane; j < D; j += 32) { float v = to_f<T>(x[base + (long long)r * D + j]); q += v * v; }
    q = warp_sum(q);
    acc += sqrtf(q);
  }
  if (lane == 0) s_part[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < nwarp; ++w) s += s_part[w];
    float mean = s / (float)rows_per_group;
    s_scale = mean > 0.f ? target / mean : 1.f;
  }
  __syncthreads();
  float sc = s_scale;
  for (long long i = threadIdx.x; i < (long long)rows_per_group * D; i += blockDim.x)
    y[base + i] = from_f<T>(to_f<T>(x[base + i]) * sc);
}

// ---- legacy rigid decomposition: cat(fg*w0, bg*w1, amb*w2), w = softmax(hierarchy_weights) -------
template <typename T>
__global__ void legacy_combine_kernel(const T* __restrict__ fg, const T* __restrict__ bg, const T* __restrict__ am,
                                      const float* __restrict__ hw, T* __restrict__ out, int B, int nf, int nb, int na,
                                      int D) {
  float m = fmaxf(hw[0], fmaxf(hw[1], hw[2]));
  float e0 = expf(hw[0] - m), e1 = expf(hw[1] - m), e2 = expf(hw[2] - m);
  float inv = 1.f / (e0 + e1 + e2);
  float w0 = e0 * inv, w1 = e1 * inv, w2 = e2 * inv;
  int K = nf + nb + na;
  long long n = (long long)B * K * D;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int dd = (int)(i % D);
    int k = (int)((i / D) % K);
    int b = (int)(i / ((long long)D * K));
    float v;
    if (k < nf) v = to_f<T>(fg[((long long)b * nf + k) * D + dd]) * w0;
    else if (k < nf + nb) v = to_f<T>(bg[((long long)b * nb + (k - nf)) * D + dd]) * w1;
    else v = to_f<T>(am[((long long)b * na + (k - nf - nb)) * D + dd]) * w2;
    out[i] = from_f<T>(v);
  }
}

}  // namespace c2d

using namespace c2d;

#define DISPATCH_T(dtype, ...)                                   \
  if ((dtype) == C2D_F32) { typedef float T; __VA_ARGS__ }       \
  else if ((dtype) == C2D_BF16) { typedef bf16 T; __VA_ARGS__ }  \
  else { set_error("bad dtype %d", (int)(dtype)); return C2D_ERR_ARG; }

extern "C" {

int c2d_bcast_add(const void* a, const void* b, void* y, int B, int K, int D, int a_mode, int b_mode, int dtype,
                  void* stream) {
  C2D_REQUIRE(a && b && y && B > 0 && K > 0 && D > 0, "bcast_add: bad args");
  C2D_REQUIRE(a_mode >= 0 && a_mode <= 2 && b_mode >= 0 && b_mode <= 2, "bcast_add: bad modes");
  long long n = (long long)B * K * D;
  int grid = (int)((n + 255) / 256);
  if (grid > num_sms() * 16) grid = num_sms() * 16;
  DISPATCH_T(dtype, bcast_add_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)a, (const T*)b, (T*)y, B, K, D, a_mode, b_mode);)
  return check_launch("bcast_add");
}

int c2d_hier_assign(const void* tokens, const void* anchors, const void* w1, const float* b1, const void* w2,
                    const float* b2, const float* temperature, float* assign, int rows, int D, int L, int Hg, int dtype,
                    void* stream) {
  C2D_REQUIRE(tokens && anchors && w1 && b1 && w2 && b2 && temperature && assign, "hier_assign: null pointer");
  C2D_REQUIRE(rows > 0 && D > 0 && L > 0 && L <= 8 && Hg > 0 && Hg <= 16, "hier_assign: bad dims (L<=8, Hg<=16)");
  int grid = ceil_div(rows * 32, 128);
  DISPATCH_T(dtype, hier_assign_kernel<T><<<grid, 128, 0, (cudaStream_t)stream>>>(
                        (const T*)tokens, (const T*)anchors, (const T*)w1, b1, (const T*)w2, b2, temperature, assign, rows, D, L, Hg);)
  return check_launch("hier_assign");
}

int c2d_hier_route(const void* tok10, const float* assign, const float* hw, const float* routing, const float* gates,
                   void* r_early, void* r_mid, void* r_late, int B, int K, int D, int dtype, void* stream) {
  C2D_REQUIRE(tok10 && assign && routing && gates && r_early && r_mid && r_late, "hier_route: null pointer");
  C2D_REQUIRE(B > 0 && K > 0 && D > 0, "hier_route: bad dims");
  DISPATCH_T(dtype, hier_route_kernel<T><<<B * K, 256, 0, (cudaStream_t)stream>>>(
                        (const T*)tok10, assign, hw, routing, gates, (T*)r_early, (T*)r_mid, (T*)r_late, K, D);)
  return check_launch("hier_route");
}

int c2d_norm_scale(const void* x, void* y, int B, int K, int D, float target, int per_sample, int dtype, void* stream) {
  C2D_REQUIRE(x && y && B > 0 && K > 0 && D > 0, "norm_scale: bad args");
  int groups = per_sample ? B : 1;
  int rows = per_sample ? K : B * K;
  DISPATCH_T(dtype, norm_scale_kernel<T><<<groups, 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, rows, D, target);)
  return check_launch("norm_scale");
}

int c2d_legacy_combine(const void* fg, const void* bg, const void* amb, const float* hierarchy_weights, void* out, int B,
                       int nf, int nb, int na, int D, int dtype, void* stream) {
  C2D_REQUIRE(fg && bg && amb && hierarchy_weights && out && B > 0 && D > 0, "legacy_combine: bad args");
  long long n = (long long)B * (nf + nb + na) * D;
  int grid = (int)((n + 255) / 256);
  if (grid > num_sms() * 16) grid = num_sms() * 16;
  DISPATCH_T(dtype, legacy_combine_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(
                        (const T*)fg, (const T*)bg, (const T*)amb, hierarchy_weights, (T*)out, B, nf, nb, na, D);)
  return check_launch("legacy_combine");
}

}  // extern "C"
