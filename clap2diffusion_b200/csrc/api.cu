// C-ABI dispatch for the dense contractions: picks the tcgen05 kernels (bf16) or the FFMA kernels (fp32
// parity mode and shapes the tensor-core kernels do not take).  No CPU path exists.
#include "common.cuh"

namespace c2d {

int linear_simt(const void*, const void*, const float*, const float*, int, const void*, void*, int, int, int, int, int,
                int, int, int, cudaStream_t);
int conv3x3_simt(const void*, const void*, const float*, const float*, const void*, void*, int, int, int, int, int, int,
                 int, int, cudaStream_t, int pad = 1);
bool linear_tc_supported(const void*, const void*, int, int, int, int);
int linear_tc(const void*, const void*, const float*, const float*, int, const void*, void*, int, int, int, int, int, int,
              int, bool, const GemmExtras*, cudaStream_t);
bool conv3x3_tc_supported(const void*, const void*, int, int, int, int, int, int, int);
int conv3x3_tc(const void*, const void*, const float*, const float*, const void*, void*, int, int, int, int, int, int,
               long long*, cudaStream_t, int pad = 1);
int attention_simt(const AttnParams&, int, int, cudaStream_t);
int attention_small(const AttnParams&, int, int, cudaStream_t);
bool attention_tc_supported(const AttnParams&, int B);
int attention_tc(const AttnParams&, int B, cudaStream_t);
bool xattn_tc_supported(int C, int heads, int Nq, int T, int T2);
long long xattn_packed_bytes(int C, int heads, int T, int T2);
int xattn_pack_kv(const void*, const void*, long long, long long, int, const void*, const void*, long long, long long, int, void*,
                  int, int, int, cudaStream_t);
int xattn_tc(const void*, long long, const void*, const float*, const long long*, const float*, float, const void*, int, int, float,
             void*, long long, int, int, int, int, float, cudaStream_t);

}  // namespace c2d

using namespace c2d;

extern "C" {

int c2d_linear(const void* x, const void* w, const float* bias, const float* rowvec, int rows_per_vec,
               const void* residual, void* y, int M, int N, int K, int ldx, int ldy, int ldr, int act, int dtype,
               int impl, void* stream) {
  C2D_REQUIRE(x && w && y, "linear: null pointer");
  C2D_REQUIRE(M > 0 && N > 0 && K > 0 && ldx >= K && ldy >= N, "linear: bad dims M=%d N=%d K=%d ldx=%d ldy=%d", M, N, K, ldx, ldy);
  C2D_REQUIRE(!residual || ldr >= N, "linear: bad residual stride %d", ldr);
  C2D_REQUIRE(act >= C2D_ACT_NONE && act <= C2D_ACT_RELU, "linear: bad act %d", act);
  C2D_REQUIRE(dtype == C2D_F32 || dtype == C2D_BF16, "linear: bad dtype %d", dtype);
  cudaStream_t s = (cudaStream_t)stream;
  bool tc_ok = dtype == C2D_BF16 && linear_tc_supported(x, w, M, N, K, ldx);
  if (impl == C2D_IMPL_TCGEN05) {
    C2D_REQUIRE(tc_ok, "linear: tcgen05 path needs bf16, K %% 8 == 0, ldx %% 8 == 0, aligned pointers");
    return linear_tc(x, w, bias, rowvec, rows_per_vec, residual, y, M, N, K, ldx, ldy, ldr, act, false, nullptr, s);
  }
  // tcgen05 from 64 rows up, and from 8 rows up when the layer is a weight-streaming one (>= 1 Mi weights: the audio
  // projector / decomposer MLPs, 512 -> 24576 at batch 8..256) -- the tile's idle rows cost nothing there, the bf16
  // weights are read exactly once; tiny layers with a handful of rows stay on the FFMA kernel (latency-bound either way)
  if (impl == C2D_IMPL_AUTO && tc_ok && (M >= 64 || (M >= 8 && (long long)N * K >= (1ll << 20))))
    return linear_tc(x, w, bias, rowvec, rows_per_vec, residual, y, M, N, K, ldx, ldy, ldr, act, false, nullptr, s);
  return linear_simt(x, w, bias, rowvec, rows_per_vec, residual, y, M, N, K, ldx, ldy, ldr, act, dtype, s);
}

int c2d_geglu_linear(const void* x, const void* w, const float* bias, void* y, int M, int F, int K, int packed, int dtype,
                     int impl, void* stream) {
  C2D_REQUIRE(x && w && y && M > 0 && F > 0 && K > 0, "geglu_linear: bad args");
  C2D_REQUIRE(packed, "geglu_linear: only the packed (c2d_pack_geglu) weight layout is fused; use c2d_linear + c2d_geglu otherwise");
  C2D_REQUIRE(dtype == C2D_BF16 && F % 64 == 0, "geglu_linear: fused path is bf16 with F %% 64 == 0");
  C2D_REQUIRE(linear_tc_supported(x, w, M, 2 * F, K, K), "geglu_linear: K %% 8 / alignment");
  (void)impl;
  return linear_tc(x, w, bias, nullptr, 1, nullptr, y, M, 2 * F, K, K, F, 0, C2D_ACT_NONE, true, nullptr, (cudaStream_t)stream);
}

int c2d_conv3x3(const void* x, const void* w, const float* bias, const float* rowvec, const void* residual, void* y,
                int B, int H, int W, int Cin, int Cout, int stride, int upsample2x, int dtype, int impl, void* stream) {
  C2D_REQUIRE(x && w && y, "conv3x3: null pointer");
  C2D_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "conv3x3: bad dims");
  C2D_REQUIRE(stride == 1 || stride == 2, "conv3x3: stride %d", stride);
  C2D_REQUIRE(!(upsample2x && stride != 1), "conv3x3: upsample2x requires stride 1");
  C2D_REQUIRE(dtype == C2D_F32 || dtype == C2D_BF16, "conv3x3: bad dtype %d", dtype);
  cudaStream_t s = (cudaStream_t)stream;
  bool tc_ok = dtype == C2D_BF16 && conv3x3_tc_supported(x, w, B, H, W, Cin, Cout, stride, upsample2x);
  if (impl == C2D_IMPL_TCGEN05) {
    C2D_REQUIRE(tc_ok, "conv3x3: tcgen05 path needs bf16, no fused upsample, pow2 output H/W, Cin %% 8 == 0");
    return conv3x3_tc(x, w, bias, rowvec, residual, y, B, H, W, Cin, Cout, stride, nullptr, s);
  }
  if (impl == C2D_IMPL_AUTO && tc_ok)
    return conv3x3_tc(x, w, bias, rowvec, residual, y, B, H, W, Cin, Cout, stride, nullptr, s);
  return conv3x3_simt(x, w, bias, rowvec, residual, y, B, H, W, Cin, Cout, stride, upsample2x, dtype, s);
}

int c2d_linear_ex(const void* x, const void* x2, int K1, int ldx2, const void* w, const float* bias, const float* rowvec,
                  int rows_per_vec, const void* residual, void* y, int M, int N, int K, int ldx, int ldy, int ldr, int act,
                  long long* chan_stats, int stats_rows, long long* row_stats_out, const long long* ln_row_stats,
                  const float* ln_colsum, float ln_eps, int dtype, void* stream) {
  C2D_REQUIRE(x && w && y, "linear_ex: null pointer");
  C2D_REQUIRE(M > 0 && N > 0 && K > 0 && ldy >= N, "linear_ex: bad dims M=%d N=%d K=%d ldy=%d", M, N, K, ldy);
  C2D_REQUIRE(!residual || ldr >= N, "linear_ex: bad residual stride %d", ldr);
  C2D_REQUIRE(act >= C2D_ACT_NONE && act <= C2D_ACT_RELU, "linear_ex: bad act %d", act);
  C2D_REQUIRE(dtype == C2D_BF16, "linear_ex: the K-concatenated / statistics-producing GEMM exists on the tcgen05 (bf16) path only");
  C2D_REQUIRE(ldx >= (x2 ? K1 : K) && (!x2 || ldx2 >= K - K1), "linear_ex: bad row strides");
  C2D_REQUIRE(linear_tc_supported(x, w, M, N, K, ldx), "linear_ex: K %% 8 / ldx %% 8 / alignment");
  if (chan_stats) C2D_REQUIRE(stats_rows > 0 && M % stats_rows == 0, "linear_ex: M=%d is not a multiple of stats_rows=%d", M, stats_rows);
  // the epilogue reduction needs whole warps (32 rows) inside one image; tiny planes take the stand-alone kernel
  const bool fused = chan_stats && stats_rows % 32 == 0;
  C2D_REQUIRE(!ln_row_stats || ln_colsum, "linear_ex: folded LayerNorm needs ln_colsum");
  GemmExtras ex = {x2, K1, ldx2, fused ? chan_stats : nullptr, stats_rows, row_stats_out, ln_row_stats, ln_colsum, ln_eps};
  int rc = linear_tc(x, w, bias, rowvec, rows_per_vec, residual, y, M, N, K, ldx, ldy, ldr, act, false, &ex, (cudaStream_t)stream);
  if (rc || !chan_stats || fused) return rc;
  C2D_REQUIRE(ldy == N, "linear_ex: channel statistics of a strided output need stats_rows %% 32 == 0");
  return c2d_channel_stats(y, chan_stats, M / stats_rows, stats_rows, N, dtype, stream);
}

int c2d_geglu_linear_ex(const void* x, const void* w, const float* bias, const long long* ln_row_stats, const float* ln_colsum,
                        float ln_eps, void* y, int M, int F, int K, int dtype, void* stream) {
  C2D_REQUIRE(x && w && y && M > 0 && F > 0 && K > 0, "geglu_linear_ex: bad args");
  C2D_REQUIRE(dtype == C2D_BF16 && F % 64 == 0, "geglu_linear_ex: bf16 with F %% 64 == 0 (packed weights, c2d_pack_geglu)");
  C2D_REQUIRE(linear_tc_supported(x, w, M, 2 * F, K, K), "geglu_linear_ex: K %% 8 / alignment");
  C2D_REQUIRE(!ln_row_stats || ln_colsum, "geglu_linear_ex: folded LayerNorm needs ln_colsum");
  GemmExtras ex = {nullptr, 0, 0, nullptr, 0, nullptr, ln_row_stats, ln_colsum, ln_eps};
  return linear_tc(x, w, bias, nullptr, 1, nullptr, y, M, 2 * F, K, K, F, 0, C2D_ACT_NONE, true, &ex, (cudaStream_t)stream);
}

int c2d_conv3x3_ex(const void* x, const void* w, const float* bias, const float* rowvec, const void* residual, void* y,
                   int B, int H, int W, int Cin, int Cout, int stride, long long* chan_stats, int dtype, void* stream) {
  C2D_REQUIRE(x && w && y, "conv3x3_ex: null pointer");
  C2D_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "conv3x3_ex: bad dims");
  C2D_REQUIRE(dtype == C2D_BF16, "conv3x3_ex: the statistics-producing convolution exists on the bf16 path only");
  C2D_REQUIRE(stride == 1 || (stride == 2 && H % 2 == 0 && W % 2 == 0), "conv3x3_ex: stride 1, or 2 with even H, W");
  const int HWo = (H / stride) * (W / stride);
  if (!conv3x3_tc_supported(x, w, B, H, W, Cin, Cout, stride, 0)) {
    // planes the tcgen05 tiler does not take (output width neither a multiple of 128 nor a power of two, e.g. 64 x 96
    // latents; Cin %% 8 != 0): the FFMA kernel + the stand-alone statistics pass -- slower, never silently wrong
    int rc = conv3x3_simt(x, w, bias, rowvec, residual, y, B, H, W, Cin, Cout, stride, 0, dtype, (cudaStream_t)stream, 1);
    if (rc || !chan_stats) return rc;
    return c2d_channel_stats(y, chan_stats, B, HWo, Cout, dtype, stream);
  }
  const bool fused = chan_stats && HWo % 32 == 0;
  int rc = conv3x3_tc(x, w, bias, rowvec, residual, y, B, H, W, Cin, Cout, stride, fused ? chan_stats : nullptr, (cudaStream_t)stream);
  if (rc || !chan_stats || fused) return rc;
  return c2d_channel_stats(y, chan_stats, B, HWo, Cout, dtype, stream);
}

int c2d_conv3x3_down(const void* x, const void* w, const float* bias, void* y, int B, int H, int W, int Cin, int Cout,
                     long long* chan_stats, int dtype, int impl, void* stream) {
  C2D_REQUIRE(x && w && y, "conv3x3_down: null pointer");
  C2D_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && H % 2 == 0 && W % 2 == 0, "conv3x3_down: bad dims (H, W must be even)");
  C2D_REQUIRE(dtype == C2D_F32 || dtype == C2D_BF16, "conv3x3_down: bad dtype %d", dtype);
  cudaStream_t s = (cudaStream_t)stream;
  const bool tc_ok = dtype == C2D_BF16 && conv3x3_tc_supported(x, w, B, H, W, Cin, Cout, 2, 0);
  if (impl == C2D_IMPL_TCGEN05) C2D_REQUIRE(tc_ok, "conv3x3_down: tcgen05 path needs bf16, pow2 output H/W, Cin %% 8 == 0");
  if (tc_ok && impl != C2D_IMPL_SIMT) {
    const int HWo = (H / 2) * (W / 2);
    const bool fused = chan_stats && HWo % 32 == 0;
    int rc = conv3x3_tc(x, w, bias, nullptr, nullptr, y, B, H, W, Cin, Cout, 2, fused ? chan_stats : nullptr, s, 0);
    if (rc || !chan_stats || fused) return rc;
    return c2d_channel_stats(y, chan_stats, B, HWo, Cout, dtype, stream);
  }
  int rc = conv3x3_simt(x, w, bias, nullptr, nullptr, y, B, H, W, Cin, Cout, 2, 0, dtype, s, 0);
  if (rc || !chan_stats) return rc;
  C2D_REQUIRE(dtype == C2D_BF16, "conv3x3_down: channel statistics exist on the bf16 path only");
  return c2d_channel_stats(y, chan_stats, B, (H / 2) * (W / 2), Cout, dtype, stream);
}

int c2d_attention(const void* q, const void* k, const void* v, void* o, int B, int heads, int Nq, int Nkv, int d,
                  long long ldq, long long ldk, long long ldv, long long ldo, long long bsq, long long bsk,
                  long long bsv, long long bso, float scale, const uint8_t* mask, int dtype, int impl, void* stream) {
  return c2d_attention_lse(q, k, v, o, B, heads, Nq, Nkv, d, ldq, ldk, ldv, ldo, bsq, bsk, bsv, bso, scale, mask, nullptr, nullptr,
                           dtype, impl, stream);
}

int c2d_attention_lse(const void* q, const void* k, const void* v, void* o, int B, int heads, int Nq, int Nkv, int d,
                      long long ldq, long long ldk, long long ldv, long long ldo, long long bsq, long long bsk,
                      long long bsv, long long bso, float scale, const uint8_t* mask, float* lse, int* lse_written, int dtype,
                      int impl, void* stream) {
  C2D_REQUIRE(q && k && v && o, "attention: null pointer");
  C2D_REQUIRE((lse == nullptr) == (lse_written == nullptr), "attention: lse and lse_written go together");
  if (lse_written) *lse_written = 0;
  C2D_REQUIRE(B > 0 && heads > 0 && Nq > 0 && Nkv > 0 && d > 0, "attention: bad dims");
  C2D_REQUIRE(dtype == C2D_F32 || dtype == C2D_BF16, "attention: bad dtype %d", dtype);
  AttnParams p = {q, k, v, o, Nq, Nkv, d, heads, ldq, ldk, ldv, ldo, bsq, bsk, bsv, bso, scale, mask, lse, lse_written};
  cudaStream_t s = (cudaStream_t)stream;
  // tiny problems / head dims beyond the flash kernels (AudioTokenGenerator single head d=768): warp-per-query kernel
  const bool vec_ok = d % 8 == 0 && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && bsq % 8 == 0 && bsk % 8 == 0 &&
                      bsv % 8 == 0 && (reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) |
                                       reinterpret_cast<uintptr_t>(v)) % 32 == 0;
  if (impl != C2D_IMPL_TCGEN05 && Nkv <= 128 && (d > 160 || !vec_ok)) return attention_small(p, B, dtype, s);
  C2D_REQUIRE(vec_ok, "attention: head_dim and strides must be multiples of 8 elements and pointers 32B-aligned "
                      "(or Nkv <= 128 for the small-attention kernel)");
  bool tc_ok = dtype == C2D_BF16 && attention_tc_supported(p, B);
  if (impl == C2D_IMPL_TCGEN05) {
    C2D_REQUIRE(tc_ok, "attention: tcgen05 path does not take this shape (d=%d Nq=%d Nkv=%d)", d, Nq, Nkv);
    return attention_tc(p, B, s);
  }
  if (impl == C2D_IMPL_AUTO && tc_ok) return attention_tc(p, B, s);
  return attention_simt(p, B, dtype, s);
}

int c2d_xattn_supported(int C, int heads, int Nq, int T, int T2, int dtype) {
  return dtype == C2D_BF16 && xattn_tc_supported(C, heads, Nq, T, T2) ? 1 : 0;
}

long long c2d_xattn_packed_bytes(int C, int heads, int T, int T2) { return xattn_packed_bytes(C, heads, T, T2); }

int c2d_xattn_pack_kv(const void* k, const void* v, long long ldkv, long long bskv, int T, const void* k2, const void* v2,
                      long long ldkv2, long long bskv2, int T2, void* packed, int B, int C, int heads, int dtype, void* stream) {
  C2D_REQUIRE(k && v && packed, "xattn_pack_kv: null pointer");
  C2D_REQUIRE(B > 0 && C > 0 && heads > 0 && T > 0 && T2 >= 0 && ldkv >= C, "xattn_pack_kv: bad dims");
  if (dtype != C2D_BF16) {
    set_error("xattn_pack_kv: the fused cross-attention kernel is bf16 / tcgen05 only");
    return C2D_ERR_UNSUPPORTED;
  }
  return xattn_pack_kv(k, v, ldkv, bskv, T, k2, v2, ldkv2, bskv2, T2, packed, B, C, heads, (cudaStream_t)stream);
}

int c2d_xattn_fwd(const void* x, long long ldx, const void* wq, const float* q_bias, const long long* ln_row_stats,
                  const float* ln_colsum, float ln_eps, const void* kv_packed, int T, int T2, float lambda2, void* o,
                  long long ldo, int B, int Nq, int C, int heads, float scale, int dtype, void* stream) {
  C2D_REQUIRE(x && wq && kv_packed && o, "xattn_fwd: null pointer");
  C2D_REQUIRE(B > 0 && Nq > 0 && C > 0 && heads > 0 && T > 0 && T2 >= 0, "xattn_fwd: bad dims");
  C2D_REQUIRE(ldx >= C && ldo >= C, "xattn_fwd: bad row strides");
  C2D_REQUIRE(!ln_row_stats || ln_colsum, "xattn_fwd: folded LayerNorm needs ln_colsum");
  if (dtype != C2D_BF16) {
    set_error("xattn_fwd: the fused kernel is bf16 / tcgen05 only; compose c2d_linear + c2d_attention in fp32 mode");
    return C2D_ERR_UNSUPPORTED;
  }
  return xattn_tc(x, ldx, wq, q_bias, ln_row_stats, ln_colsum, ln_eps, kv_packed, T, T2, lambda2, o, ldo, B, Nq, C, heads, scale,
                  (cudaStream_t)stream);
}

}  // extern "C"
