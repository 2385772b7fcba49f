/* c2d.h -- C ABI of libc2d.so: the B200 (sm_100a) compute library behind the CLAP2Diffusion
 * inference hot path.
 *
 * Boundary contract: plain `extern "C"` functions, raw DEVICE pointers + explicit
 * sizes, a dtype enum, the caller's `cudaStream_t` (passed as void*), `int` status (0 = ok).  No torch
 * types, no exceptions; the failure message is in c2d_last_error() (thread-local).  The caller
 * (PyTorch) owns every buffer including workspaces; calls are asynchronous on the given stream and
 * capturable into a CUDA graph.  There is NO CPU fallback: without a CUDA device every compute entry
 * point returns C2D_ERR_CUDA.
 *
 * Activation layout everywhere: channels-last tokens  x[b][n][c]  (n = y*W + x), i.e. NHWC.
 * `dtype` is the storage type of activations AND matmul weights (C2D_F32: fp32 FFMA path used for
 * the <=1e-4 parity mode; C2D_BF16: bf16 storage, fp32 accumulation, tcgen05 tensor cores).
 * Biases, norm gains/shifts, statistics, latents and scheduler state are always fp32.
 *
 * Each entry point names the reference code it replaces (paths relative to /root/reference; the
 * SD-1.5 UNet itself is third-party diffusers==0.23.1, requirements.txt:7, restated in oracle/sd15.py).
 */
#ifndef C2D_H_
#define C2D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define C2D_ABI_VERSION 7   /* 7: c2d_adamw_step_sched, c2d_attention_lse, have_lse argument of c2d_attention_bwd, x_stats argument of c2d_group_norm_bwd; 6: c2d_stft_frames_split, strided / typed c2d_power_spectrum, workspace argument of c2d_group_norm_bwd; 5: + backward / optimiser entry points of the stage-3 fine-tune step (c2d_*_bwd, c2d_adamw_step ...);
                             * 4: + c2d_destroy, c2d_set_workspace, c2d_splitk_workspace_bytes (the library owns no device
                             *    memory); 3: + c2d_xattn_* (fused cross-attention site), c2d_conv3x3_down */

enum { C2D_OK = 0, C2D_ERR_ARG = 1, C2D_ERR_CUDA = 2, C2D_ERR_UNSUPPORTED = 3 };
enum { C2D_F32 = 0, C2D_BF16 = 1 };
enum { C2D_ACT_NONE = 0, C2D_ACT_GELU = 1, C2D_ACT_SILU = 2, C2D_ACT_RELU = 3 };
/* kernel family selector for the dense contractions */
enum { C2D_IMPL_AUTO = 0, C2D_IMPL_SIMT = 1, C2D_IMPL_TCGEN05 = 2 };
/* AudioAttnProcessor injection modes (models/audio_attention_processor.py:24) */
enum { C2D_AUDIO_NONE = 0, C2D_AUDIO_ADD = 1, C2D_AUDIO_CONCAT = 2 };

int c2d_abi_version(void);
const char* c2d_last_error(void);
/* Per-device one-time setup (checks the device is sm_100, resolves the tensor-map driver entry point).  Does NOT
 * change the calling thread's current device and allocates NO device memory: the library's only state is a small
 * per-device context ("caller owns every buffer incl. workspace ... c2d_init / c2d_destroy"). */
int c2d_init(int device);
/* Forgets the per-device context (the registered workspace pointer and cached device facts).  Nothing to free: the
 * workspace belongs to the caller, tensor maps are built per launch and passed by value. */
int c2d_destroy(int device);
/* Split-K workspace of the small-plane 3x3 convolutions (fp32 partial slices; see c2d_conv3x3_ex).  The CALLER
 * allocates `c2d_splitk_workspace_bytes()` bytes (256-byte aligned) of device memory and registers them; without a
 * workspace those convolutions run un-split (same results, lower occupancy at 8x8 latents).  One workspace per
 * device: launches that may use it concurrently on DIFFERENT streams of one device need the caller to give each
 * stream its own (call c2d_set_workspace before enqueueing on the other stream) -- everything else in the library
 * is re-entrant across streams.  workspace = NULL removes it. */
long long c2d_splitk_workspace_bytes(void);
int c2d_set_workspace(int device, void* workspace, long long bytes);
/* Number of kernels this library has launched since load (claim for bench.py's gpu_launches). */
unsigned long long c2d_launch_count(void);
/* Name of the kernel family the calling thread launched last (e.g. "conv3x3_tc", "linear_simt"): lets
 * the benchmark attribute CUDA-event timings to the kernel C2D_IMPL_AUTO actually picked. */
const char* c2d_last_kernel(void);

/* ---- dense layer:  y[M,N] = act(x[M,K] . w[N,K]^T + bias[N] + rowvec[m / rows_per_vec][N]) + residual[M,N]
 * Replaces nn.Linear / 1x1 conv calls: attn.to_q/to_k/to_v/to_out[0]
 * (models/audio_attention_processor.py:115,120-121,134), the projector / decomposer MLPs
 * (models/audio_adapter_v4.py:43-48, models/hierarchical_audio_v4.py:110-116) and diffusers' FF / proj_in /
 * proj_out / conv_shortcut / time MLP.  ldx/ldy/ldr are row strides in elements.  bias, rowvec,
 * residual may be NULL.  rowvec is fp32 [ceil(M/rows_per_vec)][N] (the resnet time-embedding add). */
int c2d_linear(const void* x, const void* w, const float* bias, const float* rowvec, int rows_per_vec,
               const void* residual, void* y, int M, int N, int K, int ldx, int ldy, int ldr, int act,
               int dtype, int impl, void* stream);

/* ---- GEGLU feed-forward input projection fused with the gate:
 *  y[M,F] = (x.Wa^T + ba) * gelu(x.Wg^T + bg),  w = [Wa ; Wg] as stored by diffusers ff.net.0.proj ([2F,K]).
 *  `w_packed` must be the row-interleaved form produced by c2d_pack_geglu (blocks of 64 rows: a,g,a,g..)
 *  when impl resolves to TCGEN05; the SIMT path takes the original [2F,K] layout (packed = 0). */
int c2d_geglu_linear(const void* x, const void* w, const float* bias, void* y, int M, int F, int K,
                     int packed, int dtype, int impl, void* stream);

/* ---- 3x3 convolution, padding 1, NHWC, weights [Cout][3][3][Cin] (tap-major, channel-minor).
 *  y = conv(up2x?(x)) + bias + rowvec[b] + residual.   stride in {1,2}; upsample2x folds diffusers'
 *  nearest x2 Upsample2D into the gather.  H,W are the INPUT dims (before upsample). */
int c2d_conv3x3(const void* x, const void* w, const float* bias, const float* rowvec, const void* residual,
                void* y, int B, int H, int W, int Cin, int Cout, int stride, int upsample2x, int dtype,
                int impl, void* stream);

/* ---- GroupNorm (+ optional SiLU) over x[B][HW][C]; stats_ws: >= B*groups*2 doubles, zeroed by the call.
 *  Optional second source: channels [C1, C1+C2) are read from x2[B][HW][C2] (the UNet skip concat
 *  `cat([h, skip], dim=1)` is never materialised for the norm).  If raw_cat != NULL the concatenated,
 *  un-normalised input is also written there (needed by the resnet shortcut). */
int c2d_group_norm(const void* x, const void* x2, const float* gamma, const float* beta, void* y, void* raw_cat,
                   double* stats_ws, int B, int HW, int C1, int C2, int groups, float eps, int silu, int dtype,
                   void* stream);

/* ---- tcgen05 GEMM / convolution with the optional extras of the product path (bf16 only: C2D_ERR_ARG for fp32;
 *  same arithmetic as c2d_linear / c2d_conv3x3 otherwise):
 *  (1) A = [x | x2] concatenated along K (first K1 columns from x, K1 % 64 == 0): the UNet's skip concatenation
 *      `torch.cat([h, skip], 1)` feeding conv_shortcut is never materialised;
 *  (2) chan_stats != NULL: the epilogue also accumulates per-channel (sum, sum of squares) of y into
 *      chan_stats[M / stats_rows][N][2], 2^20 fixed-point 64-bit integers (caller zero-initialises), which
 *      c2d_group_norm_apply consumes -- the GroupNorm statistics pass over y disappears.
 *  (3) row_stats_out != NULL: per-row (sum, sum of squares) of y in the same fixed-point format, [M][2]
 *      (caller zero-initialises) -- the statistics a following LayerNorm needs;
 *  (4) ln_row_stats != NULL: LayerNorm folded into THIS GEMM: x is the un-normalised activation whose row
 *      statistics are ln_row_stats, w / ln_colsum / bias come from c2d_pack_lnfold, and the epilogue applies
 *      y = rstd_m (acc - mean_m colsum_n) + bias_n.  BasicTransformerBlock's norm1/2/3 kernels disappear. */
int c2d_linear_ex(const void* x, const void* x2, int K1, int ldx2, const void* w, const float* bias,
                  const float* rowvec, int rows_per_vec, const void* residual, void* y, int M, int N, int K,
                  int ldx, int ldy, int ldr, int act, long long* chan_stats, int stats_rows,
                  long long* row_stats_out, const long long* ln_row_stats, const float* ln_colsum, float ln_eps,
                  int dtype, void* stream);
/* GEGLU projection (packed weights, see c2d_geglu_linear) with the optional folded LayerNorm of (4). */
int c2d_geglu_linear_ex(const void* x, const void* w, const float* bias, const long long* ln_row_stats,
                        const float* ln_colsum, float ln_eps, void* y, int M, int F, int K, int dtype,
                        void* stream);
/* Weight preparation for (4):  w_out[n][k] = w[n][k] gamma[k] (dtype),  colsum[n] = sum_k bf16(w_out[n][k]),
 * bias_out[n] = bias[n] + sum_k beta[k] w[n][k].  w fp32 [N][K] (nn.Linear layout). */
int c2d_pack_lnfold(const float* w, const float* gamma, const float* beta, const float* bias, void* w_out,
                    float* colsum, float* bias_out, int N, int K, int dtype, void* stream);
int c2d_conv3x3_ex(const void* x, const void* w, const float* bias, const float* rowvec, const void* residual,
                   void* y, int B, int H, int W, int Cin, int Cout, int stride, long long* chan_stats, int dtype,
                   void* stream);

/* ---- diffusers Downsample2D(padding = 0) of the AutoencoderKL encoder: F.pad(x, (0, 1, 0, 1)) + conv3x3 stride 2
 *  (zeros are added on the right / bottom only): y[B][H/2][W/2][Cout], H and W even.  chan_stats as in c2d_conv3x3_ex
 *  (bf16 only, may be NULL).  SD-1.5 VAE encoder (third-party diffusers==0.23.1, requirements.txt:7; latent format
 *  (4, 64, 64) per data/audiocaps_latent_v4.py:185). */
int c2d_conv3x3_down(const void* x, const void* w, const float* bias, void* y, int B, int H, int W, int Cin, int Cout,
                     long long* chan_stats, int dtype, int impl, void* stream);

/* ---- per-channel statistics of x[B][HW][C] in the format above (for tensors not produced by an _ex call). */
int c2d_channel_stats(const void* x, long long* chan_stats, int B, int HW, int C, int dtype, void* stream);

/* ---- GroupNorm (+ optional SiLU) from channel statistics: ONE pass over x (and the skip source x2, whose
 *  statistics are stats2); same result as c2d_group_norm up to the fixed-point rounding of the sums. */
int c2d_group_norm_apply(const void* x, const void* x2, const long long* stats1, const long long* stats2,
                         const float* gamma, const float* beta, void* y, int B, int HW, int C1, int C2,
                         int groups, float eps, int silu, int dtype, void* stream);

/* ---- LayerNorm over the last dim of x[M][C]. */
int c2d_layer_norm(const void* x, const float* gamma, const float* beta, void* y, int M, int C, float eps,
                   int dtype, void* stream);

/* ---- multi-head attention core  o = softmax(q k^T * scale + mask) v.
 *  q: [B][Nq][heads*d] with row stride ldq (elements) and batch stride bsq; same for k, v, o.
 *  kv_len_valid: keys >= this index are masked out (used for padded K/V). mask: optional bool/uint8
 *  [B][Nkv] (1 = keep).  Replaces head_to_batch_dim/get_attention_scores/bmm/batch_to_head_dim
 *  (models/audio_attention_processor.py:124-131) and diffusers' SDPA self-attention. */
int c2d_attention(const void* q, const void* k, const void* v, void* o, int B, int heads, int Nq, int Nkv, int d,
                  long long ldq, long long ldk, long long ldv, long long ldo, long long bsq, long long bsk,
                  long long bsv, long long bso, float scale, const uint8_t* mask, int dtype, int impl,
                  void* stream);
/* Training forward (the attention of models/audio_attention_processor.py:124-131 / diffusers' SDPA under autograd in the reference's
 * scripts/train_stage3.py:177-180): the same call with an optional by-product for c2d_attention_bwd -- lse [B][heads][Nq] fp32, the per-row
 * log-sum-exp of the scaled scores in the log2 domain (max + log2(sum 2^(s - max)), s = scale * log2(e) * q.k).  Only the
 * long-sequence tcgen05 kernel (head_dim <= 64, Nkv > 128) produces it; *lse_written (host) says whether this call did. */
int c2d_attention_lse(const void* q, const void* k, const void* v, void* o, int B, int heads, int Nq, int Nkv, int d,
                      long long ldq, long long ldk, long long ldv, long long ldo, long long bsq, long long bsk,
                      long long bsv, long long bso, float scale, const uint8_t* mask, float* lse, int* lse_written, int dtype,
                      int impl, void* stream);

/* ---- fused cross-attention site (bf16 / tcgen05 only).  Replaces, in ONE kernel per denoising step, the reference's
 *  attn.to_q + head_to_batch_dim + get_attention_scores + bmm + batch_to_head_dim
 *  (models/audio_attention_processor.py:114-131); Q is never written to memory:
 *    q = LN?(x) Wq^T + q_bias      x [B*Nq][C] (row stride ldx), wq [C][C] (nn.Linear layout, heads*d rows);
 *                                  ln_row_stats != NULL: x is the UN-normalised stream, wq = Wq diag(gamma),
 *                                  ln_colsum / q_bias from c2d_pack_lnfold, ln_row_stats = int64 [B*Nq][2] fixed-point
 *                                  (sum, sumsq) of the rows of x (see c2d_linear_ex)
 *    o = softmax(q k^T scale) v  [ + lambda2 * softmax(q k2^T scale) v2 ]
 *  k, v [B][T][heads*d] (T <= 96): the text keys / values with the audio injected ("add": T = 77, "concat": T = 81;
 *  models/audio_attention_processor.py:85-121).  k2, v2 [B][T2][heads*d] (T2 <= 16, NULL / 0 = none): DECOUPLED audio
 *  branch -- its own softmax, scaled by lambda2 and added (the gated second branch of models/audio_adapter_v4.py:208-261
 *  when it shares the site's query).  o [B*Nq][heads*d] (row stride ldo) feeds to_out[0] (c2d_linear_ex).
 *  K / V are step-invariant: c2d_xattn_pack_kv lays them out ONCE per image as per-(batch, head) shared-memory images
 *  (c2d_xattn_packed_bytes(C, heads, T, T2) bytes per batch element, caller-owned) that the per-step kernel fetches
 *  with one bulk copy per head.  Shapes outside the kernel return C2D_ERR_UNSUPPORTED (ask c2d_xattn_supported first;
 *  c2d_xattn_packed_bytes returns 0 for them); nothing is silently routed elsewhere. */
int c2d_xattn_supported(int C, int heads, int Nq, int T, int T2, int dtype);
long long c2d_xattn_packed_bytes(int C, int heads, int T, int T2);
int c2d_xattn_pack_kv(const void* k, const void* v, long long ldkv, long long bskv, int T, const void* k2, const void* v2,
                      long long ldkv2, long long bskv2, int T2, void* packed, int B, int C, int heads, int dtype,
                      void* stream);
int c2d_xattn_fwd(const void* x, long long ldx, const void* wq, const float* q_bias, const long long* ln_row_stats,
                  const float* ln_colsum, float ln_eps, const void* kv_packed, int T, int T2, float lambda2, void* o,
                  long long ldo, int B, int Nq, int C, int heads, float scale, int dtype, void* stream);

/* ---- AudioAttnProcessor context step (models/audio_attention_processor.py:85-109), step-invariant:
 *  ehs_out = ehs + sigmoid(alpha) * mean_K(W2.gelu(W1.a + b1) + b2)            (ADD,   Tout = T)
 *  ehs_out = [ehs ; adaptive_avg_pool_{<=4}(W2.gelu(W1.a + b1) + b2)]           (CONCAT, Tout = T+min(K,4))
 *  ehs [B][T][D], audio [B][K][Da] (same dtype), w1 [Hb][Da], w2 [D][Hb], alpha: device fp32 scalar. */
int c2d_audio_context(const void* ehs, const void* audio, const void* w1, const float* b1, const void* w2,
                      const float* b2, const float* alpha, void* ehs_out, int B, int T, int D, int K, int Da,
                      int Hb, int mode, int dtype, void* stream);

/* ---- elementwise / layout ------------------------------------------------------------------- */
/* sinusoidal timestep embedding, flip_sin_to_cos, out fp32 [B][dim]: cat[cos, sin] */
int c2d_timestep_embedding(const float* t, float* out, int B, int dim, void* stream);
/* y = act(x) elementwise over n elements; x dtype_in -> y dtype_out (casts included) */
int c2d_unary(const void* x, void* y, long long n, int act, int dtype_in, int dtype_out, void* stream);
/* y = a + b (same dtype) */
int c2d_add(const void* a, const void* b, void* y, long long n, int dtype, void* stream);
/* y[M,F] = x[M,0:F] * gelu(x[M,F:2F]) */
int c2d_geglu(const void* x, void* y, int M, int F, int dtype, void* stream);
/* nearest 2x upsample of x[B][H][W][C] -> y[B][2H][2W][C] */
int c2d_upsample2x(const void* x, void* y, int B, int H, int W, int C, int dtype, void* stream);
/* y[B][HW][C1+C2] = cat(x1, x2) on channels */
int c2d_concat(const void* x1, const void* x2, void* y, long long rows, int C1, int C2, int dtype, void* stream);
/* fp32 NCHW [B][C][HW] -> dtype NHWC [B][HW][C]  and back */
int c2d_nchw_to_nhwc(const float* x, void* y, int B, int C, int HW, int dtype, void* stream);
int c2d_nhwc_to_nchw(const void* x, float* y, int B, int C, int HW, int dtype, void* stream);

/* ---- fused classifier-free guidance + scheduler update:
 *  eps2: dtype NHWC [2B][HW][4] = cat[uncond, cond];  eps = eu + g (ec - eu);
 *  x (fp32 NCHW [B][4][HW]) <- ca * x + cb * eps      (DDIM eta=0:  ca, cb from (t, t_prev);
 *                                                       Euler:       ca = 1, cb = sigma_next - sigma)
 *  then writes the next UNet input xin2 (dtype NHWC [2B][HW][4]) = cat[x, x] * in_scale
 *  (in_scale = 1 for DDIM, 1/sqrt(sigma_next^2+1) for Euler).  coef = DEVICE fp32 {ca, cb, in_scale} so one
 *  captured CUDA graph serves every step.  Optionally records x to trace (fp32).
 *  xin_cpitch >= 4: channel pitch of xin2; with 8 the four real channels are followed by zero padding the caller
 *  initialised once, which lets conv_in (Cin = 4) run as a tcgen05 implicit GEMM (16-byte TMA rows). */
int c2d_cfg_sched_step(const void* eps2, float* x, void* xin2, float* trace, int B, int HW, float guidance,
                       const float* coef, int xin_cpitch, int dtype, void* stream);

/* row softmax y = softmax(x * scale) over the last dim of x[M][N]; batched transpose y[b][C][R] = x[b][R][C].
 * Used by the materialised attention route for head dims the flash kernels do not take
 * (AutoencoderKL mid-block attention, d = 512; AudioTokenGenerator single-head d = 768,
 * models/audio_adapter_v4.py:101-108). */
int c2d_softmax_rows(const void* x, void* y, int M, int N, float scale, int dtype, void* stream);
int c2d_transpose(const void* x, void* y, int batch, int R, int C, int dtype, void* stream);

/* ---- audio-conditioning side (once per image; latency-bound small kernels) ---------------------- */
/* y[b,k,:] = a + b with per-operand broadcast: mode 0 = [B,K,D], 1 = [B,1,D], 2 = [1,K,D].
 * (token offsets / positional tables: models/hierarchical_audio_v4.py:205,479-480,490; audio_adapter_v4.py:91-93,108) */
int c2d_bcast_add(const void* a, const void* b, void* y, int B, int K, int D, int a_mode, int b_mode, int dtype,
                  void* stream);
/* SoftHierarchicalDecomposition.compute_assignments (models/hierarchical_audio_v4.py:154-182):
 * assign[r,:] = softmax((10 cos(tok_r, anchor_l) + W2 gelu(W1 tok_r + b1) + b2) / T); tokens [rows][D],
 * anchors [L][D], w1 [Hg][D], w2 [L][Hg], temperature: device fp32 scalar, assign fp32 [rows][L]. */
int c2d_hier_assign(const void* tokens, const void* anchors, const void* w1, const float* b1, const void* w2,
                    const float* b2, const float* temperature, float* assign, int rows, int D, int L, int Hg,
                    int dtype, void* stream);
/* LevelToUNetRouter.forward (models/hierarchical_audio_v4.py:325-369), 3 levels -> early/mid/late.
 * hw (fp32 [B][3]) may be NULL (no adaptive weights); routing fp32 [3][3]; gates fp32 {early,mid,late}. */
int c2d_hier_route(const void* tok10, const float* assign, const float* hw, const float* routing, const float* gates,
                   void* r_early, void* r_mid, void* r_late, int B, int K, int D, int dtype, void* stream);
/* apply_normalization (scripts/inference.py:92-99): y = x * target / mean(||x||_2); per_sample = 0 is the
 * reference's batch-coupled mean, per_sample = 1 the sharding-invariant variant (SURVEY decision D3). */
int c2d_norm_scale(const void* x, void* y, int B, int K, int D, float target, int per_sample, int dtype, void* stream);
/* HierarchicalAudioDecomposition (legacy 5-3-2, models/hierarchical_audio_v4.py:849-864):
 * out[B][nf+nb+na][D] = cat(fg * w0, bg * w1, amb * w2), w = softmax(hierarchy_weights[3]). */
int c2d_legacy_combine(const void* fg, const void* bg, const void* amb, const float* hierarchy_weights, void* out,
                       int B, int nf, int nb, int na, int D, int dtype, void* stream);

/* ==== stage-3 fine-tune step: backward + optimiser (reference scripts/train_stage3.py:132-191) ==========
 * With the SD-1.5 UNet frozen the backward pass needs ACTIVATION gradients only.  Dense layers reuse the forward
 * entry points on transformed copies of the frozen weights (c2d_linear on W^T; c2d_conv3x3 on the flipped + transposed
 * kernel, after c2d_zero_insert2x for the stride-2 convolutions); the entry points below are the pieces without a
 * forward twin.  Same dtype switch as the forward kernels, fp32 accumulation.  `add` (optional, same shape as the
 * output) is summed into the result: the fan-in of residual branches costs no extra pass. */
/* GroupNorm(+SiLU) adjoint: statistics recomputed from x;  dx = d/dx [act(gn(x))] . dy (+ add).
 * ws (optional): caller-ZEROED scratch of B * C * 32 bytes -- with it the bf16 path runs as three coalesced passes
 * (channel statistics, adjoint sums, apply) instead of one CTA per (sample, group).  x_stats (optional, with ws): the
 * [B][C][2] fixed-point channel statistics of x the forward pass already holds (c2d_channel_stats or a GEMM epilogue) --
 * the statistics pass is skipped. */
int c2d_group_norm_bwd(const void* x, const void* dy, const float* gamma, const float* beta, const void* add, void* dx, void* ws,
                       const long long* x_stats, int B, int HW, int C, int groups, float eps, int silu, int dtype, void* stream);
/* LayerNorm adjoint over the rows of x[M][C] (+ add) */
int c2d_layer_norm_bwd(const void* x, const void* dy, const float* gamma, const void* add, void* dx, int M, int C, float eps,
                       int dtype, void* stream);
/* adjoint of c2d_geglu: ag [M][2F] = [a | g], dy [M][F] -> dag [M][2F] */
int c2d_geglu_bwd(const void* ag, const void* dy, void* dag, int M, int F, int dtype, void* stream);
/* flash-attention adjoint (recompute form) of c2d_attention: dq, dk, dv from q, k, v, o, dout.  Row strides ld*, batch
 * strides bs* in elements (packed QKV views welcome); lse_ws / delta_ws: fp32 [B][heads][Nq] scratch each; d <= 160.
 * have_lse != 0: lse_ws already holds the forward pass's log-sum-exp (c2d_attention_lse reported *lse_written): the bf16
 * tensor-core kernels then skip their first sweep over the keys. */
int c2d_attention_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, void* dq, void* dk, void* dv,
                      float* lse_ws, float* delta_ws, int B, int heads, int Nq, int Nkv, int d, long long ldq, long long ldk,
                      long long ldv, long long ldo, long long lddo, long long lddq, long long lddk, long long lddv,
                      long long bsq, long long bsk, long long bsv, long long bso, long long bsdo, long long bsdq,
                      long long bsdk, long long bsdv, float scale, int have_lse, int dtype, void* stream);
/* z[B][2H][2W][C]: x at the even positions, zeros elsewhere (adjoint of the stride-2 gather);
 * y[B][H][W][C] = 2x2 block sums of x[B][2H][2W][C] (adjoint of the nearest 2x upsample);
 * y[rows][Cs] = x[rows][c0 : c0+Cs] (+ add) (adjoint of the channel concat) */
int c2d_zero_insert2x(const void* x, void* z, int B, int H, int W, int C, int dtype, void* stream);
int c2d_sumpool2x2(const void* x, void* y, int B, int H, int W, int C, int dtype, void* stream);
int c2d_slice_channels(const void* x, const void* add, void* y, long long rows, int C, int c0, int Cs, int dtype, void* stream);
/* weight * mse(pred, target) and its gradient: pred NHWC [B][HW][C] (dtype), target fp32 NCHW [B][C][HW]; *loss (device
 * double, caller-zeroed) += the loss, grad (dtype, NHWC) = weight * 2 (pred - target) / n   (train_stage3.py:166) */
int c2d_mse_loss_grad(const void* pred, const float* target, void* grad, double* loss, int B, int HW, int C, float weight,
                      int dtype, void* stream);
/* out fp32 [B][C] (+)= sum over the R rows of x[B][R][C] */
int c2d_colsum(const void* x, float* out, int B, int R, int C, int accumulate, int dtype, void* stream);
/* adjoint of  ehs' = ehs + sigmoid(alpha) * af  (audio_attention_processor.py:92-97), fp32, n = B*D:
 * daf = gate * s,  *dalpha += gate (1 - gate) <s, af>   with s = sum over the text positions of d ehs' */
int c2d_gate_bwd(const float* s, const float* af, const float* alpha, float* daf, float* dalpha, int n, void* stream);
/* dz[r][j] = dh[r / K][j] * gelu'(z[r][j]) / K: adjoint of mean-over-K-tokens of gelu(z) (fp32) */
int c2d_gelu_bwd_bcast(const float* z, const float* dh, float* dz, int rows, int H, int K, void* stream);
/* optimiser (train_stage3.py:33-38, :182-188): *out (device double, caller-zeroed) += sum x^2;
 * scale = min(1, max_norm / (sqrt(sumsq) + 1e-6)) as torch.nn.utils.clip_grad_norm_ (norm_out optional);
 * AdamW (decoupled weight decay, bias correction) on flat fp32 buffers, gradient pre-multiplied by *grad_scale. */
int c2d_sumsq(const float* x, long long n, double* out, void* stream);
int c2d_clip_scale(const double* sumsq, float max_norm, float* scale, float* norm_out, void* stream);
int c2d_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int step, const float* grad_scale, void* stream);
/* (replaces optimizer.step() + scheduler.step() of scripts/train_stage3.py:188-189 with the AdamW / CosineAnnealingLR set up at :33-46)
 * The same update with the schedule in device memory -- sched [sched_len][3] = (learning rate, 1 - beta1^(t+1), 1 - beta2^(t+1))
 * of optimiser step t, *step_dev = steps taken so far (the caller increments it) -- so that the launch can live in a CUDA graph
 * replayed once per training step. */
int c2d_adamw_step_sched(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, const float* sched,
                         int sched_len, const int* step_dev, float beta1, float beta2, float eps, float weight_decay,
                         const float* grad_scale, void* stream);

/* ---- weight packing helpers (run once at load time) ------------------------------------------ */
/* [Cout][Cin][3][3] fp32 (PyTorch/diffusers layout) -> [Cout][3][3][Cin] dtype */
int c2d_pack_conv3x3(const float* w, void* out, int Cout, int Cin, int dtype, void* stream);
/* diffusers GEGLU proj [2F][K] fp32 -> row-interleaved (64-row blocks a,g,a,g,...) dtype [2F][K];
 * bias [2F] -> interleaved the same way (fp32). */
int c2d_pack_geglu(const float* w, const float* bias, void* w_out, float* bias_out, int F, int K, int dtype,
                   void* stream);

/* ==== CLAP HTSAT audio tower (models/audio_encoder.py:164-174 -> HF ClapFeatureExtractor + ClapModel.get_audio_features) ====
 * The dense layers use c2d_linear / c2d_layer_norm; these are the remaining pieces. */
/* frames[(b*n_frames+f)][k] = wave[b][reflect(f*hop + k - n_fft/2)] * window[k]  (centered STFT framing, fp32) */
int c2d_stft_frames(const float* wave, const float* window, float* frames, int B, int T, int n_fft, int hop, int n_frames,
                    void* stream);
/* bf16 product mode: frames3 [B*n_frames][3*n_fft] bf16 = [hi | lo | hi] with x = hi + lo (hi = bf16(x), lo = bf16(x - hi)).
 * Against constant-matrix rows [HI | HI | LO] one bf16 tensor-core GEMM (c2d_linear, fp32 accumulation) gives
 * hi HI + lo HI + hi LO: the DFT at ~2^-16 operand precision instead of an fp32 FFMA GEMM.  n_fft % 8 == 0. */
int c2d_stft_frames_split(const float* wave, const float* window, void* frames3, int B, int T, int n_fft, int hop, int n_frames,
                          void* stream);
/* dft row m = [re(0..nb) at column 0 | im(0..nb) at column im_off], row pitch ld elements of `dtype` (the GEMM of the frames
 * against the constant DFT matrix) -> out fp32 [M][nb] = re^2 + im^2 */
int c2d_power_spectrum(const void* dft, float* out, long long M, int nb, int ld, int im_off, int dtype, void* stream);
/* y = 10 log10(max(x, floor)) * a[f] + b[f]: power_to_db fused with the eval-mode BatchNorm2d over mel bins */
int c2d_log_mel_affine(const float* x, const float* a, const float* b, float* y, long long M, int F, float floor_value,
                       void* stream);
/* mel [B][n_frames][64] fp32 -> patches [B*4096][16] (dtype): bicubic n_frames -> 1024, 4-chunk fold to 256x256,
 * 4x4/stride-4 patch gather (reshape_mel2img + the im2col of ClapAudioPatchEmbed.proj) */
int c2d_clap_patches(const float* mel, void* patches, int B, int n_frames, int n_mel, int dtype, void* stream);
/* Swin 8x8 window attention over a packed qkv [B][H*W][3C] with relative-position bias [heads][64][64], cyclic shift
 * and shift mask folded into the addressing; out [B][H*W][C] in the un-shifted token order */
int c2d_window_attention(const void* qkv, const float* bias, void* out, int B, int H, int W, int C, int heads, int shift,
                         float scale, int dtype, void* stream);
/* ClapAudioPatchMerging gather: x [B][H][W][C] -> out [B][H/2*W/2][4C] (order (0,0),(1,0),(0,1),(1,1)) */
int c2d_patch_merge(const void* x, void* out, int B, int H, int W, int C, int dtype, void* stream);
/* out[b][c] = mean over tokens of x [B][N][C] (fp32 result) */
int c2d_token_mean(const void* x, float* out, int B, int N, int C, int dtype, void* stream);
/* y = x / max(||x||_2, eps) per row of x [B][D] (fp32) */
int c2d_l2_normalize(const float* x, float* y, int B, int D, float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* C2D_H_ */
