// tcgen05 / TMEM / TMA GEMM for sm_100a: y = act(A . W^T + bias + rowvec) + residual, bf16 in, fp32 accumulate.
//
// One kernel, two A-operand producers:
//   plain : A = x[M][K] (row stride ldx), 2-D TMA boxes {64 k, 128 rows}
//   conv  : A = implicit im2col of an NHWC image for a 3x3 / pad-1 / stride-1 convolution.  The 128-pixel
//           M-tile is a rectangle {bw x bh x bb} of the [B][H][W][C] tensor; for filter tap (ky,kx) the
//           producer issues ONE 4-D TMA box at coordinates (c0, x0+kx-1, y0+ky-1, b0) -- the halo /
//           zero padding comes from TMA's out-of-bounds zero fill, so im2col never exists in memory.
// B = W[N][K] row-major (PyTorch Linear layout; conv weights packed [Cout][3][3][Cin]), 2-D TMA {64 k, BN rows}.
// Both operands land in 128-byte-swizzled K-major smem tiles and feed tcgen05.mma (M=128, N=BN, K=16) with
// the fp32 accumulator in TMEM.  Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc),
// warps 2..5 = epilogue (tcgen05.ld -> bias / time-embedding row vector / activation / GEGLU gate /
// residual -> bf16 global stores).  3-stage mbarrier ring; 2 CTAs per SM so one CTA's epilogue overlaps
// the other's main loop.
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace c2d {

using namespace tc;

// warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = epilogue: TWO warps per TMEM lane quadrant that take alternate
// 32-column chunks -- the short-K GEMMs of the UNet (K = 320) are epilogue-bound, a second warp per scheduler doubles the
// epilogue's issue rate and its loads in flight
constexpr int TC_BM = 128, TC_BK = 64, TC_STAGES = 3, TC_EPI_WARPS = 8, TC_THREADS = 64 + 32 * TC_EPI_WARPS;

struct TcParams {
  const float* bias;
  const float* rowvec;
  const bf16* residual;
  bf16* y;
  int M, N, K;
  long long ldy, ldr;
  int rows_per_vec;
  int act;
  int num_k_blocks;
  // conv geometry (HW, W: OUTPUT plane; cstride: convolution stride 1 | 2)
  int HW, W, Cin, cblocks, cstride;
  int cpad;        // leading zero padding of the 3x3 window: 1 (symmetric pad 1) | 0 (diffusers Downsample2D: pad right / bottom only)
  // optional per-channel statistics of y for the consuming GroupNorm: stats[b][n][{sum, sumsq}] as 2^20 fixed-point
  // 64-bit integers (integer atomics: the result does not depend on the order CTAs retire in); b = row / stats_rows
  unsigned long long* stats;
  int stats_rows;
  // plain GEMM with A = [x | x2] concatenated along K: k-blocks >= kb_split come from the second tensor map
  int kb_split;
  // LayerNorm folded into this GEMM (consumer side): x is the UN-normalised activation, w = W diag(gamma),
  // y = rstd_m (acc - mean_m colsum_n) + bias_n with (mean, rstd) from ln_stats[m] = fixed-point (sum, sumsq) of row m
  const long long* ln_stats;
  const float* ln_colsum;
  float ln_invK, ln_eps;
  // producer side of the same scheme: per-row fixed-point (sum, sumsq) of y accumulated over the N tiles
  unsigned long long* rowstats_out;
  // split-K (CTA-pair kernel, small-M convolutions): CTA (x, z) reduces k-blocks [z * kb_per_split, ...) and writes its
  // raw fp32 accumulator tile to partial[z][M][N]; splitk_finish_kernel sums the slices and applies the epilogue
  float* partial;
  int kb_per_split;
};

// raw fixed-point (sum, sumsq) of row m (zero when there is no folded LayerNorm / the row is out of range); split from
// the arithmetic so that the 16-byte load can be issued a whole tile ahead of its use
__device__ __forceinline__ longlong2 ln_row_load(const TcParams& p, int m) {
  longlong2 v = make_longlong2(0, 0);
  if (p.ln_stats && m < p.M) v = *reinterpret_cast<const longlong2*>(p.ln_stats + 2 * (long long)m);
  return v;
}
__device__ __forceinline__ void ln_row_finish(const TcParams& p, const longlong2& v, float& rstd, float& mr) {
  rstd = 1.f; mr = 0.f;
  if (p.ln_stats) {
    const float inv = p.ln_invK * (1.0f / 1048576.0f);
    const float mean = (float)v.x * inv;
    const float var = fmaxf(fmaf(-mean, mean, (float)v.y * inv), 0.f);
    rstd = rsqrtf(var + p.ln_eps);
    mr = mean * rstd;
  }
}
// (d0, d1) = (a0, a1) * (b, b) + (c0, c1) on the packed FFMA2 pipe
__device__ __forceinline__ void ffma2v(float& d0, float& d1, float a0, float a1, float b, float c0, float c1) {
  asm("{.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %4};\n\t"
      "mov.b64 rc, {%5, %6};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b), "f"(c0), "f"(c1));
}
// out[j] = acc[j] + bias[j]  |  folded LayerNorm: rstd * acc[j] - (mean * rstd) * colsum[j] + bias[j]; bias / colsum are
// 16-byte aligned smem slices read as float4 (the epilogue of the short-K GEMMs is instruction-bound)
template <bool LN>
__device__ __forceinline__ void epi_affine32(const uint32_t (&r)[32], const float* __restrict__ sb, const float* __restrict__ scs,
                                             float rstd, float nmr, float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 b4 = *reinterpret_cast<const float4*>(sb + j);
    if (LN) {
      // packed f32x2 FMAs: (colsum, colsum') * nmr + (bias, bias'), then (acc, acc') * rstd + that
      const float4 c4 = *reinterpret_cast<const float4*>(scs + j);
      float t0, t1, t2, t3;
      ffma2v(t0, t1, c4.x, c4.y, nmr, b4.x, b4.y);
      ffma2v(t2, t3, c4.z, c4.w, nmr, b4.z, b4.w);
      ffma2v(v[j], v[j + 1], __uint_as_float(r[j]), __uint_as_float(r[j + 1]), rstd, t0, t1);
      ffma2v(v[j + 2], v[j + 3], __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]), rstd, t2, t3);
    } else {
      v[j] = __uint_as_float(r[j]) + b4.x;
      v[j + 1] = __uint_as_float(r[j + 1]) + b4.y;
      v[j + 2] = __uint_as_float(r[j + 2]) + b4.z;
      v[j + 3] = __uint_as_float(r[j + 3]) + b4.w;
    }
  }
}

__device__ __forceinline__ void ln_row_coeffs(const TcParams& p, int m, float& rstd, float& mr) {
  ln_row_finish(p, ln_row_load(p, m), rstd, mr);
}

// Drain of one staged 32-row x 32-column chunk for the common case -- interior rows (all 32 valid), 16-byte aligned y
// (and residual), no statistics of any kind: lane = (row rsub of an 8-row block, 8-column group g), four 16-byte stores
// per lane, the four residual loads issued first.  The general loop below spends ~6 index / predicate / branch
// instructions per useful one (ncu on M = 2^20, N = 288, K = 96: 445 warp instructions per chunk, 70 of them data
// movement or arithmetic), which bounds every GEMM whose main loop is short (K <= 384: the CLAP tower, the UNet's
// 64 x 64-level projections).  `srow0` = this lane's first staged row at its column group, `ncol_ok` = the group lies in N.
template <bool STATS = false>
__device__ __forceinline__ void drain32_fast(const float* srow0, int pitch, bf16* yp, long long ldy, const bf16* rp, long long ldr,
                                             bool ncol_ok, float* ssum = nullptr, float* ssq = nullptr) {
  if (STATS) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { ssum[j] = 0.f; ssq[j] = 0.f; }
  }
  if (!ncol_ok) return;
  uint4 res[4];
  if (rp) {
#pragma unroll
    for (int u = 0; u < 4; ++u) res[u] = *reinterpret_cast<const uint4*>(rp + (long long)u * 8 * ldr);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float* sp = srow0 + (size_t)u * 8 * pitch;
    const float4 f0 = *reinterpret_cast<const float4*>(sp), f1 = *reinterpret_cast<const float4*>(sp + 4);
    float v[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
    if (rp) {
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&res[u]);
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); v[2 * j] += f.x; v[2 * j + 1] += f.y; }
    }
    Vec8<bf16>::store(yp + (long long)u * 8 * ldy, v);
    if (STATS) {           // per-channel (sum, sum of squares) of what was stored, for the consuming GroupNorm
#pragma unroll
      for (int j = 0; j < 8; ++j) { ssum[j] += v[j]; ssq[j] = fmaf(v[j], v[j], ssq[j]); }
    }
  }
}

// activation of 32 staged values: the switch is hoisted out of the element loop and GELU runs on the packed f32x2 form
// (the fc1 layers of the CLAP tower have K = 96 .. 768: their epilogue is the kernel)
__device__ __forceinline__ void epi_act32(float (&v)[32], int act) {
  if (act == C2D_ACT_NONE) return;
  if (act == C2D_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      float a = 1.f, b = 1.f;
      gelu_mul2(a, b, v[j], v[j + 1]);
      v[j] = a; v[j + 1] = b;
    }
  } else if (act == C2D_ACT_SILU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = silu_fast(v[j]);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
}

constexpr float STATS_SCALE = 1048576.0f;      // 2^20

// Epilogue helper.  Every lane holds sums (s) and sums of squares (q) of its 8 columns over the rows it drained;
// lanes with equal (lane & 3) own the same columns.  A transposing butterfly (8 + 4 + 2 shuffles) leaves each of
// the 8 lanes of a column group with 2 of its 16 totals, which go to the fixed-point accumulators.
__device__ __forceinline__ void stats_commit(const float (&s)[8], const float (&q)[8], int lane, unsigned long long* dst,
                                             int nvalid) {
  float w[8], x[4], y[2];
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = h16 ? s[i] : q[i];
    const float keep = h16 ? q[i] : s[i];
    w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = h8 ? w[i] : w[4 + i];
    const float keep = h8 ? w[4 + i] : w[i];
    x[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = h4 ? x[i] : x[2 + i];
    const float keep = h4 ? x[2 + i] : x[i];
    y[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const int col = (h8 ? 4 : 0) + (h4 ? 2 : 0), kind = h16 ? 1 : 0;
#pragma unroll
  for (int i = 0; i < 2; ++i)
    if (col + i < nvalid)
      atomicAdd(dst + (size_t)(col + i) * 2 + kind, (unsigned long long)__float2ll_rn(y[i] * STATS_SCALE));
}

static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

static int get_encode() {
  if (g_encode) return C2D_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    cudaGetLastError();
    set_error("cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
    return C2D_ERR_CUDA;
  }
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return C2D_OK;
}

// bf16 tensor map, 128B swizzle, zero OOB fill.  dims/box innermost-first; strides in BYTES for dims 1..rank-1.
int make_tmap_bf16_strided(CUtensorMap* m, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  int rc = get_encode();
  if (rc) return rc;
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = elem_strides ? elem_strides[i] : 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims {%llu,%llu,%llu,%llu} stride0 %llu box {%u,%u,%u,%u} base %p",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
              (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0,
              rank > 3 ? box[3] : 0, base);
    return C2D_ERR_CUDA;
  }
  return C2D_OK;
}

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  return make_tmap_bf16_strided(m, base, rank, dims, strides_bytes, box, nullptr);
}

// bf16 tensor map WITHOUT swizzle (dense shared-memory box: rows at box[0] * 2 bytes): TMA stores of epilogue tiles whose
// rows are not a multiple of the 128-byte swizzle span.
int make_tmap_bf16_plain(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box) {
  int rc = get_encode();
  if (rc) return rc;
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (no swizzle) failed (CUresult %d): rank %d dims {%llu,%llu} stride0 %llu box {%u,%u} base %p",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0], rank > 1 ? box[1] : 0, base);
    return C2D_ERR_CUDA;
  }
  return C2D_OK;
}

template <int BN>
struct TcCfg {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  static constexpr int SMEM_BYTES = TC_STAGES * STAGE_BYTES + 256 + 2 * BN * 4 + 1024;   // + barriers + bias/rowvec + slack
};

template <int BN, bool CONV, bool GEGLU, bool LNF>
__global__ void __launch_bounds__(TC_THREADS, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  using Cfg = TcCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + TC_STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + TC_STAGES;
  uint64_t* tmem_full = empty + TC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int i = 0; i < TC_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();           // everything above overlapped the predecessor's tail; global memory is touched from here on
  pdl_trigger();

  // Producer and MMA warps run their protocols CONVERGED and issue through elect_one(): the TMA / tcgen05
  // instructions are uniform-datapath instructions and must not sit in a lane-divergent region (tc_common.cuh).
  if (warp == 0) {
    // ===================== TMA producer =====================
    int b0 = 0, y0 = 0, x0 = 0;
    if (CONV) {
      b0 = m0 / p.HW;
      int rem = m0 - b0 * p.HW;
      y0 = rem / p.W;
      x0 = rem - y0 * p.W;
    }
    for (int kb = 0; kb < p.num_k_blocks; ++kb) {
      const int s = kb % TC_STAGES;
      const uint32_t ph = (uint32_t)(kb / TC_STAGES) & 1u;
      mbar_wait(&empty[s], ph ^ 1u);
      if (elect_one()) {
        uint8_t* sA = smem + s * Cfg::STAGE_BYTES;
        uint8_t* sB = sA + Cfg::A_BYTES;
        mbar_arrive_expect_tx(&full[s], Cfg::STAGE_BYTES);
        if (CONV) {
          const int tap = kb / p.cblocks;
          const int c0 = (kb - tap * p.cblocks) * TC_BK;
          const int ky = tap / 3, kx = tap - ky * 3;
          // stride-2: the tensor map traverses the input with element strides {1,2,2,1}
          tma_load_4d(sA, &tmA, &full[s], c0, x0 * p.cstride + kx - p.cpad, y0 * p.cstride + ky - p.cpad, b0);
          tma_load_2d(sB, &tmB, &full[s], tap * p.Cin + c0, n0);
        } else {
          if (kb < p.kb_split) tma_load_2d(sA, &tmA, &full[s], kb * TC_BK, m0);
          else tma_load_2d(sA, &tmA2, &full[s], (kb - p.kb_split) * TC_BK, m0);
          tma_load_2d(sB, &tmB, &full[s], kb * TC_BK, n0);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_bf16(TC_BM, BN, 0, 0);
    const uint64_t desc0 = make_desc_k_sw128(smem_u32(smem));
    for (int kb = 0; kb < p.num_k_blocks; ++kb) {
      const int s = kb % TC_STAGES;
      const uint32_t ph = (uint32_t)(kb / TC_STAGES) & 1u;
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        // descriptor address field is in 16-byte units; +2 per K = 16 step inside the 128-byte swizzle row
        const uint64_t a_desc = desc0 + (uint64_t)(s * (Cfg::STAGE_BYTES >> 4));
        const uint64_t b_desc = a_desc + (uint64_t)(Cfg::A_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k)
          umma_f16(tmem_base, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit(&empty[s]);                                   // smem slot reusable once these MMAs retire
        if (kb == p.num_k_blocks - 1) umma_commit(tmem_full);     // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    // TMEM -> registers (+bias, +row vector, activation / GEGLU gate) -> fp32 staging tile in the (now idle)
    // pipeline smem -> coalesced pass: 16-byte residual loads and bf16 stores along rows.  Each warp stages
    // and drains only its own 32 rows, so a __syncwarp is the only synchronisation.
    const int q = warp & 3;              // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;    // which of the quadrant's two warps: chunks c = half, half + 2, ...
    constexpr int NOUT = GEGLU ? BN / 2 : BN;         // output columns this CTA produces
    constexpr int PITCH = NOUT + 4;                   // floats; rows stay 16 B aligned, 16 B writes conflict-free
    // while the main loop runs: this CTA's slice of bias (+ time-embedding row vector when it is shared by the
    // whole tile) -> smem, so the drain below issues no dependent global loads for them
    float* s_bias = reinterpret_cast<float*>(smem + TC_STAGES * Cfg::STAGE_BYTES + 256);
    float* s_cs = s_bias + BN;                            // folded-LayerNorm column sums
    const bool rv_shared = p.rowvec && (m0 / p.rows_per_vec) == ((min(m0 + TC_BM, p.M) - 1) / p.rows_per_vec);
    {
      const float* rv0 = rv_shared ? p.rowvec + (long long)(m0 / p.rows_per_vec) * p.N : nullptr;
      for (int j = threadIdx.x - 64; j < BN; j += 32 * TC_EPI_WARPS) {
        const int n = n0 + j;
        float b = 0.f, cs = 0.f;
        if (n < p.N) {
          if (p.bias) b = __ldg(p.bias + n);
          if (rv0) b += __ldg(rv0 + n);
          if (p.ln_colsum) cs = __ldg(p.ln_colsum + n);
        }
        s_bias[j] = b;
        s_cs[j] = cs;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");      // epilogue warps only
    }
    float ln_rstd = 1.f, ln_mr = 0.f;
    if (LNF) ln_row_coeffs(p, m0 + (warp & 3) * 32 + lane, ln_rstd, ln_mr);
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    float* stage = reinterpret_cast<float*>(smem) + (size_t)(q * 32) * PITCH;
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    {
      const int m = m0 + q * 32 + lane;
      const float* rv = (p.rowvec && !rv_shared && m < p.M) ? p.rowvec + (long long)(m / p.rows_per_vec) * p.N : nullptr;
      float* srow = stage + (size_t)lane * PITCH;
      if (!GEGLU) {
#pragma unroll 1
        for (int c = half; c < BN / 32; c += 2) {
          uint32_t r[32];
          tmem_ld_32x32(t_row + c * 32, r);
          tmem_ld_wait();
          const int n = n0 + c * 32;
          float v[32];
          epi_affine32<LNF>(r, s_bias + c * 32, s_cs + c * 32, ln_rstd, -ln_mr, v);
          if (rv) {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (n + j < p.N) v[j] += __ldg(rv + n + j);
          }
          epi_act32(v, p.act);
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(srow + c * 32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      } else {
        // GEGLU: tile columns [0,64) = a, [64,128) = gate for output columns n0/2 + [0,64)
#pragma unroll 1
        for (int c = half; c < 2; c += 2) {
          uint32_t ra[32], rg[32];
          tmem_ld_32x32(t_row + c * 32, ra);
          tmem_ld_32x32(t_row + 64 + c * 32, rg);
          tmem_ld_wait();
          float v[32], gg[32];
          epi_affine32<LNF>(ra, s_bias + c * 32, s_cs + c * 32, ln_rstd, -ln_mr, v);
          epi_affine32<LNF>(rg, s_bias + 64 + c * 32, s_cs + 64 + c * 32, ln_rstd, -ln_mr, gg);
#pragma unroll
          for (int j = 0; j < 32; j += 2) gelu_mul2(v[j], v[j + 1], gg[j], gg[j + 1]);
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(srow + c * 32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
    }
    __syncwarp();
    {
      // coalesced drain, 32 columns at a time: lane = (row rsub of an 8-row block, 8-column group g); the four
      // residual loads of a chunk are issued before any of them is consumed
      const int ncols = GEGLU ? (p.N >> 1) : p.N;            // logical output width
      const int nbase = GEGLU ? (n0 >> 1) : n0;
      const bool vec_y = (p.ldy % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0);
      const bool vec_r = p.residual && (p.ldr % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0);
      const int rsub = lane >> 2, g = lane & 3;
      const int mrow0 = m0 + q * 32;
      float rsum[4] = {0.f, 0.f, 0.f, 0.f}, rsq[4] = {0.f, 0.f, 0.f, 0.f};      // per-row partials (rowstats_out)
      const bool fast = !p.stats && !p.rowstats_out && vec_y && (!p.residual || vec_r) && (ncols % 8 == 0) && m0 + TC_BM <= p.M;
      if (fast) {
        const long long row = mrow0 + rsub;
#pragma unroll 1
        for (int c = half; c < NOUT / 32; c += 2) {
          const int n = nbase + c * 32 + g * 8;
          drain32_fast(stage + (size_t)rsub * PITCH + c * 32 + g * 8, PITCH, p.y + row * p.ldy + n, p.ldy,
                       p.residual ? p.residual + row * p.ldr + n : nullptr, p.ldr, n < ncols);
        }
      } else
#pragma unroll 1
      for (int c = half; c < NOUT / 32; c += 2) {
        const int n = nbase + c * 32 + g * 8;
        const int nvalid = ncols - n;
        uint4 res[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int m = mrow0 + u * 8 + rsub;
          res[u] = make_uint4(0u, 0u, 0u, 0u);
          if (m < p.M && vec_r && nvalid >= 8) res[u] = *reinterpret_cast<const uint4*>(p.residual + (long long)m * p.ldr + n);
        }
        float ssum[8], ssq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { ssum[j] = 0.f; ssq[j] = 0.f; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int rr = u * 8 + rsub, m = mrow0 + rr;
          if (m >= p.M || nvalid <= 0) continue;
          const float* sp = stage + (size_t)rr * PITCH + c * 32 + g * 8;
          const float4 f0 = *reinterpret_cast<const float4*>(sp), f1 = *reinterpret_cast<const float4*>(sp + 4);
          float v[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
          bf16* yp = p.y + (long long)m * p.ldy + n;
          if (p.residual) {
            if (vec_r && nvalid >= 8) {
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&res[u]);
#pragma unroll
              for (int j = 0; j < 4; ++j) { float2 f = __bfloat1622float2(h[j]); v[2 * j] += f.x; v[2 * j + 1] += f.y; }
            } else {
              const bf16* rp = p.residual + (long long)m * p.ldr + n;
#pragma unroll
              for (int j = 0; j < 8; ++j) if (j < nvalid) v[j] += __bfloat162float(rp[j]);
            }
          }
          if (vec_y && nvalid >= 8) {
            Vec8<bf16>::store(yp, v);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) if (j < nvalid) yp[j] = __float2bfloat16_rn(v[j]);
          }
          if (p.stats) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { ssum[j] += v[j]; ssq[j] = fmaf(v[j], v[j], ssq[j]); }
          }
          if (p.rowstats_out) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j < nvalid) { rsum[u] += v[j]; rsq[u] = fmaf(v[j], v[j], rsq[u]); }
          }
        }
        if (p.stats) {      // uniform across the CTA; rows of one warp lie in one image (stats_rows % 32 == 0)
          const int bimg = mrow0 / p.stats_rows;
          stats_commit(ssum, ssq, lane, p.stats + ((size_t)bimg * ncols + (size_t)(n < ncols ? n : 0)) * 2, nvalid);
        }
      }
      if (p.rowstats_out) {
        // fold the four column-group lanes of each row, then one fixed-point atomic pair per row and N tile
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          rsum[u] += __shfl_xor_sync(0xffffffffu, rsum[u], 1); rsq[u] += __shfl_xor_sync(0xffffffffu, rsq[u], 1);
          rsum[u] += __shfl_xor_sync(0xffffffffu, rsum[u], 2); rsq[u] += __shfl_xor_sync(0xffffffffu, rsq[u], 2);
          const int m = mrow0 + u * 8 + rsub;
          if (g == 0 && m < p.M) {
            atomicAdd(p.rowstats_out + 2 * (size_t)m, (unsigned long long)__float2ll_rn(rsum[u] * STATS_SCALE));
            atomicAdd(p.rowstats_out + 2 * (size_t)m + 1, (unsigned long long)__float2ll_rn(rsq[u] * STATS_SCALE));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int BN, bool CONV, bool GEGLU, bool LNF = false>
static int launch_tc(const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmB, const TcParams& p, cudaStream_t s) {
  using Cfg = TcCfg<BN>;
  static int smem_set[C2D_MAX_DEVICES] = {};
  if (int rc = ensure_dyn_smem(gemm_tc_kernel<BN, CONV, GEGLU, LNF>, Cfg::SMEM_BYTES, smem_set, "gemm_tc")) return rc;
  dim3 grid(ceil_div(p.N, BN), ceil_div(p.M, TC_BM));
  launch_pdl(gemm_tc_kernel<BN, CONV, GEGLU, LNF>, grid, dim3(TC_THREADS), Cfg::SMEM_BYTES, s, tmA, tmA2, tmB, p);
  return check_launch(CONV ? "conv3x3_tc" : (GEGLU ? "geglu_linear_tc" : "linear_tc"));
}


// =====================================================================================================
// Persistent variant: one CTA per SM loops over output tiles; the fp32 accumulator is DOUBLE-BUFFERED in
// TMEM so the epilogue of tile i (TMEM -> regs -> smem -> coalesced global) overlaps the main loop of tile
// i+1, and the TMA producer runs ahead across tile boundaries through a deeper smem ring.  Removes the
// per-tile prologue (TMEM alloc, barrier init, pipeline fill) that dominated the short-K projections.
// =====================================================================================================
template <int BN, int STAGES, bool PAIR = false>
struct Tc2Cfg {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;
  static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * TC_BK * 2;      // CTA pair: each CTA stages half of the B rows
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int ACC_STRIDE = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);   // TMEM columns between accumulators
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int EPI_PITCH = 36;                                          // floats per staged row (32 + pad)
  static constexpr int EPI_BYTES = TC_EPI_WARPS * 32 * EPI_PITCH * 4;
  static constexpr int OFF_EPI = STAGES * STAGE_BYTES;
  static constexpr int OFF_BIAS = OFF_EPI + EPI_BYTES;                          // 2 x BN floats
  static constexpr int OFF_BAR = OFF_BIAS + 4 * BN * 4;                   // bias + folded-LN column sums, double-buffered
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
};

// PAIR: the same persistent schedule on a CTA pair (cta_group::2, cluster of two CTAs on the SMs of a TPC): a tile is 256 rows
// x BN columns, each CTA stages its own 128 A rows and HALF of the B rows and keeps its 128 x BN accumulators (double-buffered)
// in its own TMEM; the leader issues the MMAs for both.  Per SM and k-block (128 + BN / 2) x 64 operand elements come
// through L2 -> shared memory instead of (128 + BN) x 64: the K = 320 GEGLU projection is bound by exactly that traffic.
// Protocol on top of the single-CTA one: TMA bytes of both CTAs are counted on the LEADER's full[s]; the leader's commits
// are multicast to both CTAs' empty[s] / tfull[b]; both CTAs' epilogue threads arrive on the LEADER's tempty[b].
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

template <int BN, int STAGES, bool CONV, bool GEGLU, bool LNF, bool PAIR>
__device__ __forceinline__ void gemm_tc2_body(const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmB,
                                              const TcParams& p, const int num_tiles, const int num_n) {
  using Cfg = Tc2Cfg<BN, STAGES, PAIR>;
  static_assert(!PAIR || !CONV, "the persistent pair kernel serves the linears (the convolutions use gemm_tc3_kernel)");
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  // tile walk: a work item is a 128-row (PAIR: 256-row) x BN-column tile; `first` / `step` in work items
  const int first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto tile_m0 = [&](int t) { return PAIR ? ((t / num_n) * 2 + (int)rank) * TC_BM : (t / num_n) * TC_BM; };
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;       // [2] accumulator ready
  uint64_t* tempty = tfull + 2;           // [2] accumulator drained (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* s_bias = reinterpret_cast<float*>(smem + Cfg::OFF_BIAS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], (PAIR ? 2 : 1) * 32 * TC_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_2sm<Cfg::TMEM_COLS>(tmem_slot);
    else tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();       // barrier inits and the TMEM allocation are visible to the peer CTA
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();           // everything above overlapped the predecessor's tail; global memory is touched from here on
  pdl_trigger();

  if (warp == 0) {
    // ===================== TMA producer (runs ahead across tiles; converged warp, elected lane issues) =====================
    uint32_t kbc = 0;
    for (int t = first; t < num_tiles; t += step) {
      const int m0 = tile_m0(t), n0 = (t % num_n) * BN;
      int b0 = 0, y0 = 0, x0 = 0;
      if (CONV) {
        b0 = m0 / p.HW;
        const int rem = m0 - b0 * p.HW;
        y0 = rem / p.W;
        x0 = rem - y0 * p.W;
      }
      for (int kb = 0; kb < p.num_k_blocks; ++kb, ++kbc) {
        const int s = kbc % STAGES;
        mbar_wait(&empty[s], ((kbc / STAGES) & 1u) ^ 1u);
        if (elect_one()) {
          uint8_t* sA = smem + s * Cfg::STAGE_BYTES;
          uint8_t* sB = sA + Cfg::A_BYTES;
          if constexpr (PAIR) {
            if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * Cfg::STAGE_BYTES);      // both CTAs' bytes land on the leader
            const uint32_t lbar = mapa_rank0(smem_u32(&full[s]));
            if (kb < p.kb_split) tma_load_2d_2sm(sA, &tmA, lbar, kb * TC_BK, m0);
            else tma_load_2d_2sm(sA, &tmA2, lbar, (kb - p.kb_split) * TC_BK, m0);
            tma_load_2d_2sm(sB, &tmB, lbar, kb * TC_BK, n0 + (int)rank * (BN / 2));
          } else {
            mbar_arrive_expect_tx(&full[s], Cfg::STAGE_BYTES);
            if (CONV) {
              const int tap = kb / p.cblocks;
              const int c0 = (kb - tap * p.cblocks) * TC_BK;
              const int ky = tap / 3, kx = tap - ky * 3;
              tma_load_4d(sA, &tmA, &full[s], c0, x0 * p.cstride + kx - p.cpad, y0 * p.cstride + ky - p.cpad, b0);
              tma_load_2d(sB, &tmB, &full[s], tap * p.Cin + c0, n0);
            } else {
              if (kb < p.kb_split) tma_load_2d(sA, &tmA, &full[s], kb * TC_BK, m0);
              else tma_load_2d(sA, &tmA2, &full[s], (kb - p.kb_split) * TC_BK, m0);
              tma_load_2d(sB, &tmB, &full[s], kb * TC_BK, n0);
            }
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 2 * TC_BM : TC_BM, BN, 0, 0);
    const uint64_t desc0 = make_desc_k_sw128(smem_u32(smem));
    uint32_t kbc = 0, it = 0;
    if (!PAIR || rank == 0)                                      // CTA pair: the leader issues for both
    for (int t = first; t < num_tiles; t += step, ++it) {
      const uint32_t buf = it & 1u;
      mbar_wait(&tempty[buf], ((it >> 1) & 1u) ^ 1u);          // epilogue (of both CTAs) drained this accumulator
      tc_fence_after();
      const uint32_t acc = tmem_base + buf * Cfg::ACC_STRIDE;
      for (int kb = 0; kb < p.num_k_blocks; ++kb, ++kbc) {
        const int s = kbc % STAGES;
        mbar_wait(&full[s], (kbc / STAGES) & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t a_desc = desc0 + (uint64_t)(s * (Cfg::STAGE_BYTES >> 4));
          const uint64_t b_desc = a_desc + (uint64_t)(Cfg::A_BYTES >> 4);
          if constexpr (PAIR) {
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k)
              umma_f16_2sm(acc, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_2sm(&empty[s]);                               // frees the slot in both CTAs
            if (kb == p.num_k_blocks - 1) umma_commit_2sm(&tfull[buf]);
          } else {
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k)
              umma_f16(acc, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&empty[s]);
            if (kb == p.num_k_blocks - 1) umma_commit(&tfull[buf]);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (warps 2..5), overlapped with the next tile's main loop =====================
    const int q = warp & 3, half = (warp - 2) >> 2;
    float* stage = reinterpret_cast<float*>(smem + Cfg::OFF_EPI) + (size_t)(warp - 2) * 32 * Cfg::EPI_PITCH;
    const int ncols = GEGLU ? (p.N >> 1) : p.N;
    const bool vec_y = (p.ldy % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0);
    const bool vec_r = p.residual && (p.ldr % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0);
    uint32_t it = 0;
    longlong2 ln_raw = make_longlong2(0, 0);
    if (LNF && first < num_tiles) ln_raw = ln_row_load(p, tile_m0(first) + q * 32 + lane);    // first tile's row statistics
    const uint32_t tempty_leader[2] = {PAIR ? mapa_rank0(smem_u32(&tempty[0])) : 0u, PAIR ? mapa_rank0(smem_u32(&tempty[1])) : 0u};
    auto acc_drained = [&](uint32_t buf) {       // this thread has read its last accumulator column of the tile
      tc_fence_before();
      if constexpr (PAIR) mbar_arrive_cluster(tempty_leader[buf]);
      else mbar_arrive(&tempty[buf]);
    };
    for (int t = first; t < num_tiles; t += step, ++it) {
      const uint32_t buf = it & 1u;
      const int m0 = tile_m0(t), n0 = (t % num_n) * BN;
      const int nbase = GEGLU ? (n0 >> 1) : n0;
      // this tile's bias slice (+ the row vector when the whole tile shares one) -> smem
      const bool rv_shared = p.rowvec && (m0 / p.rows_per_vec) == ((min(m0 + TC_BM, p.M) - 1) / p.rows_per_vec);
      float* sb = s_bias + buf * BN;
      float* scs = s_bias + 2 * BN + buf * BN;
      {
        const float* rv0 = rv_shared ? p.rowvec + (long long)(m0 / p.rows_per_vec) * p.N : nullptr;
        for (int j = threadIdx.x - 64; j < BN; j += 32 * TC_EPI_WARPS) {
          const int n = n0 + j;
          float bv = 0.f, cs = 0.f;
          if (n < p.N) {
            if (p.bias) bv = __ldg(p.bias + n);
            if (rv0) bv += __ldg(rv0 + n);
            if (p.ln_colsum) cs = __ldg(p.ln_colsum + n);
          }
          sb[j] = bv;
          scs[j] = cs;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
      }
      float ln_rstd = 1.f, ln_mr = 0.f;
      if (LNF) {
        ln_row_finish(p, ln_raw, ln_rstd, ln_mr);
        // the next tile's statistics travel while this tile drains
        const int tn = t + step;
        if (tn < num_tiles) ln_raw = ln_row_load(p, tile_m0(tn) + q * 32 + lane);
      }
      mbar_wait(&tfull[buf], (it >> 1) & 1u);
      tc_fence_after();
      const uint32_t t_row = tmem_base + buf * Cfg::ACC_STRIDE + ((uint32_t)(q * 32) << 16);
      constexpr int NCHUNK = GEGLU ? BN / 64 : BN / 32;            // GEGLU: 32 outputs per chunk = 32 a + 32 gate columns
      const int c_last = half + 2 * ((NCHUNK - 1 - half) / 2);      // this warp's last chunk
#pragma unroll 1
      for (int c = half; c < NCHUNK; c += 2) {
        // ---- TMEM -> registers -> (+bias, activation | GEGLU gate) -> per-warp fp32 staging [32 rows][32 cols]
        float v[32];
        if (!GEGLU) {
          uint32_t r[32];
          tmem_ld_32x32(t_row + c * 32, r);
          tmem_ld_wait();
          if (c == c_last) acc_drained(buf);                                     // this warp is done with the accumulator
          epi_affine32<LNF>(r, sb + c * 32, scs + c * 32, ln_rstd, -ln_mr, v);
          if (p.rowvec && !rv_shared) {
            const int m = m0 + q * 32 + lane;
            if (m < p.M) {
              const float* rv = p.rowvec + (long long)(m / p.rows_per_vec) * p.N + n0 + c * 32;
#pragma unroll
              for (int j = 0; j < 32; ++j) if (n0 + c * 32 + j < p.N) v[j] += __ldg(rv + j);
            }
          }
          epi_act32(v, p.act);
        } else {
          // accumulator columns: 128-column groups of [a (64) | gate (64)]
          const int ca = (c >> 1) * 128 + (c & 1) * 32;
          uint32_t ra[32], rg[32];
          tmem_ld_32x32(t_row + ca, ra);
          tmem_ld_32x32(t_row + ca + 64, rg);
          tmem_ld_wait();
          if (c == c_last) acc_drained(buf);
          float gg[32];
          epi_affine32<LNF>(ra, sb + ca, scs + ca, ln_rstd, -ln_mr, v);
          epi_affine32<LNF>(rg, sb + ca + 64, scs + ca + 64, ln_rstd, -ln_mr, gg);
#pragma unroll
          for (int j = 0; j < 32; j += 2) gelu_mul2(v[j], v[j + 1], gg[j], gg[j + 1]);
        }
        float* srow = stage + (size_t)lane * Cfg::EPI_PITCH;
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(srow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        // ---- coalesced drain: 32 rows x 4 groups of 8 columns; residual loads of all 4 items issued first
        uint4 res[4];
        const int rsub = lane >> 2, g = lane & 3;
        const int mrow0 = m0 + q * 32;
        const int n = nbase + c * 32 + g * 8;
        const int nvalid = ncols - n;
        if ((GEGLU || !p.stats) && vec_y && (!p.residual || vec_r) && (ncols % 8 == 0) && m0 + TC_BM <= p.M) {
          const long long row = mrow0 + rsub;
          drain32_fast(stage + (size_t)rsub * Cfg::EPI_PITCH + g * 8, Cfg::EPI_PITCH, p.y + row * p.ldy + n, p.ldy,
                       p.residual ? p.residual + row * p.ldr + n : nullptr, p.ldr, n < ncols);
          __syncwarp();
          continue;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int m = mrow0 + u * 8 + rsub;
          res[u] = make_uint4(0u, 0u, 0u, 0u);
          if (m < p.M && vec_r && nvalid >= 8) res[u] = *reinterpret_cast<const uint4*>(p.residual + (long long)m * p.ldr + n);
        }
        float ssum[8], ssq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { ssum[j] = 0.f; ssq[j] = 0.f; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int rr = u * 8 + rsub, m = mrow0 + rr;
          if (m >= p.M || nvalid <= 0) continue;
          const float* sp = stage + (size_t)rr * Cfg::EPI_PITCH + g * 8;
          const float4 f0 = *reinterpret_cast<const float4*>(sp), f1 = *reinterpret_cast<const float4*>(sp + 4);
          float o[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
          bf16* yp = p.y + (long long)m * p.ldy + n;
          if (p.residual) {
            if (vec_r && nvalid >= 8) {
              const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&res[u]);
#pragma unroll
              for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(hh[j]); o[2 * j] += f.x; o[2 * j + 1] += f.y; }
            } else {
              const bf16* rp = p.residual + (long long)m * p.ldr + n;
#pragma unroll
              for (int j = 0; j < 8; ++j) if (j < nvalid) o[j] += __bfloat162float(rp[j]);
            }
          }
          if (vec_y && nvalid >= 8) {
            Vec8<bf16>::store(yp, o);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) if (j < nvalid) yp[j] = __float2bfloat16_rn(o[j]);
          }
          if (!GEGLU && p.stats) {      // the GEGLU projection never feeds a GroupNorm: no channel statistics
#pragma unroll
            for (int j = 0; j < 8; ++j) { ssum[j] += o[j]; ssq[j] = fmaf(o[j], o[j], ssq[j]); }
          }
        }
        if (!GEGLU && p.stats) {
          const int bimg = mrow0 / p.stats_rows;
          stats_commit(ssum, ssq, lane, p.stats + ((size_t)bimg * ncols + (size_t)(n < ncols ? n : 0)) * 2, nvalid);
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();     // the peer may still read this CTA's smem / TMEM until its last MMA has retired
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_2sm<Cfg::TMEM_COLS>(tmem_base);
    else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int BN, int STAGES, bool CONV, bool GEGLU, bool LNF>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                const __grid_constant__ CUtensorMap tmB, const TcParams p, const int num_tiles, const int num_n) {
  gemm_tc2_body<BN, STAGES, CONV, GEGLU, LNF, false>(tmA, tmA2, tmB, p, num_tiles, num_n);
}

// persistent CTA pair (see gemm_tc2_body): num_tiles counts 256-row x BN tiles
template <int BN, int STAGES, bool GEGLU, bool LNF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
gemm_tc4_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                const __grid_constant__ CUtensorMap tmB, const TcParams p, const int num_tiles, const int num_n) {
  gemm_tc2_body<BN, STAGES, false, GEGLU, LNF, true>(tmA, tmA2, tmB, p, num_tiles, num_n);
}

template <int BN, int STAGES, bool CONV, bool GEGLU, bool LNF = false>
static int launch_tc2(const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmB, const TcParams& p, cudaStream_t s) {
  using Cfg = Tc2Cfg<BN, STAGES>;
  static_assert(Cfg::SMEM_BYTES <= 227 * 1024, "persistent GEMM smem");
  static int smem_set[C2D_MAX_DEVICES] = {};
  if (int rc = ensure_dyn_smem(gemm_tc2_kernel<BN, STAGES, CONV, GEGLU, LNF>, Cfg::SMEM_BYTES, smem_set, "gemm_tc2")) return rc;
  const int num_n = ceil_div(p.N, BN);
  const int num_tiles = num_n * ceil_div(p.M, TC_BM);
  const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
  launch_pdl(gemm_tc2_kernel<BN, STAGES, CONV, GEGLU, LNF>, dim3(grid), dim3(TC_THREADS), Cfg::SMEM_BYTES, s, tmA, tmA2, tmB, p, num_tiles, num_n);
  return check_launch(CONV ? "conv3x3_tc" : (GEGLU ? "geglu_linear_tc" : "linear_tc"));
}

template <int BN, int STAGES, bool GEGLU, bool LNF = false>
static int launch_tc4(const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmB, const TcParams& p, cudaStream_t s) {
  using Cfg = Tc2Cfg<BN, STAGES, true>;
  static_assert(Cfg::SMEM_BYTES <= 227 * 1024, "persistent pair GEMM smem");
  static int smem_set[C2D_MAX_DEVICES] = {};
  if (int rc = ensure_dyn_smem(gemm_tc4_kernel<BN, STAGES, GEGLU, LNF>, Cfg::SMEM_BYTES, smem_set, "gemm_tc4")) return rc;
  const int num_n = ceil_div(p.N, BN);
  const int num_tiles = num_n * ceil_div(ceil_div(p.M, TC_BM), 2);
  const int pairs = num_sms() / 2;
  const int clusters = num_tiles < pairs ? num_tiles : pairs;
  launch_pdl(gemm_tc4_kernel<BN, STAGES, GEGLU, LNF>, dim3(2 * clusters), dim3(TC_THREADS), Cfg::SMEM_BYTES, s, tmA, tmA2, tmB, p, num_tiles, num_n);
  return check_launch(GEGLU ? "geglu_linear_tc" : "linear_tc");
}

// =====================================================================================================
// CTA-pair variant (cta_group::2): a cluster of two CTAs on the two SMs of a TPC computes a 256 x BN tile.
// Each CTA TMA-loads its own 128 A rows and HALF of the B rows (BN / 2) and keeps its own 128 x BN fp32
// accumulator in TMEM; the leader CTA (cluster rank 0) issues tcgen05.mma.cta_group::2 for both.  Per SM and
// k-block that is (128 + BN/2) x 64 operand elements through shared memory instead of (128 + BN) x 64: the
// single-CTA SS MMA saturates the 128 B/clk shared-memory port at ~55 % tensor utilisation (measured), the pair
// does not.  Pipeline protocol: every CTA's producer waits on its OWN empty[s] (armed in both CTAs by the leader's
// multicast commit), all TMA bytes of a stage are counted on the LEADER's full[s]; two clusters are co-resident
// per SM pair so one tile's epilogue overlaps the other's main loop.
// =====================================================================================================
template <int BN, int STAGES>
struct Tc3Cfg {
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;
  static constexpr int B_BYTES = (BN / 2) * TC_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);
  static constexpr int EPI_PITCH = 36;                                          // floats per staged row (32 + pad)
  static constexpr int EPI_BYTES = TC_EPI_WARPS * 32 * EPI_PITCH * 4;           // aliases the pipeline stages
  static constexpr int OFF_BIAS = STAGES * STAGE_BYTES;                         // bias[BN] then folded-LayerNorm column sums[BN]
  static constexpr int OFF_BAR = OFF_BIAS + 2 * BN * 4;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static_assert(EPI_BYTES <= STAGES * STAGE_BYTES, "epilogue staging must fit in the pipeline smem");
};

template <int BN, int STAGES, bool CONV, bool GEGLU>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 2)
gemm_tc3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                const __grid_constant__ CUtensorMap tmB, const TcParams p, const int num_n) {
  using Cfg = Tc3Cfg<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* s_bias = reinterpret_cast<float*>(smem + Cfg::OFF_BIAS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int m0 = ((pair / num_n) * 2 + (int)rank) * TC_BM, n0 = (pair % num_n) * BN;
  // split-K: this cluster's slice of the reduction (the whole K without a partial buffer)
  const int kb_begin = p.partial ? (int)blockIdx.y * p.kb_per_split : 0;
  const int nkb = p.partial ? min(p.kb_per_split, p.num_k_blocks - kb_begin) : p.num_k_blocks;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();                       // barrier inits and the TMEM allocation are visible to the peer CTA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; converged warp, elected lane issues) =====================
    int b0 = 0, y0 = 0, x0 = 0;
    if (CONV) {
      b0 = m0 / p.HW;
      int rem = m0 - b0 * p.HW;
      y0 = rem / p.W;
      x0 = rem - y0 * p.W;
    }
    const int nb = n0 + (int)rank * (BN / 2);          // this CTA's half of the B rows
    for (int it = 0; it < nkb; ++it) {
      const int kb = kb_begin + it;
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      mbar_wait(&empty[s], ph ^ 1u);
      if (elect_one()) {
        uint8_t* sA = smem + s * Cfg::STAGE_BYTES;
        uint8_t* sB = sA + Cfg::A_BYTES;
        if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * Cfg::STAGE_BYTES);      // both CTAs' bytes land here
        const uint32_t lbar = mapa_rank0(smem_u32(&full[s]));
        if (CONV) {
          const int tap = kb / p.cblocks;
          const int c0 = (kb - tap * p.cblocks) * TC_BK;
          const int ky = tap / 3, kx = tap - ky * 3;
          tma_load_4d_2sm(sA, &tmA, lbar, c0, x0 * p.cstride + kx - p.cpad, y0 * p.cstride + ky - p.cpad, b0);
          tma_load_2d_2sm(sB, &tmB, lbar, tap * p.Cin + c0, nb);
        } else {
          if (kb < p.kb_split) tma_load_2d_2sm(sA, &tmA, lbar, kb * TC_BK, m0);
          else tma_load_2d_2sm(sA, &tmA2, lbar, (kb - p.kb_split) * TC_BK, m0);
          tma_load_2d_2sm(sB, &tmB, lbar, kb * TC_BK, nb);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * TC_BM, BN, 0, 0);
      const uint64_t desc0 = make_desc_k_sw128(smem_u32(smem));
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t a_desc = desc0 + (uint64_t)(s * (Cfg::STAGE_BYTES >> 4));
          const uint64_t b_desc = a_desc + (uint64_t)(Cfg::A_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            umma_f16_2sm(tmem_base, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2sm(&empty[s]);                                   // frees the slot in both CTAs
          if (kb == nkb - 1) umma_commit_2sm(tmem_full);                // both accumulators complete
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 of both CTAs) =====================
    const int q = warp & 3, half = (warp - 2) >> 2;
    float* stage = reinterpret_cast<float*>(smem) + (size_t)(warp - 2) * 32 * Cfg::EPI_PITCH;
    const int ncols = GEGLU ? (p.N >> 1) : p.N;
    const int nbase = GEGLU ? (n0 >> 1) : n0;
    const bool vec_y = (p.ldy % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0);
    const bool vec_r = p.residual && (p.ldr % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0);
    const bool rv_shared = p.rowvec && (m0 / p.rows_per_vec) == ((min(m0 + TC_BM, p.M) - 1) / p.rows_per_vec);
    {
      const float* rv0 = rv_shared ? p.rowvec + (long long)(m0 / p.rows_per_vec) * p.N : nullptr;
      for (int j = threadIdx.x - 64; j < BN; j += 32 * TC_EPI_WARPS) {
        const int n = n0 + j;
        float bv = 0.f, cs = 0.f;
        if (n < p.N) {
          if (p.bias) bv = __ldg(p.bias + n);
          if (rv0) bv += __ldg(rv0 + n);
          if (p.ln_colsum) cs = __ldg(p.ln_colsum + n);
        }
        s_bias[j] = bv;
        s_bias[BN + j] = cs;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");
    }
    // folded LayerNorm (GEGLU projection only in this kernel): y = rstd (acc - mean colsum) + bias
    float ln_rstd = 1.f, ln_mr = 0.f;
    if (GEGLU) ln_row_coeffs(p, m0 + q * 32 + lane, ln_rstd, ln_mr);
    mbar_wait(tmem_full, 0);      // all MMAs retired: the pipeline smem of BOTH CTAs is idle (staging aliases stage 0)
    tc_fence_after();
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    constexpr int NCHUNK = GEGLU ? BN / 64 : BN / 32;
    if (!GEGLU && p.partial) {
      // split-K slice: the raw fp32 accumulator goes to partial[z][m][n] through the same per-warp staging (coalesced
      // 32-byte pieces of 8 rows per store instruction); bias / residual / statistics belong to splitk_finish_kernel
      float* part = p.partial + (size_t)blockIdx.y * (size_t)p.M * (size_t)p.N;
#pragma unroll 1
      for (int c = half; c < NCHUNK; c += 2) {
        uint32_t r[32];
        tmem_ld_32x32(t_row + c * 32, r);
        tmem_ld_wait();
        float* srow = stage + (size_t)lane * Cfg::EPI_PITCH;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(srow + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        __syncwarp();
        const int rsub = lane >> 2, g = lane & 3;
        const int n = n0 + c * 32 + g * 8;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int rr = u * 8 + rsub, m = m0 + q * 32 + rr;
          if (m < p.M && n + 8 <= p.N) {              // split-K is only selected for N % 8 == 0
            const float* sp = stage + (size_t)rr * Cfg::EPI_PITCH + g * 8;
            float* dst = part + (size_t)m * p.N + n;
            *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>(sp);
            *reinterpret_cast<float4*>(dst + 4) = *reinterpret_cast<const float4*>(sp + 4);
          }
        }
        __syncwarp();
      }
    } else
#pragma unroll 1
    for (int c = half; c < NCHUNK; c += 2) {
      // ---- TMEM -> registers -> (+bias, activation | GEGLU gate) -> per-warp fp32 staging [32 rows][32 cols]
      float v[32];
      if (!GEGLU) {
        uint32_t r[32];
        tmem_ld_32x32(t_row + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + s_bias[c * 32 + j];
        if (p.rowvec && !rv_shared) {
          const int m = m0 + q * 32 + lane;
          if (m < p.M) {
            const float* rv = p.rowvec + (long long)(m / p.rows_per_vec) * p.N + n0 + c * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) if (n0 + c * 32 + j < p.N) v[j] += __ldg(rv + j);
          }
        }
        epi_act32(v, p.act);
      } else {
        // accumulator columns: 128-column groups of [a (64) | gate (64)]; output chunk c covers 32 outputs
        const int grp = c >> 1, sub = c & 1;
        uint32_t ra[32], rg[32];
        tmem_ld_32x32(t_row + grp * 128 + sub * 32, ra);
        tmem_ld_32x32(t_row + grp * 128 + 64 + sub * 32, rg);
        tmem_ld_wait();
#pragma unroll
        {
          float gg[32];
          if (p.ln_stats) {
            epi_affine32<true>(ra, s_bias + grp * 128 + sub * 32, s_bias + BN + grp * 128 + sub * 32, ln_rstd, -ln_mr, v);
            epi_affine32<true>(rg, s_bias + grp * 128 + 64 + sub * 32, s_bias + BN + grp * 128 + 64 + sub * 32, ln_rstd, -ln_mr, gg);
          } else {
            epi_affine32<false>(ra, s_bias + grp * 128 + sub * 32, s_bias, 1.f, 0.f, v);
            epi_affine32<false>(rg, s_bias + grp * 128 + 64 + sub * 32, s_bias, 1.f, 0.f, gg);
          }
#pragma unroll
          for (int j = 0; j < 32; j += 2) gelu_mul2(v[j], v[j + 1], gg[j], gg[j + 1]);
        }
      }
      float* srow = stage + (size_t)lane * Cfg::EPI_PITCH;
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(srow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      __syncwarp();
      // ---- coalesced drain: 32 rows x 4 groups of 8 columns; residual loads of all 4 items issued first
      uint4 res[4];
      const int rsub = lane >> 2, g = lane & 3;
      const int mrow0 = m0 + q * 32;
      const int n = nbase + c * 32 + g * 8;
      const int nvalid = ncols - n;
      if (vec_y && (!p.residual || vec_r) && (ncols % 8 == 0) && m0 + TC_BM <= p.M) {
        const long long row = mrow0 + rsub;
        if (p.stats) {
          float fs[8], fq[8];
          drain32_fast<true>(stage + (size_t)rsub * Cfg::EPI_PITCH + g * 8, Cfg::EPI_PITCH, p.y + row * p.ldy + n, p.ldy,
                             p.residual ? p.residual + row * p.ldr + n : nullptr, p.ldr, n < ncols, fs, fq);
          const int bimg = mrow0 / p.stats_rows;
          stats_commit(fs, fq, lane, p.stats + ((size_t)bimg * ncols + (size_t)(n < ncols ? n : 0)) * 2, nvalid);
        } else {
          drain32_fast(stage + (size_t)rsub * Cfg::EPI_PITCH + g * 8, Cfg::EPI_PITCH, p.y + row * p.ldy + n, p.ldy,
                       p.residual ? p.residual + row * p.ldr + n : nullptr, p.ldr, n < ncols);
        }
        __syncwarp();
        continue;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int m = mrow0 + u * 8 + rsub;
        res[u] = make_uint4(0u, 0u, 0u, 0u);
        if (m < p.M && vec_r && nvalid >= 8) res[u] = *reinterpret_cast<const uint4*>(p.residual + (long long)m * p.ldr + n);
      }
      float ssum[8], ssq[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { ssum[j] = 0.f; ssq[j] = 0.f; }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int rr = u * 8 + rsub, m = mrow0 + rr;
        if (m >= p.M || nvalid <= 0) continue;
        const float* sp = stage + (size_t)rr * Cfg::EPI_PITCH + g * 8;
        const float4 f0 = *reinterpret_cast<const float4*>(sp), f1 = *reinterpret_cast<const float4*>(sp + 4);
        float o[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
        bf16* yp = p.y + (long long)m * p.ldy + n;
        if (p.residual) {
          if (vec_r && nvalid >= 8) {
            const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&res[u]);
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(hh[j]); o[2 * j] += f.x; o[2 * j + 1] += f.y; }
          } else {
            const bf16* rp = p.residual + (long long)m * p.ldr + n;
#pragma unroll
            for (int j = 0; j < 8; ++j) if (j < nvalid) o[j] += __bfloat162float(rp[j]);
          }
        }
        if (vec_y && nvalid >= 8) {
          Vec8<bf16>::store(yp, o);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) if (j < nvalid) yp[j] = __float2bfloat16_rn(o[j]);
        }
        if (p.stats) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { ssum[j] += o[j]; ssq[j] = fmaf(o[j], o[j], ssq[j]); }
        }
      }
      if (p.stats && mrow0 < p.M) {
        const int bimg = mrow0 / p.stats_rows;
        stats_commit(ssum, ssq, lane, p.stats + ((size_t)bimg * ncols + (size_t)(n < ncols ? n : 0)) * 2, nvalid);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  cluster_sync_all();             // the peer may still read this CTA's smem / TMEM until its last MMA has retired
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int BN, int STAGES, bool CONV, bool GEGLU>
static int launch_tc3(const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmB, const TcParams& p, cudaStream_t s) {
  using Cfg = Tc3Cfg<BN, STAGES>;
  static_assert(Cfg::SMEM_BYTES <= 113 * 1024, "two CTAs per SM");
  static int smem_set[C2D_MAX_DEVICES] = {};
  if (int rc = ensure_dyn_smem(gemm_tc3_kernel<BN, STAGES, CONV, GEGLU>, Cfg::SMEM_BYTES, smem_set, "gemm_tc3")) return rc;
  const int num_n = ceil_div(p.N, BN);
  const int pairs = num_n * ceil_div(ceil_div(p.M, TC_BM), 2);
  const int splits = p.partial ? ceil_div(p.num_k_blocks, p.kb_per_split) : 1;
  launch_pdl(gemm_tc3_kernel<BN, STAGES, CONV, GEGLU>, dim3(2 * pairs, splits), dim3(TC_THREADS), Cfg::SMEM_BYTES, s, tmA, tmA2, tmB, p, num_n);
  return check_launch(CONV ? "conv3x3_tc" : (GEGLU ? "geglu_linear_tc" : "linear_tc"));
}

// ---- split-K: y = sum_z partial[z] + bias + rowvec[m / rows_per_vec] + residual -> bf16, plus the per-channel
// fixed-point statistics the fused epilogue would have produced (integer atomics: order independent).
// Block = SKF_ROWS row lanes x (N / 8) column groups; thread = 8 consecutive columns, `iters` rows (stride SKF_ROWS):
// the column sums of a block's SKF_ROWS * iters rows cost ONE atomic pair per column (at M = 8192 a 4-row block spent
// its time in 1.3 M atomics per convolution).
constexpr int SKF_ROWS = 4;
__global__ void splitk_finish_kernel(const float* __restrict__ partial, int splits, const float* __restrict__ bias,
                                     const float* __restrict__ rowvec, int rows_per_vec, const bf16* __restrict__ residual,
                                     long long ldr, bf16* __restrict__ y, long long ldy, int M, int N,
                                     unsigned long long* __restrict__ stats, int stats_rows, int iters) {
  pdl_wait();
  pdl_trigger();
  const int ng = N >> 3;
  const int g = threadIdx.x % ng, rsub = threadIdx.x / ng;
  const int n = g * 8;
  const int row0 = blockIdx.x * SKF_ROWS * iters;
  float b8[8], cs[8], cq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { b8[j] = bias ? __ldg(bias + n + j) : 0.f; cs[j] = 0.f; cq[j] = 0.f; }
  for (int it = 0; it < iters; ++it) {
    const int m = row0 + it * SKF_ROWS + rsub;
    if (m >= M) break;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = b8[j];
    for (int z = 0; z < splits; ++z) {
      const float* src = partial + ((size_t)z * M + m) * N + n;
      const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    if (rowvec) {
      const float* rv = rowvec + (size_t)(m / rows_per_vec) * N + n;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += __ldg(rv + j);
    }
    if (residual) {
      float rr[8];
      Vec8<bf16>::load(residual + (size_t)m * ldr + n, rr);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += rr[j];
    }
    Vec8<bf16>::store(y + (size_t)m * ldy + n, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) { cs[j] += v[j]; cq[j] = fmaf(v[j], v[j], cq[j]); }
  }
  if (stats) {
    // column sums over the block's row lanes through shared memory (one pass per moment), then one atomic per column
    __shared__ float sh[SKF_ROWS][1280 + 8];
    const int bimg = row0 / stats_rows;                             // stats_rows % (SKF_ROWS * iters) == 0: one image per block
#pragma unroll 1
    for (int kind = 0; kind < 2; ++kind) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sh[rsub][n + j] = kind ? cq[j] : cs[j];
      __syncthreads();
      for (int col = threadIdx.x; col < N; col += blockDim.x) {
        float acc = 0.f;
#pragma unroll
        for (int r = 0; r < SKF_ROWS; ++r) acc += sh[r][col];
        atomicAdd(stats + ((size_t)bimg * N + col) * 2 + kind, (unsigned long long)__float2ll_rn(acc * STATS_SCALE));
      }
      __syncthreads();
    }
  }
}

// Split-K workspace: CALLER-owned (c2d_set_workspace), one per device; the library allocates nothing.  Without a
// workspace the small-plane convolutions simply stay un-split.
constexpr size_t SPLITK_WS_BYTES = (size_t)32 << 20;       // c2d_splitk_workspace_bytes(): fp32 slices of the largest split
float* splitk_workspace(size_t* bytes);                   // runtime.cu (per-device context)

// Kernel selection.  Measured on B200 (tools/bench_shapes.py): the persistent kernel wins on the GEGLU projection
// (N = 8C, short K: +20 %), the two-CTA-per-SM kernel wins elsewhere (two MMA issuers hide each other's
// barrier round trips).  C2D_GEMM=legacy | persistent forces one of them for A/B runs.
static int gemm_mode() {           // 0 = auto, 1 = legacy, 2 = persistent, 3 = CTA pair (cta_group::2)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("C2D_GEMM");
    v = !e ? 0 : (e[0] == 'l' ? 1 : (e[0] == 'p' ? 2 : (e[0] == '2' ? 3 : 0)));
  }
  return v;
}
// auto: the CTA-pair kernel for the 3x3 convolutions (long K, operand-bandwidth bound: +8..35 % measured), the
// two-CTA-per-SM single-CTA kernel for plain linears (short K: the pair's cluster syncs cost more than they save)
static bool use_pair(bool conv = false) {
  const int m = gemm_mode();
  return m == 3 || (m == 0 && conv);
}
// B-tile width of the pair kernel: N % 32 == 0 for M = 256 MMAs
static int pick_bn_pair(int N) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("C2D_PAIR_BN");
    forced = e ? atoi(e) : 0;
  }
  if (forced == 128 || (forced == 160 && N % 160 == 0) || (forced == 256 && N % 256 == 0)) return forced;
  if (N % 160 == 0) return 160;               // measured best on every UNet shape (BN = 256 leaves too few CTAs)
  if (N % 256 == 0) return 256;
  return 128;
}
static bool geglu_pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("C2D_GEGLU_PAIR");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}
// C2D_PPAIR=geglu | all: the persistent CTA-pair kernel (gemm_tc4_kernel) for the GEGLU projection / for every persistent linear.
// Measured at UNet batch 16 (tools/bench_shapes.py, graph-timed) against the single-CTA persistent kernel: it wins where the
// main loop is long -- GEGLU K = 1280: 81.4 -> 71.7 us (1498 TFLOP/s), linears 16384 x 640 x 1280: 28.9 -> 25.1, x 2560:
// 53.8 -> 50.2 us -- is level at K = 640 and LOSES at K = 320 (GEGLU 124 -> 139 us, QKV 66 -> 77 us): every 1.4 us tile
// pays the cross-CTA tfull / tempty round trips.  Worth ~0.07 ms of a 17.5 ms step in total, so it stays opt-in.
static int ppair_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("C2D_PPAIR");
    v = !e ? 0 : (e[0] == 'g' ? 1 : (e[0] == 'a' ? 2 : 0));
  }
  return v;
}
// auto: persistent for the GEGLU projection and for the large-M linears whose main loop is long enough (K >= 640)
// or whose output is wide enough (N >= 960) for the cross-tile prefetch to pay (measured: -5 .. -16 %)
static bool use_persistent(bool geglu = false, int M = 0, int N = 0, int K = 0) {
  const int m = gemm_mode();
  return m == 2 || (m == 0 && (geglu || (M >= 16384 && (K >= 640 || N >= 960))));
}

// C2D_SMALL_BN=0 keeps the small-M linears on the wide tiles (A/B runs)
static bool small_bn_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("C2D_SMALL_BN");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// tile width: widest tile that still yields at least one tile per SM, else 64 (more CTAs, deeper ring)
// 160 or 128 columns.  A 160-column tile drains as 5 chunks over 2 warp halves (3 + 2) against 4 (2 + 2): when the epilogue
// bounds the kernel (K <= 768) it costs 3 units per tile against 2, so it is chosen when it divides N (UNet widths) or when it
// saves a third of the tiles (N = 288: 2 tiles instead of 3; measured 419 -> 349 us at M = 2^20, K = 96, while N = 1152 went
// 89 -> 110 us on 8 x 160 instead of 9 x 128)
static int wide_bn(int N) {
  if (N % 160 == 0) return 160;
  return 3 * ceil_div(N, 160) <= 2 * ceil_div(N, 128) ? 160 : 128;
}
static int pick_bn(int M, int N) {
  const int mt = ceil_div(M, TC_BM);
  const int wide = wide_bn(N);
  if (mt * ceil_div(N, wide) >= num_sms() || N <= 64) return wide;
  return 64;
}

// GEGLU projection on the persistent kernel: 128 x 256 tiles.  One N = 256 MMA per k-step costs 162 issue cycles against
// 2 x 98 for two N = 128 ones, and a tile pays its fixed epilogue / hand-over latencies once per 256 columns (the K = 320
// sites are NOT tensor bound: knocking out 3 of 4 MMAs did not change their time; an A-stationary walk that cut their
// operand traffic by a third did not either -- DESIGN.md section 7).  Measured at UNet batch 16 (tools/bench_shapes.py), BN 128 -> 256:
// 143 -> 128, 113 -> 91, 103 -> 82, 31 -> 26 us for the four levels.  C2D_GEGLU_BN=128 forces the narrow tile (A/B runs).
static int pick_bn_geglu(int M, int N) {
  (void)M;
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("C2D_GEGLU_BN");
    forced = e ? atoi(e) : 0;
  }
  return (N % 256 == 0 && forced != 128) ? 256 : 128;
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

bool linear_tc_supported(const void* x, const void* w, int M, int N, int K, int ldx) {
  return M >= 1 && N >= 1 && K >= 8 && K % 8 == 0 && ldx % 8 == 0 && al16(x) && al16(w);
}

int linear_tc(const void* x, const void* w, const float* bias, const float* rowvec, int rows_per_vec,
              const void* residual, void* y, int M, int N, int K, int ldx, int ldy, int ldr, int act, bool geglu,
              const GemmExtras* ex, cudaStream_t s) {
  C2D_REQUIRE(linear_tc_supported(x, w, M, N, K, ldx),
              "linear_tc: needs bf16, K %% 8 == 0, ldx %% 8 == 0, 16B-aligned x/w (M=%d N=%d K=%d ldx=%d)", M, N, K, ldx);
  const bool cat = ex && ex->x2;
  const int K1 = cat ? ex->K1 : K;
  if (cat)
    C2D_REQUIRE(K1 > 0 && K1 < K && K1 % TC_BK == 0 && ex->ldx2 % 8 == 0 && al16(ex->x2),
                "linear_tc: K-concatenated input needs K1 %% 64 == 0 (K1=%d K=%d ldx2=%d)", K1, K, ex->ldx2);
  if (ex && ex->stats)
    C2D_REQUIRE(!geglu && ex->stats_rows > 0 && ex->stats_rows % 32 == 0 && M % ex->stats_rows == 0,
                "linear_tc: channel statistics need rows-per-image %% 32 == 0 (M=%d stats_rows=%d)", M, ex->stats_rows);
  const bool lnx = ex && (ex->ln_stats || ex->rowstats_out);      // only the single-CTA kernels carry these epilogues
  // C2D_GEGLU_PAIR=1: the GEGLU projection (with or without the folded LayerNorm) on CTA-pair 256 x 256 tiles, which
  // halve the operand re-streaming from L2.  Measured on B200 (same box, UNet batch 16): 2.16-2.20 ms per step against
  // 2.07-2.17 ms for the persistent single-CTA kernel (its cross-tile epilogue overlap is worth as much) -> off by default.
  const bool geglu_pair = geglu && geglu_pair_enabled() && N % 256 == 0 && M >= 256 && !(ex && (ex->rowstats_out || ex->stats || ex->x2));
  const bool pairk = (use_pair() && !lnx) || geglu_pair;
  const bool ln_or_rs = lnx;       // these epilogues exist in the BN = 160 / 128 single-CTA kernels only
  const bool persist = use_persistent(geglu, M, N, K) && !(ex && ex->rowstats_out);
  (void)ln_or_rs;
  const bool ppair = !pairk && persist && M >= 512 && ((geglu && ppair_mode() >= 1 && N % 256 == 0) || (!geglu && ppair_mode() == 2));
  int BN = pairk ? (geglu ? (N % 256 == 0 ? 256 : 128) : pick_bn_pair(N))
                 : (geglu ? (persist ? pick_bn_geglu(M, N) : 128) : ((persist && !(ex && ex->ln_stats)) ? pick_bn(M, N) : wide_bn(N)));
  // small M (the low-resolution levels, single-image latency): the wide tile leaves most SMs idle -> 64-column tiles
  if (ppair && !geglu) BN = (N % 160 == 0) ? 160 : 128;
  if (!pairk && !geglu && !persist && small_bn_enabled() && N % 64 == 0 &&
      ceil_div(M, TC_BM) * ceil_div(N, BN) * 2 <= num_sms())
    BN = 64;
  CUtensorMap tmA, tmA2, tmB;
  {
    uint64_t dims[2] = {(uint64_t)K1, (uint64_t)M};
    uint64_t st[1] = {(uint64_t)ldx * 2};
    uint32_t box[2] = {TC_BK, TC_BM};
    int rc = make_tmap_bf16(&tmA, x, 2, dims, st, box);
    if (rc) return rc;
  }
  if (cat) {
    uint64_t dims[2] = {(uint64_t)(K - K1), (uint64_t)M};
    uint64_t st[1] = {(uint64_t)ex->ldx2 * 2};
    uint32_t box[2] = {TC_BK, TC_BM};
    int rc = make_tmap_bf16(&tmA2, ex->x2, 2, dims, st, box);
    if (rc) return rc;
  } else {
    tmA2 = tmA;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t st[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {TC_BK, (uint32_t)((pairk || ppair) ? BN / 2 : BN)};
    int rc = make_tmap_bf16(&tmB, w, 2, dims, st, box);
    if (rc) return rc;
  }
  TcParams p = {};
  p.bias = bias; p.rowvec = rowvec; p.residual = reinterpret_cast<const bf16*>(residual); p.y = reinterpret_cast<bf16*>(y);
  p.M = M; p.N = N; p.K = K; p.ldy = ldy; p.ldr = ldr;
  p.rows_per_vec = rows_per_vec > 0 ? rows_per_vec : 1;
  p.act = act;
  p.num_k_blocks = ceil_div(K, TC_BK);
  p.kb_split = cat ? K1 / TC_BK : p.num_k_blocks;
  p.stats = ex ? reinterpret_cast<unsigned long long*>(ex->stats) : nullptr;
  p.stats_rows = ex && ex->stats ? ex->stats_rows : 1;
  if (ex && ex->ln_stats) {
    C2D_REQUIRE(ex->ln_colsum && !rowvec && act == C2D_ACT_NONE, "linear_tc: folded LayerNorm needs column sums, no row vector, no activation");
    p.ln_stats = ex->ln_stats; p.ln_colsum = ex->ln_colsum; p.ln_invK = 1.0f / (float)K; p.ln_eps = ex->ln_eps;
  }
  p.rowstats_out = ex ? reinterpret_cast<unsigned long long*>(ex->rowstats_out) : nullptr;
  if (pairk) {
    if (geglu) return BN == 256 ? launch_tc3<256, 3, false, true>(tmA, tmA2, tmB, p, s) : launch_tc3<128, 4, false, true>(tmA, tmA2, tmB, p, s);
    if (BN == 256) return launch_tc3<256, 3, false, false>(tmA, tmA2, tmB, p, s);
    if (BN == 160) return launch_tc3<160, 4, false, false>(tmA, tmA2, tmB, p, s);
    return launch_tc3<128, 4, false, false>(tmA, tmA2, tmB, p, s);
  }
  const bool lnf = p.ln_stats != nullptr;
  if (ppair) {
    if (geglu) return lnf ? launch_tc4<256, 5, true, true>(tmA, tmA2, tmB, p, s) : launch_tc4<256, 5, true, false>(tmA, tmA2, tmB, p, s);
    if (BN == 160) return lnf ? launch_tc4<160, 6, false, true>(tmA, tmA2, tmB, p, s) : launch_tc4<160, 6, false, false>(tmA, tmA2, tmB, p, s);
    return lnf ? launch_tc4<128, 6, false, true>(tmA, tmA2, tmB, p, s) : launch_tc4<128, 6, false, false>(tmA, tmA2, tmB, p, s);
  }
  if (geglu && lnf) {                                                                         // folded-LN GEGLU: persistent only
    if (BN == 256) return launch_tc2<256, 3, false, true, true>(tmA, tmA2, tmB, p, s);
    return launch_tc2<128, 5, false, true, true>(tmA, tmA2, tmB, p, s);
  }
  if (lnf) {
    if (persist) {
      if (BN == 160) return launch_tc2<160, 5, false, false, true>(tmA, tmA2, tmB, p, s);
      return launch_tc2<128, 5, false, false, true>(tmA, tmA2, tmB, p, s);
    }
    if (BN == 160) return launch_tc<160, false, false, true>(tmA, tmA2, tmB, p, s);
    if (BN == 64) return launch_tc<64, false, false, true>(tmA, tmA2, tmB, p, s);
    return launch_tc<128, false, false, true>(tmA, tmA2, tmB, p, s);
  }
  if (persist) {
    if (geglu) return BN == 256 ? launch_tc2<256, 3, false, true>(tmA, tmA2, tmB, p, s) : launch_tc2<128, 5, false, true>(tmA, tmA2, tmB, p, s);
    if (BN == 160) return launch_tc2<160, 5, false, false>(tmA, tmA2, tmB, p, s);
    if (BN == 64) return launch_tc2<64, 7, false, false>(tmA, tmA2, tmB, p, s);
    return launch_tc2<128, 5, false, false>(tmA, tmA2, tmB, p, s);
  }
  if (geglu) return launch_tc<128, false, true>(tmA, tmA2, tmB, p, s);
  if (BN == 160) return launch_tc<160, false, false>(tmA, tmA2, tmB, p, s);
  if (BN == 64) return launch_tc<64, false, false>(tmA, tmA2, tmB, p, s);
  return launch_tc<128, false, false>(tmA, tmA2, tmB, p, s);
}

static inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// C2D_SPLITK=0 keeps the small-M convolutions on the un-split kernel (A/B runs)
static bool splitk_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("C2D_SPLITK");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

bool conv3x3_tc_supported(const void* x, const void* w, int B, int H, int W, int Cin, int Cout, int stride, int up) {
  if (up || (stride != 1 && stride != 2)) return false;
  if (stride == 2 && ((H & 1) || (W & 1))) return false;
  const int Ho = H / stride, Wo = W / stride;
  // a tile is 128 consecutive output pixels in raster order and must be ONE TMA box (bw x bh x bb): whole multiples of
  // 128 columns, or rows of a power-of-two width <= 64 stacked 128 / Wo high (Ho a multiple of that), or whole small
  // planes (Ho * Wo dividing 128).  Ho itself need not be a power of two (e.g. 96 x 64 latents of a 768 x 512 image).
  bool geom;
  if (Wo >= 128) geom = Wo % 128 == 0;
  else if (128 % Wo) geom = false;
  else geom = Ho >= 128 / Wo ? Ho % (128 / Wo) == 0 : 128 % (Ho * Wo) == 0;
  return Cin % 8 == 0 && Cin >= 8 && Cout >= 1 && geom && al16(x) && al16(w) && (stride == 1 || Wo <= 128);
}

int conv3x3_tc(const void* x, const void* w, const float* bias, const float* rowvec, const void* residual, void* y, int B,
               int H, int W, int Cin, int Cout, int stride, long long* stats, cudaStream_t s, int pad) {
  C2D_REQUIRE(conv3x3_tc_supported(x, w, B, H, W, Cin, Cout, stride, 0),
              "conv3x3_tc: needs stride 1|2, output rows that tile into 128-pixel boxes, Cin %% 8 == 0 (H=%d W=%d Cin=%d Cout=%d)", H, W, Cin, Cout);
  const int Ho = H / stride, Wo = W / stride;
  const int bw = Wo < 128 ? Wo : 128;
  const int bh = (128 / bw) < Ho ? (128 / bw) : Ho;
  const int bb = 128 / (bw * bh);
  const bool pairk = use_pair(true);
  const int BN = pairk ? pick_bn_pair(Cout) : (gemm_mode() == 2 ? pick_bn(B * Ho * Wo, Cout) : ((Cout % 160 == 0) ? 160 : 128));
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t st[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    // box extents are in traversed input elements: with element stride 2 a box of 2*bw loads bw pixels
    uint32_t box[4] = {TC_BK, (uint32_t)(bw * stride), (uint32_t)(bh * stride), (uint32_t)bb};
    uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    int rc = make_tmap_bf16_strided(&tmA, x, 4, dims, st, box, es);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)9 * Cin, (uint64_t)Cout};
    uint64_t st[1] = {(uint64_t)9 * Cin * 2};
    uint32_t box[2] = {TC_BK, (uint32_t)(pairk ? BN / 2 : BN)};
    int rc = make_tmap_bf16(&tmB, w, 2, dims, st, box);
    if (rc) return rc;
  }
  TcParams p = {};
  p.bias = bias; p.rowvec = rowvec; p.residual = reinterpret_cast<const bf16*>(residual); p.y = reinterpret_cast<bf16*>(y);
  p.M = B * Ho * Wo; p.N = Cout; p.K = 9 * Cin; p.ldy = Cout; p.ldr = Cout;
  p.rows_per_vec = Ho * Wo;
  p.act = C2D_ACT_NONE;
  p.HW = Ho * Wo; p.W = Wo; p.Cin = Cin; p.cblocks = ceil_div(Cin, TC_BK); p.cstride = stride; p.cpad = pad;
  p.num_k_blocks = 9 * p.cblocks;
  p.kb_split = p.num_k_blocks;
  if (stats) C2D_REQUIRE((Ho * Wo) % 32 == 0, "conv3x3_tc: channel statistics need Ho*Wo %% 32 == 0 (%d)", Ho * Wo);
  p.stats = reinterpret_cast<unsigned long long*>(stats);
  p.stats_rows = Ho * Wo;
  if (pairk) {
    // Small planes (8x8 latents: M = B * 64): the tile grid does not fill the machine while K = 9 Cin is long ->
    // split the reduction over blockIdx.y (fp32 slices in the per-device workspace + one finishing pass).
    const int ctas = 2 * ceil_div(Cout, BN) * ceil_div(ceil_div(p.M, TC_BM), 2);
    size_t ws_bytes = 0;
    float* ws = splitk_workspace(&ws_bytes);
    int splits = 1;
    if (splitk_enabled() && ws && ctas * 2 <= num_sms() * 2 && Cout % 8 == 0 && Cout <= 1280 && p.num_k_blocks >= 32) {
      splits = (2 * num_sms() + ctas - 1) / ctas;                   // aim at two CTAs per SM
      if (splits > 8) splits = 8;
      while (splits > 1 && (size_t)splits * p.M * Cout * sizeof(float) > ws_bytes) --splits;
      if (p.stats_rows % SKF_ROWS) splits = 1;
    }
    if (splits > 1) {
      TcParams q = p;
      q.partial = ws;
      q.kb_per_split = ceil_div(p.num_k_blocks, splits);
      q.bias = nullptr; q.rowvec = nullptr; q.residual = nullptr; q.stats = nullptr;
      int rc = BN == 256 ? launch_tc3<256, 3, true, false>(tmA, tmA, tmB, q, s)
                         : (BN == 160 ? launch_tc3<160, 4, true, false>(tmA, tmA, tmB, q, s) : launch_tc3<128, 4, true, false>(tmA, tmA, tmB, q, s));
      if (rc) return rc;
      const int nsl = ceil_div(p.num_k_blocks, q.kb_per_split);
      const int threads = SKF_ROWS * (Cout / 8);
      int iters = 1;                                     // rows per block = 4 * iters: ~two blocks per SM, <= 32 rows, one image
      while (iters < 8 && ceil_div(p.M, SKF_ROWS * iters * 2) >= 2 * num_sms() && p.stats_rows % (SKF_ROWS * iters * 2) == 0) iters *= 2;
      launch_pdl(splitk_finish_kernel, dim3(ceil_div(p.M, SKF_ROWS * iters)), dim3(threads), 0, s, (const float*)ws, nsl, bias, rowvec,
                 p.rows_per_vec, p.residual, (long long)p.ldr, p.y, (long long)p.ldy, p.M, Cout, p.stats, p.stats_rows, iters);
      return check_launch("conv3x3_tc");
    }
    if (BN == 256) return launch_tc3<256, 3, true, false>(tmA, tmA, tmB, p, s);
    if (BN == 160) return launch_tc3<160, 4, true, false>(tmA, tmA, tmB, p, s);
    return launch_tc3<128, 4, true, false>(tmA, tmA, tmB, p, s);
  }
  if (gemm_mode() == 2) {
    if (BN == 160) return launch_tc2<160, 5, true, false>(tmA, tmA, tmB, p, s);
    if (BN == 64) return launch_tc2<64, 7, true, false>(tmA, tmA, tmB, p, s);
    return launch_tc2<128, 5, true, false>(tmA, tmA, tmB, p, s);
  }
  if (BN == 160) return launch_tc<160, true, false>(tmA, tmA, tmB, p, s);
  return launch_tc<128, true, false>(tmA, tmA, tmB, p, s);
}

long long splitk_workspace_bytes() { return (long long)SPLITK_WS_BYTES; }

int init_tc(int device) {
  (void)device;
  return get_encode();
}

}  // namespace c2d
