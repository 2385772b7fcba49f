// tcgen05 flash-attention BACKWARD for sm_100a (bf16 in / out, fp32 accumulation) -- the dominant cost of the stage-3
// fine-tune step: on the FFMA kernels of train.cu the 4096-token self-attention adjoints were 217 of
// the 245 ms of a step.
//
//   S = scale Q K^T,  P = softmax(S),  O = P V,  D_i = <dO_i, O_i>
//   dV = P^T dO,  dP = dO V^T,  dS = P o (dP - D) scale,  dQ = dS K,  dK = dS^T Q
//
// Two kernels per call, both 128 x 128 score tiles, one CTA per (tile, head, batch), warp 0 = TMA producer (4-D head-view
// tensor maps as in attn_tc.cu: head dims 40 / 80 land zero-padded in canonical 128-byte-swizzled 64-column blocks),
// warp 1 = MMA issuer, warps 2..9 = two threads per score-tile row (64 of the 128 columns each: the row phase is as long
// as the MMA phase and nothing overlaps them, S / dP being single-buffered):
//   attn_bwd_dq_tc_kernel   CTA = 128 queries, two sweeps over the keys.  Sweep 1: S only -> online (max, sum) ->
//       log-sum-exp per row (kept for the second kernel) and D = <dO, O>.  Sweep 2: S and dP into TMEM ->
//       dS (bf16) written IN PLACE over S and fed to dQ += dS K as the TMEM A operand (TS form; K tile MN-major from
//       its natural layout) -- no shared-memory round trip.
//   attn_bwd_dkv_tc_kernel  CTA = 128 keys (K, V resident), loop over query tiles: S and dP into TMEM (lanes = queries)
//       -> P and dS (bf16) to shared memory in the swizzled [query row][key] layout, which IS the MN-major A operand
//       of dV += P^T dO and dK += dS^T Q (M = keys, K = queries); dO / Q tiles serve as K-major A operands of the score
//       MMAs and as MN-major B operands of the accumulations from the same shared-memory image.
// TMEM: S 128 | dP 128 | accumulators 64 * NBLK (dQ) or 2 x 64 * NBLK (dV, dK) columns.
// Head dims 129..192 (the 1280-channel levels: d = 160) run the dQ kernel with NBLK = 3 as is; the dK / dV kernel would
// need 640 TMEM columns and 256 KB of shared memory there, so it is launched twice (MODE 1: dV only -- S, no V; MODE 2:
// dK only), one extra score GEMM on layers that hold 1 / 16 of the tokens.
// Recomputing S / dP in both kernels costs 2 extra GEMMs of 7; it avoids fp32 atomics on dQ and a dQ conversion pass.
#include <float.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace c2d {

using namespace tc;

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);

constexpr int BW_T = 128;                 // tile rows (queries or keys)
constexpr int BW_BLK = 128 * 128;         // bytes of one 128-row x 64-column bf16 block
constexpr int BW_ROW_WARPS = 8;          // two warps per TMEM lane quadrant: each takes 64 of the 128 score columns of its rows
constexpr int BW_ROW_THREADS = 32 * BW_ROW_WARPS;
constexpr int BW_THREADS = 64 + BW_ROW_THREADS;

struct BwParams {
  bf16 *dq, *dk, *dv;
  const bf16 *o, *dout;
  float *lse, *delta;                     // [B][heads][Nq], log2 domain / plain
  int Nq, Nkv, d, npv, heads;
  long long lddq, bsdq, lddk, bsdk, lddv, bsdv, ldo, bso, lddo, bsdo;
  float scale, scale_log2;
  int have_lse;                           // lse[] comes from the forward pass: the dQ kernel skips its first sweep
};

__device__ __forceinline__ float bw_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t bw_pack(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void bw_sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// TMEM accumulator [128 lanes][npv fp32 columns] -> bf16 rows in global memory (row = lane; d % 8 == 0)
// (`half`: the two warps of a lane quadrant take alternate 16-column chunks)
__device__ __forceinline__ void bw_store_acc(uint32_t taddr, bf16* row_ptr, bool row_ok, int d, int npv, int half) {
  for (int c = 16 * half; c < npv; c += 32) {
    uint32_t r[16];
    tmem_ld_32x16(taddr + (uint32_t)c, r);
    tmem_ld_wait();
    if (row_ok) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (c + g * 8 < d) {
          uint4 o4;
          o4.x = bw_pack(__uint_as_float(r[g * 8 + 0]), __uint_as_float(r[g * 8 + 1]));
          o4.y = bw_pack(__uint_as_float(r[g * 8 + 2]), __uint_as_float(r[g * 8 + 3]));
          o4.z = bw_pack(__uint_as_float(r[g * 8 + 4]), __uint_as_float(r[g * 8 + 5]));
          o4.w = bw_pack(__uint_as_float(r[g * 8 + 6]), __uint_as_float(r[g * 8 + 7]));
          *reinterpret_cast<uint4*>(row_ptr + c + g * 8) = o4;
        }
      }
    }
  }
}

// ================================================================================================ dQ (+ lse, D)
template <int NBLK>
struct BwQCfg {
  static constexpr int KS = NBLK == 1 ? 2 : 1;                 // K / V ring depth
  static constexpr int TILE = NBLK * BW_BLK;
  static constexpr int OFF_Q = 0, OFF_DO = TILE, OFF_K = 2 * TILE, OFF_V = OFF_K + KS * TILE;
  static constexpr int OFF_BAR = OFF_V + KS * TILE;
  static constexpr int OFF_XCH = OFF_BAR + 256;                 // float2[2][128]: the two column halves' (max, sum) of sweep 1
  static constexpr int SMEM_BYTES = OFF_XCH + 2 * 128 * 8 + 1024;
};

template <int NBLK>
__global__ void __launch_bounds__(BW_THREADS, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO, const BwParams p) {
  using Cfg = BwQCfg<NBLK>;
  constexpr int KS = Cfg::KS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* qo_full = bars + 0;
  uint64_t* kv_full = bars + 1;        // [2]
  uint64_t* kv_empty = bars + 3;       // [2]
  uint64_t* sd_full = bars + 5;        //      S (and dP) of the current key tile are in TMEM
  uint64_t* ds_full = bars + 6;        //      128 arrivals: S / dP consumed (sweep 2: dS written)
  uint64_t* dq_done = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BW_T, h = blockIdx.y, b = blockIdx.z;
  const int T = (p.Nkv + BW_T - 1) / BW_T;
  const int c0 = p.have_lse ? T : 0;       // first step of the (sweep 1 | sweep 2) sequence; ring / phase counters run on c - c0

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV); prefetch_tmap(&tmDO);
    mbar_init(qo_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(sd_full, 1);
    mbar_init(ds_full, BW_ROW_THREADS);
    mbar_init(dq_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + 128, tmem_acc = tmem_base + 256;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(qo_full, 2 * Cfg::TILE);
#pragma unroll
      for (int blk = 0; blk < NBLK; ++blk) {
        tma_load_4d(smem + Cfg::OFF_Q + blk * BW_BLK, &tmQ, qo_full, blk * 64, h, q0, b);
        tma_load_4d(smem + Cfg::OFF_DO + blk * BW_BLK, &tmDO, qo_full, blk * 64, h, q0, b);
      }
    }
    __syncwarp();
    for (int c = c0; c < 2 * T; ++c) {
      const int ci = c - c0;
      const int s = ci % KS, t = c < T ? c : c - T;
      const bool sweep2 = c >= T;
      mbar_wait_backoff(&kv_empty[s], ((uint32_t)(ci / KS) & 1u) ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(&kv_full[s], sweep2 ? 2 * Cfg::TILE : Cfg::TILE);
#pragma unroll
        for (int blk = 0; blk < NBLK; ++blk) {
          tma_load_4d(smem + Cfg::OFF_K + s * Cfg::TILE + blk * BW_BLK, &tmK, &kv_full[s], blk * 64, h, t * BW_T, b);
          if (sweep2) tma_load_4d(smem + Cfg::OFF_V + s * Cfg::TILE + blk * BW_BLK, &tmV, &kv_full[s], blk * 64, h, t * BW_T, b);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_s = make_idesc_bf16(128, BW_T, 0, 0);
    const uint32_t idesc_dq = make_idesc_bf16(128, p.npv, 0, 1);          // B (= K tile) is MN-major
    const uint64_t q_desc = make_desc_k_sw128(smem_u32(smem + Cfg::OFF_Q));
    const uint64_t do_desc = make_desc_k_sw128(smem_u32(smem + Cfg::OFF_DO));
    const int ksteps = (p.d + 15) >> 4;
    mbar_wait(qo_full, 0);
    for (int c = c0; c < 2 * T; ++c) {
      const int ci = c - c0;
      const int s = ci % KS;
      const bool sweep2 = c >= T;
      if (c > c0) {
        mbar_wait(ds_full, (uint32_t)(ci - 1) & 1u);      // the row warps are done with S / dP of tile c-1 (sweep 2: dS written)
        tc_fence_after();
        if (c > T) {
          // dQ += dS(c-1) K(c-1): dS is bf16 in place over S (TS form), K tile MN-major from the stage of c-1.  Issued (and
          // its stage released) BEFORE waiting for this tile's operands: with a single stage the producer needs it.
          if (elect_one()) {
            const int sp = (ci - 1) % KS;
            const uint64_t kmn = make_desc_mn_sw128(smem_u32(smem + Cfg::OFF_K + sp * Cfg::TILE), BW_BLK, 1024);
#pragma unroll
            for (int kk = 0; kk < BW_T / 16; ++kk)
              umma_f16_ts(tmem_acc, tmem_s + (uint32_t)kk * 8, kmn + (uint64_t)(kk * (2048 >> 4)), idesc_dq, (c - 1 > T || kk > 0) ? 1u : 0u);
            umma_commit(&kv_empty[sp]);
          }
          __syncwarp();
        }
      }
      mbar_wait(&kv_full[s], (uint32_t)(ci / KS) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t kd = make_desc_k_sw128(smem_u32(smem + Cfg::OFF_K + s * Cfg::TILE));
        const uint64_t vd = make_desc_k_sw128(smem_u32(smem + Cfg::OFF_V + s * Cfg::TILE));
        for (int kk = 0; kk < ksteps; ++kk) {
          const uint64_t off = (uint64_t)((kk >> 2) * (BW_BLK >> 4) + (kk & 3) * 2);
          umma_f16(tmem_s, q_desc + off, kd + off, idesc_s, kk > 0 ? 1u : 0u);
        }
        if (sweep2)
          for (int kk = 0; kk < ksteps; ++kk) {
            const uint64_t off = (uint64_t)((kk >> 2) * (BW_BLK >> 4) + (kk & 3) * 2);
            umma_f16(tmem_dp, do_desc + off, vd + off, idesc_s, kk > 0 ? 1u : 0u);
          }
        umma_commit(sd_full);
        if (!sweep2) umma_commit(&kv_empty[s]);            // sweep 1 only needs K for the score MMA
      }
      __syncwarp();
    }
    // last tile's dQ contribution
    mbar_wait(ds_full, (uint32_t)(2 * T - 1 - c0) & 1u);
    tc_fence_after();
    if (elect_one()) {
      const int sp = (2 * T - 1 - c0) % KS;
      const uint64_t kmn = make_desc_mn_sw128(smem_u32(smem + Cfg::OFF_K + sp * Cfg::TILE), BW_BLK, 1024);
#pragma unroll
      for (int kk = 0; kk < BW_T / 16; ++kk)
        umma_f16_ts(tmem_acc, tmem_s + (uint32_t)kk * 8, kmn + (uint64_t)(kk * (2048 >> 4)), idesc_dq, (T > 1 || kk > 0) ? 1u : 0u);
      umma_commit(dq_done);
    }
    __syncwarp();
  } else {
    // ===================== row warps (two threads per query row: score columns [64 half, 64 half + 64)) =====================
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int qrow = q0 + row;
    const bool row_ok = qrow < p.Nq;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const long long sidx = ((long long)b * p.heads + h) * p.Nq + qrow;
    float2* xch = reinterpret_cast<float2*>(smem + Cfg::OFF_XCH);
    const int col0 = half * 64;
    // D = <dO, O> from global memory while the first tiles are in flight
    float dl = 0.f;
    if (row_ok) {
      const bf16* orow = p.o + (long long)b * p.bso + (long long)qrow * p.ldo + (long long)h * p.d;
      const bf16* drow = p.dout + (long long)b * p.bsdo + (long long)qrow * p.lddo + (long long)h * p.d;
      for (int c = 0; c < p.d; c += 8) {
        const uint4 a = *reinterpret_cast<const uint4*>(orow + c);
        const uint4 g = *reinterpret_cast<const uint4*>(drow + c);
        const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
        const __nv_bfloat162* hg = reinterpret_cast<const __nv_bfloat162*>(&g);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 fa = __bfloat1622float2(ha[i]), fg = __bfloat1622float2(hg[i]);
          dl = fmaf(fa.x, fg.x, fmaf(fa.y, fg.y, dl));
        }
      }
      if (half == 0) p.delta[sidx] = dl;
    }
    float m = -FLT_MAX, l = 0.f, lse = 0.f;
    if (p.have_lse && row_ok) lse = p.lse[sidx];
    for (int c = c0; c < 2 * T; ++c) {
      const bool sweep2 = c >= T;
      const int t = sweep2 ? c - T : c;
      const int kvalid = min(BW_T, p.Nkv - t * BW_T);
      mbar_wait(sd_full, (uint32_t)(c - c0) & 1u);
      tc_fence_after();
      if (!sweep2) {
#pragma unroll 1
        for (int cc = col0; cc < col0 + 64; cc += 32) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_s + lane_off + (uint32_t)cc, r);
          tmem_ld_wait();
          float mx = m;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (cc + i < kvalid) mx = fmaxf(mx, __uint_as_float(r[i]) * p.scale_log2);
          float acc = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (cc + i < kvalid) acc += bw_ex2(fmaf(__uint_as_float(r[i]), p.scale_log2, -mx));
          l = l * bw_ex2(m - mx) + acc;
          m = mx;
        }
        tc_fence_before();
        mbar_arrive(ds_full);
        if (c == T - 1) {
          // fold the two column halves' (max, sum): log-sum-exp of the whole row
          xch[half * 128 + row] = make_float2(m, l);
          asm volatile("bar.sync 1, %0;" ::"n"(BW_ROW_THREADS) : "memory");
          const float2 o = xch[(half ^ 1) * 128 + row];
          const float mm = fmaxf(m, o.x);
          const float ll = (l > 0.f ? l * bw_ex2(m - mm) : 0.f) + (o.y > 0.f ? o.y * bw_ex2(o.x - mm) : 0.f);
          lse = mm + __log2f(ll);
          if (row_ok && half == 0) p.lse[sidx] = lse;
        }
      } else {
        // dS goes IN PLACE over S (packed: 32 keys -> 16 columns at cc / 2): the upper half's packed columns 32..63 lie over
        // S columns the lower-half warp of the same rows still has to read, so both warps load their whole halves first
        uint32_t pk[2][16];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int cc = col0 + u * 32;
          uint32_t rs[32], rd[32];
          tmem_ld_32x32(tmem_s + lane_off + (uint32_t)cc, rs);
          tmem_ld_32x32(tmem_dp + lane_off + (uint32_t)cc, rd);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float d0 = 0.f, d1 = 0.f;
            if (row_ok && cc + i < kvalid)
              d0 = bw_ex2(fmaf(__uint_as_float(rs[i]), p.scale_log2, -lse)) * (__uint_as_float(rd[i]) - dl) * p.scale;
            if (row_ok && cc + i + 1 < kvalid)
              d1 = bw_ex2(fmaf(__uint_as_float(rs[i + 1]), p.scale_log2, -lse)) * (__uint_as_float(rd[i + 1]) - dl) * p.scale;
            pk[u][i >> 1] = bw_pack(d0, d1);
          }
        }
        tc_fence_before();
        asm volatile("bar.sync %0, 64;" ::"r"(2 + quad) : "memory");       // the two warps of this lane quadrant
        tc_fence_after();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int cc = col0 + u * 32;
          uint32_t lo[8], hi[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { lo[i] = pk[u][i]; hi[i] = pk[u][8 + i]; }
          tmem_st_32x8(tmem_s + lane_off + (uint32_t)(cc >> 1), lo);
          tmem_st_32x8(tmem_s + lane_off + (uint32_t)(cc >> 1) + 8, hi);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(ds_full);
      }
    }
    mbar_wait(dq_done, 0);
    tc_fence_after();
    bw_store_acc(tmem_acc + lane_off, p.dq + (long long)b * p.bsdq + (long long)qrow * p.lddq + (long long)h * p.d, row_ok, p.d, p.npv, half);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ================================================================================================ dK, dV
template <int NBLK, int MODE>                                  // MODE 0: dV and dK, 1: dV only, 2: dK only
struct BwKCfg {
  static constexpr bool HAS_V = MODE != 1, HAS_P = MODE != 2, HAS_DS = MODE != 1;
  static constexpr int QS = NBLK == 1 ? 2 : 1;                 // Q / dO ring depth
  static constexpr int TILE = NBLK * BW_BLK;
  static constexpr int OFF_K = 0, OFF_V = TILE, OFF_Q = HAS_V ? 2 * TILE : TILE, OFF_DO = OFF_Q + QS * TILE;
  static constexpr int OFF_P = OFF_DO + QS * TILE, OFF_DS = OFF_P + (HAS_P ? 2 * BW_BLK : 0);
  static constexpr int OFF_BAR = OFF_DS + (HAS_DS ? 2 * BW_BLK : 0);
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static_assert(SMEM_BYTES <= 232448, "attention backward: shared memory");
  static_assert(256 + (MODE == 0 ? 2 : 1) * 64 * NBLK <= 512, "attention backward: TMEM columns");
};

template <int NBLK, int MODE>
__global__ void __launch_bounds__(BW_THREADS, 1)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO, const BwParams p) {
  using Cfg = BwKCfg<NBLK, MODE>;
  constexpr int QS = Cfg::QS;
  constexpr bool HAS_V = Cfg::HAS_V, HAS_P = Cfg::HAS_P, HAS_DS = Cfg::HAS_DS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* kv_full = bars + 0;
  uint64_t* qo_full = bars + 1;        // [2]
  uint64_t* qo_empty = bars + 3;       // [2]
  uint64_t* sd_full = bars + 5;
  uint64_t* pds_full = bars + 6;       //      128 arrivals: P and dS of the tile are in shared memory
  uint64_t* pds_free = bars + 7;       //      the accumulation MMAs have read them
  uint64_t* acc_done = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * BW_T, h = blockIdx.y, b = blockIdx.z;
  const int T = (p.Nq + BW_T - 1) / BW_T;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV); prefetch_tmap(&tmDO);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&qo_full[i], 1); mbar_init(&qo_empty[i], 1); }
    mbar_init(sd_full, 1);
    mbar_init(pds_full, BW_ROW_THREADS);
    mbar_init(pds_free, 1);
    mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + 128;
  const uint32_t tmem_dv = tmem_base + 256, tmem_dk = tmem_base + 256 + (MODE == 0 ? 64 * NBLK : 0);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(kv_full, (HAS_V ? 2 : 1) * Cfg::TILE);
#pragma unroll
      for (int blk = 0; blk < NBLK; ++blk) {
        tma_load_4d(smem + Cfg::OFF_K + blk * BW_BLK, &tmK, kv_full, blk * 64, h, k0, b);
        if (HAS_V) tma_load_4d(smem + Cfg::OFF_V + blk * BW_BLK, &tmV, kv_full, blk * 64, h, k0, b);
      }
    }
    __syncwarp();
    for (int i = 0; i < T; ++i) {
      const int s = i % QS;
      mbar_wait_backoff(&qo_empty[s], ((uint32_t)(i / QS) & 1u) ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(&qo_full[s], 2 * Cfg::TILE);
#pragma unroll
        for (int blk = 0; blk < NBLK; ++blk) {
          tma_load_4d(smem + Cfg::OFF_Q + s * Cfg::TILE + blk * BW_BLK, &tmQ, &qo_full[s], blk * 64, h, i * BW_T, b);
          tma_load_4d(smem + Cfg::OFF_DO + s * Cfg::TILE + blk * BW_BLK, &tmDO, &qo_full[s], blk * 64, h, i * BW_T, b);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_s = make_idesc_bf16(128, BW_T, 0, 0);
    const uint32_t idesc_acc = make_idesc_bf16(128, p.npv, 1, 1);         // A (P^T / dS^T) and B (dO / Q) both MN-major
    const uint64_t k_desc = make_desc_k_sw128(smem_u32(smem + Cfg::OFF_K));
    const uint64_t v_desc = make_desc_k_sw128(smem_u32(smem + Cfg::OFF_V));
    const uint64_t p_desc = make_desc_mn_sw128(smem_u32(smem + Cfg::OFF_P), BW_BLK, 1024);
    const uint64_t ds_desc = make_desc_mn_sw128(smem_u32(smem + Cfg::OFF_DS), BW_BLK, 1024);
    const int ksteps = (p.d + 15) >> 4;
    auto accumulate = [&](int i) {        // dV += P(i)^T dO(i),  dK += dS(i)^T Q(i)   (M = keys, N = d, K = 128 queries)
      const int s = i % QS;
      const uint64_t do_mn = make_desc_mn_sw128(smem_u32(smem + Cfg::OFF_DO + s * Cfg::TILE), BW_BLK, 1024);
      const uint64_t q_mn = make_desc_mn_sw128(smem_u32(smem + Cfg::OFF_Q + s * Cfg::TILE), BW_BLK, 1024);
      if (HAS_P) {
#pragma unroll
        for (int kk = 0; kk < BW_T / 16; ++kk) {
          const uint64_t off = (uint64_t)(kk * (2048 >> 4));
          umma_f16(tmem_dv, p_desc + off, do_mn + off, idesc_acc, (i > 0 || kk > 0) ? 1u : 0u);
        }
      }
      if (HAS_DS) {
#pragma unroll
        for (int kk = 0; kk < BW_T / 16; ++kk) {
          const uint64_t off = (uint64_t)(kk * (2048 >> 4));
          umma_f16(tmem_dk, ds_desc + off, q_mn + off, idesc_acc, (i > 0 || kk > 0) ? 1u : 0u);
        }
      }
      umma_commit(pds_free);
      umma_commit(&qo_empty[s]);
    };
    mbar_wait(kv_full, 0);
    for (int i = 0; i < T; ++i) {
      const int s = i % QS;
      if (i > 0) {
        mbar_wait(pds_full, (uint32_t)(i - 1) & 1u);       // P / dS of tile i-1 written, its S / dP consumed
        tc_fence_after();
        if (elect_one()) accumulate(i - 1);
        __syncwarp();
      }
      mbar_wait(&qo_full[s], (uint32_t)(i / QS) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t qd = make_desc_k_sw128(smem_u32(smem + Cfg::OFF_Q + s * Cfg::TILE));
        const uint64_t dd = make_desc_k_sw128(smem_u32(smem + Cfg::OFF_DO + s * Cfg::TILE));
        for (int kk = 0; kk < ksteps; ++kk) {
          const uint64_t off = (uint64_t)((kk >> 2) * (BW_BLK >> 4) + (kk & 3) * 2);
          umma_f16(tmem_s, qd + off, k_desc + off, idesc_s, kk > 0 ? 1u : 0u);
        }
        if (HAS_DS)
          for (int kk = 0; kk < ksteps; ++kk) {
            const uint64_t off = (uint64_t)((kk >> 2) * (BW_BLK >> 4) + (kk & 3) * 2);
            umma_f16(tmem_dp, dd + off, v_desc + off, idesc_s, kk > 0 ? 1u : 0u);
          }
        umma_commit(sd_full);
      }
      __syncwarp();
    }
    mbar_wait(pds_full, (uint32_t)(T - 1) & 1u);
    tc_fence_after();
    if (elect_one()) {
      accumulate(T - 1);
      umma_commit(acc_done);
    }
    __syncwarp();
  } else {
    // ===================== row warps (two threads per query row of the current tile, 64 keys each; key row in the epilogue) =====
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const int kvalid = min(BW_T, p.Nkv - k0);
    const uint32_t p_row = smem_u32(smem + Cfg::OFF_P) + (uint32_t)row * 128u;
    const uint32_t ds_row = smem_u32(smem + Cfg::OFF_DS) + (uint32_t)row * 128u;
    const uint32_t sw = (uint32_t)(row & 7);
    for (int i = 0; i < T; ++i) {
      const int qrow = i * BW_T + row;
      const bool row_ok = qrow < p.Nq;
      float lse = 0.f, dl = 0.f;
      if (row_ok) {
        const long long sidx = ((long long)b * p.heads + h) * p.Nq + qrow;
        lse = p.lse[sidx];
        dl = p.delta[sidx];
      }
      mbar_wait(sd_full, (uint32_t)i & 1u);
      if (i > 0) mbar_wait(pds_free, (uint32_t)(i - 1) & 1u);      // the accumulation MMAs of tile i-1 have read P / dS
      tc_fence_after();
#pragma unroll 1
      for (int cc = half * 64; cc < half * 64 + 64; cc += 32) {
        uint32_t rs[32], rd[32];
        tmem_ld_32x32(tmem_s + lane_off + (uint32_t)cc, rs);
        if (HAS_DS) tmem_ld_32x32(tmem_dp + lane_off + (uint32_t)cc, rd);
        tmem_ld_wait();
        uint32_t pp[16], pd[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float p0 = 0.f, p1 = 0.f;
          if (row_ok && cc + j < kvalid) p0 = bw_ex2(fmaf(__uint_as_float(rs[j]), p.scale_log2, -lse));
          if (row_ok && cc + j + 1 < kvalid) p1 = bw_ex2(fmaf(__uint_as_float(rs[j + 1]), p.scale_log2, -lse));
          if (HAS_P) pp[j >> 1] = bw_pack(p0, p1);
          if (HAS_DS) pd[j >> 1] = bw_pack(p0 * (__uint_as_float(rd[j]) - dl) * p.scale, p1 * (__uint_as_float(rd[j + 1]) - dl) * p.scale);
        }
        // keys [cc, cc + 32) of this query row: 64-key block cc / 64, 16-byte chunks (cc % 64) / 8 .. + 3, swizzled by the row
        const uint32_t blk = (uint32_t)(cc >> 6) * (uint32_t)BW_BLK;
        const uint32_t ch0 = (uint32_t)(cc & 63) >> 3;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t off = blk + (((ch0 + q) ^ sw) << 4);
          if (HAS_P) bw_sts128(p_row + off, pp[q * 4], pp[q * 4 + 1], pp[q * 4 + 2], pp[q * 4 + 3]);
          if (HAS_DS) bw_sts128(ds_row + off, pd[q * 4], pd[q * 4 + 1], pd[q * 4 + 2], pd[q * 4 + 3]);
        }
      }
      tc_fence_before();
      fence_proxy_async();             // generic-proxy stores -> visible to the tensor core's async-proxy reads
      mbar_arrive(pds_full);
    }
    mbar_wait(acc_done, 0);
    tc_fence_after();
    const int krow = k0 + row;
    const bool ok = krow < p.Nkv;
    if (HAS_P)
      bw_store_acc(tmem_dv + lane_off, p.dv + (long long)b * p.bsdv + (long long)krow * p.lddv + (long long)h * p.d, ok, p.d, p.npv, half);
    if (HAS_DS)
      bw_store_acc(tmem_dk + lane_off, p.dk + (long long)b * p.bsdk + (long long)krow * p.lddk + (long long)h * p.d, ok, p.d, p.npv, half);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ================================================================================================ host
static inline bool bw_al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

static int bw_head_tmap(CUtensorMap* m, const void* base, int d, int heads, int N, int B, long long ld, long long bs) {
  uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)N, (uint64_t)B};
  uint64_t st[3] = {(uint64_t)d * 2, (uint64_t)ld * 2, (uint64_t)(B > 1 ? bs : (long long)N * ld) * 2};
  uint32_t box[4] = {64, 1, (uint32_t)BW_T, 1};
  return make_tmap_bf16(m, base, 4, dims, st, box);
}

struct AttnBwdArgs {      // mirrors c2d_attention_bwd
  const void *q, *k, *v, *o, *dout;
  void *dq, *dk, *dv;
  float *lse, *delta;
  int B, heads, Nq, Nkv, d;
  long long ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv, bsq, bsk, bsv, bso, bsdo, bsdq, bsdk, bsdv;
  float scale;
  int have_lse;
};

bool attention_bwd_tc_supported(const AttnBwdArgs& a) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("C2D_ATTN_BWD_TC");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  const long long strides[] = {a.ldq, a.ldk, a.ldv, a.ldo, a.lddo, a.lddq, a.lddk, a.lddv, a.bsq, a.bsk, a.bsv, a.bso, a.bsdo, a.bsdq, a.bsdk, a.bsdv};
  for (long long s : strides)
    if (s % 8) return false;
  return enabled && a.d % 8 == 0 && a.d >= 16 && a.d <= 192 && bw_al16(a.q) && bw_al16(a.k) && bw_al16(a.v) && bw_al16(a.o) &&
         bw_al16(a.dout) && bw_al16(a.dq) && bw_al16(a.dk) && bw_al16(a.dv) && (a.B == 1 || (a.bsq > 0 && a.bsk > 0 && a.bsv > 0 && a.bsdo > 0));
}

template <int NBLK>
static int attention_bwd_tc_n(const AttnBwdArgs& a, const BwParams& p, cudaStream_t s) {
  CUtensorMap tq, tk, tv, td;
  int rc = bw_head_tmap(&tq, a.q, a.d, a.heads, a.Nq, a.B, a.ldq, a.bsq);
  if (rc) return rc;
  if ((rc = bw_head_tmap(&tk, a.k, a.d, a.heads, a.Nkv, a.B, a.ldk, a.bsk))) return rc;
  if ((rc = bw_head_tmap(&tv, a.v, a.d, a.heads, a.Nkv, a.B, a.ldv, a.bsv))) return rc;
  if ((rc = bw_head_tmap(&td, a.dout, a.d, a.heads, a.Nq, a.B, a.lddo, a.bsdo))) return rc;
  static int set_q[C2D_MAX_DEVICES] = {}, set_k[C2D_MAX_DEVICES] = {}, set_k2[C2D_MAX_DEVICES] = {};
  if ((rc = ensure_dyn_smem(attn_bwd_dq_tc_kernel<NBLK>, BwQCfg<NBLK>::SMEM_BYTES, set_q, "attention_bwd_tc"))) return rc;
  attn_bwd_dq_tc_kernel<NBLK><<<dim3(ceil_div(a.Nq, BW_T), a.heads, a.B), BW_THREADS, BwQCfg<NBLK>::SMEM_BYTES, s>>>(tq, tk, tv, td, p);
  if ((rc = check_launch("attention_bwd_tc"))) return rc;
  const dim3 kgrid(ceil_div(a.Nkv, BW_T), a.heads, a.B);
  if constexpr (NBLK <= 2) {
    if ((rc = ensure_dyn_smem(attn_bwd_dkv_tc_kernel<NBLK, 0>, BwKCfg<NBLK, 0>::SMEM_BYTES, set_k, "attention_bwd_tc"))) return rc;
    attn_bwd_dkv_tc_kernel<NBLK, 0><<<kgrid, BW_THREADS, BwKCfg<NBLK, 0>::SMEM_BYTES, s>>>(tq, tk, tv, td, p);
  } else {
    if ((rc = ensure_dyn_smem(attn_bwd_dkv_tc_kernel<NBLK, 1>, BwKCfg<NBLK, 1>::SMEM_BYTES, set_k, "attention_bwd_tc"))) return rc;
    if ((rc = ensure_dyn_smem(attn_bwd_dkv_tc_kernel<NBLK, 2>, BwKCfg<NBLK, 2>::SMEM_BYTES, set_k2, "attention_bwd_tc"))) return rc;
    attn_bwd_dkv_tc_kernel<NBLK, 1><<<kgrid, BW_THREADS, BwKCfg<NBLK, 1>::SMEM_BYTES, s>>>(tq, tk, tv, td, p);
    if ((rc = check_launch("attention_bwd_tc"))) return rc;
    attn_bwd_dkv_tc_kernel<NBLK, 2><<<kgrid, BW_THREADS, BwKCfg<NBLK, 2>::SMEM_BYTES, s>>>(tq, tk, tv, td, p);
  }
  return check_launch("attention_bwd_tc");
}

int attention_bwd_tc(const AttnBwdArgs& a, cudaStream_t s) {
  BwParams p;
  p.dq = reinterpret_cast<bf16*>(a.dq); p.dk = reinterpret_cast<bf16*>(a.dk); p.dv = reinterpret_cast<bf16*>(a.dv);
  p.o = reinterpret_cast<const bf16*>(a.o); p.dout = reinterpret_cast<const bf16*>(a.dout);
  p.lse = a.lse; p.delta = a.delta;
  p.Nq = a.Nq; p.Nkv = a.Nkv; p.d = a.d; p.npv = (a.d + 15) & ~15; p.heads = a.heads;
  p.lddq = a.lddq; p.bsdq = a.bsdq; p.lddk = a.lddk; p.bsdk = a.bsdk; p.lddv = a.lddv; p.bsdv = a.bsdv;
  p.ldo = a.ldo; p.bso = a.bso; p.lddo = a.lddo; p.bsdo = a.bsdo;
  p.scale = a.scale; p.scale_log2 = a.scale * 1.4426950408889634f;
  p.have_lse = a.have_lse;
  if (a.d <= 64) return attention_bwd_tc_n<1>(a, p, s);
  return a.d <= 128 ? attention_bwd_tc_n<2>(a, p, s) : attention_bwd_tc_n<3>(a, p, s);
}

}  // namespace c2d
