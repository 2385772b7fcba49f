// tcgen05 flash-attention forward for sm_100a (bf16 in, fp32 softmax / accumulate).
//   o[b, n, h*d + :] = softmax(q[b,n,h,:] . k[b,:,h,:]^T * scale) v[b,:,h,:]
// Used for the UNet's spatial self-attention (N = 4096/1024/256/64, d = 40/80/160, packed QKV views) and
// for the text/audio cross-attention core (Nkv = 77/81, cached K/V views).
//
// One CTA = 128 queries of one (batch, head); keys are streamed in tiles of 64.
//   warp 0      TMA producer: Q once, K and V tiles through separate mbarrier rings.  Each operand is a 4-D
//               tensor map [B][N][heads][d]; boxes are 64 elements wide so columns >= d are zero-filled by TMA --
//               head dims 40 / 80 / 160 need no padding in memory and land as canonical 128B-swizzled tiles.
//   warp 1      MMA issuer:  S[t&1] = Q K_t^T  (M=128, N=64, K = ceil16(d); both K-major)          -> TMEM
//                            O    += P_t V_t   (M=128, N = ceil16(d), K=64; P K-major smem, V MN-major) -> TMEM
//               S is double-buffered in TMEM and QK runs two tiles ahead of the softmax.
//   warps 2..5  softmax: thread = query row.  ONE pass over S in the common case: p = exp2(s*scale - m_ref)
//               against a LAGGING reference max; the largest probability is tracked on the packed bf16 bits,
//               and only when it exceeds 2^8 is the tile redone with a new reference and the O accumulator
//               rescaled in TMEM.  P (bf16) is written back IN PLACE over its S buffer in TMEM and is the A operand
//               of the P V MMA (TS form): no shared-memory round trip, no proxy fence; the score MMA of tile t+2 is
//               issued behind P V (t) (tcgen05.mma retires in issue order).  Final 1/l normalisation and 16-byte
//               bf16 stores.
// For d <= 128 (S 2x64 + O <= 128 TMEM columns, <= 97 KB smem) two CTAs share an SM.
// Measured alternatives that were slower on B200 (kept out): S triple-buffering, row sums on the tensor core
// (P x ones), 8 softmax warps with a per-key-half split -- the kernel is bounded by TMEM-read + MUFU.EX2
// throughput (ncu: pipe_tc ~58 %, xu ~50 %), not by warp-level latency hiding.
#include <float.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace c2d {

using namespace tc;

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);

constexpr int AT_BQ = 128, AT_BK = 64, AT_THREADS = 192;
constexpr int AT_QTILE = 128 * 128;     // 128 rows x 128 B
constexpr int AT_KTILE = 64 * 128;      // 64 rows  x 128 B
constexpr int AT_SB = 2;                // S ring depth in TMEM

struct AttnTcParams {
  bf16* o;
  int Nq, Nkv, d, npv;          // npv = ceil16(d): N extent of the PV MMA
  int tmem_cols;                // power of two >= 2*64 + npv
  long long ldo, bso;
  float scale_log2;             // softmax scale * log2(e)
};

template <int NBLK>
struct AtCfg {
  static constexpr int KS = NBLK >= 2 ? 2 : 3;               // K ring depth
  static constexpr int VS = NBLK >= 2 ? 2 : 3;               // V ring depth
  static constexpr int Q_BYTES = NBLK * AT_QTILE;
  static constexpr int KV_BYTES = NBLK * AT_KTILE;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + Q_BYTES;
  static constexpr int OFF_V = OFF_K + KS * KV_BYTES;
  static constexpr int OFF_BAR = OFF_V + VS * KV_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int NBLK>
__global__ void __launch_bounds__(AT_THREADS, NBLK <= 2 ? 2 : 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
  using Cfg = AtCfg<NBLK>;
  constexpr int KS = Cfg::KS, VS = Cfg::VS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;      // [3]
  uint64_t* k_empty = bars + 4;     // [3]
  uint64_t* v_full = bars + 7;      // [3]
  uint64_t* v_empty = bars + 10;    // [3]
  uint64_t* s_full = bars + 13;     // [2]
  uint64_t* p_full = bars + 15;     // [2]
  uint64_t* p_empty = bars + 17;    // [2]  PV(t) retired (P[t&1] reusable, O up to date)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BQ, h = blockIdx.y, b = blockIdx.z;
  const int ntiles = (p.Nkv + AT_BK - 1) / AT_BK;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 3; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128); mbar_init(&p_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_n(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();
  const uint32_t tmem_o = tmem_base + AT_SB * AT_BK;

  // producer and MMA warps run converged and issue through elect_one() (see tc_common.cuh)
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
      for (int blk = 0; blk < NBLK; ++blk) tma_load_4d(smem + Cfg::OFF_Q + blk * AT_QTILE, &tmQ, q_full, blk * 64, h, q0, b);
    }
    __syncwarp();
    for (int t = 0; t < ntiles; ++t) {
      {
        const int s = t % KS;
        mbar_wait(&k_empty[s], ((uint32_t)(t / KS) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&k_full[s], Cfg::KV_BYTES);
#pragma unroll
          for (int blk = 0; blk < NBLK; ++blk)
            tma_load_4d(smem + Cfg::OFF_K + (s * NBLK + blk) * AT_KTILE, &tmK, &k_full[s], blk * 64, h, t * AT_BK, b);
        }
        __syncwarp();
      }
      {
        const int s = t % VS;
        mbar_wait(&v_empty[s], ((uint32_t)(t / VS) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&v_full[s], Cfg::KV_BYTES);
#pragma unroll
          for (int blk = 0; blk < NBLK; ++blk)
            tma_load_4d(smem + Cfg::OFF_V + (s * NBLK + blk) * AT_KTILE, &tmV, &v_full[s], blk * 64, h, t * AT_BK, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_qk = make_idesc_bf16(128, AT_BK, 0, 0);
    const uint32_t idesc_pv = make_idesc_bf16(128, p.npv, 0, 1);          // B (= V) is MN-major
    const uint64_t q_desc = make_desc_k_sw128(smem_u32(smem + Cfg::OFF_Q));
    const uint64_t k_desc = make_desc_k_sw128(smem_u32(smem + Cfg::OFF_K));
    const uint64_t v_desc = make_desc_mn_sw128(smem_u32(smem + Cfg::OFF_V), AT_KTILE, 1024);
    const int ksteps = (p.d + 15) >> 4;
    auto issue_qk = [&](int t) {
      const int s = t % KS;
      mbar_wait(&k_full[s], (uint32_t)(t / KS) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t tmem_s = tmem_base + (uint32_t)(t & 1) * AT_BK;
        const uint64_t kd = k_desc + (uint64_t)(s * (Cfg::KV_BYTES >> 4));
        for (int kk = 0; kk < ksteps; ++kk) {
          // 64-column blocks are AT_QTILE / AT_KTILE bytes apart; +32 B per K = 16 step inside the swizzle row
          const uint64_t qo = (uint64_t)((kk >> 2) * (AT_QTILE >> 4) + (kk & 3) * 2);
          const uint64_t ko = (uint64_t)((kk >> 2) * (AT_KTILE >> 4) + (kk & 3) * 2);
          umma_f16(tmem_s, q_desc + qo, kd + ko, idesc_qk, kk > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[t & 1]);
        umma_commit(&k_empty[s]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_qk(0);
    if (ntiles > 1) issue_qk(1);
    for (int j = 0; j < ntiles; ++j) {
      const int pb = j & 1, vs = j % VS;
      mbar_wait(&p_full[pb], (uint32_t)(j >> 1) & 1u);     // P(j) written, S(j) consumed
      mbar_wait(&v_full[vs], (uint32_t)(j / VS) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t tmem_p = tmem_base + (uint32_t)pb * AT_BK;          // bf16 P, in place over S[pb]
        const uint64_t vd = v_desc + (uint64_t)(vs * (Cfg::KV_BYTES >> 4));
#pragma unroll
        for (int kk = 0; kk < AT_BK / 16; ++kk) {
          // 16 keys = 8 TMEM columns of P; V tile: [64 keys][64-col blocks]; 16 keys = 2 groups of 8 rows (SBO = 1024 B),
          // col blocks 8 KB apart (LBO)
          umma_f16_ts(tmem_o, tmem_p + (uint32_t)kk * 8, vd + (uint64_t)(kk * (2048 >> 4)), idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&p_empty[pb]);
        umma_commit(&v_empty[vs]);
      }
      __syncwarp();
      // refill the S buffer that softmax(j) released.  (After PV(j), never before: the producer may need
      // v_empty from PV(j) before it can reach K(j+2).)
      if (j + AT_SB < ntiles) issue_qk(j + AT_SB);
    }
  } else {
    // ===================== softmax / correction / epilogue (warps 2..5) =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    float m_ref = -INFINITY, l_run = 0.f;
    for (int j = 0; j < ntiles; ++j) {
      const int pb = j & 1;
      const int kvalid = min(AT_BK, p.Nkv - j * AT_BK);
      const uint32_t tmem_s = tmem_base + (uint32_t)pb * AT_BK + lane_off;
      mbar_wait(&s_full[pb], (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      if (j == 0) {
        // first tile: the reference max must be known before any exponential is taken
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_s + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < kvalid) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
        m_ref = mx * p.scale_log2;
      }
      // S(j) -> registers once: P overwrites these columns in place, a redone tile works from the registers
      uint32_t ra[32], rb[32];
      tmem_ld_32x32(tmem_s, ra);
      tmem_ld_32x32(tmem_s + 32, rb);
      tmem_ld_wait();
      float sum;
      float alpha = 1.f;
      bool redo = false;
      for (;;) {
        sum = 0.f;
        uint32_t pmax2 = 0u;           // running max of the packed bf16 probabilities (p >= 0: bit patterns are ordered)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t (&r)[32] = c ? rb : ra;
          uint32_t packed[16];
          if (kvalid == AT_BK) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float p0 = ex2f(fmaf(__uint_as_float(r[i]), p.scale_log2, -m_ref));
              const float p1 = ex2f(fmaf(__uint_as_float(r[i + 1]), p.scale_log2, -m_ref));
              sum += p0 + p1;
              __nv_bfloat162 hb = __floats2bfloat162_rn(p0, p1);
              packed[i >> 1] = *reinterpret_cast<uint32_t*>(&hb);
              pmax2 = __vmaxu2(pmax2, packed[i >> 1]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float p0 = (c * 32 + i < kvalid) ? ex2f(fmaf(__uint_as_float(r[i]), p.scale_log2, -m_ref)) : 0.f;
              const float p1 = (c * 32 + i + 1 < kvalid) ? ex2f(fmaf(__uint_as_float(r[i + 1]), p.scale_log2, -m_ref)) : 0.f;
              sum += p0 + p1;
              __nv_bfloat162 hb = __floats2bfloat162_rn(p0, p1);
              packed[i >> 1] = *reinterpret_cast<uint32_t*>(&hb);
              pmax2 = __vmaxu2(pmax2, packed[i >> 1]);
            }
          }
          // 32 keys = 16 packed columns at [c * 16, c * 16 + 16) of this S buffer
          uint32_t lo[8], hi[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) { lo[q] = packed[q]; hi[q] = packed[8 + q]; }
          tmem_st_32x8(tmem_s + (uint32_t)c * 16, lo);
          tmem_st_32x8(tmem_s + (uint32_t)c * 16 + 8, hi);
        }
        if (redo) break;
        // largest probability of the row, as bf16 bits: > 2^8 means the row max ran ahead of the reference
        const uint32_t pm = max(pmax2 & 0xFFFFu, pmax2 >> 16);
        const bool need = pm > 0x4380u;                      // bf16(256.0)
        if (!__any_sync(0xffffffffu, need)) break;
        // new reference for the rows that moved: m_ref + log2(pmax); if the exponential overflowed, take the exact max
        float mnew = m_ref + __log2f(__uint_as_float(pm << 16));
        if (__any_sync(0xffffffffu, need && pm >= 0x7F80u)) {
          float mx = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (i < kvalid) mx = fmaxf(mx, __uint_as_float(ra[i]));
            if (32 + i < kvalid) mx = fmaxf(mx, __uint_as_float(rb[i]));
          }
          mnew = mx * p.scale_log2;
        }
        if (need) {
          alpha = ex2f(m_ref - mnew);
          m_ref = mnew;
        }
        redo = true;
      }
      if (redo && j > 0) {
        // warp-collective rescale of the O accumulator (rows that did not move use alpha = 1); every PV issued so
        // far must have retired first
        mbar_wait(&p_empty[(j - 1) & 1], (uint32_t)((j - 1) >> 1) & 1u);
        tc_fence_after();
        for (int c = 0; c < p.npv; c += 16) {
          uint32_t o[16];
          tmem_ld_32x16(tmem_o + lane_off + c, o);
          tmem_ld_wait();
          uint32_t lo[8], hi[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            lo[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            hi[i] = __float_as_uint(__uint_as_float(o[8 + i]) * alpha);
          }
          tmem_st_32x8(tmem_o + lane_off + c, lo);
          tmem_st_32x8(tmem_o + lane_off + c + 8, hi);
        }
        tmem_st_wait();
      }
      l_run = l_run * alpha + sum;
      tmem_st_wait();                 // P (and a rescaled O) are in TMEM
      tc_fence_before();
      mbar_arrive(&p_full[pb]);       // also tells the MMA warp that S[pb] has been consumed
    }
    // ---- epilogue: O / l -> bf16 -> global
    mbar_wait(&p_empty[(ntiles - 1) & 1], (uint32_t)((ntiles - 1) >> 1) & 1u);
    tc_fence_after();
    const int qrow = q0 + row;
    const float inv = 1.f / l_run;
    bf16* orow = p.o + (long long)b * p.bso + (long long)qrow * p.ldo + (long long)h * p.d;
    for (int c = 0; c < p.npv; c += 16) {
      uint32_t r[16];
      tmem_ld_32x16(tmem_o + lane_off + c, r);
      tmem_ld_wait();
      if (qrow < p.Nq) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (c + g * 8 < p.d) {       // d % 8 == 0: whole 8-element groups are valid or not
            uint4 o4;
            __nv_bfloat162* hb = reinterpret_cast<__nv_bfloat162*>(&o4);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              hb[i] = __floats2bfloat162_rn(__uint_as_float(r[g * 8 + 2 * i]) * inv, __uint_as_float(r[g * 8 + 2 * i + 1]) * inv);
            *reinterpret_cast<uint4*>(orow + c + g * 8) = o4;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_n(tmem_base, (uint32_t)p.tmem_cols);
  }
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

bool attention_tc_supported(const AttnParams& p, int B) {
  return p.mask == nullptr && p.d % 8 == 0 && p.d >= 16 && p.d <= 192 && p.ldq % 8 == 0 && p.ldk % 8 == 0 &&
         p.ldv % 8 == 0 && p.ldo % 8 == 0 && p.bsq % 8 == 0 && p.bsk % 8 == 0 && p.bsv % 8 == 0 && p.bso % 8 == 0 &&
         (B == 1 || (p.bsq > 0 && p.bsk > 0 && p.bsv > 0)) && al16(p.q) && al16(p.k) && al16(p.v) && al16(p.o) &&
         p.Nq >= 1 && p.Nkv >= 1;
}

static int make_head_tmap(CUtensorMap* m, const void* base, int d, int heads, int N, int B, long long ld, long long bs,
                          int box_rows) {
  uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)N, (uint64_t)B};
  uint64_t st[3] = {(uint64_t)d * 2, (uint64_t)ld * 2, (uint64_t)(B > 1 ? bs : (long long)N * ld) * 2};
  uint32_t box[4] = {64, 1, (uint32_t)box_rows, 1};
  return make_tmap_bf16(m, base, 4, dims, st, box);
}

template <int NBLK>
static int launch_attn_tc(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnTcParams& ap,
                          int heads, int B, cudaStream_t s) {
  using Cfg = AtCfg<NBLK>;
  static int smem_set[C2D_MAX_DEVICES] = {};
  if (int rc = ensure_dyn_smem(attn_tc_kernel<NBLK>, Cfg::SMEM_BYTES, smem_set, "attention_tc")) return rc;
  dim3 grid(ceil_div(ap.Nq, AT_BQ), heads, B);
  launch_pdl(attn_tc_kernel<NBLK>, grid, dim3(AT_THREADS), (size_t)Cfg::SMEM_BYTES, s, tq, tk, tv, ap);
  return check_launch("attn_tc");
}

bool attention_tc2_supported(const AttnParams& p, int B);
int attention_tc2(const AttnParams& p, int B, cudaStream_t s);

// C2D_ATTN=legacy keeps every shape on the first-generation kernel (A/B runs)
static bool attn_legacy() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("C2D_ATTN");
    v = (e && e[0] == 'l') ? 1 : 0;
  }
  return v == 1;
}

int attention_tc(const AttnParams& p, int B, cudaStream_t s) {
  if (!attn_legacy() && attention_tc2_supported(p, B)) return attention_tc2(p, B, s);
  CUtensorMap tq, tk, tv;
  int rc = make_head_tmap(&tq, p.q, p.d, p.heads, p.Nq, B, p.ldq, p.bsq, AT_BQ);
  if (rc) return rc;
  rc = make_head_tmap(&tk, p.k, p.d, p.heads, p.Nkv, B, p.ldk, p.bsk, AT_BK);
  if (rc) return rc;
  rc = make_head_tmap(&tv, p.v, p.d, p.heads, p.Nkv, B, p.ldv, p.bsv, AT_BK);
  if (rc) return rc;
  AttnTcParams ap;
  ap.o = reinterpret_cast<bf16*>(p.o);
  ap.Nq = p.Nq; ap.Nkv = p.Nkv; ap.d = p.d; ap.npv = (p.d + 15) & ~15;
  ap.ldo = p.ldo; ap.bso = p.bso;
  const int need = AT_SB * AT_BK + ap.npv;
  ap.tmem_cols = need <= 256 ? 256 : 512;
  ap.scale_log2 = p.scale * 1.4426950408889634f;
  const int nblk = (p.d + 63) / 64;
  if (nblk == 1) return launch_attn_tc<1>(tq, tk, tv, ap, p.heads, B, s);
  if (nblk == 2) return launch_attn_tc<2>(tq, tk, tv, ap, p.heads, B, s);
  return launch_attn_tc<3>(tq, tk, tv, ap, p.heads, B, s);
}

}  // namespace c2d
