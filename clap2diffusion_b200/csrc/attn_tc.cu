// tcgen05 flash-attention forward for sm_100a (bf16 in, fp32 softmax / accumulate).
//   o[b, n, h*d + :] = softmax(q[b,n,h,:] . k[b,:,h,:]^T * scale) v[b,:,h,:]
// Used for the UNet's spatial self-attention (N = 4096/1024/256/64, d = 40/80/160, packed QKV views) and
// for the text/audio cross-attention core (Nkv = 77/81, cached K/V views).
//
// One CTA = 128 queries of one (batch, head); keys are streamed in tiles of 128.
//   warp 0      TMA producer: Q once, K (2 stages) and V (1 stage) tiles.  Each operand is a 4-D tensor map
//               [B][N][heads][d]; boxes are 64 elements wide so columns >= d are zero-filled by TMA -- head
//               dims 40 / 80 / 160 need no padding in memory and feed canonical 128B-swizzled smem tiles.
//   warp 1      MMA issuer:  S = Q K^T   (M=128, N=128, K = ceil16(d); both operands K-major)      -> TMEM
//                            O += P V    (M=128, N = ceil16(d), K=128; P K-major from smem, V MN-major) -> TMEM
//   warps 2..5  softmax: thread = query row.  Two passes over S in TMEM (row max, then exp2 / row sum /
//               bf16 P into swizzled smem), lazy rescale of the O accumulator (only when the running max
//               moves by more than 2^8), final 1/l normalisation and bf16 store.
// S (128 cols) and O (<=160 cols) live in TMEM; P goes through smem (generic-proxy writes + proxy fence).
// For d <= 64 two CTAs fit per SM so one CTA's softmax overlaps the other's MMAs.
#include <float.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace c2d {

using namespace tc;

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);

constexpr int AT_BQ = 128, AT_BK = 128, AT_THREADS = 192, AT_TILE = 128 * 128;   // 16 KB: 128 rows x 128 B

struct AttnTcParams {
  bf16* o;
  int Nq, Nkv, d, npv;          // npv = ceil16(d): N extent of the PV MMA
  long long ldo, bso;
  float scale_log2;             // softmax scale * log2(e)
};

template <int NBLK>
struct AtCfg {
  static constexpr int Q_BYTES = NBLK * AT_TILE;
  static constexpr int K_STAGES = 2;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + Q_BYTES;
  static constexpr int OFF_V = OFF_K + K_STAGES * Q_BYTES;
  static constexpr int OFF_P = OFF_V + Q_BYTES;
  static constexpr int OFF_BAR = OFF_P + 2 * AT_TILE;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static constexpr int TMEM_COLS = NBLK <= 2 ? 256 : 512;
  static constexpr int O_COL = 128;
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int NBLK>
__global__ void __launch_bounds__(AT_THREADS)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
  using Cfg = AtCfg<NBLK>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;
  uint64_t* v_empty = bars + 6;
  uint64_t* s_full = bars + 7;
  uint64_t* s_empty = bars + 8;
  uint64_t* p_full = bars + 9;
  uint64_t* p_empty = bars + 10;  // = PV(j) retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BQ, h = blockIdx.y, b = blockIdx.z;
  const int ntiles = (p.Nkv + AT_BK - 1) / AT_BK;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    mbar_init(v_full, 1); mbar_init(v_empty, 1);
    mbar_init(s_full, 1); mbar_init(s_empty, 128);
    mbar_init(p_full, 128); mbar_init(p_empty, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + Cfg::O_COL;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
      for (int blk = 0; blk < NBLK; ++blk) tma_load_4d(smem + Cfg::OFF_Q + blk * AT_TILE, &tmQ, q_full, blk * 64, h, q0, b);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j & 1;
        const uint32_t u = (uint32_t)(j >> 1);
        mbar_wait(&k_empty[s], (u & 1u) ^ 1u);
        mbar_arrive_expect_tx(&k_full[s], Cfg::Q_BYTES);
#pragma unroll
        for (int blk = 0; blk < NBLK; ++blk)
          tma_load_4d(smem + Cfg::OFF_K + (s * NBLK + blk) * AT_TILE, &tmK, &k_full[s], blk * 64, h, j * AT_BK, b);
        mbar_wait(v_empty, ((uint32_t)j & 1u) ^ 1u);
        mbar_arrive_expect_tx(v_full, Cfg::Q_BYTES);
#pragma unroll
        for (int blk = 0; blk < NBLK; ++blk)
          tma_load_4d(smem + Cfg::OFF_V + blk * AT_TILE, &tmV, v_full, blk * 64, h, j * AT_BK, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      const uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_pv = make_idesc_bf16(128, p.npv, 0, 1);          // B (= V) is MN-major
      const uint32_t q_addr = smem_u32(smem + Cfg::OFF_Q);
      const uint32_t p_addr = smem_u32(smem + Cfg::OFF_P);
      const uint32_t v_addr = smem_u32(smem + Cfg::OFF_V);
      const int ksteps = (p.d + 15) >> 4;
      mbar_wait(q_full, 0);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j & 1;
        mbar_wait(&k_full[s], (uint32_t)(j >> 1) & 1u);
        if (j > 0) mbar_wait(s_empty, (uint32_t)(j - 1) & 1u);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(smem + Cfg::OFF_K + s * NBLK * AT_TILE);
        for (int kk = 0; kk < ksteps; ++kk) {
          const uint32_t off = (uint32_t)(kk >> 2) * AT_TILE + (uint32_t)(kk & 3) * 32;
          umma_f16(tmem_s, make_desc_k_sw128(q_addr + off), make_desc_k_sw128(k_addr + off), idesc_qk, kk > 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        umma_commit(&k_empty[s]);
        // ---- O += P V
        mbar_wait(p_full, (uint32_t)j & 1u);
        mbar_wait(v_full, (uint32_t)j & 1u);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < AT_BK / 16; ++kk) {
          const uint64_t a_desc = make_desc_k_sw128(p_addr + (uint32_t)(kk >> 2) * AT_TILE + (uint32_t)(kk & 3) * 32);
          // V tile: [128 keys][64-col blocks]; 16 keys = 2 groups of 8 rows (SBO = 1024 B), col blocks 16 KB apart (LBO)
          const uint64_t b_desc = make_desc_mn_sw128(v_addr + (uint32_t)kk * 2048, AT_TILE, 1024);
          umma_f16(tmem_o, a_desc, b_desc, idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(p_empty);
        umma_commit(v_empty);
      }
    }
  } else {
    // ===================== softmax / correction / epilogue (warps 2..5) =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    uint8_t* p_row = smem + Cfg::OFF_P + row * 128;
    const int rsw = row & 7;
    float m_ref = -INFINITY, l_run = 0.f;
    for (int j = 0; j < ntiles; ++j) {
      const int kvalid = min(AT_BK, p.Nkv - j * AT_BK);
      mbar_wait(s_full, (uint32_t)j & 1u);
      tc_fence_after();
      if (j == 0) {
        // first tile: the reference max must be known before any exponential is taken
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_s + lane_off + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < kvalid) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
        m_ref = mx * p.scale_log2;
      } else {
        mbar_wait(p_empty, (uint32_t)(j - 1) & 1u);     // PV(j-1) retired: P smem free, O complete
      }
      // ---- single pass in the common case: p = exp2(s * scale_log2 - m_ref) against the LAGGING reference max,
      //      tracking the tile max on the side.  Only if some row's max moved by more than 2^8 is the tile redone
      //      with the new reference and the O accumulator rescaled (rare after the first tiles).
      float sum, mx;
      float alpha = 1.f;
      bool redo = false;
      for (;;) {
        sum = 0.f;
        mx = -INFINITY;
        uint32_t ra[32], rb[32];
        tmem_ld_32x32(tmem_s + lane_off, ra);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t (&r)[32] = (c & 1) ? rb : ra;
          uint32_t (&rn)[32] = (c & 1) ? ra : rb;
          tmem_ld_wait();
          if (c < 3) tmem_ld_32x32(tmem_s + lane_off + (c + 1) * 32, rn);    // prefetch next chunk
          uint32_t packed[16];
          if (kvalid == AT_BK) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float s0 = __uint_as_float(r[i]), s1 = __uint_as_float(r[i + 1]);
              mx = fmaxf(mx, fmaxf(s0, s1));
              __nv_bfloat162 hb = __floats2bfloat162_rn(ex2f(fmaf(s0, p.scale_log2, -m_ref)), ex2f(fmaf(s1, p.scale_log2, -m_ref)));
              sum += __low2float(hb) + __high2float(hb);      // what the tensor core will actually see
              packed[i >> 1] = *reinterpret_cast<uint32_t*>(&hb);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const bool v0 = c * 32 + i < kvalid, v1 = c * 32 + i + 1 < kvalid;
              const float s0 = v0 ? __uint_as_float(r[i]) : -INFINITY, s1 = v1 ? __uint_as_float(r[i + 1]) : -INFINITY;
              mx = fmaxf(mx, fmaxf(s0, s1));
              __nv_bfloat162 hb = __floats2bfloat162_rn(v0 ? ex2f(fmaf(s0, p.scale_log2, -m_ref)) : 0.f,
                                                        v1 ? ex2f(fmaf(s1, p.scale_log2, -m_ref)) : 0.f);
              sum += __low2float(hb) + __high2float(hb);
              packed[i >> 1] = *reinterpret_cast<uint32_t*>(&hb);
            }
          }
          // 32 keys = 4 chunks of 16 B; key block (64 keys) = c >> 1, chunk index within the 128 B row = (c & 1) * 4 + q
          uint8_t* blk = p_row + (c >> 1) * AT_TILE;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int chunk = ((c & 1) * 4 + q) ^ rsw;
            *reinterpret_cast<uint4*>(blk + (chunk << 4)) =
                make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
          }
        }
        if (redo) break;
        mx *= p.scale_log2;
        const bool need = mx > m_ref + 8.0f;
        if (!__any_sync(0xffffffffu, need)) break;
        if (need) {
          alpha = ex2f(m_ref - mx);
          m_ref = mx;
        }
        redo = true;
      }
      if (redo && j > 0) {
        // warp-collective rescale of the O accumulator (rows that did not move use alpha = 1)
        for (int c = 0; c < p.npv; c += 16) {
          uint32_t r[16];
          tmem_ld_32x16(tmem_o + lane_off + c, r);
          tmem_ld_wait();
          uint32_t lo[8], hi[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            lo[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            hi[i] = __float_as_uint(__uint_as_float(r[8 + i]) * alpha);
          }
          tmem_st_32x8(tmem_o + lane_off + c, lo);
          tmem_st_32x8(tmem_o + lane_off + c + 8, hi);
        }
        tmem_st_wait();
      }
      l_run = l_run * alpha + sum;
      tc_fence_before();
      mbar_arrive(s_empty);           // S(j) fully read: QK(j+1) may overwrite it
      fence_proxy_async();            // P writes (generic proxy) -> visible to tcgen05.mma (async proxy)
      mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> bf16 -> global
    mbar_wait(p_empty, (uint32_t)(ntiles - 1) & 1u);
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
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

bool attention_tc_supported(const AttnParams& p, int B) {
  (void)B;
  return p.mask == nullptr && p.d % 8 == 0 && p.d >= 16 && p.d <= 192 && p.ldq % 8 == 0 && p.ldk % 8 == 0 &&
         p.ldv % 8 == 0 && p.ldo % 8 == 0 && p.bsq % 8 == 0 && p.bsk % 8 == 0 && p.bsv % 8 == 0 && p.bso % 8 == 0 &&
         (B == 1 || (p.bsq > 0 && p.bsk > 0 && p.bsv > 0)) && al16(p.q) && al16(p.k) && al16(p.v) && al16(p.o) &&
         p.Nq >= 1 && p.Nkv >= 1;
}

static int make_head_tmap(CUtensorMap* m, const void* base, int d, int heads, int N, int B, long long ld, long long bs) {
  uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)N, (uint64_t)B};
  uint64_t st[3] = {(uint64_t)d * 2, (uint64_t)ld * 2, (uint64_t)(B > 1 ? bs : (long long)N * ld) * 2};
  uint32_t box[4] = {64, 1, 128, 1};
  return make_tmap_bf16(m, base, 4, dims, st, box);
}

template <int NBLK>
static int launch_attn_tc(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnTcParams& ap,
                          int heads, int B, cudaStream_t s) {
  using Cfg = AtCfg<NBLK>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel<NBLK>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("attention_tc: cudaFuncSetAttribute(%d B) failed: %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return C2D_ERR_CUDA;
    }
    attr_done = true;
  }
  dim3 grid(ceil_div(ap.Nq, AT_BQ), heads, B);
  attn_tc_kernel<NBLK><<<grid, AT_THREADS, Cfg::SMEM_BYTES, s>>>(tq, tk, tv, ap);
  return check_launch("attn_tc");
}

int attention_tc(const AttnParams& p, int B, cudaStream_t s) {
  CUtensorMap tq, tk, tv;
  int rc = make_head_tmap(&tq, p.q, p.d, p.heads, p.Nq, B, p.ldq, p.bsq);
  if (rc) return rc;
  rc = make_head_tmap(&tk, p.k, p.d, p.heads, p.Nkv, B, p.ldk, p.bsk);
  if (rc) return rc;
  rc = make_head_tmap(&tv, p.v, p.d, p.heads, p.Nkv, B, p.ldv, p.bsv);
  if (rc) return rc;
  AttnTcParams ap;
  ap.o = reinterpret_cast<bf16*>(p.o);
  ap.Nq = p.Nq; ap.Nkv = p.Nkv; ap.d = p.d; ap.npv = (p.d + 15) & ~15;
  ap.ldo = p.ldo; ap.bso = p.bso;
  ap.scale_log2 = p.scale * 1.4426950408889634f;
  const int nblk = (p.d + 63) / 64;
  if (nblk == 1) return launch_attn_tc<1>(tq, tk, tv, ap, p.heads, B, s);
  if (nblk == 2) return launch_attn_tc<2>(tq, tk, tv, ap, p.heads, B, s);
  return launch_attn_tc<3>(tq, tk, tv, ap, p.heads, B, s);
}

}  // namespace c2d
