// Second-generation fused cross-attention site for sm_100a: PERSISTENT, warp-specialised, weight-stationary.
// Same contract as xattn_tc.cu (the per-step half of AudioAttnProcessor.__call__, reference
// models/audio_attention_processor.py:114-131: to_q with the preceding LayerNorm folded in, then the multi-head
// softmax(Q K^T) V against the cached text(+audio) keys / values), used for the sites whose to_q slice fits in shared
// memory (SD-1.5: the five 4096-token, C = 320 sites, which are 40 % of the cross-attention time).
//
// Why a second kernel.  The first-generation kernel is one CTA per (128 rows, head group): every CTA re-streams its
// slice of Wq (120 KB at C = 320) next to 80 KB of x and 80 KB of K/V, and runs projection -> convert -> 4 x (score,
// softmax, P V) -> epilogue as ONE serial chain; two co-resident CTAs are the only overlap (ncu, round 1: tensor pipe
// 23 %, warps waiting on mbarriers).  Here:
//   * one CTA per SM for the whole launch; its head group's Wq slice [G*dp rows x C] is loaded ONCE and stays in
//     shared memory (120 KB), so a work item (128 rows x G heads) moves 80 KB of x + 80 KB of K/V through the SM
//     instead of 280 KB;
//   * software pipeline ACROSS items: the projection of item i+1 runs on its own issuing warp while item i is in its
//     attention phase; the fp32 accumulator of item i+1 (QACC) and the bf16 Q of item i (QBF) live in different TMEM
//     columns;
//   * roles: warp 0 = producer (the one-time Wq load, then three independent rings -- x stages, K slots, V slots --
//     served by non-blocking mbarrier polls: a K slot is free as soon as its score MMA has retired, a V slot only
//     after P V), warp 1 = projection MMA issuer, warp 2 = attention MMA issuer (score / P V), warps 4-7 / 8-11 = two
//     softmax warpgroups on alternate heads (each with its own score buffer), warps 12-15 = Q convert
//     (folded-LayerNorm affine, fp32 -> bf16, QACC -> QBF) and the output epilogue (registers -> dense staging tile ->
//     one TMA store per head).  setmaxnreg moves registers from the control warps to the softmax warpgroups.
//   TMEM (480 of 512 columns): QACC [0, NG) | QBF [NG, 3NG/2) | two score buffers of 96 columns: S / P (in place) at
//   [0, NS), the head's output accumulator above the bf16 P at [48, 96) -- P V of a head never waits for the epilogue
//   of the previous one, only the NEXT score MMA on the same buffer does (o_free).
//
// Measured (B200, batch 16 x 4096 tokens, C = 320; tools/bench_xattn.py, tools/xattn_p_profile.py, tools/microbench/
// mma_chain.cu):
//   * 48.9 us per launch against 50.8 us for the first-generation kernel -- NOT the 3x the tile arithmetic promised.
//   * tcgen05.mma has a fixed cost per instruction that the N/2-cycle model hides: a K = 16 MMA at M = 128 takes
//     130 cycles at N = 192 (SS form: +34 cycles for the 4 KB A read), 68 at N = 80 and 64 at N = 48 (TS form), and it
//     does not matter whether consecutive MMAs share an accumulator.  An item is 20 + 12 + 20 such instructions:
//     ~4700 tensor cycles, not the 2880 of the N/2 model -- 33 k cycles (17 us) per launch is the tensor floor of
//     ANY kernel with these tile shapes, i.e. the 60 % tensor-pipe target needs the whole launch in <= 29 us.
//   * with every MMA, softmax, convert and epilogue switched off (C2D_XATTN_SKIP=31: loads and barrier handshakes
//     only) the launch still takes 24.8 us: ~6.7 k cycles per item of pure signalling -- each of the ~16 barrier
//     round trips per item (tcgen05.commit -> mbarrier -> try_wait wake-up -> arrive -> wake-up) costs 150-300
//     cycles and they are chained through the attention issuer.  Softmax off: -4 us, convert off: -7 us, epilogue off:
//     -10 us, all three: 33 us.  The limiter is the control path, not MUFU (2560 cycles per item), not the tensor
//     pipe, not L2 -> SM bandwidth (166 MB per launch).
//   Next step (not built): batch two heads per handshake (score MMAs of both buffers behind one K-pair barrier, one
//   256-arrival P barrier, one V-pair barrier) and let each softmax warpgroup drain the OTHER group's output while it
//   waits, so the convert warpgroup only converts.
#include <float.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace c2d {

using namespace tc;

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);
int make_tmap_bf16_plain(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                         const uint32_t* box);

constexpr int XP_THREADS = 512;
constexpr int XP_BM = 128, XP_BK = 64;
constexpr int XP_XSTAGES = 3, XP_KSLOTS = 2, XP_VSLOTS = 2;
constexpr int XP_X_BYTES = XP_BM * XP_BK * 2;

struct XpParams {
  bf16* o;
  long long ldo;
  const long long* ln_stats;     // [M][2] fixed-point (sum, sumsq) of the rows of x, or null
  const float* colsum;           // [heads*d] column sums of the gamma-scaled weight (null without LayerNorm fold)
  const float* qbias;            // [heads*d] or null
  float ln_invK, ln_eps;
  int M, Nq;
  int G, d, dp, NG;              // heads per CTA, head dim, ceil16(d), G * dp
  int n1, NS;                    // keys, score columns (NCH * 16)
  float scale_log2;
  int num_kb;                    // C / 64
  int ngroups, n_tiles, heads;
  const uint8_t* kvp;            // packed K / V cache (xattn_pack_kv_kernel): [B][heads] images of slot_bytes = K block | V block
  int slot_bytes, kvblk_bytes;
  int off_w, off_x, off_k, off_v, off_o, off_bar, w_blk_bytes;
  int col_qbf, col_s0, col_s1, o_off;   // score buffers of sbw columns: S / P at [0, NS), the head's O accumulator at [o_off, o_off + dp)
  int skip;                      // timing experiments only (C2D_XATTN_SKIP bit mask): 1 softmax, 2 convert, 4 epilogue, 8 projection MMAs, 16 attention MMAs
  long long* dbg;                // optional per-role wait-cycle counters of CTA 0: [16 warps][16] (C2D_XATTN_DBG = device pointer)
};

__device__ __forceinline__ void xp_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ float xp_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x for x <= 0 on the FMA pipe: Cody-Waite split x = n + f (f in [-0.5, 0.5]), degree-3 minimax polynomial for 2^f
// (coefficients of attn_tc2.cu: max relative error 7.5e-5, below the bf16 rounding of P), exponent by integer add.
// Inputs below -125 (masked keys, -inf) clamp to 2^-125.
__device__ __forceinline__ float xp_ex2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float r = x + 12582912.f;                 // 1.5 * 2^23: round to nearest integer in the low mantissa bits
  const float n = r - 12582912.f;
  const float f = x - n;
  float pl = fmaf(f, 0.05517167f, 0.24261113f);
  pl = fmaf(pl, f, 0.69326097f);
  pl = fmaf(pl, f, 0.99992806f);
  return __int_as_float(__float_as_int(pl) + (__float_as_int(r) << 23));
}
__device__ __forceinline__ void xp_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void xp_ffma2(float& d0, float& d1, float a0, float a1, float b, float c) {
  asm("{.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %4};\n\t"
      "mov.b64 rc, {%5, %5};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
__device__ __forceinline__ void xp_ffma2v(float& d0, float& d1, float a0, float a1, float b, float c0, float c1) {
  asm("{.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %4};\n\t"
      "mov.b64 rc, {%5, %6};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void xp_fadd2(float& d0, float& d1, float a0, float a1) {
  asm("{.reg .b64 ra, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rd, {%0, %1};\n\t"
      "add.rn.f32x2 rd, rd, ra;\n\t"
      "mov.b64 {%0, %1}, rd;}"
      : "+f"(d0), "+f"(d1)
      : "f"(a0), "f"(a1));
}
__device__ __forceinline__ void xp_fmul2(float& d0, float& d1, float a0, float a1, float b) {
  asm("{.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %4};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b));
}
__device__ __forceinline__ float4 xp_lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float xp_max3(float a, float b, float c) {
  float m;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(m) : "f"(a), "f"(b), "f"(c));
  return m;
}
__device__ __forceinline__ uint32_t xp_pack(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <int REGS> __device__ __forceinline__ void xp_reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS> __device__ __forceinline__ void xp_reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }

// wait-cycle accounting (debug launches only: p.dbg != null)
#define XP_WAIT(id, stmt)                                   \
  do {                                                      \
    if (dbg_on) {                                           \
      const uint32_t t0_ = (uint32_t)clock();               \
      stmt;                                                 \
      dbg_acc[id] += (uint32_t)clock() - t0_;               \
    } else {                                                \
      stmt;                                                 \
    }                                                       \
  } while (0)

__device__ __forceinline__ bool xp_mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void xp_tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void xp_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void xp_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void xp_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// barrier slots (uint64_t each) at smem + off_bar
enum { XB_W = 0, XB_XF = 1, XB_XE = 4, XB_KF = 7, XB_KE = 9, XB_VF = 11, XB_VE = 13, XB_QDONE = 15, XB_QBF = 16, XB_QKDONE = 17,
       XB_SF = 18, XB_PF = 20, XB_OF = 22, XB_OFREE = 24, XB_TMEM = 26 };

// POLY: every 4th exponential as an FMA-pipe polynomial instead of MUFU.EX2
template <int NCH, bool POLY>
__global__ void __launch_bounds__(XP_THREADS, 1)
xattn_p_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmO, const XpParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + XB_TMEM);
  float* s_cs = reinterpret_cast<float*>(smem + p.off_bar + 256);          // [NG] column sums, head-padded
  float* s_qb = s_cs + 256;                                                // [NG] bias, head-padded

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = blockIdx.x % p.ngroups;
  const int tile0 = blockIdx.x / p.ngroups, tstride = gridDim.x / p.ngroups;
  const int head0 = grp * p.G;
  const int nitems = tile0 < p.n_tiles ? (p.n_tiles - tile0 + tstride - 1) / tstride : 0;
  const bool dbg_on = p.dbg != nullptr && blockIdx.x == 0;
  uint32_t dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const uint32_t dbg_t0 = dbg_on ? (uint32_t)clock() : 0u;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX); prefetch_tmap(&tmW); prefetch_tmap(&tmO);
    mbar_init(&bars[XB_W], 1);
    for (int i = 0; i < XP_XSTAGES; ++i) { mbar_init(&bars[XB_XF + i], 1); mbar_init(&bars[XB_XE + i], 1); }
    for (int i = 0; i < XP_KSLOTS; ++i) { mbar_init(&bars[XB_KF + i], 1); mbar_init(&bars[XB_KE + i], 1); }
    for (int i = 0; i < XP_VSLOTS; ++i) { mbar_init(&bars[XB_VF + i], 1); mbar_init(&bars[XB_VE + i], 1); }
    mbar_init(&bars[XB_QDONE], 1);
    mbar_init(&bars[XB_QBF], 128);
    mbar_init(&bars[XB_QKDONE], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[XB_SF + i], 1); mbar_init(&bars[XB_PF + i], 128);
      mbar_init(&bars[XB_OF + i], 1); mbar_init(&bars[XB_OFREE + i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();             // x / row statistics come from the kernel just before this one; nothing above touches global memory
  pdl_trigger();

  if (warp < 4) {
    xp_reg_dec<56>();
    if (warp == 0) {
      // ===================== producer: the resident Wq slice once, then three independent rings (x stages, K slots, V
      // slots) served by non-blocking polls, so a full ring never holds up the other two =====================
      if (nitems > 0) {
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars[XB_W], (uint32_t)(p.num_kb * p.w_blk_bytes));
          for (int kb = 0; kb < p.num_kb; ++kb)
            tma_load_3d(smem + p.off_w + kb * p.w_blk_bytes, &tmW, &bars[XB_W], kb * XP_BK, 0, head0);
        }
        __syncwarp();
        uint32_t xi = 0, ki = 0, vi = 0;                       // ring positions
        int xn = 0, xkb = 0, kn = 0, kh = 0, vn = 0, vh = 0;   // (item, K block) / (item, head) of the next request
        uint32_t idle = 0;
        while (xn < nitems || kn < nitems || vn < nitems) {
          bool any = false;
          if (xn < nitems) {
            const uint32_t s_ = xi % XP_XSTAGES;
            if (xp_mbar_test(&bars[XB_XE + s_], ((xi / XP_XSTAGES) & 1u) ^ 1u)) {
              if (elect_one()) {
                mbar_arrive_expect_tx(&bars[XB_XF + s_], (uint32_t)XP_X_BYTES);
                tma_load_2d(smem + p.off_x + s_ * XP_X_BYTES, &tmX, &bars[XB_XF + s_], xkb * XP_BK, (tile0 + xn * tstride) * XP_BM);
              }
              __syncwarp();
              ++xi; any = true;
              if (++xkb == p.num_kb) { xkb = 0; ++xn; }
            }
          }
          if (kn < nitems) {
            const uint32_t s_ = ki % XP_KSLOTS;
            if (xp_mbar_test(&bars[XB_KE + s_], ((ki / XP_KSLOTS) & 1u) ^ 1u)) {
              if (elect_one()) {
                const int b = ((tile0 + kn * tstride) * XP_BM) / p.Nq;
                const uint8_t* src = p.kvp + ((size_t)b * p.heads + head0 + kh) * (size_t)p.slot_bytes;
                mbar_arrive_expect_tx(&bars[XB_KF + s_], (uint32_t)p.kvblk_bytes);
                xp_bulk_load(smem + p.off_k + s_ * p.kvblk_bytes, src, (uint32_t)p.kvblk_bytes, &bars[XB_KF + s_]);
              }
              __syncwarp();
              ++ki; any = true;
              if (++kh == p.G) { kh = 0; ++kn; }
            }
          }
          if (vn < nitems) {
            const uint32_t s_ = vi % XP_VSLOTS;
            if (xp_mbar_test(&bars[XB_VE + s_], ((vi / XP_VSLOTS) & 1u) ^ 1u)) {
              if (elect_one()) {
                const int b = ((tile0 + vn * tstride) * XP_BM) / p.Nq;
                const uint8_t* src = p.kvp + ((size_t)b * p.heads + head0 + vh) * (size_t)p.slot_bytes + p.kvblk_bytes;
                mbar_arrive_expect_tx(&bars[XB_VF + s_], (uint32_t)p.kvblk_bytes);
                xp_bulk_load(smem + p.off_v + s_ * p.kvblk_bytes, src, (uint32_t)p.kvblk_bytes, &bars[XB_VF + s_]);
              }
              __syncwarp();
              ++vi; any = true;
              if (++vh == p.G) { vh = 0; ++vn; }
            }
          }
          if (!any) {
            __nanosleep(32);
            if (++idle > (1u << 24)) __trap();        // protocol bug: fail instead of hanging the GPU
          } else {
            idle = 0;
          }
        }
      }
    } else if (warp == 1) {
      // ===================== projection MMA issuer: QACC(n) = x_tile(n) . Wq_g^T =====================
      // A second issuing thread: the attention issuer's chain of barrier waits / commits (about 30 per item) would
      // otherwise delay these MMAs; the two streams write disjoint TMEM columns, ordered by qbf_ready / q_done only.
      const uint32_t idesc_q = make_idesc_bf16(128, p.NG, 0, 0);
      uint32_t xc = 0;
      if (nitems > 0) XP_WAIT(6, mbar_wait(&bars[XB_W], 0));
      for (int n = 0; n < nitems; ++n) {
        if (n > 0) XP_WAIT(5, mbar_wait(&bars[XB_QBF], (uint32_t)(n - 1) & 1u));   // convert(n-1) has read QACC
        for (int kb = 0; kb < p.num_kb; ++kb, ++xc) {
          const uint32_t s_ = xc % XP_XSTAGES;
          XP_WAIT(0, mbar_wait(&bars[XB_XF + s_], (xc / XP_XSTAGES) & 1u));
          tc_fence_after();
          if (elect_one()) {
            const uint64_t ad = make_desc_k_sw128(smem_u32(smem + p.off_x + s_ * XP_X_BYTES));
            const uint64_t bd = make_desc_k_sw128(smem_u32(smem + p.off_w + kb * p.w_blk_bytes));
            if (!(p.skip & 8)) {
#pragma unroll
              for (int kk = 0; kk < XP_BK / 16; ++kk)
                umma_f16(tmem_base, ad + (uint64_t)(2 * kk), bd + (uint64_t)(2 * kk), idesc_q, (kb > 0 || kk > 0) ? 1u : 0u);
            }
            umma_commit(&bars[XB_XE + s_]);
            if (kb == p.num_kb - 1) umma_commit(&bars[XB_QDONE]);
          }
          __syncwarp();
        }
      }
    } else if (warp == 2) {
      // ===================== attention MMA issuer =====================
      const uint32_t idesc_qk = make_idesc_bf16(128, p.NS, 0, 0);
      const uint32_t idesc_pv = make_idesc_bf16(128, p.dp, 0, 1);           // B (= V) is MN-major
      const uint32_t tmem_s0 = tmem_base + (uint32_t)p.col_s0, tmem_s1 = tmem_base + (uint32_t)p.col_s1;
      const uint32_t tmem_qbf = tmem_base + (uint32_t)p.col_qbf;
      const int ksteps_qk = p.dp >> 4, ksteps_pv = p.NS >> 4;
      uint32_t kc = 0, vc = 0, use0 = 0, use1 = 0;       // use*: score MMAs issued into buffer 0 / 1 so far
      // S_h = Q_h K_h^T into score buffer (h & 1).  The buffer also holds P (in place over S) and, in its upper columns,
      // the output accumulator of the head: a score MMA needs (a) the previous head's P V on this buffer issued -- same
      // thread, tcgen05.mma retires in issue order -- and (b) its O drained by the epilogue (o_free).
      auto issue_qk = [&](int hh, bool last) {
        const uint32_t sl = kc % XP_KSLOTS;
        const int w = hh & 1;
        const uint32_t u = w ? use1 : use0;
        if (u > 0) XP_WAIT(3, mbar_wait(&bars[XB_OFREE + w], (u - 1) & 1u));
        XP_WAIT(1, mbar_wait(&bars[XB_KF + sl], (kc / XP_KSLOTS) & 1u));
        tc_fence_after();
        if (elect_one()) {
          const uint64_t kd = make_desc_k_sw128(smem_u32(smem + p.off_k + sl * p.kvblk_bytes));
          const uint32_t tq = tmem_qbf + (uint32_t)(hh * (p.dp >> 1));
          for (int kk = 0; kk < ksteps_qk && !(p.skip & 16); ++kk)
            umma_f16_ts(w ? tmem_s1 : tmem_s0, tq + (uint32_t)kk * 8, kd + (uint64_t)(kk * 2), idesc_qk, kk > 0 ? 1u : 0u);
          umma_commit(&bars[XB_SF + w]);
          umma_commit(&bars[XB_KE + sl]);
          if (last) umma_commit(&bars[XB_QKDONE]);        // QBF of this item is dead: the next item's Q may be written
        }
        __syncwarp();
        ++kc;
        if (w) ++use1; else ++use0;
      };
      auto issue_pv = [&](int hh) {
        const uint32_t sl = vc % XP_VSLOTS;
        const int w = hh & 1;
        const uint32_t u = (w ? use1 : use0) - 1;          // this head's use index of buffer w
        XP_WAIT(2, mbar_wait(&bars[XB_PF + w], u & 1u));
        XP_WAIT(4, mbar_wait(&bars[XB_VF + sl], (vc / XP_VSLOTS) & 1u));
        tc_fence_after();
        if (elect_one()) {
          const uint64_t vd = make_desc_mn_sw128(smem_u32(smem + p.off_v + sl * p.kvblk_bytes), (uint32_t)p.kvblk_bytes, 1024);
          const uint32_t ts = w ? tmem_s1 : tmem_s0;
          for (int kk = 0; kk < ksteps_pv && !(p.skip & 16); ++kk)            // 16 keys = 8 TMEM columns of P = 2 KB of V rows
            umma_f16_ts(ts + (uint32_t)p.o_off, ts + (uint32_t)kk * 8, vd + (uint64_t)(kk * (2048 >> 4)), idesc_pv, kk > 0 ? 1u : 0u);
          umma_commit(&bars[XB_OF + w]);
          umma_commit(&bars[XB_VE + sl]);
        }
        __syncwarp();
        ++vc;
      };
      for (int n = 0; n < nitems; ++n) {
        XP_WAIT(5, mbar_wait(&bars[XB_QBF], (uint32_t)n & 1u));
        tc_fence_after();
        issue_qk(0, p.G == 1);
        if (p.G > 1) issue_qk(1, p.G == 2);
        for (int hh = 0; hh < p.G; ++hh) {
          issue_pv(hh);
          if (hh + 2 < p.G) issue_qk(hh + 2, hh + 2 == p.G - 1);
        }
      }
    }
  } else if (warp < 12) {
    // ===================== softmax warpgroups (warps 4-7: even heads, 8-11: odd heads; thread = row) =====================
    xp_reg_inc<160>();
    const int wg = (warp - 4) >> 2;
    const int quad = warp & 3;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const uint32_t tmem_s = tmem_base + (uint32_t)(wg ? p.col_s1 : p.col_s0) + lane_off;
    const float sc = p.scale_log2;
    const int nv_last = p.n1 - (NCH - 1) * 16;          // valid keys of the last 16-column chunk (1..16)
    uint32_t use = 0;
    for (int n = 0; n < nitems; ++n) {
      for (int hh = wg; hh < p.G; hh += 2, ++use) {
        XP_WAIT(0, mbar_wait(&bars[XB_SF + wg], use & 1u));
        const uint32_t tcmp = dbg_on ? (uint32_t)clock() : 0u;
        tc_fence_after();
        if (p.skip & 1) { tc_fence_before(); mbar_arrive(&bars[XB_PF + wg]); continue; }
        uint32_t s[NCH * 16];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          uint32_t t[16];
          tmem_ld_32x16(tmem_s + (uint32_t)c * 16, t);
#pragma unroll
          for (int i = 0; i < 16; ++i) s[c * 16 + i] = t[i];
        }
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (i >= nv_last) s[(NCH - 1) * 16 + i] = 0xff800000u;      // keys past the end: -inf (exp2 gives exactly 0)
        float ma = __uint_as_float(s[0]), mb = __uint_as_float(s[1]), mc = __uint_as_float(s[2]), md = __uint_as_float(s[3]);
#pragma unroll
        for (int i = 4; i + 7 < NCH * 16; i += 8) {
          ma = xp_max3(ma, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
          mb = xp_max3(mb, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
          mc = xp_max3(mc, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
          md = xp_max3(md, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
        }
        ma = xp_max3(ma, __uint_as_float(s[NCH * 16 - 4]), __uint_as_float(s[NCH * 16 - 3]));
        mb = xp_max3(mb, __uint_as_float(s[NCH * 16 - 2]), __uint_as_float(s[NCH * 16 - 1]));
        const float m1 = fmaxf(xp_max3(ma, mb, mc), md);
        const float nm1 = -m1 * sc;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int i = 0; i < NCH * 16; i += 4) {
          float x0, x1, x2, x3;
          xp_ffma2(x0, x1, __uint_as_float(s[i]), __uint_as_float(s[i + 1]), sc, nm1);
          xp_ffma2(x2, x3, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]), sc, nm1);
          x0 = xp_ex2(x0); x1 = xp_ex2(x1); x2 = xp_ex2(x2);
          x3 = POLY ? xp_ex2_poly(x3) : xp_ex2(x3);
          xp_fadd2(a0, a1, x0, x1);
          xp_fadd2(a2, a3, x2, x3);
          s[i] = __float_as_uint(x0); s[i + 1] = __float_as_uint(x1); s[i + 2] = __float_as_uint(x2); s[i + 3] = __float_as_uint(x3);
        }
        const float r1 = 1.f / ((a0 + a1) + (a2 + a3));
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float y0, y1;
            xp_fmul2(y0, y1, __uint_as_float(s[c * 16 + i]), __uint_as_float(s[c * 16 + i + 1]), r1);
            pk[i >> 1] = xp_pack(y0, y1);
          }
          tmem_st_32x8(tmem_s + (uint32_t)c * 8, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars[XB_PF + wg]);
        if (dbg_on) dbg_acc[1] += (uint32_t)clock() - tcmp;
      }
    }
  } else {
    // ===================== Q convert + output epilogue (warps 12-15; thread = row) =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const int ct = threadIdx.x - 384;
    for (int nn = ct; nn < p.NG; nn += 128) {
      const int hh = nn / p.dp, j = nn - hh * p.dp;
      const int src = (head0 + hh) * p.d + j;
      s_cs[nn] = (j < p.d && p.colsum) ? __ldg(p.colsum + src) : 0.f;
      s_qb[nn] = (j < p.d && p.qbias) ? __ldg(p.qbias + src) : 0.f;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const uint32_t cs_addr = smem_u32(s_cs), qb_addr = smem_u32(s_qb);
    const uint32_t tmem_qacc = tmem_base + lane_off, tmem_qbf = tmem_base + (uint32_t)p.col_qbf + lane_off;
    const uint32_t tmem_o0 = tmem_base + (uint32_t)(p.col_s0 + p.o_off) + lane_off;
    const uint32_t tmem_o1 = tmem_base + (uint32_t)(p.col_s1 + p.o_off) + lane_off;
    uint8_t* ostage = smem + p.off_o;                 // [128][d] bf16, dense: source of the TMA store of one head's output
    const uint32_t ost_row = smem_u32(ostage) + (uint32_t)row * (uint32_t)(p.d * 2);
    // QACC (fp32 projection of item n) -> folded-LayerNorm affine -> bf16 -> QBF
    auto convert = [&](int n) {
      const int m = (tile0 + n * tstride) * XP_BM + row;
      longlong2 st = make_longlong2(0, 0);
      if (p.ln_stats) st = *reinterpret_cast<const longlong2*>(p.ln_stats + 2 * (long long)m);
      float rstd = 1.f, nmr = 0.f;
      if (p.ln_stats) {
        const float inv = p.ln_invK * (1.0f / 1048576.0f);
        const float mean = (float)st.x * inv;
        const float var = fmaxf(fmaf(-mean, mean, (float)st.y * inv), 0.f);
        rstd = rsqrtf(var + p.ln_eps);
        nmr = -mean * rstd;
      }
      XP_WAIT(0, mbar_wait(&bars[XB_QDONE], (uint32_t)n & 1u));                      // projection of item n complete
      if (n > 0) XP_WAIT(1, mbar_wait(&bars[XB_QKDONE], (uint32_t)(n - 1) & 1u));    // every score MMA of item n-1 has read QBF
      const uint32_t tcmp = dbg_on ? (uint32_t)clock() : 0u;
      tc_fence_after();
      for (int c = 0; c < p.NG && !(p.skip & 2); c += 64) {
        const bool two = c + 32 < p.NG;
        uint32_t ra[32], rb[32];
        tmem_ld_32x32(tmem_qacc + (uint32_t)c, ra);
        if (two) tmem_ld_32x32(tmem_qacc + (uint32_t)(c + 32), rb);
        tmem_ld_wait();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (half == 0 || two) {
            uint32_t (&r)[32] = half ? rb : ra;
            const int cc = c + half * 32;
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 c4 = xp_lds128(cs_addr + (uint32_t)(cc + j) * 4);
              const float4 b4 = xp_lds128(qb_addr + (uint32_t)(cc + j) * 4);
              float t0, t1, t2, t3, v0, v1, v2, v3;
              xp_ffma2v(t0, t1, c4.x, c4.y, nmr, b4.x, b4.y);
              xp_ffma2v(t2, t3, c4.z, c4.w, nmr, b4.z, b4.w);
              xp_ffma2v(v0, v1, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), rstd, t0, t1);
              xp_ffma2v(v2, v3, __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]), rstd, t2, t3);
              pk[j >> 1] = xp_pack(v0, v1);
              pk[(j >> 1) + 1] = xp_pack(v2, v3);
            }
            xp_st16(tmem_qbf + (uint32_t)(cc >> 1), pk);
          }
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&bars[XB_QBF]);
      if (dbg_on) dbg_acc[3] += (uint32_t)clock() - tcmp;
    };
    uint32_t ou0 = 0, ou1 = 0;
    if (nitems > 0) convert(0);
    for (int n = 0; n < nitems; ++n) {
      const int m0 = (tile0 + n * tstride) * XP_BM;
      for (int hh = 0; hh < p.G; ++hh) {
        // the next item's Q as early as its inputs allow: its projection runs on the other issuer while this item is in
        // its attention phase, this item's last score MMA is issued right behind P V (G-3)
        if (p.G > 2 && hh == p.G - 2 && n + 1 < nitems) convert(n + 1);
        const int w = hh & 1;
        const uint32_t u = w ? ou1 : ou0;
        if (w) ++ou1; else ++ou0;
        XP_WAIT(2, mbar_wait(&bars[XB_OF + w], u & 1u));
        const uint32_t tepi = dbg_on ? (uint32_t)clock() : 0u;
        tc_fence_after();
        if (p.skip & 4) { tc_fence_before(); mbar_arrive(&bars[XB_OFREE + w]); continue; }
        uint32_t r[48];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          uint32_t t[16];
          if (c * 16 < p.dp) tmem_ld_32x16((w ? tmem_o1 : tmem_o0) + (uint32_t)c * 16, t);
#pragma unroll
          for (int i = 0; i < 16; ++i) r[c * 16 + i] = t[i];
        }
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&bars[XB_OFREE + w]);          // O is in registers: the buffer's next score MMA may overwrite it
        // registers -> dense [128][d] staging tile -> one TMA store per head (the row-strided 16-byte global stores of the
        // first-generation kernel cost ~860 cycles per head)
        if (threadIdx.x == 384) xp_store_wait_read();        // the previous head's store has read the staging tile
        asm volatile("bar.sync 2, 128;" ::: "memory");
#pragma unroll
        for (int g = 0; g < 6; ++g) {
          if (g * 8 < p.d) {                       // d % 8 == 0: whole 8-element groups are valid or not
            const uint32_t v0 = xp_pack(__uint_as_float(r[g * 8 + 0]), __uint_as_float(r[g * 8 + 1]));
            const uint32_t v1 = xp_pack(__uint_as_float(r[g * 8 + 2]), __uint_as_float(r[g * 8 + 3]));
            const uint32_t v2 = xp_pack(__uint_as_float(r[g * 8 + 4]), __uint_as_float(r[g * 8 + 5]));
            const uint32_t v3 = xp_pack(__uint_as_float(r[g * 8 + 6]), __uint_as_float(r[g * 8 + 7]));
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ost_row + (uint32_t)g * 16), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
          }
        }
        fence_proxy_async();
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (threadIdx.x == 384) {
          xp_tma_store_2d(&tmO, ostage, (head0 + hh) * p.d, m0);
          xp_store_commit();
        }
        if (dbg_on) dbg_acc[4] += (uint32_t)clock() - tepi;
      }
      if (p.G <= 2 && n + 1 < nitems) convert(n + 1);
    }
    if (threadIdx.x == 384) xp_store_wait_all();
  }
  if (dbg_on && lane == 0) {
    for (int i = 0; i < 7; ++i) p.dbg[warp * 8 + i] = dbg_acc[i];
    p.dbg[warp * 8 + 7] = (uint32_t)clock() - dbg_t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

struct XpPlan {
  int G, dp, NG, NS, kvblk_bytes, slot_bytes, w_blk_bytes, off_w, off_x, off_k, off_v, off_o, off_bar, smem_bytes;
  int col_qbf, col_s0, col_s1, o_off;
};

// The weight-stationary kernel takes a site when: one 64-column K / V block per head (d <= 64), a single key segment
// whose last 16-column chunk is the only ragged one, 128-row tiles, and Wq slice + rings within 227 KB / 512 TMEM columns.
static bool xp_plan(int C, int heads, int Nq, int T, int T2, XpPlan& pl) {
  if (heads <= 0 || C % heads || C % XP_BK || T2 != 0) return false;
  const int d = C / heads;
  if (d % 8 || d < 16 || d > 64) return false;
  if (Nq % XP_BM) return false;
  pl.dp = (d + 15) & ~15;
  pl.NS = (T + 15) & ~15;
  if (pl.NS < 80) pl.NS = 80;
  if (pl.NS != 80 || T <= 64) return false;                  // instance for 5 score chunks (65..80 keys)
  pl.G = 0;
  for (int g = heads; g >= 1; --g)
    if (heads % g == 0 && g * pl.dp <= 256 && (g * pl.dp) % 64 == 0) { pl.G = g; break; }
  if (!pl.G) return false;
  pl.NG = pl.G * pl.dp;
  pl.kvblk_bytes = pl.NS * 128;
  pl.slot_bytes = 2 * pl.kvblk_bytes;                        // must equal xattn_tc.cu's packed image (nblk = 1)
  pl.w_blk_bytes = pl.NG * 128;
  pl.off_w = 0;
  pl.off_x = pl.off_w + (C / XP_BK) * pl.w_blk_bytes;
  pl.off_k = pl.off_x + XP_XSTAGES * XP_X_BYTES;
  pl.off_v = pl.off_k + XP_KSLOTS * pl.kvblk_bytes;
  pl.off_o = pl.off_v + XP_VSLOTS * pl.kvblk_bytes;
  pl.off_bar = pl.off_o + ((XP_BM * d * 2 + 1023) & ~1023);
  pl.smem_bytes = pl.off_bar + 256 + 2 * 256 * 4 + 1024;
  if (pl.smem_bytes > 227 * 1024) return false;
  // score buffer = S / P columns [0, NS) and, above the bf16 P (NS / 2 columns), the head's output accumulator
  pl.o_off = ((pl.NS >> 1) + 15) & ~15;
  const int sbw = (pl.o_off + pl.dp > pl.NS ? pl.o_off + pl.dp : pl.NS);
  pl.col_qbf = pl.NG;
  pl.col_s0 = pl.col_qbf + (pl.NG >> 1);
  pl.col_s1 = pl.col_s0 + sbw;
  return pl.col_s1 + sbw <= 512 && d <= 48;
}

bool xattn_p_supported(int C, int heads, int Nq, int T, int T2) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("C2D_XATTN_P");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  XpPlan pl;
  return enabled && xp_plan(C, heads, Nq, T, T2, pl);
}

static inline bool xp_al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

int xattn_p(const void* x, long long ldx, const void* wq, const float* qbias, const long long* ln_stats, const float* ln_colsum,
            float ln_eps, const void* kv_packed, int T, void* o, long long ldo, int B, int Nq, int C, int heads, float scale,
            cudaStream_t s) {
  XpPlan pl;
  if (!xp_plan(C, heads, Nq, T, 0, pl)) {
    set_error("xattn_p: shape outside the weight-stationary kernel (C=%d heads=%d Nq=%d T=%d)", C, heads, Nq, T);
    return C2D_ERR_UNSUPPORTED;
  }
  C2D_REQUIRE(ldx % 8 == 0 && ldo % 8 == 0 && xp_al16(x) && xp_al16(wq) && xp_al16(kv_packed) && xp_al16(o),
              "xattn: strides must be multiples of 8 elements and pointers 16-byte aligned");
  const int d = C / heads;
  const long long M = (long long)B * Nq;
  CUtensorMap tx, tw, to;
  {
    uint64_t dims[2] = {(uint64_t)C, (uint64_t)M};
    uint64_t st[1] = {(uint64_t)ldx * 2};
    uint32_t box[2] = {XP_BK, XP_BM};
    int rc = make_tmap_bf16(&tx, x, 2, dims, st, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)C, (uint64_t)M};
    uint64_t st[1] = {(uint64_t)ldo * 2};
    uint32_t box[2] = {(uint32_t)d, XP_BM};
    int rc = make_tmap_bf16_plain(&to, o, 2, dims, st, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)C, (uint64_t)d, (uint64_t)heads};
    uint64_t st[2] = {(uint64_t)C * 2, (uint64_t)d * C * 2};
    uint32_t box[3] = {XP_BK, (uint32_t)pl.dp, (uint32_t)pl.G};
    int rc = make_tmap_bf16(&tw, wq, 3, dims, st, box);
    if (rc) return rc;
  }
  XpParams p;
  p.o = reinterpret_cast<bf16*>(o);
  p.ldo = ldo;
  p.ln_stats = ln_stats; p.colsum = ln_stats ? ln_colsum : nullptr; p.qbias = qbias;
  p.ln_invK = 1.0f / (float)C; p.ln_eps = ln_eps;
  p.M = (int)M; p.Nq = Nq;
  p.G = pl.G; p.d = d; p.dp = pl.dp; p.NG = pl.NG;
  p.n1 = T; p.NS = pl.NS;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.num_kb = C / XP_BK;
  p.ngroups = heads / pl.G; p.n_tiles = (int)(M / XP_BM); p.heads = heads;
  p.kvp = reinterpret_cast<const uint8_t*>(kv_packed);
  p.slot_bytes = pl.slot_bytes; p.kvblk_bytes = pl.kvblk_bytes;
  p.off_w = pl.off_w; p.off_x = pl.off_x; p.off_k = pl.off_k; p.off_v = pl.off_v; p.off_o = pl.off_o; p.off_bar = pl.off_bar; p.w_blk_bytes = pl.w_blk_bytes;
  p.col_qbf = pl.col_qbf; p.col_s0 = pl.col_s0; p.col_s1 = pl.col_s1; p.o_off = pl.o_off;
  p.dbg = nullptr;
  p.skip = 0;
  if (const char* e = getenv("C2D_XATTN_SKIP")) p.skip = atoi(e);
  if (const char* e = getenv("C2D_XATTN_DBG")) p.dbg = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
  static int poly = -1;
  if (poly < 0) {
    const char* e = getenv("C2D_XATTN_POLY");
    poly = (e && e[0] == '1') ? 1 : 0;       // measured: 50.0 us with the polynomial share vs 48.9 us without (the MUFU is not the limiter)
  }
  // one CTA per SM, every head group on the same number of SMs
  int ctas = (num_sms() / p.ngroups) * p.ngroups;
  const int want = p.n_tiles * p.ngroups;
  if (ctas > want) ctas = want;
  auto kern = poly ? xattn_p_kernel<5, true> : xattn_p_kernel<5, false>;
  static int smem_set[2][C2D_MAX_DEVICES] = {};
  if (int rc = ensure_dyn_smem(kern, pl.smem_bytes, smem_set[poly], "xattn_p")) return rc;
  launch_pdl(kern, dim3((unsigned)ctas), dim3(XP_THREADS), (size_t)pl.smem_bytes, s, tx, tw, to, p);
  return check_launch("xattn_p");
}

}  // namespace c2d
