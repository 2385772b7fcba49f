// Fused cross-attention site for sm_100a (the per-step half of AudioAttnProcessor.__call__,
// models/audio_attention_processor.py:114-131): to_q projection -- with the preceding LayerNorm folded in -- and the
// multi-head softmax(Q K^T) V against the cached text(+audio) keys / values, plus an optional DECOUPLED audio branch
// (second key / value set with its own softmax, scaled by lambda and added).  Q never leaves the SM:
//
//   CTA = 128 rows of the residual stream x one group of G heads (G * dp <= 256 accumulator columns; dp = ceil16(d)).
//   phase 1  Qacc[128, G*dp] = X[128, C] . Wq_g^T      tcgen05.mma SS, operands by TMA through a 2-stage ring (the first two
//            stages are issued before the TMEM allocation / CTA sync); the weight box is 3-D {64 k, dp rows, G heads} over
//            Wq[heads][d][C]: rows d..dp-1 of every head are out of bounds and arrive as zeros, so the head-padded layout
//            the later MMAs need never exists in memory.
//   phase 2  thread = row: q = rstd (acc - mean colsum) + bias -> bf16, written back IN PLACE into TMEM (two per column),
//            where it is the A operand of the score MMA (TS form).  No shared-memory round trip.
//   phase 3  per head: S = Q_h K_h^T (TS, K-major K tile) -> single-pass softmax over <= 112 keys in registers (segment 1 =
//            text keys, segment 2 = decoupled audio keys, each normalised on its own; segment 2 times lambda) -> bf16 P in
//            place over S -> O_h = P V_h (TS, V MN-major straight from its natural layout) -> bf16 -> global.
//            K / V of a head is ONE contiguous pre-swizzled image in the packed cache (xattn_pack_kv_kernel, once per
//            image): a single cp.async.bulk per head into the (now idle) phase-1 ring memory.  Row-granular tensor boxes
//            over the natural [B][T][2C] cache cost the TMA unit ~25 cycles per 80-byte row (measured: 3400-4300 cycles per
//            head, the whole head pipeline waited on them).
//            G >= 2: two score buffers, heads alternate -- the score MMA of head h+1 retires during the softmax of head h,
//            and the softmax warps run softmax(h+1) before the epilogue of head h, so P V (h) is off their critical path.
//            The output accumulator ALIASES the bf16 Q of heads 0 / 1 (dead once their score MMAs are issued; tcgen05.mma
//            retires in issue order), which is what lets Q (G*dp/2) + 2 x 80 score columns fit 256 TMEM columns.
//   warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2..5 = convert / softmax / epilogue.
//   Two CTAs per SM (<= 256 TMEM columns, <= 112 KB smem each) overlap one CTA's projection (tensor pipe) with the other's
//   softmax (MUFU).
//   Measured dead ends (B200): per-column masking / segment selects inside the softmax (if-converted: 3.5x the
//   instructions); generic-pointer smem reads of the folded-LayerNorm vectors (LD.E instead of LDS: +550 cycles per 64
//   columns, see align_smem_1024); half of the exponentials as an FMA-pipe polynomial (no gain: the warps wait on
//   mbarriers, not on MUFU); four-chunk epilogue loads (register spills under the 168-register cap of 2 CTAs / SM);
//   transposing the output rows through shared memory for row-major 16-byte stores (58.8 vs 53.4 us: the extra shared-
//   memory pass and index arithmetic cost more than the 32-line store instructions they replace).
#include <float.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace c2d {

using namespace tc;

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);

constexpr int XA_BM = 128, XA_BK = 64, XA_STAGES = 2, XA_THREADS = 192;
constexpr int XA_A_BYTES = XA_BM * XA_BK * 2;
constexpr int XA_MAX_SLOTS = 3;

struct XaParams {
  bf16* o;
  long long ldo;
  const long long* ln_stats;     // [M][2] fixed-point (sum, sumsq) of the rows of x, or null (plain projection)
  const float* colsum;           // [heads*d] column sums of the gamma-scaled weight (null without LayerNorm fold)
  const float* qbias;            // [heads*d] or null
  float ln_invK, ln_eps;
  int M, Nq, rpt;                // rows, rows per sample, rows per tile (min(128, Nq))
  int G, d, dp, NG;              // heads per CTA, head dim, ceil16(d), G * dp
  int n1, s2, n2, NS;            // text keys, first column of segment 2 (= ceil16(n1)), audio keys, S columns
  float scale_log2, lambda2;
  int num_kb;                    // C / 64
  int region_bytes, stage_bytes, slots, slot_bytes, kvblk_bytes;
  int tmem_cols, col_s, col_s1, col_o;   // col_s1 != col_s: two score buffers, heads alternate (G >= 2)
  int heads;
  const uint8_t* kvp;            // packed K / V cache: [B][heads] images of slot_bytes (see xattn_pack_kv_kernel)
  long long* dbg;                // optional timeline of one mid-grid CTA: [role][64] clock64 stamps (C2D_XATTN_DBG)
};

#define XA_STAMP()                                           \
  do {                                                       \
    if (dbg_on && lane == 0 && dbgi < 64) dbg_row[dbgi++] = clock64(); \
  } while (0)

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ float xa_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void xa_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void xa_st1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
// (d0, d1) = (a0, a1) * (b, b) + (c, c)
__device__ __forceinline__ void xa_ffma2(float& d0, float& d1, float a0, float a1, float b, float c) {
  asm("{.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %4};\n\t"
      "mov.b64 rc, {%5, %5};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
// (d0, d1) = (a0, a1) * (b, b) + (c0, c1)
__device__ __forceinline__ void xa_ffma2v(float& d0, float& d1, float a0, float a1, float b, float c0, float c1) {
  asm("{.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %4};\n\t"
      "mov.b64 rc, {%5, %6};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void xa_fadd2(float& d0, float& d1, float a0, float a1) {
  asm("{.reg .b64 ra, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rd, {%0, %1};\n\t"
      "add.rn.f32x2 rd, rd, ra;\n\t"
      "mov.b64 {%0, %1}, rd;}"
      : "+f"(d0), "+f"(d1)
      : "f"(a0), "f"(a1));
}
__device__ __forceinline__ void xa_fmul2(float& d0, float& d1, float a0, float a1, float b) {
  asm("{.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %4};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b));
}
// explicit shared-space 16-byte load (the kernel's smem base pointer is a re-aligned generic pointer: plain C++
// dereferences compile to generic LD.E, which take the global-memory path before they are routed to shared memory)
__device__ __forceinline__ float4 xa_lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float xa_max3(float a, float b, float c) {
  float m;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(m) : "f"(a), "f"(b), "f"(c));
  return m;
}
__device__ __forceinline__ uint32_t xa_pack(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// columns >= nv of the 16-column chunk C get a score of -inf (exp2 gives exactly 0)
template <int C, int N>
__device__ __forceinline__ void xa_mask_chunk(uint32_t (&s)[N], int nv) {
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (i >= nv) s[C * 16 + i] = 0xff800000u;
}
// NBLK = ceil(d / 64) 64-column blocks per K / V tile; NCH = NS / 16 score chunks (5: <= 80 keys, 6: <= 96 or 80 + audio,
// 7: 96 + audio).  With a decoupled audio branch its 16 columns are always the LAST chunk.
template <int NBLK, int NCH>
__global__ void __launch_bounds__(XA_THREADS, 2)
xattn_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const XaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.region_bytes);
  uint64_t* full = bars + 0;                 // [2]  phase-1 ring
  uint64_t* empty = bars + 2;                // [2]
  uint64_t* q_done = bars + 4;               //      all projection MMAs retired (accumulator complete, ring memory idle)
  uint64_t* qbf_ready = bars + 5;            //      bf16 Q written back to TMEM (128 arrivals)
  uint64_t* kv_full = bars + 6;              // [3]
  uint64_t* kv_empty = bars + 9;             // [3]
  uint64_t* s_full = bars + 12;              // [2]  one per score buffer
  uint64_t* p_full = bars + 14;              // [2]  128 arrivals
  uint64_t* o_full = bars + 16;
  uint64_t* o_free = bars + 17;              //      128 arrivals: O of the previous head is in registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
  float* s_cs = reinterpret_cast<float*>(smem + p.region_bytes + 256);     // [NG] column sums, head-padded
  float* s_qb = s_cs + 256;                                                // [NG] bias, head-padded

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = blockIdx.x;
  const int m0 = blockIdx.y * p.rpt;
  const int b = m0 / p.Nq;
  const int head0 = grp * p.G;
  const bool dbg_on = p.dbg && blockIdx.x == 0 && blockIdx.y == (gridDim.y >> 1) && (warp <= 2);
  long long* dbg_row = p.dbg + warp * 64;
  int dbgi = 0;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX); prefetch_tmap(&tmW);
    for (int i = 0; i < XA_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(q_done, 1);
    mbar_init(qbf_ready, 128);
    for (int i = 0; i < XA_MAX_SLOTS; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128); }
    mbar_init(o_full, 1);
    mbar_init(o_free, 128);
    fence_barrier_init();
    pdl_wait();           // x / row statistics come from the kernel just before this one
    // the first ring stages go out before the TMEM allocation and the CTA-wide sync: their latency overlaps the setup
    for (int kb = 0; kb < XA_STAGES && kb < p.num_kb; ++kb) {
      uint8_t* st = smem + kb * p.stage_bytes;
      mbar_arrive_expect_tx(&full[kb], (uint32_t)p.stage_bytes);
      tma_load_2d(st, &tmX, &full[kb], kb * XA_BK, m0);
      tma_load_3d(st + XA_A_BYTES, &tmW, &full[kb], kb * XA_BK, 0, head0);
    }
  }
  if (warp == 1) tmem_alloc_n(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();             // (no-op for thread 0, which already waited) nothing above touches global memory
  pdl_trigger();

  if (warp == 0) {
    // ===================== TMA producer =====================
    XA_STAMP();
    for (int kb = XA_STAGES; kb < p.num_kb; ++kb) {
      const int s = kb % XA_STAGES;
      mbar_wait(&empty[s], ((uint32_t)(kb / XA_STAGES) & 1u) ^ 1u);
      XA_STAMP();
      if (elect_one()) {
        uint8_t* st = smem + s * p.stage_bytes;
        mbar_arrive_expect_tx(&full[s], (uint32_t)p.stage_bytes);
        tma_load_2d(st, &tmX, &full[s], kb * XA_BK, m0);
        tma_load_3d(st + XA_A_BYTES, &tmW, &full[s], kb * XA_BK, 0, head0);
      }
      __syncwarp();
    }
    mbar_wait_backoff(q_done, 0);            // the ring memory is idle from here on: it becomes the K / V slots
    XA_STAMP();
    // K / V of one head = ONE contiguous, pre-swizzled image in the packed cache (c2d_xattn_pack_kv): a single bulk copy
    // instead of row-granular tensor boxes (80-byte rows at a 2C stride cost the TMA unit ~25 cycles per row).
    const uint8_t* kv_src = p.kvp + ((size_t)b * p.heads + head0) * (size_t)p.slot_bytes;
    for (int hh = 0; hh < p.G; ++hh) {
      const int sl = hh % p.slots;
      mbar_wait_backoff(&kv_empty[sl], ((uint32_t)(hh / p.slots) & 1u) ^ 1u);
      XA_STAMP();
      if (elect_one()) {
        mbar_arrive_expect_tx(&kv_full[sl], (uint32_t)p.slot_bytes);
        bulk_load_1d(smem + sl * p.slot_bytes, kv_src + (size_t)hh * p.slot_bytes, (uint32_t)p.slot_bytes, &kv_full[sl]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_q = make_idesc_bf16(128, p.NG, 0, 0);
    const uint32_t idesc_qk = make_idesc_bf16(128, p.NS, 0, 0);
    const uint32_t idesc_pv = make_idesc_bf16(128, p.dp, 0, 1);           // B (= V) is MN-major
    for (int kb = 0; kb < p.num_kb; ++kb) {
      const int s = kb % XA_STAGES;
      mbar_wait(&full[s], (uint32_t)(kb / XA_STAGES) & 1u);
      XA_STAMP();
      tc_fence_after();
      if (elect_one()) {
        const uint64_t ad = make_desc_k_sw128(smem_u32(smem + s * p.stage_bytes));
        const uint64_t bd = make_desc_k_sw128(smem_u32(smem + s * p.stage_bytes + XA_A_BYTES));
#pragma unroll
        for (int kk = 0; kk < XA_BK / 16; ++kk)
          umma_f16(tmem_base, ad + (uint64_t)(2 * kk), bd + (uint64_t)(2 * kk), idesc_q, (kb > 0 || kk > 0) ? 1u : 0u);
        umma_commit(&empty[s]);
        if (kb == p.num_kb - 1) umma_commit(q_done);
      }
      __syncwarp();
    }
    mbar_wait(qbf_ready, 0);
    XA_STAMP();
    tc_fence_after();
    const bool dual = p.col_s1 != p.col_s;
    const uint32_t tmem_s0 = tmem_base + (uint32_t)p.col_s, tmem_s1 = tmem_base + (uint32_t)p.col_s1;
    const uint32_t tmem_o = tmem_base + (uint32_t)p.col_o;
    const int ksteps_qk = p.dp >> 4, ksteps_pv = p.NS >> 4;
    // S_h = Q_h K_h^T into score buffer (h & 1) [dual] / the single buffer.  tcgen05.mma retires in issue order, so a
    // score MMA may be issued right behind the P V MMA whose P columns (same buffer) it overwrites, and P V of head 0
    // -- whose accumulator aliases the bf16 Q of heads 0 and 1 in dual mode -- right behind the score MMA of head 1.
    auto issue_qk = [&](int hh) {
      const int sl = hh % p.slots, bi = dual ? (hh & 1) : 0;
      mbar_wait(&kv_full[sl], (uint32_t)(hh / p.slots) & 1u);
      XA_STAMP();
      tc_fence_after();
      if (elect_one()) {
        const uint64_t kd = make_desc_k_sw128(smem_u32(smem + sl * p.slot_bytes));
        const uint32_t tq = tmem_base + (uint32_t)(hh * (p.dp >> 1));
        for (int kk = 0; kk < ksteps_qk; ++kk)
          umma_f16_ts(bi ? tmem_s1 : tmem_s0, tq + (uint32_t)kk * 8, kd + (uint64_t)((kk >> 2) * (p.kvblk_bytes >> 4) + (kk & 3) * 2), idesc_qk,
                      kk > 0 ? 1u : 0u);
        umma_commit(&s_full[bi]);
      }
      __syncwarp();
    };
    issue_qk(0);
    if (dual && p.G > 1) issue_qk(1);
    for (int hh = 0; hh < p.G; ++hh) {
      const int sl = hh % p.slots, bi = dual ? (hh & 1) : 0;
      const uint32_t use = dual ? (uint32_t)(hh >> 1) : (uint32_t)hh;      // how often this buffer has been used before
      mbar_wait(&p_full[bi], use & 1u);
      if (hh > 0) mbar_wait(o_free, (uint32_t)(hh - 1) & 1u);
      XA_STAMP();
      tc_fence_after();
      if (elect_one()) {
        const uint64_t vd = make_desc_mn_sw128(smem_u32(smem + sl * p.slot_bytes) + NBLK * p.kvblk_bytes, (uint32_t)p.kvblk_bytes, 1024);
        for (int kk = 0; kk < ksteps_pv; ++kk)            // 16 keys = 8 TMEM columns of P = 2 KB of V rows
          umma_f16_ts(tmem_o, (bi ? tmem_s1 : tmem_s0) + (uint32_t)kk * 8, vd + (uint64_t)(kk * (2048 >> 4)), idesc_pv, kk > 0 ? 1u : 0u);
        umma_commit(o_full);
        umma_commit(&kv_empty[sl]);
      }
      __syncwarp();
      const int nxt = dual ? hh + 2 : hh + 1;
      if (nxt < p.G) issue_qk(nxt);
    }
  } else {
    // ===================== convert / softmax / epilogue (warps 2..5; thread = row) =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int m = m0 + row;
    const bool row_ok = row < p.rpt && m < p.M;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const int ct = threadIdx.x - 64;
    // row statistics of the folded LayerNorm: issue the load before anything else
    longlong2 st = make_longlong2(0, 0);
    if (p.ln_stats && m < p.M) st = *reinterpret_cast<const longlong2*>(p.ln_stats + 2 * (long long)m);
    for (int n = ct; n < p.NG; n += 128) {
      const int hh = n / p.dp, j = n - hh * p.dp;
      const int src = (head0 + hh) * p.d + j;
      s_cs[n] = (j < p.d && p.colsum) ? __ldg(p.colsum + src) : 0.f;
      s_qb[n] = (j < p.d && p.qbias) ? __ldg(p.qbias + src) : 0.f;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const uint32_t cs_addr = smem_u32(s_cs), qb_addr = smem_u32(s_qb);
    float rstd = 1.f, nmr = 0.f;
    if (p.ln_stats) {
      const float inv = p.ln_invK * (1.0f / 1048576.0f);
      const float mean = (float)st.x * inv;
      const float var = fmaxf(fmaf(-mean, mean, (float)st.y * inv), 0.f);
      rstd = rsqrtf(var + p.ln_eps);
      nmr = -mean * rstd;
    }
    // ---- phase 2: Q fp32 -> affine -> bf16, in place (columns [c, c+32) -> [c/2, c/2+16); ascending c never overtakes).
    //      Two 32-column chunks per TMEM round trip.
    XA_STAMP();
    mbar_wait(q_done, 0);
    XA_STAMP();
    tc_fence_after();
    for (int c = 0; c < p.NG; c += 64) {
      const bool two = c + 32 < p.NG;
      uint32_t ra[32], rb[32];
      tmem_ld_32x32(tmem_base + lane_off + (uint32_t)c, ra);
      if (two) tmem_ld_32x32(tmem_base + lane_off + (uint32_t)(c + 32), rb);
      tmem_ld_wait();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (half == 0 || two) {
          uint32_t (&r)[32] = half ? rb : ra;
          const int cc = c + half * 32;
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 c4 = xa_lds128(cs_addr + (uint32_t)(cc + j) * 4);
            const float4 b4 = xa_lds128(qb_addr + (uint32_t)(cc + j) * 4);
            float t0, t1, t2, t3, v0, v1, v2, v3;
            xa_ffma2v(t0, t1, c4.x, c4.y, nmr, b4.x, b4.y);
            xa_ffma2v(t2, t3, c4.z, c4.w, nmr, b4.z, b4.w);
            xa_ffma2v(v0, v1, __uint_as_float(r[j]), __uint_as_float(r[j + 1]), rstd, t0, t1);
            xa_ffma2v(v2, v3, __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]), rstd, t2, t3);
            pk[j >> 1] = xa_pack(v0, v1);
            pk[(j >> 1) + 1] = xa_pack(v2, v3);
          }
          xa_st16(tmem_base + lane_off + (uint32_t)(cc >> 1), pk);
        }
      }
      XA_STAMP();
    }
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(qbf_ready);
    XA_STAMP();

    // ---- phase 3.  Softmax in registers, 16-column chunks, no per-column decisions: the few columns between the last
    //      key and the chunk boundary are overwritten with -inf in TMEM first (exp2 gives exactly 0).
    const bool dual = p.col_s1 != p.col_s;
    const uint32_t tmem_s0 = tmem_base + (uint32_t)p.col_s + lane_off, tmem_s1 = tmem_base + (uint32_t)p.col_s1 + lane_off;
    const uint32_t tmem_o = tmem_base + (uint32_t)p.col_o + lane_off;
    const bool has2 = p.n2 > 0;
    constexpr int NC1 = NCH - 1;                      // chunks that always belong to segment 1
    const float sc = p.scale_log2;
    // software pipeline: softmax(h) then the epilogue of head h-1, whose P V ran meanwhile
    for (int step = 0; step <= p.G; ++step) {
     if (step < p.G) {
      const int hh = step, bi = dual ? (hh & 1) : 0;
      const uint32_t use = dual ? (uint32_t)(hh >> 1) : (uint32_t)hh;
      const uint32_t tmem_s = bi ? tmem_s1 : tmem_s0;
      mbar_wait(&s_full[bi], use & 1u);
      XA_STAMP();
      tc_fence_after();
      // keys past the end of a segment: in the common case only the LAST chunk of a segment is ragged and is masked in
      // registers below; shorter key sets (several ragged chunks) overwrite the columns in TMEM first
      const int last1 = has2 ? NCH - 2 : NCH - 1;                 // last chunk of segment 1
      const bool reg_mask = p.n1 > last1 * 16;
      if (!reg_mask) {
        for (int c = p.n1; c < p.s2; ++c) xa_st1(tmem_s + (uint32_t)c, 0xff800000u);
        tmem_st_wait();
      }
      uint32_t s[NCH * 16];
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        uint32_t t[16];
        tmem_ld_32x16(tmem_s + (uint32_t)c * 16, t);
#pragma unroll
        for (int i = 0; i < 16; ++i) s[c * 16 + i] = t[i];
      }
      tmem_ld_wait();
      if (has2) {
        if (reg_mask) xa_mask_chunk<(NCH >= 2 ? NCH - 2 : 0)>(s, p.n1 - (NCH - 2) * 16);
        xa_mask_chunk<NCH - 1>(s, p.n2);
      } else if (reg_mask) {
        xa_mask_chunk<NCH - 1>(s, p.n1 - (NCH - 1) * 16);
      }
      XA_STAMP();
      // segment-1 maximum over the first NC1 chunks (four FMNMX3 chains); the last chunk joins it or is segment 2
      float ma = __uint_as_float(s[0]), mb = __uint_as_float(s[1]), mc = __uint_as_float(s[2]), md = __uint_as_float(s[3]);
#pragma unroll
      for (int i = 4; i + 7 < NC1 * 16; i += 8) {
        ma = xa_max3(ma, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
        mb = xa_max3(mb, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
        mc = xa_max3(mc, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
        md = xa_max3(md, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
      }
      ma = xa_max3(ma, __uint_as_float(s[NC1 * 16 - 4]), __uint_as_float(s[NC1 * 16 - 3]));
      mb = xa_max3(mb, __uint_as_float(s[NC1 * 16 - 2]), __uint_as_float(s[NC1 * 16 - 1]));
      float m1 = fmaxf(xa_max3(ma, mb, mc), md);
      float la = xa_max3(__uint_as_float(s[NC1 * 16]), __uint_as_float(s[NC1 * 16 + 1]), __uint_as_float(s[NC1 * 16 + 2]));
      float lb = xa_max3(__uint_as_float(s[NC1 * 16 + 3]), __uint_as_float(s[NC1 * 16 + 4]), __uint_as_float(s[NC1 * 16 + 5]));
#pragma unroll
      for (int i = 6; i + 3 < 16; i += 4) {
        la = xa_max3(la, __uint_as_float(s[NC1 * 16 + i]), __uint_as_float(s[NC1 * 16 + i + 1]));
        lb = xa_max3(lb, __uint_as_float(s[NC1 * 16 + i + 2]), __uint_as_float(s[NC1 * 16 + i + 3]));
      }
      la = xa_max3(la, __uint_as_float(s[NC1 * 16 + 14]), __uint_as_float(s[NC1 * 16 + 15]));
      const float mlast = fmaxf(la, lb);
      if (!has2) m1 = fmaxf(m1, mlast);
      const float nm1 = -m1 * sc;
      const float nml = has2 ? -mlast * sc : nm1;
      XA_STAMP();
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int i = 0; i < NC1 * 16; i += 4) {
        float x0, x1, x2, x3;
        xa_ffma2(x0, x1, __uint_as_float(s[i]), __uint_as_float(s[i + 1]), sc, nm1);
        xa_ffma2(x2, x3, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]), sc, nm1);
        x0 = xa_ex2(x0); x1 = xa_ex2(x1); x2 = xa_ex2(x2); x3 = xa_ex2(x3);
        xa_fadd2(a0, a1, x0, x1);
        xa_fadd2(a2, a3, x2, x3);
        s[i] = __float_as_uint(x0); s[i + 1] = __float_as_uint(x1); s[i + 2] = __float_as_uint(x2); s[i + 3] = __float_as_uint(x3);
      }
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
      for (int i = NC1 * 16; i < NCH * 16; i += 4) {
        float x0, x1, x2, x3;
        xa_ffma2(x0, x1, __uint_as_float(s[i]), __uint_as_float(s[i + 1]), sc, nml);
        xa_ffma2(x2, x3, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]), sc, nml);
        x0 = xa_ex2(x0); x1 = xa_ex2(x1); x2 = xa_ex2(x2); x3 = xa_ex2(x3);
        xa_fadd2(l0, l1, x0, x1);
        xa_fadd2(l2, l3, x2, x3);
        s[i] = __float_as_uint(x0); s[i + 1] = __float_as_uint(x1); s[i + 2] = __float_as_uint(x2); s[i + 3] = __float_as_uint(x3);
      }
      const float suml = (l0 + l1) + (l2 + l3);
      const float sum1 = (a0 + a1) + (a2 + a3) + (has2 ? 0.f : suml);
      const float r1 = 1.f / sum1;
      const float rl = has2 ? p.lambda2 / suml : r1;
      XA_STAMP();
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const float f0 = c == NC1 ? rl : r1;
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          float y0, y1;
          xa_fmul2(y0, y1, __uint_as_float(s[c * 16 + i]), __uint_as_float(s[c * 16 + i + 1]), f0);
          pk[i >> 1] = xa_pack(y0, y1);
        }
        tmem_st_32x8(tmem_s + (uint32_t)c * 8, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[bi]);
      XA_STAMP();
     }
     if (step >= 1) {
      // ---- O_h -> bf16 -> global
      const int hh = step - 1;
      mbar_wait(o_full, (uint32_t)hh & 1u);
      XA_STAMP();
      tc_fence_after();
      bf16* orow = p.o + (long long)m * p.ldo + (long long)(head0 + hh) * p.d;
      for (int c = 0; c < p.dp; c += 16) {
        uint32_t r[16];
        tmem_ld_32x16(tmem_o + (uint32_t)c, r);
        tmem_ld_wait();
        if (c + 16 >= p.dp) {                       // last chunk in registers: P V of the next head may overwrite O
          tc_fence_before();
          mbar_arrive(o_free);
        }
        if (row_ok) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (c + g * 8 < p.d) {                  // d % 8 == 0: whole 8-element groups are valid or not
              uint4 o4;
              o4.x = xa_pack(__uint_as_float(r[g * 8 + 0]), __uint_as_float(r[g * 8 + 1]));
              o4.y = xa_pack(__uint_as_float(r[g * 8 + 2]), __uint_as_float(r[g * 8 + 3]));
              o4.z = xa_pack(__uint_as_float(r[g * 8 + 4]), __uint_as_float(r[g * 8 + 5]));
              o4.w = xa_pack(__uint_as_float(r[g * 8 + 6]), __uint_as_float(r[g * 8 + 7]));
              *reinterpret_cast<uint4*>(orow + c + g * 8) = o4;
            }
          }
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

static inline bool xa_al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

struct XaPlan {
  int G, dp, NG, nblk, s2, NS, slots, slot_bytes, kvblk_bytes, stage_bytes, region_bytes, tmem_cols, col_s, col_s1, col_o, smem_bytes;
};

// Picks the head grouping and the shared / tensor memory plan; false when the shape is outside the kernel's envelope.
static bool xa_plan(int C, int heads, int T, int T2, XaPlan& pl) {
  if (heads <= 0 || C % heads || C % XA_BK) return false;
  const int d = C / heads;
  if (d % 8 || d < 16 || d > 192) return false;
  pl.dp = (d + 15) & ~15;
  pl.nblk = (d + 63) / 64;
  pl.G = 0;
  for (int g = heads; g >= 1; --g)
    if (heads % g == 0 && g * pl.dp <= 256 && (g * pl.dp) % 32 == 0) { pl.G = g; break; }
  if (!pl.G) return false;
  pl.NG = pl.G * pl.dp;
  if (T < 1 || T > 96 || T2 < 0 || T2 > 16) return false;      // <= 96 text(+concat) keys + 16 decoupled audio keys = 7 chunks
  pl.s2 = (T + 15) & ~15;
  if (pl.s2 < 80) pl.s2 = 80;                  // kernel instances exist for 5, 6, 7 score chunks (columns >= T are masked)
  pl.NS = pl.s2 + (T2 ? 16 : 0);
  pl.kvblk_bytes = pl.NS * 128;
  pl.slot_bytes = 2 * pl.nblk * pl.kvblk_bytes;
  pl.stage_bytes = XA_A_BYTES + pl.NG * 128;
  const int ring = XA_STAGES * pl.stage_bytes;
  const int budget = 110 * 1024;               // two CTAs per SM
  pl.slots = pl.G < XA_MAX_SLOTS ? pl.G : XA_MAX_SLOTS;
  while (pl.slots > 1 && pl.slots * pl.slot_bytes > (ring > budget ? ring : budget)) --pl.slots;
  pl.region_bytes = pl.slots * pl.slot_bytes > ring ? pl.slots * pl.slot_bytes : ring;
  // TMEM: bf16 Q at [0, NG/2).  G >= 2 heads: two score buffers (heads alternate; the score MMA of the next head runs
  // during this head's softmax) and the output accumulator ALIASES the bf16 Q of heads 0 and 1, which are dead once
  // their score MMAs have been issued.  G == 1: one score buffer, separate accumulator.
  pl.col_s = pl.NG >> 1;
  int cols;
  if (pl.G >= 2 && pl.slots >= 2) {
    pl.col_s1 = pl.col_s + pl.NS;
    pl.col_o = 0;
    cols = pl.col_s1 + pl.NS;
  } else {
    pl.col_s1 = pl.col_s;
    pl.col_o = pl.col_s + pl.NS;
    cols = pl.col_o + pl.dp;
  }
  if (cols > 512) return false;
  pl.tmem_cols = cols <= 256 ? 256 : 512;
  pl.smem_bytes = pl.region_bytes + 256 + 2 * 256 * 4 + 1024;
  return pl.smem_bytes <= 227 * 1024;
}

bool xattn_tc_supported(int C, int heads, int Nq, int T, int T2) {
  XaPlan pl;
  return xa_plan(C, heads, T, T2, pl) && (Nq % XA_BM == 0 || (Nq < XA_BM && Nq % 32 == 0));
}

template <int NBLK, int NCH>
static int xa_launch(const CUtensorMap& tx, const CUtensorMap& tw, const XaParams& p, int smem_bytes, dim3 grid, cudaStream_t s) {
  static int smem_set[C2D_MAX_DEVICES] = {};
  if (int rc = ensure_dyn_smem(xattn_tc_kernel<NBLK, NCH>, smem_bytes, smem_set, "xattn_tc")) return rc;
  launch_pdl(xattn_tc_kernel<NBLK, NCH>, grid, dim3(XA_THREADS), (size_t)smem_bytes, s, tx, tw, p);
  return check_launch("xattn_tc");
}

template <int NBLK>
static int xa_launch_nch(const CUtensorMap& tx, const CUtensorMap& tw, const XaParams& p, int smem_bytes, dim3 grid, cudaStream_t s) {
  switch (p.NS >> 4) {
    case 5: return xa_launch<NBLK, 5>(tx, tw, p, smem_bytes, grid, s);
    case 6: return xa_launch<NBLK, 6>(tx, tw, p, smem_bytes, grid, s);
    default: return xa_launch<NBLK, 7>(tx, tw, p, smem_bytes, grid, s);
  }
}

// ---- packed K / V cache.  Per (batch, head): [K block 0 .. NBLK-1][V block 0 .. NBLK-1], every block the exact
// shared-memory image of an NS-row x 64-column bf16 tile in the 128-byte-swizzled layout tcgen05 reads (row r at
// r * 128 B, its 16-byte chunk j at position j ^ (r & 7)); rows [0, T) = text(+audio) keys, rows [s2, s2 + T2) = the
// decoupled audio keys, everything else (rows past the keys, columns >= d) zero.
__global__ void xattn_pack_kv_kernel(const bf16* __restrict__ k, const bf16* __restrict__ v, long long ldkv, long long bskv,
                                     int T, const bf16* __restrict__ k2, const bf16* __restrict__ v2, long long ldkv2,
                                     long long bskv2, int T2, uint4* __restrict__ out, int heads, int d, int nblk, int s2,
                                     int NS) {
  const int bh = blockIdx.x, b = bh / heads, h = bh % heads;
  const int chunks = 2 * nblk * NS * 8;                 // 16-byte chunks of this (batch, head) image
  uint4* dst = out + (size_t)bh * chunks;
  for (int i = threadIdx.x; i < chunks; i += blockDim.x) {
    const int j = i & 7, r = (i >> 3) % NS, blk = (i >> 3) / NS % nblk, isv = (i >> 3) / NS / nblk;
    const int col = blk * 64 + j * 8;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (col < d) {                                      // d % 8 == 0
      const bf16* src = nullptr;
      if (r < T) src = (isv ? v : k) + (long long)b * bskv + (long long)r * ldkv;
      else if (r >= s2 && r < s2 + T2) src = (isv ? v2 : k2) + (long long)b * bskv2 + (long long)(r - s2) * ldkv2;
      if (src) val = *reinterpret_cast<const uint4*>(src + (long long)h * d + col);
    }
    dst[((isv * nblk + blk) * NS + r) * 8 + (j ^ (r & 7))] = val;
  }
}

long long xattn_packed_bytes(int C, int heads, int T, int T2) {
  XaPlan pl;
  if (!xa_plan(C, heads, T, T2, pl)) return 0;
  return (long long)pl.slot_bytes * heads;              // per batch element
}

int xattn_pack_kv(const void* k, const void* v, long long ldkv, long long bskv, int T, const void* k2, const void* v2,
                  long long ldkv2, long long bskv2, int T2, void* packed, int B, int C, int heads, cudaStream_t s) {
  XaPlan pl;
  if (!xa_plan(C, heads, T, T2, pl)) {
    set_error("xattn_pack_kv: shape outside the fused kernel (C=%d heads=%d T=%d T2=%d)", C, heads, T, T2);
    return C2D_ERR_UNSUPPORTED;
  }
  C2D_REQUIRE(ldkv % 8 == 0 && bskv % 8 == 0 && xa_al16(k) && xa_al16(v) && xa_al16(packed),
              "xattn_pack_kv: strides must be multiples of 8 elements and pointers 16-byte aligned");
  C2D_REQUIRE(!T2 || (k2 && v2 && ldkv2 % 8 == 0 && bskv2 % 8 == 0 && xa_al16(k2) && xa_al16(v2)), "xattn_pack_kv: bad second K/V set");
  xattn_pack_kv_kernel<<<B * heads, 256, 0, s>>>(reinterpret_cast<const bf16*>(k), reinterpret_cast<const bf16*>(v), ldkv, bskv, T,
                                                reinterpret_cast<const bf16*>(k2), reinterpret_cast<const bf16*>(v2), ldkv2,
                                                bskv2, T2, reinterpret_cast<uint4*>(packed), heads, C / heads, pl.nblk, pl.s2,
                                                pl.NS);
  return check_launch("xattn_pack_kv");
}

// xattn_tc2.cu: persistent weight-stationary kernel for the sites whose to_q slice fits in shared memory
bool xattn_p_supported(int C, int heads, int Nq, int T, int T2);
int xattn_p(const void* x, long long ldx, const void* wq, const float* qbias, const long long* ln_stats, const float* ln_colsum,
            float ln_eps, const void* kv_packed, int T, void* o, long long ldo, int B, int Nq, int C, int heads, float scale,
            cudaStream_t s);

int xattn_tc(const void* x, long long ldx, const void* wq, const float* qbias, const long long* ln_stats,
             const float* ln_colsum, float ln_eps, const void* kv_packed, int T, int T2, float lambda2, void* o, long long ldo,
             int B, int Nq, int C, int heads, float scale, cudaStream_t s) {
  if (xattn_p_supported(C, heads, Nq, T, T2))
    return xattn_p(x, ldx, wq, qbias, ln_stats, ln_colsum, ln_eps, kv_packed, T, o, ldo, B, Nq, C, heads, scale, s);
  XaPlan pl;
  if (!xa_plan(C, heads, T, T2, pl) || !(Nq % XA_BM == 0 || (Nq < XA_BM && Nq % 32 == 0))) {
    set_error("xattn: shape outside the fused kernel (C=%d heads=%d Nq=%d T=%d T2=%d)", C, heads, Nq, T, T2);
    return C2D_ERR_UNSUPPORTED;
  }
  C2D_REQUIRE(ldx % 8 == 0 && ldo % 8 == 0 && xa_al16(x) && xa_al16(wq) && xa_al16(kv_packed) && xa_al16(o),
              "xattn: strides must be multiples of 8 elements and pointers 16-byte aligned");
  const int d = C / heads;
  const long long M = (long long)B * Nq;
  CUtensorMap tx, tw;
  {
    uint64_t dims[2] = {(uint64_t)C, (uint64_t)M};
    uint64_t st[1] = {(uint64_t)ldx * 2};
    uint32_t box[2] = {XA_BK, XA_BM};
    int rc = make_tmap_bf16(&tx, x, 2, dims, st, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)C, (uint64_t)d, (uint64_t)heads};
    uint64_t st[2] = {(uint64_t)C * 2, (uint64_t)d * C * 2};
    uint32_t box[3] = {XA_BK, (uint32_t)pl.dp, (uint32_t)pl.G};
    int rc = make_tmap_bf16(&tw, wq, 3, dims, st, box);
    if (rc) return rc;
  }
  XaParams p;
  p.o = reinterpret_cast<bf16*>(o);
  p.ldo = ldo;
  p.ln_stats = ln_stats; p.colsum = ln_stats ? ln_colsum : nullptr; p.qbias = qbias;
  p.ln_invK = 1.0f / (float)C; p.ln_eps = ln_eps;
  p.M = (int)M; p.Nq = Nq; p.rpt = Nq < XA_BM ? Nq : XA_BM;
  p.G = pl.G; p.d = d; p.dp = pl.dp; p.NG = pl.NG;
  p.n1 = T; p.s2 = pl.s2; p.n2 = T2; p.NS = pl.NS;
  p.scale_log2 = scale * 1.4426950408889634f; p.lambda2 = lambda2;
  p.num_kb = C / XA_BK;
  p.region_bytes = pl.region_bytes; p.stage_bytes = pl.stage_bytes; p.slots = pl.slots; p.slot_bytes = pl.slot_bytes;
  p.kvblk_bytes = pl.kvblk_bytes;
  p.tmem_cols = pl.tmem_cols; p.col_s = pl.col_s; p.col_s1 = pl.col_s1; p.col_o = pl.col_o;
  p.heads = heads;
  p.kvp = reinterpret_cast<const uint8_t*>(kv_packed);
  p.dbg = nullptr;
  if (const char* e = getenv("C2D_XATTN_DBG")) p.dbg = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
  dim3 grid(heads / pl.G, (unsigned)((M + p.rpt - 1) / p.rpt));
  if (pl.nblk == 1) return xa_launch_nch<1>(tx, tw, p, pl.smem_bytes, grid, s);
  if (pl.nblk == 2) return xa_launch_nch<2>(tx, tw, p, pl.smem_bytes, grid, s);
  return xa_launch_nch<3>(tx, tw, p, pl.smem_bytes, grid, s);
}

}  // namespace c2d
