// tcgen05 flash-attention forward, second generation, for head_dim <= 64 and long key sequences (the UNet's
// 4096-token spatial self-attention, d = 40): the exponentials, not the tensor core, bound this shape
// (128 x 128 scores = 16384 MUFU.EX2 = 1024 SM cycles against ~450 tensor cycles), so the kernel is organised
// around keeping MUFU fed and everything else off the softmax warps' critical path.
//
//   CTA = 256 queries of one (batch, head) = two 128-row Q tiles, one CTA per SM, keys streamed 128 at a time.
//   warp 0        TMA producer: Q tiles once, K / V tiles through 4-deep mbarrier rings (4-D head-view tensor maps,
//                 64-wide boxes: columns >= d are zero-filled, no padded copies in HBM)
//   warp 1        MMA issuer.  S_qt = Q_qt K_t^T (SS, M128 N128 K16 x ceil(d/16)) into TMEM;
//                 O_qt += P_qt V_t (TS: A = bf16 P read straight from TMEM, B = V MN-major from smem)
//   warps 4..7    softmax of Q tile 0, warps 8..11 softmax of Q tile 1 (thread = query row):
//                 S -> registers in one go (then S is released: QK of the NEXT key tile overlaps this tile's
//                 exponentials), row max with 3-input FMNMX, conditional rescale (only when the max grew by > 2^8),
//                 exp2 with packed FFMA2 / FADD2, bf16 P -> TMEM with tcgen05.st.  P never touches shared memory.
//   TMEM (512 columns): S0 | S1 (128 each) | P0 | P1 (64 each, bf16 pairs) | O0 | O1 (64 each)
#include <float.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace c2d {

using namespace tc;

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);

constexpr int A2_BQ = 128, A2_QT = 2, A2_BK = 128, A2_THREADS = 384, A2_ST = 4;
constexpr int A2_TILE = 128 * 128;                      // bytes of one 128-row x 64-column bf16 tile
constexpr int A2_OFF_Q = 0;
constexpr int A2_OFF_K = A2_OFF_Q + A2_QT * A2_TILE;
constexpr int A2_OFF_V = A2_OFF_K + A2_ST * A2_TILE;
constexpr int A2_OFF_BAR = A2_OFF_V + A2_ST * A2_TILE;
constexpr int A2_SMEM = A2_OFF_BAR + 512 + 1024;
constexpr uint32_t A2_COL_S = 0, A2_COL_P = 256, A2_COL_O = 384;
constexpr float A2_RESCALE_LOG2 = 8.f;                  // rescale O only when the row max grew by more than 2^8

struct Attn2Params {
  bf16* o;
  float* lse;                   // optional [B][heads][Nq]: m + log2(sum) of the scaled scores (training forward)
  int Nq, Nkv, d, npv;
  long long ldo, bso;
  float scale_log2;
  int stagger_ns;               // initial delay of the second softmax warpgroup (ping-pong phase offset)
  long long* dbg;               // optional timeline of CTA (1,0,0): [role][tile][8] clock64 stamps (C2D_ATTN_DBG)
};

#define A2_STAMP(role, tile, ev)                                                                         \
  do {                                                                                                   \
    if (p.dbg && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0 && (threadIdx.x & 127) == 0)      \
      p.dbg[((role) * 64 + (tile)) * 8 + (ev)] = clock64();                                              \
  } while (0)

__device__ __forceinline__ float a2_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float a2_max3(float a, float b, float c) {
  float m;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(m) : "f"(a), "f"(b), "f"(c));
  return m;
}
// (d0, d1) = (a0, a1) * (b, b) + (c, c)
__device__ __forceinline__ void a2_ffma2(float& d0, float& d1, float a0, float a1, float b, float c) {
  asm("{.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %4};\n\t"
      "mov.b64 rc, {%5, %5};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
__device__ __forceinline__ void a2_fadd2(float& d0, float& d1, float a0, float a1) {
  asm("{.reg .b64 ra, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rd, {%0, %1};\n\t"
      "add.rn.f32x2 rd, rd, ra;\n\t"
      "mov.b64 {%0, %1}, rd;}"
      : "+f"(d0), "+f"(d1)
      : "f"(a0), "f"(a1));
}
// exp2 of two values on the FMA / ALU pipes (no MUFU): Cody-Waite split with the 1.5 * 2^23 rounding constant,
// degree-3 minimax polynomial for 2^f on [-0.5, 0.5] (max rel. error 7.5e-5, far below the bf16 rounding of P),
// integer part added into the exponent field.  10 instructions per pair (2 FMNMX, 2 FADD2, 4 FFMA2, 2 LEA);
// inputs below -125 (masked keys, -inf) clamp to 2^-125.
__device__ __forceinline__ void a2_exp2_poly2(float& p0, float& p1, float x0, float x1) {
  x0 = fmaxf(x0, -125.f);
  x1 = fmaxf(x1, -125.f);
  asm("{\n\t"
      ".reg .b64 x, t, f, p, k;\n\t"
      ".reg .b32 t0, t1, q0, q1;\n\t"
      "mov.b64 x, {%2, %3};\n\t"
      "mov.b64 k, 0x4B4000004B400000;\n\t"       // (1.5 * 2^23, 1.5 * 2^23)
      "add.rn.f32x2 t, x, k;\n\t"                // round(x) lands in the low mantissa bits
      "mov.b64 k, 0xCB400000CB400000;\n\t"
      "add.rn.f32x2 f, t, k;\n\t"                // round(x) as a float
      "mov.b64 k, 0xBF800000BF800000;\n\t"
      "fma.rn.f32x2 f, f, k, x;\n\t"             // f = x - round(x)
      "mov.b64 k, 0x3D61FBB03D61FBB0;\n\t"       // c3 = 0.05517167
      "mov.b64 p, 0x3E786F0D3E786F0D;\n\t"       // c2 = 0.24261113
      "fma.rn.f32x2 p, f, k, p;\n\t"
      "mov.b64 k, 0x3F31798D3F31798D;\n\t"       // c1 = 0.69326097
      "fma.rn.f32x2 p, p, f, k;\n\t"
      "mov.b64 k, 0x3F7FFB493F7FFB49;\n\t"       // c0 = 0.99992806
      "fma.rn.f32x2 p, p, f, k;\n\t"
      "mov.b64 {t0, t1}, t;\n\t"
      "mov.b64 {q0, q1}, p;\n\t"
      "shl.b32 t0, t0, 23;\n\t"
      "shl.b32 t1, t1, 23;\n\t"
      "add.s32 q0, q0, t0;\n\t"
      "add.s32 q1, q1, t1;\n\t"
      "mov.b32 %0, q0;\n\t"
      "mov.b32 %1, q1;\n\t"
      "}"
      : "=f"(p0), "=f"(p1)
      : "f"(x0), "f"(x1));
}
__device__ __forceinline__ void a2_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void a2_st1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
template <int N>
__device__ __forceinline__ void a2_setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void a2_setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void a2_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// NP = pairs out of every 8 whose exponential runs on the FMA pipe instead of MUFU (0..4)
template <int NP>
__global__ void __launch_bounds__(A2_THREADS, 1)
attn_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const Attn2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A2_OFF_BAR);
  uint64_t* q_full = bars;                 // [1]
  uint64_t* k_full = bars + 1;             // [A2_ST]
  uint64_t* k_empty = k_full + A2_ST;
  uint64_t* v_full = k_empty + A2_ST;
  uint64_t* v_empty = v_full + A2_ST;
  uint64_t* s_full = v_empty + A2_ST;      // [2]  QK(qt, t) retired
  uint64_t* s_free = s_full + 2;           // [2]  softmax(qt) holds S in registers (128 arrivals)
  uint64_t* p_full = s_free + 2;           // [2]  P(qt, t) written to TMEM (128 arrivals)
  uint64_t* o_done = p_full + 2;           // [2]  PV(qt, t) retired: P reusable, O up to date
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (A2_BQ * A2_QT), h = blockIdx.y, b = blockIdx.z;
  const int ntiles = (p.Nkv + A2_BK - 1) / A2_BK;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmK); prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < A2_ST; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 128); mbar_init(&p_full[i], 128); mbar_init(&o_done[i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      mbar_arrive_expect_tx(q_full, A2_QT * A2_TILE);
#pragma unroll
      for (int qt = 0; qt < A2_QT; ++qt) tma_load_4d(smem + A2_OFF_Q + qt * A2_TILE, &tmQ, q_full, 0, h, q0 + qt * A2_BQ, b);
      for (int t = 0; t < ntiles; ++t) {
        const int s = t % A2_ST;
        const uint32_t ph = ((uint32_t)(t / A2_ST) & 1u) ^ 1u;
        mbar_wait(&k_empty[s], ph);
        mbar_arrive_expect_tx(&k_full[s], A2_TILE);
        tma_load_4d(smem + A2_OFF_K + s * A2_TILE, &tmK, &k_full[s], 0, h, t * A2_BK, b);
        mbar_wait(&v_empty[s], ph);
        mbar_arrive_expect_tx(&v_full[s], A2_TILE);
        tma_load_4d(smem + A2_OFF_V + s * A2_TILE, &tmV, &v_full[s], 0, h, t * A2_BK, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the whole warp runs the protocol, one elected lane issues =====================
    const uint32_t idesc_qk = make_idesc_bf16(128, A2_BK, 0, 0);
    const uint32_t idesc_pv = make_idesc_bf16(128, p.npv, 0, 1);          // B (= V) is MN-major
    const uint64_t q_desc = make_desc_k_sw128(smem_u32(smem + A2_OFF_Q));
    const uint64_t k_desc = make_desc_k_sw128(smem_u32(smem + A2_OFF_K));
    const uint64_t v_desc = make_desc_mn_sw128(smem_u32(smem + A2_OFF_V), A2_TILE, 1024);
    const int ksteps = (p.d + 15) >> 4;
    const bool dbg_cta = p.dbg && blockIdx.x == 1 && blockIdx.y == 0 && blockIdx.z == 0;
    // descriptor address fields are in 16-byte units: tile t of a ring = + t * (A2_TILE >> 4), 16-element K step
    // inside a 128-byte swizzled row = + 2, 16 key rows of an MN-major V tile = + (2048 >> 4)
    auto issue_qk = [&](int qt, int t) {
      if (elect_one()) {
        const uint32_t tmem_s = tmem_base + A2_COL_S + (uint32_t)qt * A2_BK;
        const uint64_t qd = q_desc + (uint64_t)(qt * (A2_TILE >> 4));
        const uint64_t kd = k_desc + (uint64_t)((t % A2_ST) * (A2_TILE >> 4));
        for (int kk = 0; kk < ksteps; ++kk)
          umma_f16(tmem_s, qd + (uint64_t)(2 * kk), kd + (uint64_t)(2 * kk), idesc_qk, kk > 0 ? 1u : 0u);
        umma_commit(&s_full[qt]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    issue_qk(0, 0);
    issue_qk(1, 0);
    if (elect_one()) umma_commit(&k_empty[0]);
    __syncwarp();
    for (int j = 0; j < ntiles; ++j) {
      const uint32_t jp = (uint32_t)j & 1u;
      if (j + 1 < ntiles) {
        // S of tile j is in the softmax warps' registers: overwrite it with the next tile's scores
        const int s = (j + 1) % A2_ST;
        mbar_wait(&k_full[s], (uint32_t)((j + 1) / A2_ST) & 1u);
#pragma unroll
        for (int qt = 0; qt < A2_QT; ++qt) {
          mbar_wait(&s_free[qt], jp);
          if (dbg_cta && lane == 0) p.dbg[(2 * 64 + j) * 8 + qt] = clock64();
          tc_fence_after();
          issue_qk(qt, j + 1);
        }
        if (elect_one()) umma_commit(&k_empty[s]);
        __syncwarp();
        if (dbg_cta && lane == 0) p.dbg[(2 * 64 + j) * 8 + 2] = clock64();
      }
      const int vs = j % A2_ST;
      mbar_wait(&v_full[vs], (uint32_t)(j / A2_ST) & 1u);
      const uint64_t vd = v_desc + (uint64_t)(vs * (A2_TILE >> 4));
#pragma unroll
      for (int qt = 0; qt < A2_QT; ++qt) {
        mbar_wait(&p_full[qt], jp);
        if (dbg_cta && lane == 0) p.dbg[(2 * 64 + j) * 8 + 3 + qt] = clock64();
        tc_fence_after();
        if (elect_one()) {
          const uint32_t tmem_p = tmem_base + A2_COL_P + (uint32_t)qt * 64;
          const uint32_t tmem_o = tmem_base + A2_COL_O + (uint32_t)qt * 64;
#pragma unroll
          for (int kk = 0; kk < A2_BK / 16; ++kk)      // 16 keys = 8 TMEM columns of P = 2 KB of V rows
            umma_f16_ts(tmem_o, tmem_p + (uint32_t)kk * 8, vd + (uint64_t)(kk * (2048 >> 4)), idesc_pv, (j > 0 || kk > 0) ? 1u : 0u);
          umma_commit(&o_done[qt]);
          if (qt == A2_QT - 1) umma_commit(&v_empty[vs]);
        }
        __syncwarp();
      }
      if (dbg_cta && lane == 0) p.dbg[(2 * 64 + j) * 8 + 5] = clock64();
    }
  } else if (warp >= 4) {
    // ===================== softmax / correction / epilogue =====================
    const int qt = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const uint32_t tmem_s = tmem_base + A2_COL_S + (uint32_t)qt * A2_BK + lane_off;
    const uint32_t tmem_p = tmem_base + A2_COL_P + (uint32_t)qt * 64 + lane_off;
    const uint32_t tmem_o = tmem_base + A2_COL_O + (uint32_t)qt * 64 + lane_off;
    const float sc = p.scale_log2;
    float m_ref = -INFINITY;
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
    // ping-pong: the second warpgroup starts half a tile late so that its MUFU phase overlaps the first one's
    // TMEM load / row max / P store phase (both warpgroups share the SM's 16 MUFU lanes)
    // The two warpgroups take turns on the MUFU: named barrier 2 + qt opens warpgroup qt's exponential phase and is
    // armed by the other warpgroup when it leaves its own (256 = 128 syncing + 128 arriving threads).  While one
    // warpgroup runs its 128 x 128 exponentials at the full 16 / clk, the other does its TMEM load, row max, P store.
    const bool pingpong = p.stagger_ns == 0;
    if (pingpong && qt == 1) asm volatile("bar.arrive 2, 256;" ::: "memory");
    for (int j = 0; j < ntiles; ++j) {
      const uint32_t jp = (uint32_t)j & 1u;
      const int kvalid = min(A2_BK, p.Nkv - j * A2_BK);
      A2_STAMP(qt, j, 0);
      mbar_wait(&s_full[qt], jp);
      A2_STAMP(qt, j, 1);
      tc_fence_after();
      if (kvalid < A2_BK) {
        // ragged last tile: keys >= Nkv get a score of -inf, written into TMEM so the hot path has no selects
        for (int c = kvalid; c < A2_BK; ++c) a2_st1(tmem_s + c, 0xff800000u);
        tmem_st_wait();
      }
      uint32_t s[A2_BK];
#pragma unroll
      for (int c = 0; c < A2_BK / 32; ++c) a2_ld32(tmem_s + c * 32, &s[c * 32]);
      tmem_ld_wait();
      A2_STAMP(qt, j, 2);
      tc_fence_before();
      mbar_arrive(&s_free[qt]);                                   // QK(j+1) may overwrite S now
      // row max: four independent FMNMX3 chains
      float mxa = __uint_as_float(s[0]), mxb = __uint_as_float(s[1]), mxc = __uint_as_float(s[2]), mxd = __uint_as_float(s[3]);
#pragma unroll
      for (int i = 4; i + 7 < A2_BK; i += 8) {
        mxa = a2_max3(mxa, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
        mxb = a2_max3(mxb, __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
        mxc = a2_max3(mxc, __uint_as_float(s[i + 4]), __uint_as_float(s[i + 5]));
        mxd = a2_max3(mxd, __uint_as_float(s[i + 6]), __uint_as_float(s[i + 7]));
      }
      mxa = a2_max3(mxa, __uint_as_float(s[A2_BK - 4]), __uint_as_float(s[A2_BK - 3]));
      mxb = a2_max3(mxb, __uint_as_float(s[A2_BK - 2]), __uint_as_float(s[A2_BK - 1]));
      const float mx = fmaxf(a2_max3(mxa, mxb, mxc), mxd);
      const float msc = mx * sc;
      bool waited = false;
      if (j == 0) {
        m_ref = msc;
      } else {
        const bool need = msc > m_ref + A2_RESCALE_LOG2;
        if (__any_sync(0xffffffffu, need)) {
          // warp-collective rescale of the O rows that moved (alpha = 1 elsewhere); PV(j-1) must have retired
          const float alpha = need ? a2_ex2(m_ref - msc) : 1.f;
          if (need) m_ref = msc;
          mbar_wait(&o_done[qt], jp ^ 1u);
          tc_fence_after();
          waited = true;
          for (int c = 0; c < p.npv; c += 16) {
            uint32_t o[16];
            tmem_ld_32x16(tmem_o + c, o);
            tmem_ld_wait();
            uint32_t lo[8], hi[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              lo[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              hi[i] = __float_as_uint(__uint_as_float(o[8 + i]) * alpha);
            }
            tmem_st_32x8(tmem_o + c, lo);
            tmem_st_32x8(tmem_o + c + 8, hi);
          }
          tmem_st_wait();
          l0 *= alpha; l1 *= alpha; l2 *= alpha; l3 *= alpha;
        }
      }
      if (pingpong) {
        if (qt == 0) asm volatile("bar.sync 2, 256;" ::: "memory");
        else asm volatile("bar.sync 3, 256;" ::: "memory");
      }
      float neg_m = -m_ref;
      asm volatile("" : "+f"(neg_m));          // keep the exponentials below the barrier above
      A2_STAMP(qt, j, 3);
      uint32_t pk[A2_BK / 2];
#pragma unroll
      for (int i = 0; i < A2_BK; i += 2) {
        constexpr int dummy = 0; (void)dummy;
        const int pr = (i >> 1) & 7;
        float x0, x1, p0, p1;
        a2_ffma2(x0, x1, __uint_as_float(s[i]), __uint_as_float(s[i + 1]), sc, neg_m);
        if (((pr * NP) & 7) < NP) {
          a2_exp2_poly2(p0, p1, x0, x1);
        } else {
          p0 = a2_ex2(x0);
          p1 = a2_ex2(x1);
        }
        if (pr & 1) a2_fadd2(l2, l3, p0, p1);
        else a2_fadd2(l0, l1, p0, p1);
        __nv_bfloat162 h0 = __floats2bfloat162_rn(p0, p1);
        pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h0);
      }
      A2_STAMP(qt, j, 4);
      if (pingpong) {
        if (qt == 0) asm volatile("bar.arrive 3, 256;" ::: "memory");
        else if (j + 1 < ntiles) asm volatile("bar.arrive 2, 256;" ::: "memory");
      }
      if (j > 0 && !waited) mbar_wait(&o_done[qt], jp ^ 1u);          // PV(j-1) retired: the P buffer is free
      A2_STAMP(qt, j, 5);
      tc_fence_after();
      a2_st32(tmem_p, &pk[0]);
      a2_st32(tmem_p + 32, &pk[32]);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[qt]);
      A2_STAMP(qt, j, 6);
    }
    // ---- epilogue: O / l -> bf16 -> global
    mbar_wait(&o_done[qt], (uint32_t)(ntiles - 1) & 1u);
    tc_fence_after();
    const int qrow = q0 + qt * A2_BQ + row;
    const float lsum = (l0 + l1) + (l2 + l3);
    const float inv = 1.f / lsum;
    if (p.lse && qrow < p.Nq) p.lse[((long long)b * gridDim.y + h) * p.Nq + qrow] = m_ref + __log2f(lsum);
    bf16* orow = p.o + (long long)b * p.bso + (long long)qrow * p.ldo + (long long)h * p.d;
    for (int c = 0; c < p.npv; c += 16) {
      uint32_t r[16];
      tmem_ld_32x16(tmem_o + c, r);
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
    tmem_dealloc<512>(tmem_base);
  }
}

bool attention_tc2_supported(const AttnParams& p, int B) {
  (void)B;
  return p.d <= 64 && p.Nkv > A2_BK && p.scale > 0.f;
}

static int make_head_tmap2(CUtensorMap* m, const void* base, int d, int heads, int N, int B, long long ld, long long bs) {
  uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)N, (uint64_t)B};
  uint64_t st[3] = {(uint64_t)d * 2, (uint64_t)ld * 2, (uint64_t)(B > 1 ? bs : (long long)N * ld) * 2};
  uint32_t box[4] = {64, 1, 128, 1};
  return make_tmap_bf16(m, base, 4, dims, st, box);
}

// Caller (attention_tc) has validated strides / alignment.
int attention_tc2(const AttnParams& p, int B, cudaStream_t s) {
  CUtensorMap tq, tk, tv;
  int rc = make_head_tmap2(&tq, p.q, p.d, p.heads, p.Nq, B, p.ldq, p.bsq);
  if (rc) return rc;
  rc = make_head_tmap2(&tk, p.k, p.d, p.heads, p.Nkv, B, p.ldk, p.bsk);
  if (rc) return rc;
  rc = make_head_tmap2(&tv, p.v, p.d, p.heads, p.Nkv, B, p.ldv, p.bsv);
  if (rc) return rc;
  Attn2Params ap;
  ap.o = reinterpret_cast<bf16*>(p.o);
  ap.lse = p.lse;
  if (p.lse && p.lse_written) *p.lse_written = 1;
  ap.Nq = p.Nq; ap.Nkv = p.Nkv; ap.d = p.d; ap.npv = (p.d + 15) & ~15;
  ap.ldo = p.ldo; ap.bso = p.bso;
  ap.scale_log2 = p.scale * 1.4426950408889634f;
  static int stagger = -1;
  if (stagger < 0) {
    const char* e = getenv("C2D_ATTN_STAGGER");
    stagger = e ? atoi(e) : 0;
  }
  ap.stagger_ns = stagger;
  ap.dbg = nullptr;
  if (const char* e = getenv("C2D_ATTN_DBG")) ap.dbg = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
  static int poly = -1;
  if (poly < 0) {
    const char* e = getenv("C2D_ATTN_POLY");
    poly = e ? atoi(e) : 2;
    if (poly < 0 || poly > 4) poly = 2;
  }
  dim3 grid(ceil_div(p.Nq, A2_BQ * A2_QT), p.heads, B);
#define A2_LAUNCH(NP)                                                                                              \
  do {                                                                                                             \
    static int smem_set[C2D_MAX_DEVICES] = {};                                                                     \
    if (int rc = ensure_dyn_smem(attn_tc2_kernel<NP>, A2_SMEM, smem_set, "attention_tc2")) return rc;              \
    launch_pdl(attn_tc2_kernel<NP>, grid, dim3(A2_THREADS), (size_t)A2_SMEM, s, tq, tk, tv, ap);                    \
  } while (0)
  switch (poly) {
    case 0: A2_LAUNCH(0); break;
    case 1: A2_LAUNCH(1); break;
    case 4: A2_LAUNCH(4); break;
    case 3: A2_LAUNCH(3); break;
    default: A2_LAUNCH(2); break;
  }
#undef A2_LAUNCH
  return check_launch("attn_tc2");
}

}  // namespace c2d
