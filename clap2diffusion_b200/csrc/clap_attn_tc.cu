// Swin window attention of the CLAP HTSAT tower on tcgen05 (bf16 product mode; ClapAudioSelfAttention,
// transformers/models/clap/modeling_clap.py:376-423 inside ClapAudioLayer :584-607 -- the call the reference makes through
// /root/reference/models/audio_encoder.py:164-174).  The FFMA kernels of clap_audio.cu spend 6 ms of a 20 ms pass over 256
// clips on 163 GFLOP; the data they touch (3.9 GB) is 0.6 ms of HBM time.
//
//   CTA = 128 threads = TWO 8 x 8 windows (thread = token = TMEM lane), a group of WT_HPC heads in sequence.
//   per head:  every thread gathers its token's q / k / v (3 x 48 B; cyclic shift and window partition are index
//              arithmetic, as in the FFMA kernels) and writes them into three 128-row tiles in the canonical
//              128-byte-swizzled K-major layout (chunk c of row r at c ^ (r & 7); columns d..31 stay zero)
//              S [128 x 128] = Q K^T             2 SS MMAs (K = 32), both windows at once: only the two diagonal
//                                                64 x 64 blocks are read back
//              softmax in registers              + relative-position bias, - 100 across shifted regions, exp2
//              P (bf16, off-diagonal blocks = 0) written IN PLACE over S in TMEM (64 packed columns)
//              O [128 x 32] = P V                8 TS MMAs (A = P from TMEM, B = V MN-major from its natural rows)
//              O / sum -> bf16 -> global
//   (Measured and rejected: several window pairs per CTA with the head loop outside, so that the bias is staged once per
//   head -- 433 -> 647 us at the 64 x 64 stage: the four heads of a token share cache lines, and walking other windows
//   between them loses that L1 reuse.)
//   TMEM: 128 columns per CTA (S | P in place | O over the dead upper half of S) -> three CTAs per SM (67 KB of shared
//   memory each) hide each other's serial chain (gather -> MMA -> softmax -> MMA -> store).
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace c2d {

using namespace tc;

constexpr int WT_THREADS = 128;
constexpr int WT_HPC = 4;                       // heads per CTA (HTSAT: 4, 8, 16, 32 heads)
constexpr int WT_TILE = 128 * 128;              // bytes of one 128-row x 64-column bf16 tile
constexpr int WT_OFF_Q = 0, WT_OFF_K = WT_TILE, WT_OFF_V = 2 * WT_TILE;
constexpr int WT_OFF_REG = 3 * WT_TILE;         // int[128] shift-mask region of each token
constexpr int WT_OFF_BAR = WT_OFF_REG + 512;
constexpr int WT_BIAS_PITCH = 65;               // floats per staged bias row: row-per-thread reads hit 32 distinct banks
constexpr int WT_OFF_BIAS = WT_OFF_BAR + 64;    // float[64][65]: relative-position bias of the current head
constexpr int WT_SMEM = WT_OFF_BIAS + 64 * WT_BIAS_PITCH * 4 + 1024;
constexpr uint32_t WT_COL_O = 64;               // O accumulator over the upper half of the dead S tile

__device__ __forceinline__ void wt_st32(uint32_t taddr, const uint32_t* r) {
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
__device__ __forceinline__ float wt_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t wt_pack(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int D>                                 // head dim: 24 (every HTSAT stage) or 32
__global__ void __launch_bounds__(WT_THREADS, 3)
window_attention_tc_kernel(const bf16* __restrict__ qkv, const float* __restrict__ bias, bf16* __restrict__ out, int H, int W,
                           int C, int heads, int shift, float scale, int total_windows) {
  constexpr int NCH = D / 8;                     // 16-byte chunks per row
  static_assert(D == 24 || D == 32, "window attention: head dim");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  int* sreg = reinterpret_cast<int*>(smem + WT_OFF_REG);
  float* sbias = reinterpret_cast<float*>(smem + WT_OFF_BIAS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WT_OFF_BAR);
  uint64_t* s_full = bars;                       // QK retired
  uint64_t* p_full = bars + 1;                   // 128 arrivals: P in TMEM
  uint64_t* o_full = bars + 2;                   // PV retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int ws = tid >> 6, i = tid & 63, iy = i >> 3, ix = i & 7;
  const int nw = W >> 3, nwin = (H >> 3) * nw;
  const int gw = blockIdx.x * 2 + ws;
  const bool valid = gw < total_windows;
  const int b = valid ? gw / nwin : 0, win = valid ? gw % nwin : 0;
  const int wy = win / nw, wx = win % nw;
  const int ys = wy * 8 + iy, xs = wx * 8 + ix;                          // coordinates in the shifted image
  const int y = (ys + shift) % H, x = (xs + shift) % W;                  // source token
  const long long tok = (long long)b * H * W + (long long)y * W + x;
  int reg = 0;
  if (shift > 0) {
    const int rh = ys < H - 8 ? 0 : (ys < H - shift ? 1 : 2);
    const int rw = xs < W - 8 ? 0 : (xs < W - shift ? 1 : 2);
    reg = rh * 3 + rw;
  }
  sreg[tid] = reg;
  const uint32_t row_off = (uint32_t)tid * 128u, sw = (uint32_t)(tid & 7);
  const uint32_t q_row = smem_u32(smem + WT_OFF_Q) + row_off, k_row = smem_u32(smem + WT_OFF_K) + row_off,
                 v_row = smem_u32(smem + WT_OFF_V) + row_off;
  // Cooperative gather / scatter mapping: FOUR consecutive lanes move the (up to four) 16-byte chunks of one token's
  // head slice, a warp instruction covers 8 tokens (8 cache lines) instead of 32 tokens at one chunk each (32 lines:
  // the LSU wavefronts of the row-per-thread gather bounded the kernel -- ncu: mio-throttle 5.8 / issue).  Thread tid
  // serves chunk gc of tile rows gr0 + 32 j, j = 0..3.
  const int gc = tid & 3, gr0 = tid >> 2;
  const bool g_act = gc < NCH;
  long long gtok[4];
  bool gval[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = gr0 + 32 * j, rws = r >> 6, ri = r & 63;
    const int rgw = blockIdx.x * 2 + rws;
    gval[j] = rgw < total_windows;
    const int rb = gval[j] ? rgw / nwin : 0, rwin = gval[j] ? rgw % nwin : 0;
    const int ry = ((rwin / nw) * 8 + (ri >> 3) + shift) % H, rx = ((rwin % nw) * 8 + (ri & 7) + shift) % W;
    gtok[j] = (long long)rb * H * W + (long long)ry * W + rx;
  }
  const uint32_t g_off = (uint32_t)gr0 * 128u + (((uint32_t)gc ^ (uint32_t)(gr0 & 7)) << 4);    // (gr0 + 32 j) & 7 == gr0 & 7
  // O staging over the V tile (dead once P V has retired; its padding columns only feed unused output columns)
  const uint32_t so_w = smem_u32(smem + WT_OFF_V) + (uint32_t)tid * 64u;
  const uint32_t so_r = smem_u32(smem + WT_OFF_V) + (uint32_t)gr0 * 64u + (uint32_t)gc * 16u;
  auto sts = [](uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  };
  // columns D .. 31 of every row are zero for the whole launch (K of the score MMA is padded to 32)
  if (D < 32) {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    const uint32_t o3 = (3u ^ sw) << 4;
    sts(q_row + o3, z); sts(k_row + o3, z); sts(v_row + o3, z);
  }
  if (tid == 0) {
    mbar_init(s_full, 1);
    mbar_init(p_full, WT_THREADS);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<128>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  const uint32_t tmem_s = tmem_base + lane_off, tmem_o = tmem_base + WT_COL_O + lane_off;

  constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
  constexpr uint32_t idesc_pv = make_idesc_bf16(128, 32, 0, 1);          // B (= V) is MN-major
  const uint64_t q_desc = make_desc_k_sw128(smem_u32(smem + WT_OFF_Q));
  const uint64_t k_desc = make_desc_k_sw128(smem_u32(smem + WT_OFF_K));
  const uint64_t v_desc = make_desc_mn_sw128(smem_u32(smem + WT_OFF_V), WT_TILE, 1024);
  constexpr float LOG2E = 1.4426950408889634f;

  const int h0 = blockIdx.y * WT_HPC;
  for (int hh = 0; hh < WT_HPC; ++hh) {
    const int head = h0 + hh;
    const uint32_t ph = (uint32_t)hh & 1u;
    // ---- gather q / k / v of the head into the swizzled operand tiles
    if (g_act) {
      uint4 rq[4], rk[4], rv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        rq[j] = rk[j] = rv[j] = make_uint4(0u, 0u, 0u, 0u);
        if (gval[j]) {
          const uint4* src = reinterpret_cast<const uint4*>(qkv + gtok[j] * 3 * C + head * D) + gc;
          rq[j] = __ldg(src);
          rk[j] = __ldg(src + (C >> 3));
          rv[j] = __ldg(src + (C >> 2));
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t o = g_off + (uint32_t)j * 4096u;                   // 32 rows further
        sts(smem_u32(smem + WT_OFF_Q) + o, rq[j]);
        sts(smem_u32(smem + WT_OFF_K) + o, rk[j]);
        sts(smem_u32(smem + WT_OFF_V) + o, rv[j]);
      }
    }
    // the head's bias [64][64] -> shared memory with coalesced loads: a thread reading its own 256-byte row from global
    // memory costs 32 L1 wavefronts per load instruction, which bounded the first version of this kernel
    {
      const float4* bsrc = reinterpret_cast<const float4*>(bias + (long long)head * 4096);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int e4 = tid + k * WT_THREADS;                  // float4 index: row e4 / 16, columns (e4 % 16) * 4 ..
        const float4 bv = __ldg(bsrc + e4);
        float* dst = sbias + (e4 >> 4) * WT_BIAS_PITCH + (e4 & 15) * 4;
        dst[0] = bv.x; dst[1] = bv.y; dst[2] = bv.z; dst[3] = bv.w;
      }
    }
    fence_proxy_async();                 // generic-proxy stores -> visible to the tensor core's async-proxy reads
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        umma_f16(tmem_base, q_desc, k_desc, idesc_qk, 0u);
        umma_f16(tmem_base, q_desc + 2, k_desc + 2, idesc_qk, 1u);
        umma_commit(s_full);
      }
      __syncwarp();
    }
    // bias row of this token while the MMAs run
    float t[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) t[j] = sbias[i * WT_BIAS_PITCH + j];
    mbar_wait(s_full, ph);
    tc_fence_after();
    float mx = -INFINITY;
    {
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(tmem_s + (uint32_t)ws * 64, r0);                   // this window's diagonal 64 x 64 block
      tmem_ld_32x32(tmem_s + (uint32_t)ws * 64 + 32, r1);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        t[j] = fmaf(__uint_as_float(r0[j]), scale, t[j]);
        t[32 + j] = fmaf(__uint_as_float(r1[j]), scale, t[32 + j]);
      }
    }
    if (shift > 0) {
#pragma unroll
      for (int j = 0; j < 64; ++j)
        if (sreg[ws * 64 + j] != reg) t[j] -= 100.0f;
    }
#pragma unroll
    for (int j = 0; j < 64; ++j) mx = fmaxf(mx, t[j]);
    const float nm = -mx * LOG2E;
    float sum = 0.f;
    uint32_t pk[32];
#pragma unroll
    for (int j = 0; j < 64; j += 2) {
      const float p0 = wt_ex2(fmaf(t[j], LOG2E, nm)), p1 = wt_ex2(fmaf(t[j + 1], LOG2E, nm));
      sum += p0 + p1;
      pk[j >> 1] = wt_pack(p0, p1);
    }
    {
      // P row: 128 keys = 64 packed columns; this window's keys at [ws * 32, ws * 32 + 32), the other window's are zero
      uint32_t zr[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) zr[j] = 0u;
      wt_st32(tmem_s + (uint32_t)ws * 32, pk);
      wt_st32(tmem_s + (uint32_t)(ws ^ 1) * 32, zr);
      tmem_st_wait();
    }
    tc_fence_before();
    mbar_arrive(p_full);
    if (warp == 0) {
      mbar_wait(p_full, ph);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)           // 16 keys = 8 TMEM columns of P = 2 KB of V rows
          umma_f16_ts(tmem_base + WT_COL_O, tmem_base + (uint32_t)kk * 8, v_desc + (uint64_t)(kk * (2048 >> 4)), idesc_pv,
                      kk > 0 ? 1u : 0u);
        umma_commit(o_full);
      }
      __syncwarp();
    }
    mbar_wait(o_full, ph);
    tc_fence_after();
    {
      uint32_t r[32];
      tmem_ld_32x32(tmem_o, r);
      tmem_ld_wait();
      const float inv = 1.f / sum;
      // rows -> shared memory (over the V tile: P V has retired) -> the cooperative mapping of the gather
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        uint4 o4;
        o4.x = wt_pack(__uint_as_float(r[8 * c]) * inv, __uint_as_float(r[8 * c + 1]) * inv);
        o4.y = wt_pack(__uint_as_float(r[8 * c + 2]) * inv, __uint_as_float(r[8 * c + 3]) * inv);
        o4.z = wt_pack(__uint_as_float(r[8 * c + 4]) * inv, __uint_as_float(r[8 * c + 5]) * inv);
        o4.w = wt_pack(__uint_as_float(r[8 * c + 6]) * inv, __uint_as_float(r[8 * c + 7]) * inv);
        sts(so_w + (uint32_t)c * 16u, o4);
      }
    }
    tc_fence_before();     // the next head's score MMA overwrites S / O: ordered by the barriers below
    __syncthreads();
    if (g_act) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (gval[j]) {
          uint4 o4;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(o4.x), "=r"(o4.y), "=r"(o4.z), "=r"(o4.w)
                       : "r"(so_r + (uint32_t)j * 2048u));
          *(reinterpret_cast<uint4*>(out + gtok[j] * C + head * D) + gc) = o4;
        }
      }
    }
    __syncthreads();       // the staged rows are read: the next head's gather may overwrite the V tile
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}

bool window_attention_tc_supported(const void* qkv, const float* bias, const void* out, int C, int heads) {
  const int d = heads > 0 ? C / heads : 0;
  return (d == 24 || d == 32) && heads % WT_HPC == 0 && C % 8 == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0;
}

int window_attention_tc(const void* qkv, const float* bias, void* out, int B, int H, int W, int C, int heads, int shift, float scale,
                        cudaStream_t s) {
  const int total = B * (H / 8) * (W / 8);
  const dim3 grid(ceil_div(total, 2), heads / WT_HPC);
  static int set24[C2D_MAX_DEVICES] = {}, set32[C2D_MAX_DEVICES] = {};
  if (C / heads == 24) {
    if (int rc = ensure_dyn_smem(window_attention_tc_kernel<24>, WT_SMEM, set24, "window_attention_tc")) return rc;
    window_attention_tc_kernel<24><<<grid, WT_THREADS, WT_SMEM, s>>>((const bf16*)qkv, bias, (bf16*)out, H, W, C, heads, shift, scale, total);
  } else {
    if (int rc = ensure_dyn_smem(window_attention_tc_kernel<32>, WT_SMEM, set32, "window_attention_tc")) return rc;
    window_attention_tc_kernel<32><<<grid, WT_THREADS, WT_SMEM, s>>>((const bf16*)qkv, bias, (bf16*)out, H, W, C, heads, shift, scale, total);
  }
  return check_launch("window_attention");
}

}  // namespace c2d
