// Microbenchmark (one CTA per SM): cycles per tcgen05.mma (kind::f16, M = 128, K = 16) when consecutive MMAs accumulate into
// the SAME TMEM tile (a dependent chain, as in a GEMM main loop) versus round-robin over several independent
// accumulators, for the SS form (A, B from shared memory) and the TS form (A from TMEM), N in {48, 80, 192, 256}.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I clap2diffusion_b200/csrc -o clap2diffusion_b200/csrc/build/mma_chain tools/microbench/mma_chain.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_common.cuh"

using namespace c2d::tc;

// mode 0: SS, mode 1: TS.  nacc independent accumulators used round-robin.
__global__ void __launch_bounds__(128, 1) k_chain(long long* out, int N, int mode, int nacc, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async();
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint64_t ad = make_desc_k_sw128(smem_u32(smem));
    const uint64_t bd = make_desc_k_sw128(smem_u32(smem + 16384));
    // accumulators at columns 0, 256 (N <= 256) or 0, 128, 256, 384 (N <= 128); TS A operand at column 480..
    const uint32_t stride = nacc <= 2 ? 256 : 128;
    __syncwarp();
    t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
        const uint32_t d = tb + (uint32_t)(it % nacc) * stride;
        if (mode == 0) umma_f16(d, ad + (uint64_t)(2 * (it & 3)), bd + (uint64_t)(2 * (it & 3)), idesc, 1u);
        else umma_f16_ts(d, tb + 480 + (uint32_t)(it & 3) * 8, bd + (uint64_t)(2 * (it & 3)), idesc, 1u);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (threadIdx.x == 32 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tb); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
  const int iters = 512;
  printf("form  N   nacc  cycles/MMA   (floor 128*N/256 = N/2)\n");
  for (int mode = 0; mode < 2; ++mode)
    for (int N : {48, 80, 192, 256})
      for (int nacc : {1, 2, 4}) {
        if (nacc == 2 && N > 256) continue;
        if (nacc == 4 && N > 96) continue;        // 4 x 128 columns; the TS A operand lives at 480..511
        if (nacc <= 2 && N > 224 && mode == 1) continue;
        for (int rep = 0; rep < 2; ++rep) k_chain<<<148, 128, 70 * 1024>>>(d, N, mode, nacc, iters);
        long long c = 0;
        cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        printf("%s  %3d  %d   %8.1f   (%d)%s\n", mode ? "TS" : "SS", N, nacc, (double)c / iters, N / 2, e ? cudaGetErrorString(e) : "");
      }
  return 0;
}
