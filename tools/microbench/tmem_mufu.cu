// Microbenchmarks on one SM-resident CTA per SM: tcgen05.ld throughput (bytes / clk / SM) and MUFU.EX2 rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o clap2diffusion_b200/csrc/build/tmem_mufu tools/microbench/tmem_mufu.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(512, 1) k_tmem_ld(long long* out, int iters, int nwarps_active) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps_active) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(base + (uint32_t)(((it & 3) * 128 + c * 32) & 511))
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= r[i];
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) out[0] = 0;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

__global__ void __launch_bounds__(512, 1) k_mufu(long long* out, float* sink, int iters, int nwarps_active) {
  const int warp = threadIdx.x >> 5;
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = -0.001f * (threadIdx.x + i);
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps_active) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  if (s == 123.f) sink[0] = s;
}

int main() {
  long long* d;
  float* sink;
  cudaMalloc(&d, 148 * 8);
  cudaMalloc(&sink, 4);
  long long h[148];
  const int iters = 2000;
  for (int nw : {1, 2, 4, 8, 16}) {
    k_tmem_ld<<<148, 512>>>(d, iters, nw);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("tmem_ld failed: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double bytes = (double)nw * iters * 4 * 32 * 32 * 4;
    printf("tcgen05.ld x32, %2d warps: %lld clk, %.1f B/clk/SM\n", nw, h[5], bytes / (double)h[5]);
  }
  for (int nw : {1, 4, 8, 16}) {
    k_mufu<<<148, 512>>>(d, sink, iters, nw);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mufu failed: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double ops = (double)nw * iters * 16 * 32;
    printf("MUFU.EX2, %2d warps: %lld clk, %.2f ops/clk/SM\n", nw, h[5], ops / (double)h[5]);
  }
  return 0;
}
