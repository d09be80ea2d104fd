// smem_contention_probe.cu -- do TMA writes into shared memory slow the tensor core's operand reads down?
// warp 0 issues 288 tcgen05.mma (M=128, N=96, K=16, both operands from shared memory) and times them; warp 1 meanwhile
// streams 16 KB bulk copies from global memory into four other shared-memory slots (or idles).  Reports cycles per MMA and
// the background stream's bytes per cycle.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../medical-image-denoising-using-diffusion_b200/csrc/tc_common.cuh"
using namespace xrd;

struct P { int bg; const char* src; long long* out; };

template <int N>
__global__ void __launch_bounds__(128, 1) k_probe(P p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, bg_bar[4];
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) tc::mbar_init(&bg_bar[i], 1); stop = 0; tc::fence_barrier_init(); }
  tc::fence_async_smem();
  if (warp == 0) { tc::tmem_alloc(&slot, 512); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (warp == 0) {
    const uint32_t idesc = tc::umma_idesc(128, N, 0);
    const uint64_t adesc0 = tc::umma_desc_sw128(tc::smem_u32(smem));
    const uint64_t bdesc0 = tc::umma_desc_sw128(tc::smem_u32(smem + 80 * 1024));
    long long t_done = 0;
    const int rounds = 20;
    for (int r = 0; r < rounds; ++r) {
      long long t0 = clock64();
      if (tc::elect_one()) {
#pragma unroll
        for (int g = 0; g < 8; ++g)
#pragma unroll
          for (int t = 0; t < 9; ++t)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc::umma_f16((uint32_t)((g & 1) * N), adesc0 + (uint64_t)(((t / 3) * 130 + (t % 3)) * 8 + k * 2), bdesc0 + (uint64_t)(k * 2), idesc,
                           (g >= 2 || t > 0 || k > 0) ? 1u : 0u);
        tc::umma_commit(&bar);
      }
      __syncwarp();
      tc::mbar_wait(&bar, r & 1);
      t_done += clock64() - t0;
    }
    if (lane == 0) stop = 1;
    if (lane == 0 && blockIdx.x == 0) p.out[0] = t_done / rounds;
  } else if (warp == 1 && p.bg) {
    if (lane == 0) {
      long long t0 = clock64(), n = 0;
      uint32_t ph[4] = {0, 0, 0, 0};
      for (int i = 0; i < 4; ++i) {
        tc::mbar_expect_tx(&bg_bar[i], 16384);
        tc::bulk_load_1d(smem + 112 * 1024 + i * 16384, p.src + ((size_t)blockIdx.x * 64 + i) * 16384, 16384, &bg_bar[i]);
      }
      int i = 0;
      while (!stop) {
        tc::mbar_wait(&bg_bar[i], ph[i]); ph[i] ^= 1; ++n;
        tc::mbar_expect_tx(&bg_bar[i], 16384);
        tc::bulk_load_1d(smem + 112 * 1024 + i * 16384, p.src + ((size_t)blockIdx.x * 64 + ((n + 4) & 63)) * 16384, 16384, &bg_bar[i]);
        i = (i + 1) & 3;
      }
      for (int k = 0; k < 4; ++k) { tc::mbar_wait(&bg_bar[i], ph[i]); ph[i] ^= 1; i = (i + 1) & 3; }
      if (blockIdx.x == 0) { p.out[1] = n * 16384; p.out[2] = clock64() - t0; }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(0, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 32);
  char* src; cudaMalloc(&src, (size_t)148 * 64 * 16384); cudaMemset(src, 0, (size_t)148 * 64 * 16384);
  auto k = k_probe<96>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  const int grids[4] = {148, 148, 74, 37};
  for (int bg = 0; bg < 4; ++bg) {
    cudaMemset(d, 0, 32);
    P p; p.bg = bg > 0; p.src = src; p.out = d;
    k<<<grids[bg], 128, 200 * 1024>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("N=96, 288 MMAs, background TMA stream %s: %.1f cycles/MMA", bg ? "ON " : "off", (double)h[0] / 288);
    if (bg) printf("   background: %.1f B/clk into this SM's shared memory (%d SMs streaming, L2-resident source)", (double)h[1] / (double)h[2], grids[bg]);
    printf("\n");
  }
  return 0;
}
