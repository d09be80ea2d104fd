// commit_probe.cu -- does a tcgen05.commit between short MMA groups stall the tensor pipe?
// G groups of 6 MMAs (M=128, N=64, K=16); variants: commit after every group / only at the end; fresh accumulator
// (accumulate=0) at every group start or not; alternate between two accumulators or not.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../medical-image-denoising-using-diffusion_b200/csrc/tc_common.cuh"
using namespace xrd;

struct P { int rounds; long long* out; };

template <int N, int G, int PER, int COMMIT_EACH, int FRESH, int ALT>
__global__ void __launch_bounds__(128, 1) k_probe(P p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[G + 1];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i <= G; ++i) tc::mbar_init(&bar[i], 1); tc::fence_barrier_init(); }
  tc::fence_async_smem();
  if (warp == 0) { tc::tmem_alloc(&slot, 512); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (warp == 0) {
    const uint32_t idesc = tc::umma_idesc(128, N, 0);
    const uint64_t adesc0 = tc::umma_desc_sw128(tc::smem_u32(smem));
    const uint64_t bdesc0 = tc::umma_desc_sw128(tc::smem_u32(smem + 96 * 1024));
    long long t_issue = 0, t_done = 0;
    for (int r = 0; r < p.rounds; ++r) {
      long long t0 = clock64();
      if (tc::elect_one()) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
#pragma unroll
          for (int k = 0; k < PER; ++k)
            tc::umma_f16((uint32_t)((ALT ? (g & 1) : 0) * N), adesc0 + (uint64_t)((k & 3) * 2 + (k >> 2) * 1024), bdesc0 + (uint64_t)((k & 3) * 2), idesc,
                         (FRESH && k == 0) ? 0u : 1u);
          if (COMMIT_EACH) tc::umma_commit(&bar[g]);
        }
        tc::umma_commit(&bar[G]);
      }
      __syncwarp();
      long long t1 = clock64();
      tc::mbar_wait(&bar[G], r & 1);
      long long t2 = clock64();
      t_issue += t1 - t0; t_done += t2 - t0;
    }
    if (lane == 0 && blockIdx.x == 0) { p.out[0] = t_issue / p.rounds; p.out[1] = t_done / p.rounds; }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(0, 512); }
}

template <int N, int G, int PER, int COMMIT_EACH, int FRESH, int ALT>
void run(long long* d) {
  auto k = k_probe<N, G, PER, COMMIT_EACH, FRESH, ALT>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  P p; p.rounds = 20; p.out = d;
  k<<<148, 128, 170 * 1024>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("N=%3d groups=%2d per=%d commit_each=%d fresh=%d alt=%d | issue %7lld done %7lld  cyc/mma %6.1f\n", N, G, PER, COMMIT_EACH, FRESH, ALT, h[0],
         h[1], (double)h[1] / (G * PER));
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  run<64, 16, 6, 0, 0, 0>(d);
  run<64, 16, 6, 1, 0, 0>(d);
  run<64, 16, 6, 0, 1, 0>(d);
  run<64, 16, 6, 1, 1, 0>(d);
  run<64, 16, 6, 1, 1, 1>(d);
  run<64, 16, 6, 0, 1, 1>(d);
  run<96, 16, 4, 1, 1, 1>(d);
  run<96, 16, 4, 0, 0, 0>(d);
  run<128, 16, 6, 1, 1, 1>(d);
  return 0;
}
