// mma_probe.cu -- microbenchmark: issue rate / completion latency of tcgen05.mma (kind::f16, M=128) from shared
// memory operands, as a function of N, of the number of MMAs between commits and of the A start-address alignment.
// One CTA per SM, operands are whatever is in shared memory (zero-filled).  The issue loop is fully unrolled with
// compile-time descriptor offsets (a single thread's dependent integer chain costs ~4 cycles per instruction, so a
// runtime-indexed loop measures the loop, not the tensor core).  Build: tools/probes/build.sh
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../medical-image-denoising-using-diffusion_b200/csrc/tc_common.cuh"
using namespace xrd;

struct P { int rounds; long long* out; };

// GROUPS groups of 9 "taps" x KS k-steps; tap t reads A rows starting at OFF + (t/3)*130 + (t%3)  (halo style) or at 0
template <int N, int GROUPS, int KS, int HALO, int NACC, int MNB = 0>
__global__ void __launch_bounds__(128, 1) k_probe(P p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  tc::fence_async_smem();
  if (warp == 0) { tc::tmem_alloc(&slot, 512); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (warp == 0) {
    const uint32_t idesc = tc::umma_idesc(128, N, 0) | (MNB ? (1u << 16) : 0u);   // MNB: B operand MN-major (the P.V product of attention)
    const uint64_t adesc0 = tc::umma_desc_sw128(tc::smem_u32(smem));               // A region: 4 halo rows = 66 KB
    uint64_t bdesc0 = tc::umma_desc_sw128(tc::smem_u32(smem + 96 * 1024));   // B region: 9 taps would not fit; reuse 2
    if (MNB) {   // MN-major: 64-element N blocks 8 KB apart (LBO), 8-key groups 1024 B apart (SBO); a k-step (16 keys) is 2048 B
      bdesc0 = (uint64_t)((tc::smem_u32(smem + 96 * 1024) & 0x3FFFF) >> 4) | ((uint64_t)((8192 >> 4) & 0x3FFF) << 16) | ((uint64_t)(1024 >> 4) << 32) |
               ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    }
    long long t_issue = 0, t_done = 0;
    for (int r = 0; r < p.rounds; ++r) {
      long long t0 = clock64();
      if (tc::elect_one()) {
#pragma unroll
        for (int g = 0; g < GROUPS; ++g) {
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const int arow = HALO ? ((t / 3) + (g % NACC)) * 130 + (t % 3) : 0;
#pragma unroll
            for (int k = 0; k < KS; ++k)
              tc::umma_f16((uint32_t)((g % NACC) * N), adesc0 + (uint64_t)(arow * 8 + k * 2), bdesc0 + (uint64_t)(MNB ? (k * 128) : ((t & 1) * N * 8 + k * 2)), idesc,
                           (g >= NACC || t > 0 || k > 0) ? 1u : 0u);
          }
        }
        tc::umma_commit(&bar);
      }
      __syncwarp();
      long long t1 = clock64();
      tc::mbar_wait(&bar, r & 1);
      long long t2 = clock64();
      t_issue += t1 - t0; t_done += t2 - t0;
    }
    if (lane == 0 && blockIdx.x == 0) { p.out[0] = t_issue / p.rounds; p.out[1] = t_done / p.rounds; }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(0, 512); }
}

template <int N, int GROUPS, int KS, int HALO, int NACC, int MNB = 0>
void run(long long* d) {
  auto k = k_probe<N, GROUPS, KS, HALO, NACC, MNB>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  P p; p.rounds = 20; p.out = d;
  k<<<148, 128, 170 * 1024>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const int nmma = GROUPS * 9 * KS;
  printf("%4d%s %5d %3d %5d %5d | %10lld %10lld %9.1f %9.1f\n", N, MNB ? "T" : " ", nmma, KS, HALO, NACC, h[0], h[1], (double)h[1] / nmma, N / 2.0);
}

template <int N> void sweep(long long* d) {
  constexpr int NACC = (512 / N) >= 2 ? 2 : 1;
  run<N, 1, 1, 0, 1>(d);
  run<N, 1, 4, 0, 1>(d);
  run<N, 2, 4, 0, NACC>(d);
  run<N, 4, 4, 0, NACC>(d);
  run<N, 4, 4, 1, NACC>(d);
  run<N, 8, 4, 1, NACC>(d);
  run<N, 4, 3, 1, NACC>(d);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  printf("%4s %5s %3s %5s %5s | %10s %10s %9s %9s\n", "N", "nmma", "KS", "halo", "nacc", "issue_cyc", "done_cyc", "cyc/mma", "ideal");
  sweep<48>(d); sweep<96>(d); sweep<144>(d); sweep<192>(d); sweep<256>(d);
  run<64, 4, 4, 0, 2>(d); run<64, 4, 3, 0, 2>(d); run<128, 4, 4, 0, 2>(d);
  run<96, 4, 4, 0, 2, 1>(d); run<96, 8, 4, 0, 2, 1>(d); run<64, 4, 4, 0, 2, 1>(d); run<128, 4, 4, 0, 2, 1>(d);
  return 0;
}
