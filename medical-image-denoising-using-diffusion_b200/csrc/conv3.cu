// conv3.cu -- persistent, halo-reusing tcgen05 implicit-GEMM 3x3 convolution (stride 1, pad 1) for feature maps
// whose width is a multiple of 128.  This is the dominant kernel of the UNet's three high-resolution levels.
//
// What bounds a 3x3 conv with few output channels on B200 (measured, tools/probes/mma_probe.cu + profiles/):
//   * tcgen05.mma M=128, K=16 with both operands in shared memory costs max(N/2, 32 + N/4) cycles: below N=128 the
//     instruction is limited by the 128 B/clk shared-memory read of A (4 KB) + B (32*N B), not by the tensor pipe;
//   * one issuing thread retires ~1 dependent integer instruction per 4 cycles, so every descriptor must be a
//     compile-time offset from a per-stage base (a runtime-indexed issue loop costs >100 cycles per MMA);
//   * L2 -> SM delivers ~6.4 TB/s chip-wide, i.e. no more than HBM: operand re-fetches are as expensive as DRAM.
// Hence:
//   * one CTA per SM walks tiles of TH image rows x 128 columns (persistent, static round robin);
//   * per 64-channel chunk ONE TMA box {64 ch, 130, TH+2} brings the tile plus its halo into shared memory (zero
//     padding, image borders and channel tails by TMA out-of-bounds fill); the nine taps are nine descriptor
//     offsets into that buffer: tap (dy,dx) of tile row s starts at buffer pixel (s+dy)*130 + dx.  The start
//     address is then not 1024-byte aligned; the 128B swizzle is a function of the absolute shared-memory address
//     (as TMA wrote it), so the descriptor base-offset field stays 0 (verified bit-exact on B200);
//   * a virtual channel concat (x1 | x2) is walked as the chunks of x1 followed by the chunks of x2;
//   * weights stay resident in shared memory when all (chunk, tap) blocks fit, else they stream through a ring
//     once per tile (not once per 128 pixels);
//   * TH accumulators (one per image row) live in TMEM, double buffered when they fit, so the epilogue of tile i
//     overlaps the MMAs of tile i+1;
//   * the MMA issue loop is fully unrolled over (tap, row, k-step);
//   * epilogue: TMEM -> registers, + bias + time-embedding row + residual (prefetched), GroupNorm partial sums
//     of the OUTPUT (so the next GroupNorm needs no statistics pass), 16-bit pack, 16-byte stores of whole
//     sectors straight from registers; eight epilogue warps (a single warp retires ~1 instruction per 3-4
//     cycles, four warps cannot drain 128x48 accumulators as fast as the tensor core refills them).
#include "kernels.cuh"
#include "tc_common.cuh"

#include <stdlib.h>

namespace xrd {

struct Conv3P {
  int H, W, nimg;
  int tiles_w, tiles_h, ntiles;
  int c0, c1;               // channels of the two sources (c1 = 0: single source)
  int nchunk0, nchunk;      // 64-channel chunks of source 0 / of both
  int nb;                   // weight ring slots when streaming
  int nres;                 // streaming variant: the first nres K blocks are nevertheless resident (after the ring in shared memory)
  int dbg;
  int rowload;              // 1: one TMA instruction per halo row instead of one per (TH+2)-row box
  const float* bias;
  const float* chan_add; int chan_add_bstride;
  const void* resid;
  void* y;
  double* stats;            // optional [nimg][8][2] (sum, sum of squares) of the stored output, 8 channel groups
  const void* wsw;          // pre-swizzled weight blocks for 1-D bulk loads (null: tensor-map loads)
  const float2* in_coef;    // GN variant: [nimg][c0+c1] (0.5*scale, 0.5*shift) of the GroupNorm + SiLU applied to the input in shared memory
  long long* prof;          // optional clock64 trace of block 0
};

constexpr int kC3Threads = 320;    // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int kC3ThreadsGN = 448;  // + warps 10..13: GroupNorm + SiLU applied in place to every landed activation stage
constexpr int kC3Pitch = 136;   // smem pixels per halo row: 130 used (128 + one halo column each side); 136*128 B keeps rows 1024-aligned
constexpr int kC3Box = 130;

// clock64 trace of block 0 and the bottleneck-experiment switches (XRD_C3_DBG): compiled in only with -DXRD_TRACE -- their
// predicated stores and tests sit in the MMA issue stream, where every instruction is tensor-pipe idle time (measured: 4-7 %)
#ifdef XRD_TRACE
#define XRD_KDBG(m) (p.dbg & (m))
#define C3PROF(tile, slot) do { if (p.prof && blockIdx.x == 0 && (tile) < 64) p.prof[(tile) * 8 + (slot)] = clock64(); } while (0)
#else
#define XRD_KDBG(m) 0
#define C3PROF(tile, slot) do { } while (0)
#endif

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <typename T> __device__ __forceinline__ void unpack8(const uint4& t, float (&v)[8]);
template <> __device__ __forceinline__ void unpack8<__half>(const uint4& t, float (&v)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <> __device__ __forceinline__ void unpack8<__nv_bfloat16>(const uint4& t, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

// All MMAs of one 64-channel chunk: 9 taps x TH rows x KS k-steps, every descriptor a compile-time offset.
template <int COUT, int TH, int KS, bool WRES>
__device__ __forceinline__ void c3_issue_chunk(uint64_t adesc0, uint64_t bdesc0, uint32_t sB_addr, uint64_t* b_full, uint64_t* b_empty,
                                               uint32_t& wslot, uint32_t& wphase, uint32_t& wprobed, int nb, int kb0, int nres, uint32_t acc0,
                                               uint32_t idesc, uint32_t not_first) {
  constexpr uint32_t B_BYTES = COUT * 128;
  // streamed weights: the barrier of tap t+1 is probed BEFORE the MMAs of tap t are issued, so the probe's round trip runs under
  // them (a wait issued right in front of its MMAs stalls the issue stream, and with it the tensor pipe, for ~57 cycles per tap);
  // the probe of the next chunk's first tap is carried out of this call in `wprobed`
  bool probed = wprobed != 0u;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    uint64_t bdesc;
    bool streamed = false;
    if (WRES) {
      bdesc = bdesc0 + (uint64_t)(tap * (B_BYTES >> 4));
    } else if (kb0 + tap < nres) {        // partially resident: block kb lives behind the ring
      bdesc = tc::umma_desc_sw128(sB_addr + (uint32_t)(nb + kb0 + tap) * B_BYTES);
    } else {
      streamed = true;
      tc::mbar_wait_probed(probed, &b_full[wslot], wphase);
      bdesc = tc::umma_desc_sw128(sB_addr + wslot * B_BYTES);
    }
    uint32_t nslot = wslot + 1, nphase = wphase;
    if (nslot == (uint32_t)nb) { nslot = 0; nphase ^= 1; }
    if (streamed) probed = tc::mbar_test(&b_full[nslot], nphase);
    const int dy = tap / 3, dx = tap % 3;
#pragma unroll
    for (int s = 0; s < TH; ++s) {
#pragma unroll
      for (int k = 0; k < KS; ++k)
        tc::umma_f16(acc0 + (uint32_t)(s * COUT), adesc0 + (uint64_t)(((s + dy) * kC3Pitch + dx) * 8 + k * 2), bdesc + (uint64_t)(k * 2), idesc,
                     (tap == 0 && k == 0) ? not_first : 1u);
    }
    if (streamed) {
      tc::umma_commit(&b_empty[wslot]);
      wslot = nslot; wphase = nphase;
    }
  }
  wprobed = probed ? 1u : 0u;
}

template <typename T, int COUT, int TH, int NACC, bool WRES, bool GN>
__global__ void __launch_bounds__(GN ? kC3ThreadsGN : kC3Threads, 1)
k_conv3(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB,
        const Conv3P p) {
  constexpr uint32_t A_BYTES = (uint32_t)(((TH + 2) * kC3Pitch * 128 + 1023) & ~1023);
  constexpr uint32_t B_BYTES = COUT * 128;
  constexpr int CPG = COUT / 8;                            // channels per GroupNorm group (8 groups)
  constexpr int NBLK = COUT / 48;                          // 48-column epilogue blocks
  // the whole TMEM: a 512-column allocation always starts at column 0, which the compile-time accumulator addresses below rely on,
  // also when a CTA of another kernel shares the SM (side branches of xrd_hybrid); see conv1.cu
  constexpr uint32_t TMEM_COLS = 512;
  static_assert(COUT % 48 == 0 && NACC * TH * COUT <= 512, "accumulators must fit TMEM");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int nkb = p.nchunk * 9;
  const int nbslots = WRES ? nkb : p.nb + p.nres;
  uint8_t* sA = smem;                                       // [2][A_BYTES]
  uint8_t* sB = sA + 2 * (size_t)A_BYTES;                   // [nbslots][B_BYTES]
  float* s_badd = (float*)(sB + (size_t)nbslots * B_BYTES);   // [8 warps][COUT] bias + time-embedding row of the current image
  uint64_t* bars = (uint64_t*)(s_badd + 8 * COUT);
  uint64_t* a_full = bars;            // [2]
  uint64_t* a_empty = bars + 2;       // [2]
  uint64_t* acc_full = bars + 4;      // [2]
  uint64_t* acc_empty = bars + 6;     // [2]
  uint64_t* w_full = bars + 8;        // [1] resident weights
  uint64_t* b_full = bars + 9;        // [16]
  uint64_t* b_empty = b_full + 16;    // [16]
  uint64_t* a_ready = b_empty + 16;   // [2] GN variant: stage transformed in place, ready for the tensor core
  uint32_t* tmem_slot = (uint32_t*)(a_ready + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA0);
    tc::tma_prefetch_desc(&tmA1);
    tc::tma_prefetch_desc(&tmB);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&a_full[s], 1); tc::mbar_init(&a_empty[s], 1);
      tc::mbar_init(&acc_full[s], 1); tc::mbar_init(&acc_empty[s], 256);
      tc::mbar_init(&a_ready[s], 128);
    }
    tc::mbar_init(w_full, 1);
    for (int s = 0; s < 16; ++s) { tc::mbar_init(&b_full[s], 1); tc::mbar_init(&b_empty[s], 1); }
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, TMEM_COLS);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  // One CTA per SM (shared-memory limited) and this is its only allocation: the allocator returns column 0 / lane 0.
  // Treating it as the constant 0 keeps every tcgen05.mma operand in uniform registers.
  if (*tmem_slot != 0u) {
    if (threadIdx.x == 0) printf("libxrd: conv3 expects TMEM base 0, got %u\n", *tmem_slot);
    __trap();
  }
  constexpr uint32_t tmem_base = 0u;

  if (warp == 0) {
    // ===================== TMA producer (one elected lane issues; the warp stays converged) =====================
    if (tc::elect_one()) {
      if (WRES) {
        tc::mbar_expect_tx(w_full, (uint32_t)nkb * B_BYTES);
        for (int kb = 0; kb < nkb; ++kb) tc::tma_load_3d(sB + (size_t)kb * B_BYTES, &tmB, w_full, 0, 0, kb);
      } else if (p.nres > 0) {
        tc::mbar_expect_tx(w_full, (uint32_t)p.nres * B_BYTES);
        for (int kb = 0; kb < p.nres; ++kb) tc::tma_load_3d(sB + (size_t)(p.nb + kb) * B_BYTES, &tmB, w_full, 0, 0, kb);
      }
    }
    __syncwarp();
    uint32_t ai = 0, wslot = 0, wphase = 0;
    for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
      int r = t;
      const int txi = r % p.tiles_w; r /= p.tiles_w;
      const int tyi = r % p.tiles_h;
      const int img = r / p.tiles_h;
      const int w0 = txi * 128 - 1, h0 = tyi * TH - 1;
      for (int c = 0; c < p.nchunk; ++c, ++ai) {
        const int st = ai & 1;
        tc::mbar_wait(&a_empty[st], ((ai >> 1) & 1) ^ 1);
        if (lane == 0) C3PROF(ai, 0);
        uint32_t leader;
        if (tc::elect_one(leader)) {
          if (XRD_KDBG(4) && ai >= 2) {
            tc::mbar_arrive(&a_full[st]);
          } else {
            tc::mbar_expect_tx(&a_full[st], (uint32_t)(TH + 2) * kC3Box * 128u);
            const bool second = c >= p.nchunk0;
            const CUtensorMap* tm = second ? &tmA1 : &tmA0;
            const int cc0 = (second ? c - p.nchunk0 : c) * 64;
#pragma unroll
            for (int rr = 0; rr < TH + 2; ++rr)
              tc::tma_load_4d(sA + (size_t)st * A_BYTES + (size_t)rr * kC3Pitch * 128, tm, &a_full[st], cc0, w0, h0 + rr, img);
          }
          if (!WRES) {
            for (int tap = 0; tap < 9; ++tap) {
              if (c * 9 + tap < p.nres) continue;          // resident block
              tc::mbar_wait(&b_empty[wslot], wphase ^ 1);
              tc::mbar_expect_tx(&b_full[wslot], B_BYTES);
              if (p.wsw) tc::bulk_load_1d(sB + (size_t)wslot * B_BYTES, (const char*)p.wsw + (size_t)(c * 9 + tap) * B_BYTES, B_BYTES, &b_full[wslot]);
              else tc::tma_load_3d(sB + (size_t)wslot * B_BYTES, &tmB, &b_full[wslot], 0, 0, c * 9 + tap);
              if (++wslot == (uint32_t)p.nb) { wslot = 0; wphase ^= 1; }
            }
          }
        }
        __syncwarp();
        if (!WRES) {   // every lane tracks the ring position the leader advanced
          wslot = __shfl_sync(0xffffffffu, wslot, leader);
          wphase = __shfl_sync(0xffffffffu, wphase, leader);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (WRES || p.nres > 0) tc::mbar_wait(w_full, 0);
    const uint32_t idesc = tc::umma_idesc(128, COUT, tc::umma_fmt<T>());
    const uint32_t sB_addr = tc::smem_u32(sB);
    uint32_t ai = 0, ti = 0, wslot = 0, wphase = 0;
    uint64_t* const abar = GN ? a_ready : a_full;
    bool a_probed = false;                                   // result of the early probe of the next stage's barrier
    bool acc_probed = false;                                 // ... of the next tile's accumulator-drained barrier
    uint32_t wprobed = 0u;                                   // ... of the next streamed weight block (leader's value, broadcast)
    for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++ti) {
      const int ab = NACC == 2 ? (int)(ti & 1) : 0;
      const uint32_t use = NACC == 2 ? (ti >> 1) : ti;       // how many times this accumulator buffer was used before
      if (lane == 0) C3PROF(ti, 1);
      tc::mbar_wait_probed(acc_probed, &acc_empty[ab], (use & 1) ^ 1);
      tc::tc_fence_after();
      if (lane == 0) C3PROF(ti, 2);
      const uint32_t acc0 = tmem_base + (uint32_t)(ab * TH * COUT);
      for (int c = 0; c < p.nchunk; ++c, ++ai) {
        const int st = ai & 1;
        tc::mbar_wait_probed(a_probed, &abar[st], (ai >> 1) & 1);
        tc::tc_fence_after();
        // probe the NEXT stage now: the round trip of the barrier test runs under this chunk's MMAs (tc_common.cuh)
        a_probed = tc::mbar_test(&abar[(ai + 1) & 1], ((ai + 1) >> 1) & 1);
        if (c == p.nchunk - 1) {                               // last chunk of the tile: probe the next tile's accumulator barrier as well
          const uint32_t tn = ti + 1;
          acc_probed = tc::mbar_test(&acc_empty[NACC == 2 ? (tn & 1) : 0], ((NACC == 2 ? (tn >> 1) : tn) & 1) ^ 1);
        }
        if (c == 0 && lane == 0) C3PROF(ti, 3);
        const bool second = c >= p.nchunk0;
        const int cl = second ? c - p.nchunk0 : c;
        const int ks = min(64, (second ? p.c1 : p.c0) - cl * 64) >> 4;
        uint32_t leader;
        if (tc::elect_one(leader)) {
          const uint64_t adesc0 = tc::umma_desc_sw128(tc::smem_u32(sA + (size_t)st * A_BYTES));
          const uint64_t bdesc0 = tc::umma_desc_sw128(sB_addr + (uint32_t)(c * 9) * B_BYTES);
          const uint32_t nf = c ? 1u : 0u;
          if (!XRD_KDBG(2)) {
            switch (ks) {
              case 4: c3_issue_chunk<COUT, TH, 4, WRES>(adesc0, bdesc0, sB_addr, b_full, b_empty, wslot, wphase, wprobed, p.nb, c * 9, p.nres, acc0, idesc, nf); break;
              case 3: c3_issue_chunk<COUT, TH, 3, WRES>(adesc0, bdesc0, sB_addr, b_full, b_empty, wslot, wphase, wprobed, p.nb, c * 9, p.nres, acc0, idesc, nf); break;
              case 2: c3_issue_chunk<COUT, TH, 2, WRES>(adesc0, bdesc0, sB_addr, b_full, b_empty, wslot, wphase, wprobed, p.nb, c * 9, p.nres, acc0, idesc, nf); break;
              default: c3_issue_chunk<COUT, TH, 1, WRES>(adesc0, bdesc0, sB_addr, b_full, b_empty, wslot, wphase, wprobed, p.nb, c * 9, p.nres, acc0, idesc, nf); break;
            }
          } else if (!WRES) {   // experiment: consume the weight ring without issuing MMAs
            wprobed = 0u;
            for (int tap = 0; tap < 9; ++tap) {
              if (c * 9 + tap < p.nres) continue;
              tc::mbar_wait(&b_full[wslot], wphase);
              tc::mbar_arrive(&b_empty[wslot]);
              if (++wslot == (uint32_t)p.nb) { wslot = 0; wphase ^= 1; }
            }
          }
          tc::umma_commit(&a_empty[st]);
          if (c == p.nchunk - 1) tc::umma_commit(&acc_full[ab]);
        }
        __syncwarp();
        if (!WRES) {
          wslot = __shfl_sync(0xffffffffu, wslot, leader);
          wphase = __shfl_sync(0xffffffffu, wphase, leader);
          wprobed = __shfl_sync(0xffffffffu, wprobed, leader);
        }
      }
      if (lane == 0) C3PROF(ti, 4);
    }
  } else if (GN && warp >= 10) {
    // ===================== input transform (warps 10..13): a = SiLU(GroupNorm(x)) in place, HYB:264-265 / 269-270 =====================
    // Thread = (16-byte channel chunk j of the stage's valid channels, pixel lane).  The stage is K-major with the 128B
    // swizzle TMA applied: logical chunk j of buffer pixel sp sits at physical chunk j ^ (sp & 7).  Pixels outside the
    // image stay zero (the convolution pads the ACTIVATED tensor); the stage's channel tail is never read by the MMAs.
    const int tt = threadIdx.x - 320;
    uint32_t ai = 0;
    for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
      int r = t;
      const int txi = r % p.tiles_w; r /= p.tiles_w;
      const int tyi = r % p.tiles_h;
      const int img = r / p.tiles_h;
      const int w0 = txi * 128 - 1, h0 = tyi * TH - 1;
      for (int c = 0; c < p.nchunk; ++c, ++ai) {
        const int st = ai & 1;
        const bool second = c >= p.nchunk0;
        const int cl = second ? c - p.nchunk0 : c;
        const int nvc = min(64, (second ? p.c1 : p.c0) - cl * 64) >> 3;        // valid 8-channel chunks: 2, 4, 6 or 8
        const int npl = 128 / nvc;
        const int j = tt % nvc, plane = tt / nvc;
        float sc[8], sh[8];
        if (plane < npl) {
          const float2* cf = p.in_coef + (size_t)img * (p.c0 + p.c1) + (second ? p.c0 : 0) + cl * 64 + j * 8;
#pragma unroll
          for (int i = 0; i < 8; ++i) { const float2 v = __ldg(cf + i); sc[i] = v.x; sh[i] = v.y; }
        }
        tc::mbar_wait(&a_full[st], (ai >> 1) & 1);
        if (plane < npl) {
          const uint32_t sbase = tc::smem_u32(sA + (size_t)st * A_BYTES);
          // four pixels per iteration: the chain LDS -> FMA -> MUFU.TANH -> FMA -> STS is ~280 cycles long, a single pixel
          // in flight per thread made the transform (not the tensor core) the critical path of the tile
          constexpr int U = 4;
          for (int pix0 = plane; pix0 < (TH + 2) * kC3Box; pix0 += U * npl) {
            uint32_t addr[U];
            bool ok[U];
            uint4 q[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int pix = pix0 + u * npl;
              const int rr = pix / kC3Box, cc = pix - rr * kC3Box;
              const int ih = h0 + rr, iw = w0 + cc;
              ok[u] = pix < (TH + 2) * kC3Box && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
              const int sp = rr * kC3Pitch + cc;
              addr[u] = sbase + (uint32_t)sp * 128u + (uint32_t)((j ^ (sp & 7)) << 4);
              if (ok[u]) asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q[u].x), "=r"(q[u].y), "=r"(q[u].z), "=r"(q[u].w) : "r"(addr[u]));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
              if (!ok[u]) continue;
              float v[8];
              unpack8<T>(q[u], v);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float h = fmaf(v[i], sc[i], sh[i]);          // 0.5 * GroupNorm(x)
                float th;
                asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
                v[i] = fmaf(h, th, h);                              // x*sigmoid(x) = h*tanh(h) + h with h = x/2
              }
              q[u].x = tc::pack2<T>(v[0], v[1]); q[u].y = tc::pack2<T>(v[2], v[3]); q[u].z = tc::pack2<T>(v[4], v[5]); q[u].w = tc::pack2<T>(v[6], v[7]);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr[u]), "r"(q[u].x), "r"(q[u].y), "r"(q[u].z), "r"(q[u].w) : "memory");
            }
          }
        }
        tc::fence_async_smem();                        // generic-proxy writes -> visible to the tensor core's async-proxy reads
        tc::mbar_arrive(&a_ready[st]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..9, two groups of four) =====================
    // Group g = (warp-2)/4 drains the tile rows s with s % 2 == g; inside a group warp w owns TMEM lanes
    // 32*(w%4).., i.e. 32 consecutive output pixels of that image row.  Every thread holds ONE pixel: its COUT
    // channels are COUT*2 contiguous bytes of the NHWC output (whole 32-byte sectors), stored straight from
    // registers with 16-byte stores -- no staging buffer, no CTA-wide barrier.
    const int quad = warp & 3;
    const int grp = (warp - 2) >> 2;
    T* yp = (T*)p.y;
    const T* rp = (const T*)p.resid;
    float* badd = s_badd + (warp - 2) * COUT;                         // per-warp copy of bias + time-embedding row
    float gs[8], gq[8];                                               // GroupNorm partial sums of this thread's pixels
#pragma unroll
    for (int g = 0; g < 8; ++g) { gs[g] = 0.f; gq[g] = 0.f; }
    auto flush_stats = [&](int img) {
      if (!p.stats || img < 0) return;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          gs[g] += __shfl_xor_sync(0xffffffffu, gs[g], o);
          gq[g] += __shfl_xor_sync(0xffffffffu, gq[g], o);
        }
      }
      if (lane < 16) {
        float v = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) { if (lane == 2 * g) v = gs[g]; if (lane == 2 * g + 1) v = gq[g]; }
        atomicAdd(p.stats + (size_t)img * 16 + lane, (double)v);
      }
#pragma unroll
      for (int g = 0; g < 8; ++g) { gs[g] = 0.f; gq[g] = 0.f; }
    };
    uint32_t ti = 0;
    int cur_img = -1;
    for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++ti) {
      int r = t;
      const int txi = r % p.tiles_w; r /= p.tiles_w;
      const int tyi = r % p.tiles_h;
      const int img = r / p.tiles_h;
      const int ab = NACC == 2 ? (int)(ti & 1) : 0;
      const uint32_t use = NACC == 2 ? (ti >> 1) : ti;
      if (img != cur_img) {
        flush_stats(cur_img);
        __syncwarp();
        for (int cc = lane; cc < COUT; cc += 32)
          badd[cc] = (p.bias ? __ldg(p.bias + cc) : 0.f) + (p.chan_add ? __ldg(p.chan_add + (int64_t)img * p.chan_add_bstride + cc) : 0.f);
        cur_img = img;
        __syncwarp();
      }
      const int ow = txi * 128 + quad * 32 + lane;
      // residual of this group's first (row, block): issued before the accumulator wait so its latency hides behind the MMAs
      uint4 rcur[6], rnext[6];
      const int64_t pix0 = ((int64_t)img * p.H + (int64_t)tyi * TH + grp) * p.W + ow;
      if (rp && tyi * TH + grp < p.H) {
#pragma unroll
        for (int j = 0; j < 6; ++j) rcur[j] = __ldg(reinterpret_cast<const uint4*>(rp + pix0 * COUT) + j);
      }
      if (warp == 2 && lane == 0) C3PROF(ti, 5);
      tc::mbar_wait(&acc_full[ab], use & 1);
      tc::tc_fence_after();
      if (warp == 2 && lane == 0) C3PROF(ti, 6);
#pragma unroll
      for (int s0 = 0; s0 < TH; s0 += 2) {
        const int s = s0 + grp;
        if (s < TH) {
          const int oh = tyi * TH + s;
          const bool row_ok = oh < p.H;
          const int64_t opix = pix0 + (int64_t)s0 * p.W;
          const uint32_t tacc = tmem_base + (uint32_t)((ab * TH + s) * COUT) + ((uint32_t)(quad * 32) << 16);
#pragma unroll
          for (int cb = 0; cb < NBLK; ++cb) {
            uint32_t v[48];
            uint4 pk_even;
            if (!XRD_KDBG(8)) {
              tmem_ld16_nowait(tacc + (uint32_t)(cb * 48), *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
              tmem_ld16_nowait(tacc + (uint32_t)(cb * 48 + 16), *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
              tmem_ld16_nowait(tacc + (uint32_t)(cb * 48 + 32), *reinterpret_cast<uint32_t(*)[16]>(&v[32]));
            }
            // prefetch the residual of the next (row, block) this thread will process in this tile
            const bool has_next = rp && (cb + 1 < NBLK || (s + 2 < TH && oh + 2 < p.H));
            if (has_next) {
              const T* nsrc = (cb + 1 < NBLK) ? rp + opix * COUT + (cb + 1) * 48 : rp + (opix + 2 * (int64_t)p.W) * COUT;
#pragma unroll
              for (int j = 0; j < 6; ++j) rnext[j] = __ldg(reinterpret_cast<const uint4*>(nsrc) + j);
            }
            if (!XRD_KDBG(8)) tmem_wait_ld();
#pragma unroll
            for (int h8 = 0; h8 < 6; ++h8) {
              const int co = cb * 48 + h8 * 8;
              const float4 b0 = *reinterpret_cast<const float4*>(badd + co), b1 = *reinterpret_cast<const float4*>(badd + co + 4);
              float r8[8];
              r8[0] = __uint_as_float(v[h8 * 8 + 0]) + b0.x; r8[1] = __uint_as_float(v[h8 * 8 + 1]) + b0.y;
              r8[2] = __uint_as_float(v[h8 * 8 + 2]) + b0.z; r8[3] = __uint_as_float(v[h8 * 8 + 3]) + b0.w;
              r8[4] = __uint_as_float(v[h8 * 8 + 4]) + b1.x; r8[5] = __uint_as_float(v[h8 * 8 + 5]) + b1.y;
              r8[6] = __uint_as_float(v[h8 * 8 + 6]) + b1.z; r8[7] = __uint_as_float(v[h8 * 8 + 7]) + b1.w;
              if (rp && row_ok) {
                float q8[8];
                unpack8<T>(rcur[h8], q8);
#pragma unroll
                for (int j = 0; j < 8; ++j) r8[j] += q8[j];
              }
              if (p.stats && row_ok) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const int g = (co + j) / CPG;
                  gs[g] += r8[j];
                  gq[g] = fmaf(r8[j], r8[j], gq[g]);
                }
              }
              uint4 pk;
              pk.x = tc::pack2<T>(r8[0], r8[1]); pk.y = tc::pack2<T>(r8[2], r8[3]);
              pk.z = tc::pack2<T>(r8[4], r8[5]); pk.w = tc::pack2<T>(r8[6], r8[7]);
              // two 16-byte halves -> one 32-byte store of a whole sector
              if (h8 & 1) { if (row_ok && !XRD_KDBG(1)) tc::st_global_v8(yp + opix * COUT + co - 8, pk_even, pk); } else pk_even = pk;
            }
            if (has_next) {
#pragma unroll
              for (int j = 0; j < 6; ++j) rcur[j] = rnext[j];
            }
          }
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&acc_empty[ab]);
      if (warp == 2 && lane == 0) C3PROF(ti, 7);
    }
    flush_stats(cur_img);
  }
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

static int c3_env(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

bool conv3_supported(const Tens& x1, const Tens* x2, const ConvW& w, const ConvEpi& e) {
  static const int enabled = c3_env("XRD_CONV3", 1);
  if (!enabled) return false;
  if (x1.dt == DT_F32) return false;
  if (!(w.kh == 3 && w.kw == 3 && w.stride == 1 && w.pad == 1) || w.d2s) return false;
  if (x1.c % 16 != 0 || (x2 && x2->c % 16 != 0)) return false;
  if (!(w.cout == 48 || w.cout == 96 || w.cout == 144)) return false;
  if (x1.w % 128 != 0) return false;
  if (e.in_scale || e.out_scale || e.act != ACT_NONE) return false;
  if (e.in_coef && e.in_act != ACT_SILU) return false;
  return true;
}

namespace {
struct C3Cfg { int th, nacc; bool wres; int nb, nres; size_t smem; };

// shared-memory plan: two A stages + weights (resident or ring) + per-warp bias rows + barriers
C3Cfg c3_plan(int cout, int nkb) {
  const size_t budget = 227 * 1024 - 1024 /*alignment slack*/;
  const size_t bb = (size_t)cout * 128;
  auto fixed = [&](int th) {
    const size_t a = (((size_t)(th + 2) * kC3Pitch * 128 + 1023) & ~(size_t)1023);
    return 2 * a + 8 * cout * 4 + (11 + 32) * 8 + 64;
  };
  C3Cfg c{};
  c.th = 2;
  c.nacc = cout <= 96 ? 2 : 1;
  size_t f = fixed(c.th);
  c.wres = f + (size_t)nkb * bb <= budget;
  if (c3_env("XRD_C3_WRES", 1) == 0) c.wres = false;
  c.nb = 0; c.nres = 0;
  if (!c.wres) {
    // Streamed weights: measured (tools/probes/smem_contention_probe.cu) every SM ingests at most ~30 B/clk into shared memory,
    // chip load or not, and the TMA fill does not slow the MMAs down; a streamed-weight tile is therefore bound by
    // (weight bytes + halo bytes) / 30 B/clk.  Taller tiles amortise the weight pass over more rows:
    //   cout 48: 4 rows, still double-buffered accumulators (2*4*48 = 384 columns): 96->48 @512^2 x16 measured 607 vs 750 us
    //   cout 96: 3 rows would need a single accumulator set (288 columns): measured slower (254 vs 239 us), kept at 2 rows
    const int tall = c3_env("XRD_C3_TALL", 1);
    if (tall && cout == 48) { c.th = 4; c.nacc = 2; }
    else if (tall >= 2 && cout == 96) { c.th = 3; c.nacc = 1; }
    f = fixed(c.th);
    const int fit = (int)((budget - f) / bb);          // weight blocks that fit beside the two activation stages
    XRD_REQUIRE(fit >= 2, "conv3: shared memory budget exceeded (cout=%d)", cout);
    // A deep ring, nothing resident.  Keeping the first blocks resident behind a short ring (XRD_C3_PARTIAL=1) measured no
    // faster: the shorter ring costs what the saved bytes gain.
    if (c3_env("XRD_C3_PARTIAL", 0)) {
      c.nb = std::min(fit, c3_env("XRD_C3_RING", 4));
      c.nres = std::max(0, std::min(nkb - 1, fit - c.nb));
    } else {
      c.nb = std::min(fit, 12);
    }
  }
  c.smem = 1024 + f + (size_t)(c.wres ? nkb : c.nb + c.nres) * bb;
  return c;
}

template <typename T, int COUT, int TH, int NACC, bool WRES>
void c3_launch(Ctx& c, int grid, size_t smem, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const Conv3P& p) {
  if (p.in_coef) ensure_dyn_smem(k_conv3<T, COUT, TH, NACC, WRES, true>, 227 * 1024);
  else ensure_dyn_smem(k_conv3<T, COUT, TH, NACC, WRES, false>, 227 * 1024);
  if (p.in_coef) XRD_LAUNCH(c, (k_conv3<T, COUT, TH, NACC, WRES, true>), grid, kC3ThreadsGN, smem, a0, a1, b, p);
  else XRD_LAUNCH(c, (k_conv3<T, COUT, TH, NACC, WRES, false>), grid, kC3Threads, smem, a0, a1, b, p);
}

template <typename T>
void c3_dispatch(Ctx& c, int cout, const C3Cfg& g, int grid, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                 const Conv3P& p) {
  if (cout == 48) {
    if (g.wres) c3_launch<T, 48, 2, 2, true>(c, grid, g.smem, a0, a1, b, p);
    else if (g.th == 4) c3_launch<T, 48, 4, 2, false>(c, grid, g.smem, a0, a1, b, p);
    else c3_launch<T, 48, 2, 2, false>(c, grid, g.smem, a0, a1, b, p);
  } else if (cout == 96) {
    if (g.wres) c3_launch<T, 96, 2, 2, true>(c, grid, g.smem, a0, a1, b, p);
    else if (g.th == 3) c3_launch<T, 96, 3, 1, false>(c, grid, g.smem, a0, a1, b, p);
    else c3_launch<T, 96, 2, 2, false>(c, grid, g.smem, a0, a1, b, p);
  } else {
    if (g.wres) c3_launch<T, 144, 2, 1, true>(c, grid, g.smem, a0, a1, b, p);
    else c3_launch<T, 144, 2, 1, false>(c, grid, g.smem, a0, a1, b, p);
  }
}

void c3_encode_act(CUtensorMap* m, const Tens& x, int th) {
  const cuuint64_t dims[4] = {(cuuint64_t)x.c, (cuuint64_t)x.w, (cuuint64_t)x.h, (cuuint64_t)x.n};
  const cuuint64_t strides[3] = {(cuuint64_t)x.c * 2, (cuuint64_t)x.w * x.c * 2, (cuuint64_t)x.h * x.w * x.c * 2};
  const cuuint32_t box[4] = {64, (cuuint32_t)kC3Box, 1, 1};   // one halo row per TMA instruction
  (void)th;
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = get_encode_tiled()(m, tmap_dtype(x.dt), 4, x.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail(XRD_ERR_CUDA, "cuTensorMapEncodeTiled(conv3 activations) failed: %d", (int)r);
}
}  // namespace

void conv3(Ctx& c, const Tens& x1, const Tens* x2, ConvW& w, const ConvEpi& e, Tens& y) {
  XRD_REQUIRE(conv3_supported(x1, x2, w, e), "conv3: unsupported configuration");
  const int c0 = x1.c, c1 = x2 ? x2->c : 0;
  XRD_REQUIRE(c0 + c1 == w.cin && y.n == x1.n && y.h == x1.h && y.w == x1.w && y.c == w.cout && y.dt == x1.dt, "conv3: shape mismatch");
  if (x2) XRD_REQUIRE(x2->n == x1.n && x2->h == x1.h && x2->w == x1.w && x2->dt == x1.dt, "conv3: source mismatch");
  if (e.resid.p) XRD_REQUIRE(e.resid.dt == y.dt && e.resid.numel() == y.numel(), "conv3: residual mismatch");
  if (c.dry) return;
  if (!w.wtc[x1.dt] || w.tc_c1 != c0) conv_tc_pack(c.s, w, x1.dt, c0);
  Conv3P p;
  p.H = x1.h; p.W = x1.w; p.nimg = x1.n;
  p.c0 = c0; p.c1 = c1;
  p.nchunk0 = (c0 + 63) / 64;
  p.nchunk = p.nchunk0 + (c1 + 63) / 64;
  const int nkb = p.nchunk * 9;
  XRD_REQUIRE(nkb == w.tc_nkb && w.tc_npad == w.cout, "conv3: packed weights out of date");
  const C3Cfg g = c3_plan(w.cout, nkb);
  p.nb = g.nb; p.nres = g.nres;
  p.tiles_w = x1.w / 128;
  p.tiles_h = cdiv(x1.h, g.th);
  p.ntiles = p.tiles_w * p.tiles_h * x1.n;
  p.dbg = c3_env("XRD_C3_DBG", 0);
  p.rowload = 1;
  p.bias = w.bias;
  p.chan_add = e.chan_add; p.chan_add_bstride = e.chan_add_bstride;
  p.resid = e.resid.p; p.y = y.p;
  p.stats = e.stats_out;
  p.in_coef = e.in_coef;
  p.wsw = c3_env("XRD_WBULK", 1) ? w.wtc_swz(x1.dt) : nullptr;
  p.prof = nullptr;
  static long long* prof_buf = nullptr;
  const int want_prof = c3_env("XRD_C3_PROF", 0);
  if (want_prof) {
    if (!prof_buf) { XRD_CUDA(cudaMalloc(&prof_buf, 64 * 8 * sizeof(long long))); }
    XRD_CUDA(cudaMemsetAsync(prof_buf, 0, 64 * 8 * sizeof(long long), c.s));
    p.prof = prof_buf;
  }

  alignas(64) CUtensorMap tmA0, tmA1, tmB;
  c3_encode_act(&tmA0, x1, g.th);
  if (x2) c3_encode_act(&tmA1, *x2, g.th); else tmA1 = tmA0;
  {
    const cuuint64_t dims[3] = {64, (cuuint64_t)w.cout, (cuuint64_t)nkb};
    const cuuint64_t strides[2] = {128, (cuuint64_t)w.cout * 128};
    const cuuint32_t box[3] = {64, (cuuint32_t)w.cout, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = get_encode_tiled()(&tmB, tmap_dtype(x1.dt), 3, w.wtc[x1.dt], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail(XRD_ERR_CUDA, "cuTensorMapEncodeTiled(conv3 weights) failed: %d", (int)r);
  }
  const int nsm = sm_count();
  const int grid = std::min(p.ntiles, nsm);
  if (x1.dt == DT_BF16) c3_dispatch<__nv_bfloat16>(c, w.cout, g, grid, tmA0, tmA1, tmB, p);
  else c3_dispatch<__half>(c, w.cout, g, grid, tmA0, tmA1, tmB, p);

  if (want_prof == 2) {   // dump: per tile, cycles relative to the first sample
    XRD_CUDA(cudaStreamSynchronize(c.s));
    long long h[64 * 8];
    XRD_CUDA(cudaMemcpy(h, prof_buf, sizeof(h), cudaMemcpyDeviceToHost));
    long long t0 = h[0];
    for (int i = 0; i < 64 * 8; ++i) if (h[i] && h[i] < t0) t0 = h[i];
    fprintf(stderr, "CONV3 PROF cin=%d+%d cout=%d W=%d th=%d wres=%d nb=%d nacc=%d dbg=%d: tile prodAempty mmaTop mmaAccFree mmaAfull mmaDone epiTop epiAccFull epiDone\n",
            c0, c1, w.cout, x1.w, g.th, (int)g.wres, g.nb, g.nacc, p.dbg);
    for (int t = 0; t < 24; ++t) {
      fprintf(stderr, "  %2d:", t);
      for (int j = 0; j < 8; ++j) fprintf(stderr, " %8lld", h[t * 8 + j] ? h[t * 8 + j] - t0 : -1);
      fprintf(stderr, "\n");
    }
  }
}

}  // namespace xrd
