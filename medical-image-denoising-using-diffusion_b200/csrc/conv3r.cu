// conv3r.cu -- row-ring tcgen05 implicit-GEMM 3x3 convolution (stride 1, pad 1) for layers with at most 64 input
// channels on wide maps (W % 128 == 0): the 48-channel layers of the UNet's full-resolution level, which are bound by
// memory and by the TMA request rate (one <=128-byte pixel row per ~8 cycles per SM), not by the tensor pipe.
//
//   * work item = (image, 128-column block, segment of up to 32 image rows); one persistent CTA per SM walks items;
//   * every input row of the segment (+ one halo row above and below) is fetched ONCE by one TMA box {64 ch, 130 px}
//     into a ring of row slots; output row r is the nine taps over ring rows r, r+1, r+2 (three runtime base
//     descriptors, compile-time dx / k offsets): activations are read 1.06x instead of the 2x of a 2-row halo tile;
//   * GroupNorm + SiLU of the input (HYB:264-265, 269-270) can be applied in place to each landed row by four extra
//     warps (GN variant): with the deep row ring the transform of row r+3 runs under the MMAs of row r, and each row is
//     transformed once -- the activated tensor is never written to HBM;
//   * all nine weight blocks stay resident; one TMEM accumulator per output row in a ring of 8 (COUT=48) / 4 (COUT=96);
//   * 8-warp register epilogue as in conv3.cu: + bias + time-embedding row + residual, GroupNorm sums of the output.
#include "kernels.cuh"
#include "tc_common.cuh"

#include <algorithm>
#include <mutex>
#include <type_traits>
#include <vector>
#include <stdlib.h>

namespace xrd {

struct Conv3RP {
  int H, W, nimg;
  int ncb, nseg, nitems;    // column blocks per row, row segments per image, work items
  int cin;                  // <= 64, multiple of 16
  const float* bias;
  const float* chan_add; int chan_add_bstride;
  const void* resid;
  void* y;
  double* stats;
  const float2* in_coef;    // GN variant: [nimg][cin] (0.5*scale, 0.5*shift)
  int dbg;                  // bottleneck experiments: 1 no output stores, 2 no MMAs, 4 no row loads after the first ring fill
  long long* prof;          // optional clock64 trace of block 0: [64][8]
};

// clock64 trace of block 0: compiled in only with -DXRD_TRACE (the predicated stores and clock reads sit in the MMA issue stream,
// where every instruction is tensor-pipe idle time)
#ifdef XRD_TRACE
#define XRD_KDBG(m) (p.dbg & (m))          // bottleneck-experiment switches (XRD_C3_DBG / XRD_C3R_DBG), trace builds only
#define RPROF(idx, slot_) do { if (p.prof && blockIdx.x == 0 && (idx) < 64) p.prof[(idx) * 8 + (slot_)] = clock64(); } while (0)
#else
#define XRD_KDBG(m) 0
#define RPROF(idx, slot_) do { } while (0)
#endif

constexpr int kRThreads = 320, kRThreadsGN = 448;
constexpr int kRSeg = 32;                    // output rows per work item
constexpr int kRBox = 130;                   // pixels fetched per row (128 + halo column each side)
constexpr uint32_t kRSlot = 136 * 128;       // ring slot: 17 KB keeps every row 1024-byte aligned

__device__ __forceinline__ void r3_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
template <typename T> __device__ __forceinline__ void r3_unpack8(const uint4& t, float (&v)[8]);
template <> __device__ __forceinline__ void r3_unpack8<__half>(const uint4& t, float (&v)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <> __device__ __forceinline__ void r3_unpack8<__nv_bfloat16>(const uint4& t, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

// the nine taps of one output row: A bases of the three ring rows are runtime, everything else compile time
template <int COUT, int KS>
__device__ __forceinline__ void r3_issue_row(uint64_t a0, uint64_t a1, uint64_t a2, uint64_t bdesc0, uint32_t acc, uint32_t idesc) {
  constexpr uint32_t B16 = (COUT * 128) >> 4;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = tap / 3, dx = tap % 3;
    const uint64_t ad = (dy == 0 ? a0 : (dy == 1 ? a1 : a2)) + (uint64_t)(dx * 8);
#pragma unroll
    for (int k = 0; k < KS; ++k)
      tc::umma_f16(acc, ad + (uint64_t)(k * 2), bdesc0 + (uint64_t)(tap * B16 + k * 2), idesc, (tap | k) ? 1u : 0u);
  }
}

template <typename T, int COUT, int KS, int R, int NA, bool GN>     // KS = input channels / 16: a template constant, no dispatch in the issue stream
__global__ void __launch_bounds__(GN ? kRThreadsGN : kRThreads, 1)
k_conv3r(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Conv3RP p) {
  constexpr uint32_t B_BYTES = COUT * 128;
  constexpr int CPG = COUT / 8;
  constexpr int NBLK = COUT / 48;
  static_assert(COUT % 48 == 0 && NA * COUT <= 512 && (NA & 1) == 0, "accumulator ring must fit TMEM and be even");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                      // [R][17 KB]
  uint8_t* sB = sA + (size_t)R * kRSlot;                   // [9][B_BYTES]
  float* s_badd = (float*)(sB + 9 * (size_t)B_BYTES);      // [8 warps][COUT]
  uint64_t* bars = (uint64_t*)(s_badd + 8 * COUT);
  uint64_t* r_full = bars;                 // [R]  TMA landed
  uint64_t* r_ready = bars + R;            // [R]  GN variant: transformed
  uint64_t* r_empty = bars + 2 * R;        // [R]  MMAs that read the row have completed
  uint64_t* acc_full = bars + 3 * R;       // [NA]
  uint64_t* acc_empty = acc_full + NA;     // [NA]
  uint64_t* w_full = acc_empty + NA;
  uint32_t* tmem_slot = (uint32_t*)(w_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmB);
    for (int s = 0; s < R; ++s) { tc::mbar_init(&r_full[s], 1); tc::mbar_init(&r_ready[s], 128); tc::mbar_init(&r_empty[s], 1); }
    for (int s = 0; s < NA; ++s) { tc::mbar_init(&acc_full[s], 1); tc::mbar_init(&acc_empty[s], 128); }
    tc::mbar_init(w_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, 512);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (*tmem_slot != 0u) {      // one CTA per SM (shared memory) and its only allocation: base 0 keeps MMA operands uniform
    if (threadIdx.x == 0) printf("libxrd: conv3r expects TMEM base 0, got %u\n", *tmem_slot);
    __trap();
  }

  // item q -> (image, column block, row segment)
  auto item = [&](int q, int& img, int& cb, int& r0, int& rows) {
    const int seg = q % p.nseg; q /= p.nseg;
    cb = q % p.ncb; img = q / p.ncb;
    r0 = seg * kRSeg;
    rows = min(kRSeg, p.H - r0);
  };

  if (warp == 0) {
    // ===================== TMA producer: one box per input row =====================
    if (tc::elect_one()) {
      tc::mbar_expect_tx(w_full, 9u * B_BYTES);
      for (int kb = 0; kb < 9; ++kb) tc::tma_load_3d(sB + (size_t)kb * B_BYTES, &tmB, w_full, 0, 0, kb);
      uint32_t slot = 0, phase = 0;
      for (int q = blockIdx.x; q < p.nitems; q += gridDim.x) {
        int img, cb, r0, rows;
        item(q, img, cb, r0, rows);
        for (int j = 0; j < rows + 2; ++j) {
          tc::mbar_wait(&r_empty[slot], phase ^ 1);
          if (XRD_KDBG(4) && phase) {
            tc::mbar_arrive(&r_full[slot]);
          } else {
            tc::mbar_expect_tx(&r_full[slot], (uint32_t)kRBox * 128u);
            tc::tma_load_4d(sA + (size_t)slot * kRSlot, &tmA, &r_full[slot], 0, cb * 128 - 1, r0 - 1 + j, img);   // rows outside the image: zero fill
          }
          if (q == (int)blockIdx.x && (j & 1) == 0) RPROF(j >> 1, 6);
          if (++slot == R) { slot = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    tc::mbar_wait(w_full, 0);
    const uint32_t idesc = tc::umma_idesc(128, COUT, tc::umma_fmt<T>());
    const uint32_t sA_addr = tc::smem_u32(sA);
    const uint64_t bdesc0 = tc::umma_desc_sw128(tc::smem_u32(sB));
    // Two output rows per iteration: the per-iteration skeleton (barrier waits, fences, commits: ~900 cycles measured) is not
    // hidden by the shallow MMA queue, so it is paid once per 54 MMAs instead of once per 27; the waits themselves run in
    // parallel on different lanes of the (converged) warp.
    uint32_t islot = 0, iphase = 0;        // ring position of the first row of the current item
    uint32_t o = 0;                        // running output-row counter (accumulator ring)
    for (int q = blockIdx.x; q < p.nitems; q += gridDim.x) {
      int img, cb, r0, rows;
      item(q, img, cb, r0, rows);
      auto slot_of = [&](int i) { const uint32_t t = islot + (uint32_t)i; return t % R; };
      auto phase_of = [&](int i) { const uint32_t t = islot + (uint32_t)i; return (iphase ^ ((t / R) & 1u)); };
      int landed = 0;                      // item rows already waited for
      bool probed = false;                 // this lane's barrier of the coming iteration was already seen complete (early probe)
      for (int r = 0; r < rows;) {
        const int nrow = (rows - r >= 2) ? 2 : 1;
        if (lane == 0) RPROF(o >> 1, 0);
        const int need = r + nrow + 2 - landed;                    // new input rows this iteration (<= 4)
        if (lane < need) {
          const int i = landed + lane;
          tc::mbar_wait_probed(probed, GN ? &r_ready[slot_of(i)] : &r_full[slot_of(i)], phase_of(i));
        } else if (lane >= 8 && lane < 8 + nrow) {
          const uint32_t oo = o + (uint32_t)(lane - 8);
          tc::mbar_wait_probed(probed, &acc_empty[oo % NA], ((oo / NA) & 1) ^ 1);
        }
        __syncwarp();
        landed += need;
        if (lane == 0) RPROF(o >> 1, 7);
        tc::tc_fence_after();
        // Probe the barriers of the NEXT iteration now, lane by lane as it will wait for them: the probes' round trips run under
        // the MMAs issued below (a satisfied wait in front of the MMAs idles the tensor pipe for its whole latency, mio_probe.cu).
        probed = false;
        {
          const int r2 = r + nrow;
          if (r2 < rows) {
            const int nrow2 = (rows - r2 >= 2) ? 2 : 1;
            const int need2 = r2 + nrow2 + 2 - landed;
            if (lane < need2) {
              const int i = landed + lane;
              probed = tc::mbar_test(GN ? &r_ready[slot_of(i)] : &r_full[slot_of(i)], phase_of(i));
            } else if (lane >= 8 && lane < 8 + nrow2) {
              const uint32_t oo = o + (uint32_t)nrow + (uint32_t)(lane - 8);
              probed = tc::mbar_test(&acc_empty[oo % NA], ((oo / NA) & 1) ^ 1);
            }
          }
        }
        if (lane == 0) RPROF(o >> 1, 1);
        if (tc::elect_one()) {
          const uint32_t s0 = slot_of(r), s1 = slot_of(r + 1), s2 = slot_of(r + 2), s3 = slot_of(r + 3);
          const uint64_t d0 = tc::umma_desc_sw128(sA_addr + s0 * kRSlot), d1 = tc::umma_desc_sw128(sA_addr + s1 * kRSlot),
                         d2 = tc::umma_desc_sw128(sA_addr + s2 * kRSlot), d3 = tc::umma_desc_sw128(sA_addr + s3 * kRSlot);
          const uint32_t a0 = o % NA, a1 = (o + 1) % NA;
          if (!XRD_KDBG(2)) {
            r3_issue_row<COUT, KS>(d0, d1, d2, bdesc0, a0 * COUT, idesc);
            if (nrow == 2) r3_issue_row<COUT, KS>(d1, d2, d3, bdesc0, a1 * COUT, idesc);
          }
          tc::umma_commit(&acc_full[a0]);
          if (nrow == 2) tc::umma_commit(&acc_full[a1]);
          tc::umma_commit(&r_empty[s0]);                 // input rows r (and r+1) are not needed by any later output row
          if (nrow == 2) tc::umma_commit(&r_empty[s1]);
          if (r + nrow == rows) {                        // end of the segment: release its last two rows as well
            tc::umma_commit(&r_empty[slot_of(rows)]);
            tc::umma_commit(&r_empty[slot_of(rows + 1)]);
          }
        }
        __syncwarp();
        if (lane == 0) RPROF(o >> 1, 2);
        o += (uint32_t)nrow;
        r += nrow;
      }
      const uint32_t t = islot + (uint32_t)(rows + 2);
      iphase ^= (t / R) & 1u;
      islot = t % R;
    }
  } else if (GN && warp >= 10) {
    // ===================== input transform (warps 10..13): a = SiLU(GroupNorm(x)) in place, once per landed row =====================
    // thread = (16-byte chunk j of the valid channels, pixel lane); logical chunk j of ring pixel sp sits at physical
    // chunk j ^ (sp & 7) (128B swizzle on absolute addresses; slots are 1024-aligned).  Out-of-image pixels stay zero.
    const int tt = threadIdx.x - 320;
    const int nvc = p.cin >> 3, npl = 128 / nvc;
    const int j8 = tt % nvc, plane = tt / nvc;
    const bool active = plane < npl;
    float sc[8], sh[8];
    int cur_img = -1;
    uint32_t slot = 0, phase = 0;
    for (int q = blockIdx.x; q < p.nitems; q += gridDim.x) {
      int img, cb, r0, rows;
      item(q, img, cb, r0, rows);
      if (img != cur_img && active) {
        const float2* cf = p.in_coef + (size_t)img * p.cin + j8 * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float2 v = __ldg(cf + i); sc[i] = v.x; sh[i] = v.y; }
      }
      cur_img = img;
      const int w0 = cb * 128 - 1;
      for (int j = 0; j < rows + 2; ++j) {
        tc::mbar_wait(&r_full[slot], phase);
        const int ih = r0 - 1 + j;
        if (active && ih >= 0 && ih < p.H) {
          const uint32_t sbase = tc::smem_u32(sA + (size_t)slot * kRSlot);
          constexpr int U = 4;
          for (int c0 = plane; c0 < kRBox; c0 += U * npl) {
            uint32_t addr[U]; bool ok[U]; uint4 qv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int cc = c0 + u * npl;
              const int iw = w0 + cc;
              ok[u] = cc < kRBox && iw >= 0 && iw < p.W;
              addr[u] = sbase + (uint32_t)cc * 128u + (uint32_t)((j8 ^ (cc & 7)) << 4);
              if (ok[u]) asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(qv[u].x), "=r"(qv[u].y), "=r"(qv[u].z), "=r"(qv[u].w) : "r"(addr[u]));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
              if (!ok[u]) continue;
              float v[8];
              r3_unpack8<T>(qv[u], v);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float h = fmaf(v[i], sc[i], sh[i]);          // 0.5 * GroupNorm(x)
                float th;
                asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
                v[i] = fmaf(h, th, h);                              // x*sigmoid(x) = h*tanh(h) + h, h = x/2
              }
              qv[u].x = tc::pack2<T>(v[0], v[1]); qv[u].y = tc::pack2<T>(v[2], v[3]); qv[u].z = tc::pack2<T>(v[4], v[5]); qv[u].w = tc::pack2<T>(v[6], v[7]);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr[u]), "r"(qv[u].x), "r"(qv[u].y), "r"(qv[u].z), "r"(qv[u].w) : "memory");
            }
          }
        }
        tc::fence_async_smem();
        tc::mbar_arrive(&r_ready[slot]);
        if (++slot == R) { slot = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9): group g drains the output rows whose running index is == g mod 2 =====================
    const int quad = warp & 3;
    const int grp = (warp - 2) >> 2;
    T* yp = (T*)p.y;
    const T* rp = (const T*)p.resid;
    float* badd = s_badd + (warp - 2) * COUT;
    float gs[8], gq[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) { gs[g] = 0.f; gq[g] = 0.f; }
    auto flush_stats = [&](int img) {
      if (!p.stats || img < 0) return;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) {
          gs[g] += __shfl_xor_sync(0xffffffffu, gs[g], of);
          gq[g] += __shfl_xor_sync(0xffffffffu, gq[g], of);
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
    uint32_t o = 0;
    int cur_img = -1;
    for (int q = blockIdx.x; q < p.nitems; q += gridDim.x) {
      int img, cb, r0, rows;
      item(q, img, cb, r0, rows);
      if (img != cur_img) {
        flush_stats(cur_img);
        __syncwarp();
        for (int cc = lane; cc < COUT; cc += 32)
          badd[cc] = (p.bias ? __ldg(p.bias + cc) : 0.f) + (p.chan_add ? __ldg(p.chan_add + (int64_t)img * p.chan_add_bstride + cc) : 0.f);
        cur_img = img;
        __syncwarp();
      }
      for (int r = 0; r < rows; ++r, ++o) {
        if ((int)(o & 1) != grp) continue;
        const uint32_t a = o % NA, use = o / NA;
        const int64_t pix = ((int64_t)img * p.H + r0 + r) * p.W + cb * 128 + quad * 32 + lane;
        uint4 rcur[6], rnext[6];
        if (rp) {
#pragma unroll
          for (int j = 0; j < 6; ++j) rcur[j] = __ldg(reinterpret_cast<const uint4*>(rp + pix * COUT) + j);
        }
        if (warp == 2 && lane == 0) RPROF(o >> 1, 3);
        tc::mbar_wait(&acc_full[a], use & 1);
        tc::tc_fence_after();
        if (warp == 2 && lane == 0) RPROF(o >> 1, 4);
        const uint32_t tacc = a * COUT + ((uint32_t)(quad * 32) << 16);
#pragma unroll
        for (int cbk = 0; cbk < NBLK; ++cbk) {
          uint32_t v[48];
          uint4 pk_even;
          r3_tmem_ld16(tacc + (uint32_t)(cbk * 48), *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
          r3_tmem_ld16(tacc + (uint32_t)(cbk * 48 + 16), *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
          r3_tmem_ld16(tacc + (uint32_t)(cbk * 48 + 32), *reinterpret_cast<uint32_t(*)[16]>(&v[32]));
          const bool has_next = rp && cbk + 1 < NBLK;
          if (has_next) {
#pragma unroll
            for (int j = 0; j < 6; ++j) rnext[j] = __ldg(reinterpret_cast<const uint4*>(rp + pix * COUT + (cbk + 1) * 48) + j);
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int h8 = 0; h8 < 6; ++h8) {
            const int co = cbk * 48 + h8 * 8;
            const float4 b0 = *reinterpret_cast<const float4*>(badd + co), b1 = *reinterpret_cast<const float4*>(badd + co + 4);
            float r8[8];
            r8[0] = __uint_as_float(v[h8 * 8 + 0]) + b0.x; r8[1] = __uint_as_float(v[h8 * 8 + 1]) + b0.y;
            r8[2] = __uint_as_float(v[h8 * 8 + 2]) + b0.z; r8[3] = __uint_as_float(v[h8 * 8 + 3]) + b0.w;
            r8[4] = __uint_as_float(v[h8 * 8 + 4]) + b1.x; r8[5] = __uint_as_float(v[h8 * 8 + 5]) + b1.y;
            r8[6] = __uint_as_float(v[h8 * 8 + 6]) + b1.z; r8[7] = __uint_as_float(v[h8 * 8 + 7]) + b1.w;
            if (rp) {
              float q8[8];
              r3_unpack8<T>(rcur[h8], q8);
#pragma unroll
              for (int j = 0; j < 8; ++j) r8[j] += q8[j];
            }
            if (p.stats) {
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
            // two 16-byte halves -> one 32-byte store of a whole sector (partial-sector stores doubled the L2 write requests)
            if (h8 & 1) { if (!XRD_KDBG(1)) tc::st_global_v8(yp + pix * COUT + co - 8, pk_even, pk); } else pk_even = pk;
          }
          if (has_next) {
#pragma unroll
            for (int j = 0; j < 6; ++j) rcur[j] = rnext[j];
          }
        }
        tc::tc_fence_before();
        tc::mbar_arrive(&acc_empty[a]);
        if (warp == 2 && lane == 0) RPROF(o >> 1, 5);
      }
    }
    flush_stats(cur_img);
  }
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(0u, 512);
  }
}

bool conv3r_supported(const Tens& x1, const Tens* x2, const ConvW& w, const ConvEpi& e) {
  static const int enabled = getenv("XRD_CONV3R") ? atoi(getenv("XRD_CONV3R")) : 1;
  if (!enabled) return false;
  if (x1.dt == DT_F32 || x2) return false;
  if (!(w.kh == 3 && w.kw == 3 && w.stride == 1 && w.pad == 1) || w.d2s) return false;
  if (x1.c % 16 != 0 || x1.c > 64) return false;
  if (!(w.cout == 48 || w.cout == 96)) return false;
  if (x1.w % 128 != 0) return false;
  if (e.in_scale || e.out_scale || e.act != ACT_NONE) return false;
  if (e.in_coef && e.in_act != ACT_SILU) return false;
  return true;
}

void conv3r(Ctx& c, const Tens& x, ConvW& w, const ConvEpi& e, Tens& y) {
  XRD_REQUIRE(conv3r_supported(x, nullptr, w, e), "conv3r: unsupported configuration");
  XRD_REQUIRE(x.c == w.cin && y.n == x.n && y.h == x.h && y.w == x.w && y.c == w.cout && y.dt == x.dt, "conv3r: shape mismatch");
  if (e.resid.p) XRD_REQUIRE(e.resid.dt == y.dt && e.resid.numel() == y.numel(), "conv3r: residual mismatch");
  if (c.dry) return;
  if (!w.wtc[x.dt] || w.tc_c1 != x.c) conv_tc_pack(c.s, w, x.dt, x.c);
  XRD_REQUIRE(w.tc_nkb == 9 && w.tc_npad == w.cout, "conv3r: packed weights out of date");
  Conv3RP p;
  p.H = x.h; p.W = x.w; p.nimg = x.n;
  p.ncb = x.w / 128;
  p.nseg = cdiv(x.h, kRSeg);
  p.nitems = p.ncb * p.nseg * x.n;
  p.cin = x.c;
  p.bias = w.bias;
  p.chan_add = e.chan_add; p.chan_add_bstride = e.chan_add_bstride;
  p.resid = e.resid.p; p.y = y.p;
  p.stats = e.stats_out;
  p.in_coef = e.in_coef;
  p.dbg = getenv("XRD_C3R_DBG") ? atoi(getenv("XRD_C3R_DBG")) : 0;
  p.prof = nullptr;
  static long long* prof_buf = nullptr;
  const int want_prof = getenv("XRD_C3R_PROF") ? atoi(getenv("XRD_C3R_PROF")) : 0;
  if (want_prof) {
    if (!prof_buf) XRD_CUDA(cudaMalloc(&prof_buf, 64 * 8 * sizeof(long long)));
    XRD_CUDA(cudaMemsetAsync(prof_buf, 0, 64 * 8 * sizeof(long long), c.s));
    p.prof = prof_buf;
  }

  alignas(64) CUtensorMap tmA, tmB;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)x.c, (cuuint64_t)x.w, (cuuint64_t)x.h, (cuuint64_t)x.n};
    const cuuint64_t strides[3] = {(cuuint64_t)x.c * 2, (cuuint64_t)x.w * x.c * 2, (cuuint64_t)x.h * x.w * x.c * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)kRBox, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = get_encode_tiled()(&tmA, tmap_dtype(x.dt), 4, x.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail(XRD_ERR_CUDA, "cuTensorMapEncodeTiled(conv3r activations) failed: %d", (int)r);
  }
  {
    const cuuint64_t dims[3] = {64, (cuuint64_t)w.cout, 9};
    const cuuint64_t strides[2] = {128, (cuuint64_t)w.cout * 128};
    const cuuint32_t box[3] = {64, (cuuint32_t)w.cout, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = get_encode_tiled()(&tmB, tmap_dtype(x.dt), 3, w.wtc[x.dt], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail(XRD_ERR_CUDA, "cuTensorMapEncodeTiled(conv3r weights) failed: %d", (int)r);
  }
  const int R = w.cout == 48 ? 8 : 6, NA = w.cout == 48 ? 8 : 4;
  const size_t smem = 1024 + (size_t)R * kRSlot + 9 * (size_t)w.cout * 128 + 8 * w.cout * 4 + (3 * R + 2 * NA + 1) * 8 + 64;
  const int nsm = sm_count();
  const int grid = std::min(p.nitems, nsm);
  const bool gn = e.in_coef != nullptr;
  auto launch = [&](auto kern) {
    ensure_dyn_smem(kern, 227 * 1024);
    XRD_LAUNCH(c, kern, grid, gn ? kRThreadsGN : kRThreads, smem, tmA, tmB, p);
  };
  auto pick = [&](auto tag, auto ks_tag) {
    using T = decltype(tag);
    constexpr int KS = decltype(ks_tag)::value;
    if (w.cout == 48) { if (gn) launch(k_conv3r<T, 48, KS, 8, 8, true>); else launch(k_conv3r<T, 48, KS, 8, 8, false>); }
    else { if (gn) launch(k_conv3r<T, 96, KS, 6, 4, true>); else launch(k_conv3r<T, 96, KS, 6, 4, false>); }
  };
  auto pick_ks = [&](auto tag) {
    switch (x.c >> 4) {
      case 4: pick(tag, std::integral_constant<int, 4>()); break;
      case 3: pick(tag, std::integral_constant<int, 3>()); break;
      case 2: pick(tag, std::integral_constant<int, 2>()); break;
      default: pick(tag, std::integral_constant<int, 1>()); break;
    }
  };
  if (x.dt == DT_BF16) pick_ks(__nv_bfloat16()); else pick_ks(__half());
  if (want_prof == 2) {
    XRD_CUDA(cudaStreamSynchronize(c.s));
    long long h[64 * 8];
    XRD_CUDA(cudaMemcpy(h, prof_buf, sizeof(h), cudaMemcpyDeviceToHost));
    long long t0 = 0;
    for (int i = 0; i < 64 * 8; ++i) if (h[i] && (!t0 || h[i] < t0)) t0 = h[i];
    fprintf(stderr, "CONV3R PROF cin=%d cout=%d W=%d gn=%d: pair | mmaTop mmaWaited mmaIssued | epiTop epiAccFull epiDone | tmaRowIssue(item0)\n", x.c, w.cout, x.w, (int)gn);
    for (int t = 0; t < 28; ++t) {
      fprintf(stderr, "  %2d:", t);
      for (int j = 0; j < 8; ++j) fprintf(stderr, " %8lld", h[t * 8 + j] ? h[t * 8 + j] - t0 : -1);
      fprintf(stderr, "\n");
    }
  }
}

}  // namespace xrd
