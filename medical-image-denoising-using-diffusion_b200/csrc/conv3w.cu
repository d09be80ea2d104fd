// conv3w.cu -- persistent tcgen05 implicit-GEMM 3x3 convolution (stride 1, pad 1) for feature maps exactly 64 pixels
// wide: the UNet's lowest level (64x64 at 512^2 input; HYB:327-351, mid blocks and the first up stage).
//
// A UMMA A-operand needs its 128 rows at a uniform shared-memory stride, so a tile row of 64 pixels plus halo
// columns cannot form one M=128 operand out of two image rows.  Instead the tile's rows are staged three times, once
// per horizontal tap: copy dx is the TMA box {64 ch, 64 px starting at x = dx-1, TH+2 rows} (the out-of-image column
// arrives as TMA zero fill), so its rows are exactly 8 KB apart and rows (r, r+1) of ONE copy are a contiguous
// [128 x 64] K-major operand; tap (dy, dx) of output rows (2a, 2a+1) is copy dx at row offset 2a+dy.
// Compared with the per-tap kernel this fetches the activations 3*(6/4) = 4.5x instead of 9x and the weights once per
// 256 pixels instead of once per 128 (the level is bound by L2 -> SM operand traffic, not by the tensor pipe).
//   * tile = 4 image rows x 64 columns of one image = two M=128 accumulators in TMEM (single buffered: the epilogue of
//     a tile is ~7 % of its MMA time);
//   * A stage = one (64-channel chunk, dx) copy, 48 KB, three stages; weight blocks stream through a small ring in the
//     order (chunk, dx, dy);
//   * issue loop with compile-time descriptor offsets, 8-warp register epilogue with bias + time-embedding row +
//     residual + GroupNorm sums, as in conv3.cu;
//   * small batches (the served shape is ONE image: 16 tiles for 148 SMs, each CTA streaming the whole 192 x 192 x 9 weight
//     block): the 192 output channels are split over blockIdx.y into two halves of 96 -- twice the CTAs, half the weight
//     stream and half the MMA columns per CTA; a half covers GroupNorm groups 4*y .. 4*y+3 exactly, so the sums need no care.
#include "kernels.cuh"
#include "tc_common.cuh"

#include <algorithm>
#include <mutex>
#include <vector>
#include <stdlib.h>

namespace xrd {

struct Conv3WP {
  int H, nimg, tiles_h, ntiles;
  int c0, c1, nchunk0, nchunk;
  int nb;                   // weight ring slots
  const float* bias;
  const float* chan_add; int chan_add_bstride;
  const void* resid;
  void* y;
  double* stats;
  const void* wsw;          // pre-swizzled weight blocks for 1-D bulk loads (null: tensor-map loads)
  int ldc;                  // channels of the output / residual tensors (= COUT of the kernel unless the outputs are split)
  uint32_t wblk;            // bytes of one whole packed weight block (ldc rows of 128 bytes)
};

constexpr int kW3Threads = 320;
constexpr int kW3TH = 4;                              // image rows per tile
constexpr uint32_t kW3Row = 64 * 128;                 // one 64-pixel row of a copy: 8 KB
constexpr uint32_t kW3ABytes = (kW3TH + 2) * kW3Row;  // 48 KB
constexpr int kW3Stages = 3;

__device__ __forceinline__ void w3_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
template <typename T> __device__ __forceinline__ void w3_unpack8(const uint4& t, float (&v)[8]);
template <> __device__ __forceinline__ void w3_unpack8<__half>(const uint4& t, float (&v)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <> __device__ __forceinline__ void w3_unpack8<__nv_bfloat16>(const uint4& t, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

// the three vertical taps of one (chunk, dx) copy: 3 weight blocks x 2 accumulators x KS k-steps
template <int COUT, int KS>
__device__ __forceinline__ void w3_issue_dx(uint64_t adesc0, uint32_t sB_addr, uint64_t* b_full, uint64_t* b_empty, uint32_t& wslot,
                                            uint32_t& wphase, uint32_t& bprobed, int nb, uint32_t idesc, uint32_t not_first) {
  constexpr uint32_t B_BYTES = COUT * 128;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    // the barrier of this weight block was probed before the previous block's MMAs were issued (tc_common.cuh: mbar_wait_probed)
    tc::mbar_wait_probed(bprobed != 0u, &b_full[wslot], wphase);
    const uint64_t bdesc = tc::umma_desc_sw128(sB_addr + wslot * B_BYTES);
    uint32_t nslot = wslot + 1, nphase = wphase;
    if (nslot == (uint32_t)nb) { nslot = 0; nphase ^= 1; }
    bprobed = tc::mbar_test(&b_full[nslot], nphase) ? 1u : 0u;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
#pragma unroll
      for (int k = 0; k < KS; ++k)
        tc::umma_f16((uint32_t)(a * COUT), adesc0 + (uint64_t)((((2 * a + dy) * kW3Row) >> 4) + k * 2), bdesc + (uint64_t)(k * 2), idesc,
                     (dy == 0 && k == 0) ? not_first : 1u);
    }
    tc::umma_commit(&b_empty[wslot]);
    wslot = nslot; wphase = nphase;
  }
}

// COUT: output channels this CTA computes = columns [blockIdx.y * COUT, +COUT) of the layer's p.ldc; CPG: channels per GroupNorm
// group of the whole layer (p.ldc / 8).
template <typename T, int COUT, int CPG = COUT / 8>
__global__ void __launch_bounds__(kW3Threads, 1)
k_conv3w(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB,
         const Conv3WP p) {
  constexpr uint32_t B_BYTES = COUT * 128;
  constexpr int NBLK = COUT / 48;
  constexpr int GL = COUT / CPG;            // GroupNorm groups this CTA's columns cover (8, or 4 for a half)
  static_assert(COUT % 48 == 0 && 2 * COUT <= 512, "two accumulators must fit TMEM");
  static_assert(COUT % CPG == 0 && GL <= 8, "an output split must cover whole GroupNorm groups");
  const int coff = blockIdx.y * COUT;       // first output channel of this CTA

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                          // [kW3Stages][48 KB]
  uint8_t* sB = sA + (size_t)kW3Stages * kW3ABytes;            // [nb][B_BYTES]
  float* s_badd = (float*)(sB + (size_t)p.nb * B_BYTES);       // [8 warps][COUT]
  uint64_t* bars = (uint64_t*)(s_badd + 8 * COUT);
  uint64_t* a_full = bars;                 // [3]
  uint64_t* a_empty = bars + 3;            // [3]
  uint64_t* acc_full = bars + 6;
  uint64_t* acc_empty = bars + 7;
  uint64_t* b_full = bars + 8;             // [8]
  uint64_t* b_empty = bars + 16;           // [8]
  uint32_t* tmem_slot = (uint32_t*)(bars + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA0);
    tc::tma_prefetch_desc(&tmA1);
    tc::tma_prefetch_desc(&tmB);
    for (int s = 0; s < kW3Stages; ++s) { tc::mbar_init(&a_full[s], 1); tc::mbar_init(&a_empty[s], 1); }
    tc::mbar_init(acc_full, 1); tc::mbar_init(acc_empty, 256);
    for (int s = 0; s < 8; ++s) { tc::mbar_init(&b_full[s], 1); tc::mbar_init(&b_empty[s], 1); }
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
    if (threadIdx.x == 0) printf("libxrd: conv3w expects TMEM base 0, got %u\n", *tmem_slot);
    __trap();
  }

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (tc::elect_one()) {
      uint32_t st = 0, ph = 0, ws = 0, wph = 0;
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        const int tyi = t % p.tiles_h, img = t / p.tiles_h;
        const int h0 = tyi * kW3TH - 1;
        for (int c = 0; c < p.nchunk; ++c) {
          const bool second = c >= p.nchunk0;
          const CUtensorMap* tm = second ? &tmA1 : &tmA0;
          const int cc0 = (second ? c - p.nchunk0 : c) * 64;
          for (int dx = 0; dx < 3; ++dx) {
            tc::mbar_wait(&a_empty[st], ph ^ 1);
            tc::mbar_expect_tx(&a_full[st], kW3ABytes);
            tc::tma_load_4d(sA + (size_t)st * kW3ABytes, tm, &a_full[st], cc0, dx - 1, h0, img);
            if (++st == kW3Stages) { st = 0; ph ^= 1; }
            for (int dy = 0; dy < 3; ++dy) {
              tc::mbar_wait(&b_empty[ws], wph ^ 1);
              tc::mbar_expect_tx(&b_full[ws], B_BYTES);
              // rows [coff, coff + COUT) of the block: whole 8-row swizzle atoms (COUT % 8 == 0), so a slab of the pre-swizzled copy is
              // itself a correctly swizzled [COUT x 64] operand
              if (p.wsw) tc::bulk_load_1d(sB + (size_t)ws * B_BYTES, (const char*)p.wsw + (size_t)(c * 9 + dy * 3 + dx) * p.wblk + (size_t)coff * 128, B_BYTES, &b_full[ws]);
              else tc::tma_load_3d(sB + (size_t)ws * B_BYTES, &tmB, &b_full[ws], 0, coff, c * 9 + dy * 3 + dx);
              if (++ws == (uint32_t)p.nb) { ws = 0; wph ^= 1; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = tc::umma_idesc(128, COUT, tc::umma_fmt<T>());
    const uint32_t sA_addr = tc::smem_u32(sA), sB_addr = tc::smem_u32(sB);
    uint32_t st = 0, ph = 0, ws = 0, wph = 0, ti = 0;
    uint32_t bprobed = 0u;                   // early probe of the next weight block's barrier (carried across the issue blocks)
    bool aprobed = false;                    // early probe of the next activation stage's barrier
    for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++ti) {
      tc::mbar_wait(acc_empty, (ti & 1) ^ 1);
      tc::tc_fence_after();
      for (int c = 0; c < p.nchunk; ++c) {
        const bool second = c >= p.nchunk0;
        const int cl = second ? c - p.nchunk0 : c;
        const int ks = min(64, (second ? p.c1 : p.c0) - cl * 64) >> 4;
        for (int dx = 0; dx < 3; ++dx) {
          tc::mbar_wait_probed(aprobed, &a_full[st], ph);
          tc::tc_fence_after();
          uint32_t nst = st + 1, nph = ph;
          if (nst == kW3Stages) { nst = 0; nph ^= 1; }
          aprobed = tc::mbar_test(&a_full[nst], nph);      // its round trip runs under the MMAs issued below
          uint32_t leader;
          if (tc::elect_one(leader)) {
            const uint64_t adesc0 = tc::umma_desc_sw128(sA_addr + st * kW3ABytes);
            const uint32_t nf = (c | dx) ? 1u : 0u;
            switch (ks) {
              case 4: w3_issue_dx<COUT, 4>(adesc0, sB_addr, b_full, b_empty, ws, wph, bprobed, p.nb, idesc, nf); break;
              case 3: w3_issue_dx<COUT, 3>(adesc0, sB_addr, b_full, b_empty, ws, wph, bprobed, p.nb, idesc, nf); break;
              case 2: w3_issue_dx<COUT, 2>(adesc0, sB_addr, b_full, b_empty, ws, wph, bprobed, p.nb, idesc, nf); break;
              default: w3_issue_dx<COUT, 1>(adesc0, sB_addr, b_full, b_empty, ws, wph, bprobed, p.nb, idesc, nf); break;
            }
            tc::umma_commit(&a_empty[st]);
            if (c == p.nchunk - 1 && dx == 2) tc::umma_commit(acc_full);
          }
          __syncwarp();
          ws = __shfl_sync(0xffffffffu, ws, leader);
          wph = __shfl_sync(0xffffffffu, wph, leader);
          bprobed = __shfl_sync(0xffffffffu, bprobed, leader);
          st = nst; ph = nph;
        }
      }
    }
  } else {
    // ===================== epilogue: group g = (warp-2)/4 drains accumulator g = image rows (2g, 2g+1) of the tile =====================
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
        for (int o = 16; o > 0; o >>= 1) {
          gs[g] += __shfl_xor_sync(0xffffffffu, gs[g], o);
          gq[g] += __shfl_xor_sync(0xffffffffu, gq[g], o);
        }
      }
      if (lane < 2 * GL) {                   // this CTA's columns are groups coff / CPG .. + GL - 1 of the layer
        float v = 0.f;
#pragma unroll
        for (int g = 0; g < GL; ++g) { if (lane == 2 * g) v = gs[g]; if (lane == 2 * g + 1) v = gq[g]; }
        atomicAdd(p.stats + (size_t)img * 16 + 2 * (coff / CPG) + lane, (double)v);
      }
#pragma unroll
      for (int g = 0; g < 8; ++g) { gs[g] = 0.f; gq[g] = 0.f; }
    };
    uint32_t ti = 0;
    int cur_img = -1;
    for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++ti) {
      const int tyi = t % p.tiles_h, img = t / p.tiles_h;
      if (img != cur_img) {
        flush_stats(cur_img);
        __syncwarp();
        for (int cc = lane; cc < COUT; cc += 32)
          badd[cc] = (p.bias ? __ldg(p.bias + coff + cc) : 0.f) + (p.chan_add ? __ldg(p.chan_add + (int64_t)img * p.chan_add_bstride + coff + cc) : 0.f);
        cur_img = img;
        __syncwarp();
      }
      // M row m of accumulator g <-> image row tyi*4 + 2g + (m >> 6), column m & 63: 128 consecutive NHWC pixels
      const int64_t pix = ((int64_t)img * p.H + (int64_t)tyi * kW3TH + 2 * grp) * 64 + quad * 32 + lane;
      uint4 rcur[6], rnext[6];
      if (rp) {
#pragma unroll
        for (int j = 0; j < 6; ++j) rcur[j] = __ldg(reinterpret_cast<const uint4*>(rp + pix * p.ldc + coff) + j);
      }
      tc::mbar_wait(acc_full, ti & 1);
      tc::tc_fence_after();
      const uint32_t tacc = (uint32_t)(grp * COUT) + ((uint32_t)(quad * 32) << 16);
#pragma unroll
      for (int cb = 0; cb < NBLK; ++cb) {
        uint32_t v[48];
        uint4 pk_even;
        w3_tmem_ld16(tacc + (uint32_t)(cb * 48), *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
        w3_tmem_ld16(tacc + (uint32_t)(cb * 48 + 16), *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
        w3_tmem_ld16(tacc + (uint32_t)(cb * 48 + 32), *reinterpret_cast<uint32_t(*)[16]>(&v[32]));
        const bool has_next = rp && cb + 1 < NBLK;
        if (has_next) {
#pragma unroll
          for (int j = 0; j < 6; ++j) rnext[j] = __ldg(reinterpret_cast<const uint4*>(rp + pix * p.ldc + coff + (cb + 1) * 48) + j);
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int h8 = 0; h8 < 6; ++h8) {
          const int co = cb * 48 + h8 * 8;
          const float4 b0 = *reinterpret_cast<const float4*>(badd + co), b1 = *reinterpret_cast<const float4*>(badd + co + 4);
          float r8[8];
          r8[0] = __uint_as_float(v[h8 * 8 + 0]) + b0.x; r8[1] = __uint_as_float(v[h8 * 8 + 1]) + b0.y;
          r8[2] = __uint_as_float(v[h8 * 8 + 2]) + b0.z; r8[3] = __uint_as_float(v[h8 * 8 + 3]) + b0.w;
          r8[4] = __uint_as_float(v[h8 * 8 + 4]) + b1.x; r8[5] = __uint_as_float(v[h8 * 8 + 5]) + b1.y;
          r8[6] = __uint_as_float(v[h8 * 8 + 6]) + b1.z; r8[7] = __uint_as_float(v[h8 * 8 + 7]) + b1.w;
          if (rp) {
            float q8[8];
            w3_unpack8<T>(rcur[h8], q8);
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
          if (h8 & 1) tc::st_global_v8(yp + pix * p.ldc + coff + co - 8, pk_even, pk); else pk_even = pk;   // 32-byte sector stores
        }
        if (has_next) {
#pragma unroll
          for (int j = 0; j < 6; ++j) rcur[j] = rnext[j];
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(acc_empty);
    }
    flush_stats(cur_img);
  }
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(0u, 512);
  }
}

bool conv3w_supported(const Tens& x1, const Tens* x2, const ConvW& w, const ConvEpi& e) {
  static const int enabled = getenv("XRD_CONV3W") ? atoi(getenv("XRD_CONV3W")) : 1;
  if (!enabled) return false;
  if (x1.dt == DT_F32) return false;
  if (!(w.kh == 3 && w.kw == 3 && w.stride == 1 && w.pad == 1) || w.d2s) return false;
  if (x1.c % 16 != 0 || (x2 && x2->c % 16 != 0)) return false;
  if (!(w.cout == 144 || w.cout == 192)) return false;
  if (x1.w != 64 || x1.h % kW3TH != 0) return false;
  if (e.in_scale || e.out_scale || e.act != ACT_NONE) return false;
  return true;
}

void conv3w(Ctx& c, const Tens& x1, const Tens* x2, ConvW& w, const ConvEpi& e, Tens& y) {
  XRD_REQUIRE(conv3w_supported(x1, x2, w, e), "conv3w: unsupported configuration");
  const int c0 = x1.c, c1 = x2 ? x2->c : 0;
  XRD_REQUIRE(c0 + c1 == w.cin && y.n == x1.n && y.h == x1.h && y.w == x1.w && y.c == w.cout && y.dt == x1.dt, "conv3w: shape mismatch");
  if (x2) XRD_REQUIRE(x2->n == x1.n && x2->h == x1.h && x2->w == x1.w && x2->dt == x1.dt, "conv3w: source mismatch");
  if (e.resid.p) XRD_REQUIRE(e.resid.dt == y.dt && e.resid.numel() == y.numel(), "conv3w: residual mismatch");
  if (c.dry) return;
  if (!w.wtc[x1.dt] || w.tc_c1 != c0) conv_tc_pack(c.s, w, x1.dt, c0);
  Conv3WP p;
  p.H = x1.h; p.nimg = x1.n;
  p.tiles_h = x1.h / kW3TH;
  p.ntiles = p.tiles_h * x1.n;
  p.c0 = c0; p.c1 = c1;
  p.nchunk0 = (c0 + 63) / 64;
  p.nchunk = p.nchunk0 + (c1 + 63) / 64;
  XRD_REQUIRE(p.nchunk * 9 == w.tc_nkb && w.tc_npad == w.cout, "conv3w: packed weights out of date");
  const int nsm = sm_count();
  // output split for small batches (see the header): only where a half is whole GroupNorm groups and whole 48-column epilogue
  // blocks, i.e. 192 outputs, and only while twice the tiles still fit the machine
  static const int split_on = getenv("XRD_C3W_SPLIT") ? atoi(getenv("XRD_C3W_SPLIT")) : 1;
  const int nsplit = (split_on && w.cout == 192 && 2 * p.ntiles <= nsm) ? 2 : 1;
  const int co = w.cout / nsplit;
  p.ldc = w.cout;
  p.wblk = (uint32_t)w.cout * 128u;
  const size_t bb = (size_t)co * 128;
  const size_t fixed = (size_t)kW3Stages * kW3ABytes + 8 * co * 4 + 32 * 8 + 64;
  p.nb = (int)std::min<size_t>(8, (227 * 1024 - 1024 - fixed) / bb);
  XRD_REQUIRE(p.nb >= 2, "conv3w: shared memory budget exceeded");
  p.bias = w.bias;
  p.chan_add = e.chan_add; p.chan_add_bstride = e.chan_add_bstride;
  p.resid = e.resid.p; p.y = y.p;
  p.stats = e.stats_out;
  p.wsw = (getenv("XRD_WBULK") && atoi(getenv("XRD_WBULK")) == 0) ? nullptr : w.wtc_swz(x1.dt);

  auto encode_act = [&](CUtensorMap* m, const Tens& x) {
    const cuuint64_t dims[4] = {(cuuint64_t)x.c, (cuuint64_t)x.w, (cuuint64_t)x.h, (cuuint64_t)x.n};
    const cuuint64_t strides[3] = {(cuuint64_t)x.c * 2, (cuuint64_t)x.w * x.c * 2, (cuuint64_t)x.h * x.w * x.c * 2};
    const cuuint32_t box[4] = {64, 64, (cuuint32_t)(kW3TH + 2), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = get_encode_tiled()(m, tmap_dtype(x.dt), 4, x.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail(XRD_ERR_CUDA, "cuTensorMapEncodeTiled(conv3w activations) failed: %d", (int)r);
  };
  alignas(64) CUtensorMap tmA0, tmA1, tmB;
  encode_act(&tmA0, x1);
  if (x2) encode_act(&tmA1, *x2); else tmA1 = tmA0;
  {
    const cuuint64_t dims[3] = {64, (cuuint64_t)w.cout, (cuuint64_t)w.tc_nkb};
    const cuuint64_t strides[2] = {128, (cuuint64_t)w.cout * 128};
    const cuuint32_t box[3] = {64, (cuuint32_t)co, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = get_encode_tiled()(&tmB, tmap_dtype(x1.dt), 3, w.wtc[x1.dt], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail(XRD_ERR_CUDA, "cuTensorMapEncodeTiled(conv3w weights) failed: %d", (int)r);
  }
  const size_t smem = 1024 + fixed + (size_t)p.nb * bb;
  const dim3 grid(std::min(p.ntiles, nsm), nsplit);
  auto launch = [&](auto kern) {
    ensure_dyn_smem(kern, 227 * 1024);
    XRD_LAUNCH(c, kern, grid, kW3Threads, smem, tmA0, tmA1, tmB, p);
  };
  if (x1.dt == DT_BF16) {
    if (w.cout == 144) launch(k_conv3w<__nv_bfloat16, 144>);
    else if (nsplit == 2) launch(k_conv3w<__nv_bfloat16, 96, 24>);
    else launch(k_conv3w<__nv_bfloat16, 192>);
  } else {
    if (w.cout == 144) launch(k_conv3w<__half, 144>);
    else if (nsplit == 2) launch(k_conv3w<__half, 96, 24>);
    else launch(k_conv3w<__half, 192>);
  }
}

}  // namespace xrd
