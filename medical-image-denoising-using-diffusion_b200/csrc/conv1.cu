// conv1.cu -- persistent tcgen05 GEMM for the UNet's 1x1 convolutions (res_conv HYB:274, qkv / proj HYB:289-290).
//
//   Y[pixel, n0 + 0..COUT) = [X1 | X2][pixel, :] * W[n0 + 0..COUT, :]^T (+ bias, + residual), pixels = N*H*W flattened
//
// The per-tap kernel (conv_tc.cu) launches one CTA per 128 pixels; for a 1x1 conv that is 3..6 K blocks of work
// behind ~10 us of per-CTA set-up (TMEM allocation, barrier init, first TMA round trip, a serial epilogue):
// measured 170 us for qkv 192->576 @64^2 x16 whose operands move in ~20 us.  Here:
//   * one CTA per SM walks 128-pixel tiles (static round robin); the [COUT x K] weight slice stays in shared memory;
//   * activations stream through a ring of [128 px x 64 ch] TMA tiles (two sources walked back to back = the
//     virtual channel concat of the up path, HYB:383), no halo, no image geometry: the tensor is a 2-D matrix;
//   * two TMEM accumulators: epilogue group g (4 warps) drains accumulator g while the tensor core fills the other;
//   * the MMA issue loop uses compile-time descriptor offsets (see conv3.cu);
//   * epilogue: + bias, + residual (prefetched), optional GroupNorm sums of the output, 16-byte stores of whole
//     sectors; an output wider than 192 channels is covered by blockIdx.y (qkv: 3 slices of 192).
// The same kernel serves the NAFBlock 1x1 convolutions (HYB:152-169; channel counts 32..1024, powers of two): slices of
// 32/64/128/256 output channels drained in 32-column blocks, the per-channel beta/gamma scale of HYB:161,169 in the
// epilogue, and for conv4 (C -> 2C, 2C <= 256) SimpleGate (HYB:119-121: first half x second half) applied to the
// accumulator row before it is stored, so the 2C-channel tensor never reaches HBM.
#include "kernels.cuh"
#include "tc_common.cuh"

#include <algorithm>
#include <mutex>
#include <vector>
#include <stdlib.h>

namespace xrd {

struct Conv1P {
  int64_t npix;             // N*H*W
  int hw;                   // pixels per image (statistics only)
  int ntiles;
  int c0, c1;               // channels of the two sources (c1 = 0: single source)
  int nchunk0, nchunk;      // 64-channel chunks of source 0 / of both
  int ldc;                  // output channels per pixel (row stride of Y and of the residual)
  const float* bias;
  const float* out_scale;   // optional per-output-channel factor on (acc + bias)
  const void* resid;
  void* y;
  double* stats;            // optional [nimg][8][2]; requires ldc == COUT and hw % 128 == 0
};

constexpr int kC1Threads = 320;    // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two groups of four)
constexpr int kC1Stages = 4;       // activation ring
constexpr uint32_t kC1ATile = 128 * 128;

__device__ __forceinline__ void c1_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

template <typename T> __device__ __forceinline__ void c1_unpack8(const uint4& t, float (&v)[8]);
template <> __device__ __forceinline__ void c1_unpack8<__half>(const uint4& t, float (&v)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <> __device__ __forceinline__ void c1_unpack8<__nv_bfloat16>(const uint4& t, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

template <int KS>
__device__ __forceinline__ void c1_issue(uint32_t acc, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t not_first) {
#pragma unroll
  for (int k = 0; k < KS; ++k) tc::umma_f16(acc, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, k == 0 ? not_first : 1u);
}

template <typename T, int COUT, bool GATE>
__global__ void __launch_bounds__(kC1Threads, 1)
k_conv1(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB,
        const Conv1P p) {
  constexpr uint32_t B_BYTES = COUT * 128;                 // one 64-channel K block of the weight slice
  constexpr int OC = GATE ? COUT / 2 : COUT;               // channels stored per pixel by this CTA
  constexpr int EB = (OC % 48 == 0) ? 48 : 32;             // epilogue block: columns drained per TMEM round trip
  constexpr int CPG = OC / 8;
  constexpr int NBLK = OC / EB;
  // The whole TMEM, whatever the two accumulators need: the MMA and epilogue addresses below are compile-time constants relative
  // to base 0, and only a 512-column allocation is guaranteed to start there once CTAs of OTHER kernels may share the SM (the
  // hybrid's side branches run NAFNet / the router on their own streams).  A partial allocation by a co-resident CTA makes this
  // one wait in tcgen05.alloc until that CTA has released its columns -- no CTA of this library allocates twice, so nobody holds
  // columns while waiting for more.
  constexpr uint32_t TMEM_COLS = 512;
  static_assert(OC % EB == 0 && COUT % 16 == 0 && COUT <= 256 && 2 * COUT <= 512, "two accumulators must fit TMEM");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                        // [kC1Stages][16 KB]
  uint8_t* sB = sA + (size_t)kC1Stages * kC1ATile;           // [nchunk][B_BYTES] resident weight slice
  float* s_bias = (float*)(sB + (size_t)p.nchunk * B_BYTES); // [COUT]
  float* s_scale = s_bias + COUT;                            // [COUT]
  uint64_t* bars = (uint64_t*)(s_scale + COUT);
  uint64_t* a_full = bars;                   // [kC1Stages]
  uint64_t* a_empty = bars + kC1Stages;      // [kC1Stages]
  uint64_t* acc_full = bars + 2 * kC1Stages; // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint64_t* w_full = acc_empty + 2;
  uint32_t* tmem_slot = (uint32_t*)(w_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * COUT;           // first output channel of this CTA's slice (GATE: a single slice)

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA0);
    tc::tma_prefetch_desc(&tmA1);
    tc::tma_prefetch_desc(&tmB);
    for (int s = 0; s < kC1Stages; ++s) { tc::mbar_init(&a_full[s], 1); tc::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&acc_full[s], 1); tc::mbar_init(&acc_empty[s], 128); }
    tc::mbar_init(w_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, TMEM_COLS);
    tc::tmem_relinquish();
  }
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) {
    s_bias[i] = p.bias ? p.bias[n0 + i] : 0.f;
    s_scale[i] = (p.out_scale && i < OC) ? p.out_scale[n0 + i] : 1.f;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (*tmem_slot != 0u) {      // one CTA per SM (shared memory) and its only allocation: base 0 keeps MMA operands uniform
    if (threadIdx.x == 0) printf("libxrd: conv1 expects TMEM base 0, got %u\n", *tmem_slot);
    __trap();
  }

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (tc::elect_one()) {
      tc::mbar_expect_tx(w_full, (uint32_t)p.nchunk * B_BYTES);
      for (int kb = 0; kb < p.nchunk; ++kb) tc::tma_load_3d(sB + (size_t)kb * B_BYTES, &tmB, w_full, 0, n0, kb);
      uint32_t st = 0, ph = 0;
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        for (int c = 0; c < p.nchunk; ++c) {
          tc::mbar_wait(&a_empty[st], ph ^ 1);
          tc::mbar_expect_tx(&a_full[st], kC1ATile);
          const bool second = c >= p.nchunk0;
          tc::tma_load_2d(sA + (size_t)st * kC1ATile, second ? &tmA1 : &tmA0, &a_full[st], (second ? c - p.nchunk0 : c) * 64, t * 128);
          if (++st == kC1Stages) { st = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    tc::mbar_wait(w_full, 0);
    const uint32_t idesc = tc::umma_idesc(128, COUT, tc::umma_fmt<T>());
    const uint32_t sA_addr = tc::smem_u32(sA), sB_addr = tc::smem_u32(sB);
    uint32_t st = 0, ph = 0, ti = 0;
    bool a_probed = false, acc_probed = false;     // early probes of the next stage / next accumulator barrier (tc_common.cuh)
    for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++ti) {
      const uint32_t ab = ti & 1, use = ti >> 1;
      tc::mbar_wait_probed(acc_probed, &acc_empty[ab], (use & 1) ^ 1);
      tc::tc_fence_after();
      for (int c = 0; c < p.nchunk; ++c) {
        tc::mbar_wait_probed(a_probed, &a_full[st], ph);
        tc::tc_fence_after();
        {
          uint32_t nst = st + 1, nph = ph;
          if (nst == kC1Stages) { nst = 0; nph ^= 1; }
          a_probed = tc::mbar_test(&a_full[nst], nph);            // the round trip runs under the MMAs issued below
          if (c == p.nchunk - 1) acc_probed = tc::mbar_test(&acc_empty[(ti + 1) & 1], (((ti + 1) >> 1) & 1) ^ 1);
        }
        const bool second = c >= p.nchunk0;
        const int cl = second ? c - p.nchunk0 : c;
        const int ks = min(64, (second ? p.c1 : p.c0) - cl * 64) >> 4;
        if (tc::elect_one()) {
          const uint64_t adesc = tc::umma_desc_sw128(sA_addr + st * kC1ATile);
          const uint64_t bdesc = tc::umma_desc_sw128(sB_addr + (uint32_t)c * B_BYTES);
          const uint32_t acc = ab * COUT, nf = c ? 1u : 0u;
          switch (ks) {
            case 4: c1_issue<4>(acc, adesc, bdesc, idesc, nf); break;
            case 3: c1_issue<3>(acc, adesc, bdesc, idesc, nf); break;
            case 2: c1_issue<2>(acc, adesc, bdesc, idesc, nf); break;
            default: c1_issue<1>(acc, adesc, bdesc, idesc, nf); break;
          }
          tc::umma_commit(&a_empty[st]);
          if (c == p.nchunk - 1) tc::umma_commit(&acc_full[ab]);
        }
        __syncwarp();
        if (++st == kC1Stages) { st = 0; ph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: group g = (warp-2)/4 drains accumulator g (tiles with ti % 2 == g) =====================
    const int quad = warp & 3;
    const int grp = (warp - 2) >> 2;
    T* yp = (T*)p.y;
    const T* rp = (const T*)p.resid;
    float gs[8], gq[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) { gs[g] = 0.f; gq[g] = 0.f; }
    int cur_img = -1;
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
    for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++ti) {
      if ((int)(ti & 1) != grp) continue;
      const uint32_t use = ti >> 1;
      const int64_t pix = (int64_t)t * 128 + quad * 32 + lane;
      const bool ok = pix < p.npix;
      if (p.stats) {                                   // hw % 128 == 0: the whole tile belongs to one image
        const int img = (int)(((int64_t)t * 128) / p.hw);
        if (img != cur_img) { flush_stats(cur_img); cur_img = img; }
      }
      uint4 rcur[EB / 8], rnext[EB / 8];
      if (rp && ok) {
#pragma unroll
        for (int j = 0; j < EB / 8; ++j) rcur[j] = __ldg(reinterpret_cast<const uint4*>(rp + pix * p.ldc + n0) + j);
      }
      tc::mbar_wait(&acc_full[grp], use & 1);
      tc::tc_fence_after();
      const uint32_t tacc = (uint32_t)(grp * COUT) + ((uint32_t)(quad * 32) << 16);
#pragma unroll
      for (int cb = 0; cb < NBLK; ++cb) {
        uint32_t v[EB], v2[GATE ? EB : 1];
        uint4 pk_even;
#pragma unroll
        for (int j = 0; j < EB / 16; ++j) c1_tmem_ld16(tacc + (uint32_t)(cb * EB + j * 16), *reinterpret_cast<uint32_t(*)[16]>(&v[j * 16]));
        if (GATE) {
#pragma unroll
          for (int j = 0; j < EB / 16; ++j)
            c1_tmem_ld16(tacc + (uint32_t)(OC + cb * EB + j * 16), *reinterpret_cast<uint32_t(*)[16]>(&v2[GATE ? j * 16 : 0]));
        }
        const bool has_next = rp && ok && cb + 1 < NBLK;
        if (has_next) {
#pragma unroll
          for (int j = 0; j < EB / 8; ++j) rnext[j] = __ldg(reinterpret_cast<const uint4*>(rp + pix * p.ldc + n0 + (cb + 1) * EB) + j);
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int h8 = 0; h8 < EB / 8; ++h8) {
          const int co = cb * EB + h8 * 8;
          const float4 b0 = *reinterpret_cast<const float4*>(s_bias + co), b1 = *reinterpret_cast<const float4*>(s_bias + co + 4);
          float r8[8];
          r8[0] = __uint_as_float(v[h8 * 8 + 0]) + b0.x; r8[1] = __uint_as_float(v[h8 * 8 + 1]) + b0.y;
          r8[2] = __uint_as_float(v[h8 * 8 + 2]) + b0.z; r8[3] = __uint_as_float(v[h8 * 8 + 3]) + b0.w;
          r8[4] = __uint_as_float(v[h8 * 8 + 4]) + b1.x; r8[5] = __uint_as_float(v[h8 * 8 + 5]) + b1.y;
          r8[6] = __uint_as_float(v[h8 * 8 + 6]) + b1.z; r8[7] = __uint_as_float(v[h8 * 8 + 7]) + b1.w;
          if (GATE) {                                    // SimpleGate: channel co times channel OC + co (HYB:119-121)
            const float4 g0 = *reinterpret_cast<const float4*>(s_bias + OC + co), g1 = *reinterpret_cast<const float4*>(s_bias + OC + co + 4);
            r8[0] *= __uint_as_float(v2[GATE ? h8 * 8 + 0 : 0]) + g0.x; r8[1] *= __uint_as_float(v2[GATE ? h8 * 8 + 1 : 0]) + g0.y;
            r8[2] *= __uint_as_float(v2[GATE ? h8 * 8 + 2 : 0]) + g0.z; r8[3] *= __uint_as_float(v2[GATE ? h8 * 8 + 3 : 0]) + g0.w;
            r8[4] *= __uint_as_float(v2[GATE ? h8 * 8 + 4 : 0]) + g1.x; r8[5] *= __uint_as_float(v2[GATE ? h8 * 8 + 5 : 0]) + g1.y;
            r8[6] *= __uint_as_float(v2[GATE ? h8 * 8 + 6 : 0]) + g1.z; r8[7] *= __uint_as_float(v2[GATE ? h8 * 8 + 7 : 0]) + g1.w;
          }
          if (p.out_scale) {
            const float4 s0 = *reinterpret_cast<const float4*>(s_scale + co), s1 = *reinterpret_cast<const float4*>(s_scale + co + 4);
            r8[0] *= s0.x; r8[1] *= s0.y; r8[2] *= s0.z; r8[3] *= s0.w; r8[4] *= s1.x; r8[5] *= s1.y; r8[6] *= s1.z; r8[7] *= s1.w;
          }
          if (rp && ok) {
            float q8[8];
            c1_unpack8<T>(rcur[h8], q8);
#pragma unroll
            for (int j = 0; j < 8; ++j) r8[j] += q8[j];
          }
          if (p.stats && ok) {
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
          if (h8 & 1) { if (ok) tc::st_global_v8(yp + pix * p.ldc + (GATE ? 0 : n0) + co - 8, pk_even, pk); } else pk_even = pk;   // 32-byte sector stores
        }
        if (has_next) {
#pragma unroll
          for (int j = 0; j < EB / 8; ++j) rcur[j] = rnext[j];
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&acc_empty[grp]);
    }
    flush_stats(cur_img);
  }
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(0u, TMEM_COLS);
  }
}

// output channels per CTA slice: the UNet's multiples of 48 (<= 192) or NAFNet's powers of two (<= 256), weight slice resident
static int c1_slice(int cout, int nchunk, bool gate) {
  auto fits = [&](int sl) {
    return 1024 + (size_t)kC1Stages * kC1ATile + (size_t)nchunk * sl * 128 + (size_t)sl * 8 + 256 <= (size_t)227 * 1024;
  };
  if (gate) return (cout == 64 || cout == 128 || cout == 256) && fits(cout) ? cout : 0;
  if (cout % 48 == 0) {
    const int sl = cout <= 192 ? cout : 192;
    return (sl == 48 || sl == 96 || sl == 144 || sl == 192) && cout % sl == 0 && fits(sl) ? sl : 0;
  }
  for (int sl : {256, 128, 64, 32})
    if (cout % sl == 0 && fits(sl)) return sl;
  return 0;
}

bool conv1_supported(const Tens& x1, const Tens* x2, const ConvW& w, const ConvEpi& e) {
  static const int enabled = getenv("XRD_CONV1") ? atoi(getenv("XRD_CONV1")) : 1;
  if (!enabled) return false;
  if (x1.dt == DT_F32) return false;
  if (!(w.kh == 1 && w.kw == 1 && w.stride == 1 && w.pad == 0) || w.d2s) return false;
  if (x1.c % 16 != 0 || (x2 && x2->c % 16 != 0)) return false;
  if (e.in_scale || e.chan_add || e.act != ACT_NONE || e.in_coef) return false;
  const int nchunk = (x1.c + 63) / 64 + (x2 ? (x2->c + 63) / 64 : 0);
  const int slice = c1_slice(w.cout, nchunk, e.gate);
  if (!slice) return false;
  if (e.stats_out && (e.gate || w.cout != slice || ((int64_t)x1.h * x1.w) % 128 != 0)) return false;
  return true;
}

void conv1(Ctx& c, const Tens& x1, const Tens* x2, ConvW& w, const ConvEpi& e, Tens& y) {
  XRD_REQUIRE(conv1_supported(x1, x2, w, e), "conv1: unsupported configuration");
  const int c0 = x1.c, c1 = x2 ? x2->c : 0;
  XRD_REQUIRE(c0 + c1 == w.cin && y.n == x1.n && y.h == x1.h && y.w == x1.w && y.c == (e.gate ? w.cout / 2 : w.cout) && y.dt == x1.dt,
              "conv1: shape mismatch");
  if (x2) XRD_REQUIRE(x2->n == x1.n && x2->h == x1.h && x2->w == x1.w && x2->dt == x1.dt, "conv1: source mismatch");
  if (e.resid.p) XRD_REQUIRE(e.resid.dt == y.dt && e.resid.numel() == y.numel(), "conv1: residual mismatch");
  if (c.dry) return;
  if (!w.wtc[x1.dt] || w.tc_c1 != c0) conv_tc_pack(c.s, w, x1.dt, c0);
  const int slice = c1_slice(w.cout, (c0 + 63) / 64 + (c1 + 63) / 64, e.gate);
  Conv1P p;
  p.npix = (int64_t)x1.n * x1.h * x1.w;
  p.hw = x1.h * x1.w;
  p.ntiles = (int)cdiv64(p.npix, 128);
  p.c0 = c0; p.c1 = c1;
  p.nchunk0 = (c0 + 63) / 64;
  p.nchunk = p.nchunk0 + (c1 + 63) / 64;
  XRD_REQUIRE(p.nchunk == w.tc_nkb, "conv1: packed weights out of date");
  p.ldc = y.c;
  p.bias = w.bias;
  p.out_scale = e.out_scale;
  p.resid = e.resid.p; p.y = y.p;
  p.stats = e.stats_out;

  auto encode_act = [&](CUtensorMap* m, const Tens& x) {
    const cuuint64_t dims[2] = {(cuuint64_t)x.c, (cuuint64_t)p.npix};
    const cuuint64_t strides[1] = {(cuuint64_t)x.c * 2};
    const cuuint32_t box[2] = {64, 128};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode_tiled()(m, tmap_dtype(x.dt), 2, x.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail(XRD_ERR_CUDA, "cuTensorMapEncodeTiled(conv1 activations) failed: %d", (int)r);
  };
  alignas(64) CUtensorMap tmA0, tmA1, tmB;
  encode_act(&tmA0, x1);
  if (x2) encode_act(&tmA1, *x2); else tmA1 = tmA0;
  {
    const cuuint64_t dims[3] = {64, (cuuint64_t)w.tc_npad, (cuuint64_t)p.nchunk};
    const cuuint64_t strides[2] = {128, (cuuint64_t)w.tc_npad * 128};
    const cuuint32_t box[3] = {64, (cuuint32_t)slice, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = get_encode_tiled()(&tmB, tmap_dtype(x1.dt), 3, w.wtc[x1.dt], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail(XRD_ERR_CUDA, "cuTensorMapEncodeTiled(conv1 weights) failed: %d", (int)r);
  }
  const size_t smem = 1024 + (size_t)kC1Stages * kC1ATile + (size_t)p.nchunk * slice * 128 + (size_t)slice * 8 + 256;
  const int nsm = sm_count();
  const int nslices = w.cout / slice;
  dim3 grid(std::max(1, std::min(p.ntiles, nsm / nslices)), nslices);
  auto launch = [&](auto kern) {
    ensure_dyn_smem(kern, 227 * 1024);
    XRD_LAUNCH(c, kern, grid, kC1Threads, smem, tmA0, tmA1, tmB, p);
  };
  auto pick = [&](auto tag) {
    using T = decltype(tag);
    if (e.gate) {
      if (slice == 64) launch(k_conv1<T, 64, true>); else if (slice == 128) launch(k_conv1<T, 128, true>); else launch(k_conv1<T, 256, true>);
      return;
    }
    switch (slice) {
      case 48: launch(k_conv1<T, 48, false>); break;
      case 96: launch(k_conv1<T, 96, false>); break;
      case 144: launch(k_conv1<T, 144, false>); break;
      case 192: launch(k_conv1<T, 192, false>); break;
      case 32: launch(k_conv1<T, 32, false>); break;
      case 64: launch(k_conv1<T, 64, false>); break;
      case 128: launch(k_conv1<T, 128, false>); break;
      default: launch(k_conv1<T, 256, false>); break;
    }
  };
  if (x1.dt == DT_BF16) pick(__nv_bfloat16()); else pick(__half());
}

}  // namespace xrd
