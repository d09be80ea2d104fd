// conv_halo.cu -- persistent, halo-reusing tcgen05 implicit-GEMM 3x3 convolution for wide feature maps.
//
// The per-tap kernel in conv_tc.cu re-fetches the activation tile once per filter tap (9x) and the weight tile once
// per 128 pixels; at N = Cout <= 96 that is far more L2->SM traffic than the tensor core can hide.  Here:
//   * one CTA per SM loops over tiles of TH image rows x 128 columns (persistent, static round robin);
//   * per 64-channel chunk ONE TMA box {64 ch, 130, TH+2} brings the tile plus its halo into shared memory
//     (zero padding and channel tails again by TMA out-of-bounds fill);
//   * the nine taps are nine *descriptor offsets* into that buffer: tap (dy,dx) of image row s starts at
//     buffer pixel (s+dy)*130 + dx.  The start address is then no longer 1024-byte aligned; measured on B200 this
//     needs NO descriptor base offset: the 128B swizzle is applied on absolute shared-memory address bits, exactly
//     as TMA wrote it (setting base_offset = (addr>>7)&7 gives wrong results, leaving it 0 is bit-correct);
//   * weights stay resident in shared memory for the whole kernel when all taps x chunks fit, otherwise they are
//     streamed through a small ring, once per tile instead of once per 128 pixels;
//   * TH accumulators (one per image row) live in TMEM, double buffered when 2*TH*N <= 512 columns so the epilogue
//     of tile i overlaps the MMAs of tile i+1.
#include "kernels.cuh"
#include "tc_common.cuh"

#include <stdlib.h>

namespace xrd {

struct ConvHaloP {
  int H, W, nimg;
  int th;                   // image rows per tile (== accumulators)
  int tiles_w, tiles_h, ntiles;
  int cin, nchunk;          // input channels, 64-channel chunks
  int bn, cout;             // N tile (== padded cout, <= 256) and valid output channels
  int wres;                 // 1: all weight K blocks resident in smem
  int nb;                   // weight ring stages when streaming
  int nacc;                 // TMEM accumulator buffers (1 or 2)
  int use_bo;               // descriptor base-offset field
  int dbg;                  // bottleneck experiments: 1 no output store, 2 no MMA, 4 no activation loads
  uint32_t out_bytes;       // staging for one output row of the tile: 128 * cout * 2, 1024-aligned
  uint32_t a_bytes;         // one halo buffer, 1024-aligned
  uint32_t b_bytes;         // one weight K block: bn * 128
  uint32_t idesc, tmem_cols;
  const float* bias;
  const float* chan_add; int chan_add_bstride;
  const void* resid;
  void* y;
};

constexpr int kHaloThreads = 192;
constexpr int kHaloPitch = 130;   // 128 output columns + one halo column on each side

__device__ __forceinline__ uint64_t umma_desc_sw128_bo(uint32_t saddr, int use_bo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  if (use_bo) d |= (uint64_t)((saddr >> 7) & 7) << 49;   // swizzle phase of the first row
  d |= (uint64_t)2 << 61;
  return d;
}

template <typename T>
__global__ void __launch_bounds__(kHaloThreads, 1)
k_conv_halo(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvHaloP p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int nkb = p.nchunk * 9;
  const int nbslots = p.wres ? nkb : p.nb;
  uint8_t* sA = smem;                                   // [2][a_bytes]
  uint8_t* sB = sA + 2 * (size_t)p.a_bytes;             // [nbslots][b_bytes]
  uint8_t* sOut = sB + (size_t)nbslots * p.b_bytes;     // [128][cout] 16-bit staging for the bulk store
  float* s_badd = (float*)(sOut + (size_t)p.out_bytes);  // [bn] bias + time-embedding row of the current image
  uint64_t* bars = (uint64_t*)(s_badd + 4 * 256);
  uint64_t* a_full = bars;            // [2]
  uint64_t* a_empty = bars + 2;       // [2]
  uint64_t* acc_full = bars + 4;      // [2]
  uint64_t* acc_empty = bars + 6;     // [2]
  uint64_t* w_full = bars + 8;        // [1] resident weights
  uint64_t* b_full = bars + 9;        // [nb]
  uint64_t* b_empty = b_full + 16;    // [nb]
  uint32_t* tmem_slot = (uint32_t*)(b_empty + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmB);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&a_full[s], 1); tc::mbar_init(&a_empty[s], 1);
      tc::mbar_init(&acc_full[s], 1); tc::mbar_init(&acc_empty[s], 128);
    }
    tc::mbar_init(w_full, 1);
    for (int s = 0; s < 16; ++s) { tc::mbar_init(&b_full[s], 1); tc::mbar_init(&b_empty[s], 1); }
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, p.tmem_cols);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  // One CTA per SM (smem-limited) and this is its only allocation, so the allocator returns column 0 / lane 0.
  // Treating it as the constant 0 keeps every tcgen05.mma operand in uniform registers: with a value loaded from
  // shared memory the compiler wraps EACH mma in an ELECT/R2UR.BROADCAST loop (~80 cycles per instruction measured),
  // which is more than the 24-48 cycles of tensor work of an N<=96 instruction.
  if (*tmem_slot != 0u) {
    if (threadIdx.x == 0) printf("libxrd: conv_halo expects TMEM base 0, got %u\n", *tmem_slot);
    __trap();
  }
  constexpr uint32_t tmem_base = 0u;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      if (p.wres) {
        tc::mbar_expect_tx(w_full, (uint32_t)nkb * p.b_bytes);
        for (int kb = 0; kb < nkb; ++kb) tc::tma_load_3d(sB + (size_t)kb * p.b_bytes, &tmB, w_full, 0, 0, kb);
      }
      uint32_t ai = 0, bi = 0;   // running A-buffer and B-slot counters
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        int r = t;
        const int txi = r % p.tiles_w; r /= p.tiles_w;
        const int tyi = r % p.tiles_h;
        const int img = r / p.tiles_h;
        const int w0 = txi * 128 - 1, h0 = tyi * p.th - 1;
        for (int c = 0; c < p.nchunk; ++c, ++ai) {
          const int st = ai & 1;
          tc::mbar_wait(&a_empty[st], ((ai >> 1) & 1) ^ 1);
          if ((p.dbg & 4) && ai >= 2) {
            tc::mbar_arrive(&a_full[st]);
          } else {
            tc::mbar_expect_tx(&a_full[st], (uint32_t)(p.th + 2) * kHaloPitch * 128u);
            tc::tma_load_4d(sA + (size_t)st * p.a_bytes, &tmA, &a_full[st], c * 64, w0, h0, img);
          }
          if (!p.wres) {
            for (int tap = 0; tap < 9; ++tap, ++bi) {
              const int bs = bi % p.nb;
              tc::mbar_wait(&b_empty[bs], ((bi / p.nb) & 1) ^ 1);
              tc::mbar_expect_tx(&b_full[bs], p.b_bytes);
              tc::tma_load_3d(sB + (size_t)bs * p.b_bytes, &tmB, &b_full[bs], 0, 0, c * 9 + tap);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      if (p.wres) tc::mbar_wait(w_full, 0);
      uint32_t ai = 0, bi = 0, ti = 0;
      for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++ti) {
        const int ab = p.nacc == 2 ? (ti & 1) : 0;
        const uint32_t use = p.nacc == 2 ? (ti >> 1) : ti;       // how many times this accumulator buffer was used before
        tc::mbar_wait(&acc_empty[ab], (use & 1) ^ 1);
        tc::tc_fence_after();
        const uint32_t acc0 = tmem_base + (uint32_t)(ab * p.th * p.bn);
        for (int c = 0; c < p.nchunk; ++c, ++ai) {
          const int st = ai & 1;
          tc::mbar_wait(&a_full[st], (ai >> 1) & 1);
          tc::tc_fence_after();
          const int ksteps = min(64, p.cin - c * 64) >> 4;
          const uint32_t abase = tc::smem_u32(sA + (size_t)st * p.a_bytes);
          for (int tap = 0; tap < 9; ++tap) {
            uint32_t bbase;
            int bs = 0;
            if (p.wres) {
              bbase = tc::smem_u32(sB + (size_t)(c * 9 + tap) * p.b_bytes);
            } else {
              bs = bi % p.nb;
              tc::mbar_wait(&b_full[bs], (bi / p.nb) & 1);
              tc::tc_fence_after();
              bbase = tc::smem_u32(sB + (size_t)bs * p.b_bytes);
              ++bi;
            }
            const int dy = tap / 3, dx = tap - dy * 3;
            const uint64_t bdesc = tc::umma_desc_sw128(bbase);
            for (int s = 0; s < p.th; ++s) {
              const uint32_t arow = abase + (uint32_t)((s + dy) * kHaloPitch + dx) * 128u;
              const uint64_t adesc = umma_desc_sw128_bo(arow, p.use_bo);
              if (!(p.dbg & 2) || (c | tap) == 0)
                for (int k = 0; k < ksteps; ++k)
                  tc::umma_f16(acc0 + (uint32_t)(s * p.bn), adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), p.idesc,
                               (c | tap | k) ? 1u : 0u);
            }
            if (!p.wres) tc::umma_commit(&b_empty[bs]);
          }
          tc::umma_commit(&a_empty[st]);
        }
        tc::umma_commit(&acc_full[ab]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    // Each warp owns 32 accumulator rows == 32 consecutive output pixels of an image row, i.e. ONE contiguous
    // 32*cout*2-byte span of the NHWC output.  TMEM -> registers (+bias +time-embedding row +residual) -> 16-bit ->
    // the warp's private staging rows in smem -> one bulk copy per (warp, image row).  No CTA-wide barrier.
    const int quad = warp & 3;
    T* yp = (T*)p.y;
    const T* rp = (const T*)p.resid;
    const uint32_t row_bytes = (uint32_t)p.cout * 2u;
    uint8_t* wstage = sOut + (size_t)quad * 32 * row_bytes;          // this warp's 32 staging rows
    uint8_t* my_row = wstage + (size_t)lane * row_bytes;
    float* badd = s_badd + quad * 256;                               // per-warp copy of bias + time-embedding row
    uint32_t ti = 0;
    int cur_img = -1;
    bool pending = false;
    for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++ti) {
      int r = t;
      const int txi = r % p.tiles_w; r /= p.tiles_w;
      const int tyi = r % p.tiles_h;
      const int img = r / p.tiles_h;
      const int ab = p.nacc == 2 ? (ti & 1) : 0;
      const uint32_t use = p.nacc == 2 ? (ti >> 1) : ti;
      if (img != cur_img) {
        __syncwarp();
        for (int cc = lane; cc < p.cout; cc += 32)
          badd[cc] = (p.bias ? __ldg(p.bias + cc) : 0.f) + (p.chan_add ? __ldg(p.chan_add + (int64_t)img * p.chan_add_bstride + cc) : 0.f);
        cur_img = img;
        __syncwarp();
      }
      tc::mbar_wait(&acc_full[ab], use & 1);
      tc::tc_fence_after();
      const int ow = txi * 128 + quad * 32 + lane;
      for (int s = 0; s < p.th; ++s) {
        const int oh = tyi * p.th + s;
        const bool row_ok = oh < p.H;
        const int64_t opix = ((int64_t)img * p.H + oh) * p.W + ow;
        if (lane == 0 && pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging rows reusable
        __syncwarp();
        const uint32_t tacc = tmem_base + (uint32_t)((ab * p.th + s) * p.bn) + ((uint32_t)(quad * 32) << 16);
        for (int c0 = 0; c0 < ((p.dbg & 8) ? 0 : p.bn); c0 += 32) {
          float v[32];
          if (c0 + 32 <= p.bn) {
            tc::tmem_ld32(tacc + (uint32_t)c0, v);
          } else {                                   // 16-column tail (bn is a multiple of 16)
            float v16[16];
            tc::tmem_ld16(tacc + (uint32_t)c0, v16);
#pragma unroll
            for (int j = 0; j < 16; ++j) { v[j] = v16[j]; v[16 + j] = 0.f; }
          }
#pragma unroll
          for (int h8 = 0; h8 < 4; ++h8) {
            const int co = c0 + h8 * 8;
            if (co < p.cout) {
              float r8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) r8[j] = v[h8 * 8 + j] + badd[co + j];
              if (rp && row_ok) {
                float q8[8];
                tc::ld8<T>(rp + opix * p.cout + co, q8);
#pragma unroll
                for (int j = 0; j < 8; ++j) r8[j] += q8[j];
              }
              uint4 pk;
              pk.x = tc::pack2<T>(r8[0], r8[1]); pk.y = tc::pack2<T>(r8[2], r8[3]);
              pk.z = tc::pack2<T>(r8[4], r8[5]); pk.w = tc::pack2<T>(r8[6], r8[7]);
              *reinterpret_cast<uint4*>(my_row + co * 2) = pk;
            }
          }
        }
        if (!(p.dbg & 16)) tc::fence_async_smem();
        __syncwarp();
        if (lane == 0 && row_ok && !(p.dbg & 1)) {
          const T* dst = yp + (((int64_t)img * p.H + oh) * p.W + (int64_t)txi * 128 + quad * 32) * p.cout;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"((uint64_t)dst), "r"(tc::smem_u32(wstage)),
                       "r"(32u * row_bytes)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          pending = true;
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&acc_empty[ab]);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

static int halo_env(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

bool conv_halo_supported(const Tens& x1, const Tens* x2, const ConvW& w, const ConvEpi& e) {
  static const int enabled = halo_env("XRD_CONV_HALO", 1);
  if (!enabled) return false;
  if (x2 || x1.dt == DT_F32) return false;
  if (!(w.kh == 3 && w.kw == 3 && w.stride == 1 && w.pad == 1) || w.d2s) return false;
  if (x1.c % 16 != 0 || w.cout % 8 != 0) return false;
  if (x1.w % 128 != 0) return false;
  if (e.in_scale || e.out_scale || e.act != ACT_NONE) return false;
  const int npad = (w.cout + 15) & ~15;
  if (npad > 256) return false;
  return true;
}

void conv_halo(Ctx& c, const Tens& x, ConvW& w, const ConvEpi& e, Tens& y) {
  XRD_REQUIRE(conv_halo_supported(x, nullptr, w, e), "conv_halo: unsupported configuration");
  XRD_REQUIRE(x.c == w.cin && y.n == x.n && y.h == x.h && y.w == x.w && y.c == w.cout && y.dt == x.dt, "conv_halo: shape mismatch");
  if (e.resid.p) XRD_REQUIRE(e.resid.dt == y.dt && e.resid.numel() == y.numel(), "conv_halo: residual mismatch");
  if (c.dry) return;
  if (!w.wtc[x.dt] || w.tc_c1 != x.c) conv_tc_pack(c.s, w, x.dt, x.c);
  ConvHaloP p;
  p.H = x.h; p.W = x.w; p.nimg = x.n;
  p.cin = x.c; p.nchunk = (x.c + 63) / 64;
  p.bn = w.tc_npad; p.cout = w.cout;
  const int nkb = p.nchunk * 9;
  XRD_REQUIRE(nkb == w.tc_nkb, "conv_halo: packed weights out of date");
  p.b_bytes = (uint32_t)p.bn * 128u;
  // rows per tile: as many accumulators as TMEM allows, at most 4, and the smem budget below
  int th = halo_env("XRD_HALO_TH", 0);
  if (th <= 0) th = p.bn <= 64 ? 2 : (p.bn <= 128 ? 2 : 1);
  while (th > 1 && th * p.bn > 512) --th;
  p.out_bytes = (uint32_t)(((size_t)128 * w.cout * 2 + 1023) & ~(size_t)1023);
  const size_t budget = 225 * 1024 - 2048 - p.out_bytes - 4096;
  for (;; --th) {
    p.th = th;
    p.a_bytes = (uint32_t)(((size_t)(th + 2) * kHaloPitch * 128 + 1023) & ~(size_t)1023);
    const size_t fixed = 2 * (size_t)p.a_bytes + 1024;
    p.wres = (fixed + (size_t)nkb * p.b_bytes <= budget) ? 1 : 0;
    if (halo_env("XRD_HALO_WRES", 1) == 0) p.wres = 0;
    p.nb = 0;
    if (!p.wres) {
      p.nb = (int)std::min<size_t>(9, (budget - std::min(budget, fixed)) / p.b_bytes);
      if (p.nb > 16) p.nb = 16;
    }
    if (p.wres || p.nb >= 2 || th == 1) break;
  }
  XRD_REQUIRE(p.wres || p.nb >= 2, "conv_halo: shared memory budget exceeded (cin=%d cout=%d)", x.c, w.cout);
  p.nacc = (2 * p.th * p.bn <= 512) ? 2 : 1;
  uint32_t cols = 32;
  while ((int)cols < p.nacc * p.th * p.bn) cols <<= 1;
  p.tmem_cols = cols;
  p.tiles_w = x.w / 128;
  p.tiles_h = cdiv(x.h, p.th);
  p.ntiles = p.tiles_w * p.tiles_h * x.n;
  p.dbg = halo_env("XRD_HALO_DBG", 0);
  p.use_bo = halo_env("XRD_HALO_BO", 0);   // measured on B200: the swizzle is a function of the absolute smem address; the field must stay 0
  p.idesc = tc::umma_idesc(128, p.bn, x.dt == DT_BF16 ? 1 : 0);
  p.bias = w.bias;
  p.chan_add = e.chan_add; p.chan_add_bstride = e.chan_add_bstride;
  p.resid = e.resid.p; p.y = y.p;

  alignas(64) CUtensorMap tmA, tmB;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)x.c, (cuuint64_t)x.w, (cuuint64_t)x.h, (cuuint64_t)x.n};
    const cuuint64_t strides[3] = {(cuuint64_t)x.c * 2, (cuuint64_t)x.w * x.c * 2, (cuuint64_t)x.h * x.w * x.c * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)kHaloPitch, (cuuint32_t)(p.th + 2), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = get_encode_tiled()(&tmA, tmap_dtype(x.dt), 4, x.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail(XRD_ERR_CUDA, "cuTensorMapEncodeTiled(halo activations) failed: %d", (int)r);
  }
  {
    const cuuint64_t dims[3] = {64, (cuuint64_t)p.bn, (cuuint64_t)nkb};
    const cuuint64_t strides[2] = {128, (cuuint64_t)p.bn * 128};
    const cuuint32_t box[3] = {64, (cuuint32_t)p.bn, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = get_encode_tiled()(&tmB, tmap_dtype(x.dt), 3, w.wtc[x.dt], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail(XRD_ERR_CUDA, "cuTensorMapEncodeTiled(halo weights) failed: %d", (int)r);
  }
  const size_t smem = 1024 + 2 * (size_t)p.a_bytes + (size_t)(p.wres ? nkb : p.nb) * p.b_bytes + p.out_bytes + 4096 + (9 + 32) * 8 + 16;
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  const int grid = std::min(p.ntiles, nsm);
  if (x.dt == DT_BF16) {
    static bool attr = false;
    if (!attr) { XRD_CUDA(cudaFuncSetAttribute(k_conv_halo<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr = true; }
    XRD_LAUNCH(c, k_conv_halo<__nv_bfloat16>, grid, kHaloThreads, smem, tmA, tmB, p);
  } else {
    static bool attr = false;
    if (!attr) { XRD_CUDA(cudaFuncSetAttribute(k_conv_halo<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr = true; }
    XRD_LAUNCH(c, k_conv_halo<__half>, grid, kHaloThreads, smem, tmA, tmB, p);
  }
}

}  // namespace xrd
