// conv_tc.cu -- implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
//   D[128 pixels, bn out-channels] (fp32, TMEM) += A[128 pixels, 64 in-channels] * B[bn, 64]^T   per K block
//
// * activations are NHWC bf16/f16; an A tile is one TMA 4-D box {64 ch, tw, th, 1 image} whose
//   coordinates are shifted by the filter tap, so zero padding, image borders and channel tails
//   (C not a multiple of 64) are all produced by TMA out-of-bounds zero fill -- no im2col buffer;
// * weights are pre-packed [kblock][cout][64] so a B tile is one TMA 3-D box;
// * both land in 128B-swizzled K-major shared-memory tiles that tcgen05.mma consumes directly;
// * warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM owner,
//   warps 2..5 = epilogue (tcgen05.ld -> bias/time-embedding/scale/residual/activation -> NHWC store);
// * up to two A sources are walked back to back (virtual channel concat), so torch.cat never exists.
#include "kernels.cuh"
#include "tc_common.cuh"

#include <algorithm>
#include <mutex>
#include <vector>
#include <stdlib.h>

namespace xrd {

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  });
  if (!fn) fail(XRD_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable from the CUDA driver");
  return fn;
}

struct ConvTcP {
  int Ho, Wo;               // output spatial size (GEMM pixel grid)
  int tiles_w, tiles_h;     // tiles per image
  int th, tw;               // tile = th x tw pixels (th*tw == 128)
  int stride, pad, kw, ntaps;
  int c_src0, c_src1;       // channels of the two sources
  int nchunk0, nchunk1;     // 64-channel chunks per source
  int nkb;                  // total K blocks
  int bn;                   // output-channel tile (<=256, multiple of 16)
  int cout;                 // valid output channels
  int stages;
  uint32_t idesc;
  uint32_t tmem_cols;
  // epilogue
  const float* bias;
  const float* chan_add; int chan_add_bstride;
  const float* out_scale;
  const void* resid;
  void* y;
  int act, d2s;
  double* stats;            // optional [nimg][8][2] sums of the stored output (stats kernel variant only)
};

constexpr int kTcThreads = 192;
constexpr int kABytes = 128 * 128;  // 128 pixels x 64 ch x 2 B

// BNS == 0: generic epilogue (runtime bn).  BNS in {48,96,144,192}: bn == cout == BNS, no depth-to-space; the
// epilogue additionally accumulates the GroupNorm partial sums (8 groups of BNS/8 channels) of what it stores.
template <typename T, int BNS>
__global__ void __launch_bounds__(kTcThreads)
k_conv_tc(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
          const __grid_constant__ CUtensorMap tmB, const ConvTcP p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][A 16 KB][B bn*128] then barriers
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t b_bytes = (uint32_t)p.bn * 128u;
  const uint32_t stage_bytes = kABytes + ((b_bytes + 1023u) & ~1023u);
  uint64_t* full_bar = (uint64_t*)(smem + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full = empty_bar + p.stages;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile coordinates
  int t = blockIdx.x;
  const int tx_ = t % p.tiles_w; t /= p.tiles_w;
  const int ty_ = t % p.tiles_h;
  const int img = t / p.tiles_h;
  const int ow0 = tx_ * p.tw, oh0 = ty_ * p.th;
  const int n0 = blockIdx.y * p.bn;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA0);
    if (p.nchunk1) tc::tma_prefetch_desc(&tmA1);
    tc::tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(tmem_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, p.tmem_cols);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const int nkb0 = p.nchunk0 * p.ntaps;
      for (int kb = 0; kb < p.nkb; ++kb) {
        tc::mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + (size_t)stage * stage_bytes;
        uint8_t* sb = sa + kABytes;
        tc::mbar_expect_tx(&full_bar[stage], kABytes + b_bytes);
        const bool second = kb >= nkb0;
        const int k2 = second ? kb - nkb0 : kb;
        const int chunk = k2 / p.ntaps, tap = k2 - chunk * p.ntaps;
        const int ky = tap / p.kw, kx = tap - ky * p.kw;
        const int iw = ow0 * p.stride - p.pad + kx, ih = oh0 * p.stride - p.pad + ky;
        tc::tma_load_4d(sa, second ? &tmA1 : &tmA0, &full_bar[stage], chunk * 64, iw, ih, img);
        tc::tma_load_3d(sb, &tmB, &full_bar[stage], 0, n0, kb);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The warp stays converged; one elected lane issues with compile-time k offsets (issuing from an `if (lane == 0)`
    // branch makes ptxas wrap every tcgen05.mma in an ELECT/BRA loop).
    int stage = 0; uint32_t phase = 0;
    const int nkb0 = p.nchunk0 * p.ntaps;
    for (int kb = 0; kb < p.nkb; ++kb) {
      tc::mbar_wait(&full_bar[stage], phase);
      tc::tc_fence_after();
      const bool second = kb >= nkb0;
      const int k2 = second ? kb - nkb0 : kb;
      const int chunk = k2 / p.ntaps;
      const int cs = second ? p.c_src1 : p.c_src0;
      const int ksteps = min(64, cs - chunk * 64) >> 4;
      if (tc::elect_one()) {
        const uint32_t sa = tc::smem_u32(smem + (size_t)stage * stage_bytes);
        const uint64_t adesc = tc::umma_desc_sw128(sa);
        const uint64_t bdesc = tc::umma_desc_sw128(sa + kABytes);
        // +32 B per 16-element K step inside the 128 B swizzle row (start-address field is in 16 B units)
        tc::umma_f16(tmem_base, adesc, bdesc, p.idesc, kb ? 1u : 0u);
        if (ksteps > 1) tc::umma_f16(tmem_base, adesc + 2, bdesc + 2, p.idesc, 1u);
        if (ksteps > 2) tc::umma_f16(tmem_base, adesc + 4, bdesc + 4, p.idesc, 1u);
        if (ksteps > 3) tc::umma_f16(tmem_base, adesc + 6, bdesc + 6, p.idesc, 1u);
        tc::umma_commit(&empty_bar[stage]);   // frees the smem stage once these MMAs retire
        if (kb == p.nkb - 1) tc::umma_commit(tmem_full);   // accumulator complete
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    tc::mbar_wait(tmem_full, 0);
    tc::tc_fence_after();
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int m = quad * 32 + lane;            // accumulator row == pixel within the tile
    const int dy = m / p.tw, dx = m - dy * p.tw;
    const int oh = oh0 + dy, ow = ow0 + dx;
    const bool pix_ok = oh < p.Ho && ow < p.Wo;
    const int cf = p.d2s ? (p.cout >> 2) : p.cout;
    const int64_t opix = ((int64_t)img * p.Ho + oh) * p.Wo + ow;
    T* yp = (T*)p.y;
    const T* rp = (const T*)p.resid;
    if (BNS > 0) {
      constexpr int CPG = BNS > 0 ? BNS / 8 : 1;
      float gs[8], gq[8];
#pragma unroll
      for (int g = 0; g < 8; ++g) { gs[g] = 0.f; gq[g] = 0.f; }
#pragma unroll
      for (int c0 = 0; c0 < BNS; c0 += 16) {
        float v[16];
        tc::tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8) {
          const int co = c0 + h8 * 8;
          float r8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float tv = v[h8 * 8 + j];
            if (p.bias) tv += __ldg(p.bias + co + j);
            if (p.chan_add) tv += __ldg(p.chan_add + (int64_t)img * p.chan_add_bstride + co + j);
            if (p.out_scale) tv *= __ldg(p.out_scale + co + j);
            r8[j] = tv;
          }
          if (pix_ok) {
            const int64_t o = opix * BNS + co;
            if (rp) {
              float q8[8];
              tc::ld8<T>(rp + o, q8);
#pragma unroll
              for (int j = 0; j < 8; ++j) r8[j] += q8[j];
            }
            if (p.act != ACT_NONE) {
#pragma unroll
              for (int j = 0; j < 8; ++j) r8[j] = act_apply(r8[j], p.act);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int g = (co + j) / CPG;
              gs[g] += r8[j];
              gq[g] = fmaf(r8[j], r8[j], gq[g]);
            }
            uint4 pk;
            pk.x = tc::pack2<T>(r8[0], r8[1]); pk.y = tc::pack2<T>(r8[2], r8[3]);
            pk.z = tc::pack2<T>(r8[4], r8[5]); pk.w = tc::pack2<T>(r8[6], r8[7]);
            *reinterpret_cast<uint4*>(yp + o) = pk;
          }
        }
      }
      if (p.stats) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            gs[g] += __shfl_xor_sync(0xffffffffu, gs[g], o);
            gq[g] += __shfl_xor_sync(0xffffffffu, gq[g], o);
          }
        }
        if (lane < 16) {
          float sv = 0.f;
#pragma unroll
          for (int g = 0; g < 8; ++g) { if (lane == 2 * g) sv = gs[g]; if (lane == 2 * g + 1) sv = gq[g]; }
          atomicAdd(p.stats + (size_t)img * 16 + lane, (double)sv);
        }
      }
    } else
    for (int c0 = 0; c0 < p.bn; c0 += 16) {
      float v[16];
      __syncwarp();
      tc::tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
      const int col = n0 + c0;
      if (pix_ok && col < p.cout) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int co = col + j;   // cout is a multiple of 8; columns >= cout are never stored
          if (co < p.cout) {
            float tv = v[j];
            if (p.bias) tv += __ldg(p.bias + co);
            if (p.chan_add) tv += __ldg(p.chan_add + (int64_t)img * p.chan_add_bstride + co);
            if (p.out_scale) tv *= __ldg(p.out_scale + co);
            v[j] = tv;
          }
        }
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8) {
          const int co = col + h8 * 8;
          if (co < p.cout) {
            int64_t o;
            if (p.d2s) {
              const int q = co / cf, cc = co - q * cf;
              o = (((int64_t)img * (2 * p.Ho) + (2 * oh + (q >> 1))) * (2 * p.Wo) + (2 * ow + (q & 1))) * cf + cc;
            } else {
              o = opix * p.cout + co;
            }
            float r8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) r8[j] = v[h8 * 8 + j];
            if (rp) {
              float a4[4], b4[4];
              ld4<T>(rp + o, a4);
              ld4<T>(rp + o + 4, b4);
#pragma unroll
              for (int j = 0; j < 4; ++j) { r8[j] += a4[j]; r8[4 + j] += b4[j]; }
            }
            if (p.act != ACT_NONE) {
#pragma unroll
              for (int j = 0; j < 8; ++j) r8[j] = act_apply(r8[j], p.act);
            }
            float lo[4] = {r8[0], r8[1], r8[2], r8[3]}, hi[4] = {r8[4], r8[5], r8[6], r8[7]};
            st4<T>(yp + o, lo);
            st4<T>(yp + o + 4, hi);
          }
        }
      }
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// -------------------------------------------------------------------------------------------------
// weight packing for the tensor-core path: [kb][npad][64] with K blocks ordered (source, chunk, tap)
// -------------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_pack_tc(const float* __restrict__ w /*[ntaps][cin][cout]*/, T* __restrict__ out, int ntaps, int cin, int cout,
                          int c1, int npad, int nkb) {
  const int c2 = cin - c1;
  const int nchunk0 = (c1 + 63) / 64;
  const int64_t total = (int64_t)nkb * npad * 64;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i & 63);
    int64_t r = i >> 6;
    const int n = (int)(r % npad);
    const int kb = (int)(r / npad);
    const int nkb0 = nchunk0 * ntaps;
    const bool second = kb >= nkb0;
    const int k2 = second ? kb - nkb0 : kb;
    const int chunk = k2 / ntaps, tap = k2 - chunk * ntaps;
    const int cs = second ? c2 : c1;
    const int cl = chunk * 64 + k;
    float v = 0.f;
    if (cl < cs && n < cout) {
      const int ci = second ? c1 + cl : cl;
      v = w[((int64_t)tap * cin + ci) * cout + n];
    }
    stf<T>(out + i, v);
    // second copy, pre-swizzled exactly as TMA's 128B swizzle would place the block in (1024-aligned) shared memory: the
    // 16-byte chunk index is XORed with the row index mod 8.  A streamed weight block can then be fetched with ONE 1-D bulk
    // copy instead of one TMA request per 128-byte row (the TMA request rate, ~6.4 cycles per row per SM, bounds the
    // streamed-weight convolutions).
    const int64_t blk = (int64_t)kb * npad * 64;
    stf<T>(out + total + blk + (int64_t)n * 64 + ((((k >> 3) ^ (n & 7)) << 3) | (k & 7)), v);
  }
}

void conv_tc_pack(cudaStream_t s, ConvW& w, DType dt, int c1) {
  XRD_REQUIRE(dt != DT_F32, "conv_tc_pack: tensor-core weights are 16-bit");
  const int ntaps = w.kh * w.kw;
  const int c2 = w.cin - c1;
  const int nkb = ((c1 + 63) / 64 + (c2 + 63) / 64) * ntaps;
  const int npad = (w.cout + 15) & ~15;
  if (w.wtc[dt] && w.tc_c1 == c1) return;
  if (w.tc_c1 != c1) {
    for (int i = 0; i < 3; ++i)
      if (w.wtc[i]) { cudaFree(w.wtc[i]); w.wtc[i] = nullptr; }
  }
  size_t bytes = (size_t)nkb * npad * 64 * 2;
  XRD_CUDA(cudaMalloc(&w.wtc[dt], 2 * bytes));     // plain copy (tensor-map loads) + pre-swizzled copy (1-D bulk loads)
  w.tc_c1 = c1; w.tc_nkb = nkb; w.tc_npad = npad;
  int64_t total = (int64_t)nkb * npad * 64;
  int blocks = (int)std::min<int64_t>(cdiv64(total, 256), 148 * 16);
  if (dt == DT_BF16)
    k_pack_tc<__nv_bfloat16><<<blocks, 256, 0, s>>>(w.w, (__nv_bfloat16*)w.wtc[dt], ntaps, w.cin, w.cout, c1, npad, nkb);
  else
    k_pack_tc<__half><<<blocks, 256, 0, s>>>(w.w, (__half*)w.wtc[dt], ntaps, w.cin, w.cout, c1, npad, nkb);
  XRD_CUDA(cudaPeekAtLastError());
}

bool conv_tc_stats_supported(const ConvW& w) {
  return !w.d2s && (w.cout == 48 || w.cout == 96 || w.cout == 144 || w.cout == 192);
}

bool conv_tc_supported(const Tens& x1, const Tens* x2, const ConvW& w, const ConvEpi& e) {
  if (x1.dt == DT_F32) return false;
  if (e.stats_out && !conv_tc_stats_supported(w)) return false;
  if (x1.c % 16 != 0 || (x2 && x2->c % 16 != 0)) return false;
  if (w.cout % 8 != 0) return false;
  if (e.in_scale) return false;
  const bool k3 = w.kh == 3 && w.kw == 3 && w.pad == 1 && (w.stride == 1 || w.stride == 2);
  const bool k1 = w.kh == 1 && w.kw == 1 && w.pad == 0 && w.stride == 1;
  const bool k2 = w.kh == 2 && w.kw == 2 && w.pad == 0 && w.stride == 2;
  if (!(k3 || k1 || k2)) return false;
  if (w.d2s && (w.cout / 4) % 8 != 0) return false;
  return true;
}

static void encode_act_map(CUtensorMap* m, const Tens& x, int box_w, int box_h, int stride) {
  const cuuint64_t dims[4] = {(cuuint64_t)x.c, (cuuint64_t)x.w, (cuuint64_t)x.h, (cuuint64_t)x.n};
  const cuuint64_t strides[3] = {(cuuint64_t)x.c * 2, (cuuint64_t)x.w * x.c * 2, (cuuint64_t)x.h * x.w * x.c * 2};
  const cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = get_encode_tiled()(m, tmap_dtype(x.dt), 4, x.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    fail(XRD_ERR_CUDA, "cuTensorMapEncodeTiled(activations c=%d w=%d h=%d n=%d box=%dx%d stride=%d) failed: %d", x.c, x.w, x.h, x.n,
         box_w, box_h, stride, (int)r);
}

void conv_tc(Ctx& c, const Tens& x1, const Tens* x2, ConvW& w, const ConvEpi& e, Tens& y) {
  XRD_REQUIRE(conv_tc_supported(x1, x2, w, e), "conv_tc: unsupported configuration");
  const int c1 = x1.c, c2 = x2 ? x2->c : 0;
  XRD_REQUIRE(c1 + c2 == w.cin, "conv_tc: channels %d+%d != %d", c1, c2, w.cin);
  if (x2) XRD_REQUIRE(x2->n == x1.n && x2->h == x1.h && x2->w == x1.w && x2->dt == x1.dt, "conv_tc: source mismatch");
  ConvTcP p;
  p.Ho = (x1.h + 2 * w.pad - w.kh) / w.stride + 1;
  p.Wo = (x1.w + 2 * w.pad - w.kw) / w.stride + 1;
  const int cf = w.d2s ? w.cout / 4 : w.cout;
  const int eh = w.d2s ? 2 * p.Ho : p.Ho, ew = w.d2s ? 2 * p.Wo : p.Wo;
  XRD_REQUIRE(y.n == x1.n && y.h == eh && y.w == ew && y.c == cf && y.dt == x1.dt, "conv_tc: output shape/dtype mismatch");
  if (e.resid.p) XRD_REQUIRE(e.resid.dt == y.dt && e.resid.numel() == y.numel(), "conv_tc: residual mismatch");
  if (c.dry) return;
  if (!w.wtc[x1.dt] || w.tc_c1 != c1) conv_tc_pack(c.s, w, x1.dt, c1);

  p.tw = p.Wo >= 16 ? 16 : 8;
  p.th = 128 / p.tw;
  p.tiles_w = cdiv(p.Wo, p.tw);
  p.tiles_h = cdiv(p.Ho, p.th);
  p.stride = w.stride; p.pad = w.pad; p.kw = w.kw; p.ntaps = w.kh * w.kw;
  p.c_src0 = c1; p.c_src1 = c2;
  p.nchunk0 = (c1 + 63) / 64; p.nchunk1 = (c2 + 63) / 64;
  p.nkb = (p.nchunk0 + p.nchunk1) * p.ntaps;
  XRD_REQUIRE(p.nkb == w.tc_nkb, "conv_tc: packed weights out of date");
  const int npad = w.tc_npad;
  // output-channel tile: the largest multiple of 16 that divides npad and is <= 256
  int bn = 0;
  for (int cand = std::min(npad, 256); cand >= 16; cand -= 16)
    if (npad % cand == 0) { bn = cand; break; }
  XRD_REQUIRE(bn > 0, "conv_tc: no output tile for cout=%d", w.cout);
  p.bn = bn;
  p.cout = w.cout;
  p.idesc = tc::umma_idesc(128, bn, x1.dt == DT_BF16 ? 1 : 0);
  uint32_t cols = 32;
  while ((int)cols < bn) cols <<= 1;
  p.tmem_cols = cols;
  const uint32_t b_bytes = (uint32_t)bn * 128u;
  const uint32_t stage_bytes = kABytes + ((b_bytes + 1023u) & ~1023u);
  // Several CTAs per SM overlap one tile's epilogue/prologue with another's main loop (TMEM: ctas * cols <= 512).
  static const int env_ctas = getenv("XRD_TC_CTAS") ? atoi(getenv("XRD_TC_CTAS")) : 0;
  static const int env_stages = getenv("XRD_TC_STAGES") ? atoi(getenv("XRD_TC_STAGES")) : 0;
  int ctas = env_ctas > 0 ? env_ctas : (bn <= 64 ? 3 : 2);
  while (ctas > 1 && (int)cols * ctas > 512) --ctas;
  int stages = (int)((220 * 1024 / ctas - 2048) / stage_bytes);
  if (env_stages > 0) stages = env_stages;
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  if (stages > p.nkb) stages = std::max(1, p.nkb);
  p.stages = stages;
  p.bias = w.bias;
  p.chan_add = e.chan_add; p.chan_add_bstride = e.chan_add_bstride;
  p.out_scale = e.out_scale;
  p.resid = e.resid.p; p.y = y.p; p.act = e.act; p.d2s = w.d2s;

  alignas(64) CUtensorMap tmA0, tmA1, tmB;
  encode_act_map(&tmA0, x1, p.tw * w.stride, p.th * w.stride, w.stride);
  if (x2) encode_act_map(&tmA1, *x2, p.tw * w.stride, p.th * w.stride, w.stride);
  else tmA1 = tmA0;
  {
    const cuuint64_t dims[3] = {64, (cuuint64_t)npad, (cuuint64_t)p.nkb};
    const cuuint64_t strides[2] = {128, (cuuint64_t)npad * 128};
    const cuuint32_t box[3] = {64, (cuuint32_t)bn, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = get_encode_tiled()(&tmB, tmap_dtype(x1.dt), 3, w.wtc[x1.dt], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail(XRD_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
  }
  size_t smem = 1024 + (size_t)stages * stage_bytes + (2 * stages + 1) * 8 + 16;
  dim3 grid(p.tiles_w * p.tiles_h * x1.n, npad / bn);
  // statistics variant: whole cout in one N tile, plain NHWC store
  const bool want_stats = e.stats_out != nullptr;
  if (want_stats) XRD_REQUIRE(!w.d2s && bn == w.cout && (bn == 48 || bn == 96 || bn == 144 || bn == 192), "conv_tc: no statistics epilogue for cout=%d", w.cout);
  p.stats = e.stats_out;
  auto launch = [&](auto kern) {
    ensure_dyn_smem(kern, 227 * 1024);
    XRD_LAUNCH(c, kern, grid, kTcThreads, smem, tmA0, tmA1, tmB, p);
  };
  if (x1.dt == DT_BF16) {
    using T = __nv_bfloat16;
    if (!want_stats) launch(k_conv_tc<T, 0>);
    else if (bn == 48) launch(k_conv_tc<T, 48>);
    else if (bn == 96) launch(k_conv_tc<T, 96>);
    else if (bn == 144) launch(k_conv_tc<T, 144>);
    else launch(k_conv_tc<T, 192>);
  } else {
    using T = __half;
    if (!want_stats) launch(k_conv_tc<T, 0>);
    else if (bn == 48) launch(k_conv_tc<T, 48>);
    else if (bn == 96) launch(k_conv_tc<T, 96>);
    else if (bn == 144) launch(k_conv_tc<T, 144>);
    else launch(k_conv_tc<T, 192>);
  }
}

}  // namespace xrd
