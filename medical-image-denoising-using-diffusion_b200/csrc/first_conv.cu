// first_conv.cu -- the UNet's in_conv (HYB:335, 362-363: Conv2d(2, mc, 3, padding=1) on cat([x, condition])) in ONE pass over
// HBM: the two fp32 planes in, the 16-bit NHWC activation (+ its GroupNorm sums) out.
//
// Round 1/2 ran this layer as k_im2col_3x3_2ch (fp32 planes -> [pixel][32] 16-bit rows, 64 B per pixel) followed by the
// persistent tcgen05 1x1 GEMM: 33 MB read, 268 MB written, 268 MB read again, 403 MB written for a batch of sixteen 512x512
// images (227 us).  The contraction is tiny (K = 18 real taps x channels, N = 48), so it does not need tcgen05 operands in
// shared memory at all: here every warp forms the im2col A fragments of warp-level mma.sync tiles straight from a 16-bit
// copy of the input tile -- one 32-bit shared-memory word per pixel holds (x, condition), which is exactly one k-pair
// (k = 2*tap + channel) of an A-fragment register, so an A register is ONE shared load at the tap's offset -- keeps the
// whole weight matrix as B fragments in registers, and writes the result through a per-warp staging row in 16-byte,
// fully coalesced stores.  Traffic: 33 MB + 403 MB, the layer's algorithmic bytes.
//
// Arithmetic is that of the path it replaces: operands rounded to the mode's 16-bit type, fp32 accumulation, fp32 bias,
// GroupNorm sums of the unrounded fp32 results (fp32 partial sums per block, fp64 atomics), saturating 16-bit store.
#include "kernels.cuh"
#include "tc_common.cuh"

namespace xrd {

namespace {

template <typename T> struct WarpMma;
template <> struct WarpMma<__half> {
  // c = a x b + (i0, i1, i0, i1): the accumulator starts from the bias pair of this thread's two columns
  static __device__ __forceinline__ void k16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, float i0, float i1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%10,%11};"
        : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(i0), "f"(i1));
  }
  static __device__ __forceinline__ void k8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(b0));
  }
};
template <> struct WarpMma<__nv_bfloat16> {
  // c = a x b + (i0, i1, i0, i1): the accumulator starts from the bias pair of this thread's two columns
  static __device__ __forceinline__ void k16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, float i0, float i1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%10,%11};"
        : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(i0), "f"(i1));
  }
  static __device__ __forceinline__ void k8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(b0));
  }
};

constexpr int kFcTH = 8;      // output rows per block: one per warp
constexpr int kFcTW = 256;    // output columns per block: sixteen 16-pixel MMA tiles per warp (the block's prologue -- weights,
                              // input staging, one barrier -- is ~1 us of exposed latency: ncu r02 showed it as the top stall)
constexpr int kFcSR = 268;    // words per staged input row (>= TW + 2; 268 % 32 == 12 keeps the four taps a quad reads in
                              // one A-fragment load on distinct banks: offsets {0, 1, 2, 12} and {13, 14, 24, 25} + pixel 0..7)
constexpr int kFcStage = ((kFcTH + 2) * (kFcTW + 2) + 255) / 256;   // staging elements per thread

// GEMM view: M = pixels, K = 32 with k = 2*tap + channel (tap = ky*3 + kx; k >= 18 is zero padding), N = COUT.
// mma.m16n8k16 fragments (g = lane / 4, t = lane % 4):  A: a0 = (pixel g, k 2t..2t+1), a1 = (pixel g+8, same k), a2 / a3 = the
// same pixels at k + 8;  B: b0 = (k 2t..2t+1, n g), b1 = k + 8;  C: c0,c1 = (pixel g, n 2t..2t+1), c2,c3 = pixel g+8.
// k-step 0 (m16n8k16) covers taps 0..7: thread t reads tap t and tap t+4; the m16n8k8 tail covers tap 8 (k = 16, 17: t == 0).
template <typename T, int COUT>
__global__ void __launch_bounds__(256, 2) k_first_conv_mma(const float* __restrict__ xa, const float* __restrict__ xb, const float* __restrict__ w,
                                                           const float* __restrict__ bias, T* __restrict__ y, double* __restrict__ stats, int H,
                                                           int W) {
  constexpr int NT = COUT / 8;              // 8-column MMA tiles
  constexpr int CPG = COUT / 8;             // channels per GroupNorm group
  constexpr int OSTR = COUT * 2 + 16;       // bytes per staged output pixel (112 for 48 channels: conflict-free 4-byte writes)
  static_assert(COUT % 8 == 0 && OSTR % 16 == 0, "first conv: COUT must be a multiple of 8");
  __shared__ uint32_t s_in[(kFcTH + 2) * kFcSR];
  __shared__ __align__(16) uint8_t s_out[kFcTH][16 * OSTR];
  __shared__ float s_sum[2 * COUT];

  const int n = blockIdx.z, y0 = blockIdx.y * kFcTH, x0 = blockIdx.x * kFcTW;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const float* pa = xa + (int64_t)n * H * W;
  const float* pb = xb + (int64_t)n * H * W;

  // input tile with a one-pixel halo, zero outside the image (padding = 1), rounded to the operand type.  All loads of a thread
  // are issued before the first is converted (one exposed latency, not kFcStage of them).
  {
    float va[kFcStage], vb[kFcStage];
#pragma unroll
    for (int k = 0; k < kFcStage; ++k) {
      const int i = tid + k * 256;
      const int r = i / (kFcTW + 2), cc = i - r * (kFcTW + 2);
      const int yy = y0 + r - 1, xx = x0 + cc - 1;
      const bool ok = i < (kFcTH + 2) * (kFcTW + 2) && yy >= 0 && yy < H && xx >= 0 && xx < W;
      va[k] = ok ? __ldg(pa + (int64_t)yy * W + xx) : 0.f;
      vb[k] = ok ? __ldg(pb + (int64_t)yy * W + xx) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < kFcStage; ++k) {
      const int i = tid + k * 256;
      const int r = i / (kFcTW + 2), cc = i - r * (kFcTW + 2);
      if (i < (kFcTH + 2) * (kFcTW + 2)) s_in[r * kFcSR + cc] = tc::pack2<T>(va[k], vb[k]);
    }
  }
  if (tid < 2 * COUT) s_sum[tid] = 0.f;

  // the whole weight matrix as B fragments: w is [18][COUT] fp32, row k = 2*tap + channel
  uint32_t b0[NT], b1[NT], b2[NT];
  float bs[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int col = nt * 8 + g;
    b0[nt] = tc::pack2<T>(__ldg(w + (2 * t) * COUT + col), __ldg(w + (2 * t + 1) * COUT + col));
    b1[nt] = tc::pack2<T>(__ldg(w + (2 * t + 8) * COUT + col), __ldg(w + (2 * t + 9) * COUT + col));
    b2[nt] = t == 0 ? tc::pack2<T>(__ldg(w + 16 * COUT + col), __ldg(w + 17 * COUT + col)) : 0u;
    bs[nt][0] = bias ? __ldg(bias + nt * 8 + 2 * t) : 0.f;
    bs[nt][1] = bias ? __ldg(bias + nt * 8 + 2 * t + 1) : 0.f;
  }
  __syncthreads();

  float cs[NT][2], cq[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) { cs[nt][0] = cs[nt][1] = cq[nt][0] = cq[nt][1] = 0.f; }

  const int oy = y0 + warp;
  const int o0 = (t / 3) * kFcSR + (t % 3);                 // tap t
  const int o1 = ((t + 4) / 3) * kFcSR + ((t + 4) % 3);     // tap t + 4
  const int o2 = 2 * kFcSR + 2;                             // tap 8
  if (oy < H) {
    uint8_t* so = s_out[warp];
    for (int mt = 0; mt < kFcTW / 16; ++mt) {
      const int c0 = mt * 16;
      if (x0 + c0 >= W) break;                              // warp-uniform
      const uint32_t* base = s_in + warp * kFcSR + c0 + g;
      const uint32_t a[4] = {base[o0], base[8 + o0], base[o1], base[8 + o1]};
      const uint32_t a8lo = t == 0 ? base[o2] : 0u, a8hi = t == 0 ? base[8 + o2] : 0u;
      float c[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        WarpMma<T>::k16(c[nt], a, b0[nt], b1[nt], bs[nt][0], bs[nt][1]);
        WarpMma<T>::k8(c[nt], a8lo, a8hi, b2[nt]);
      }
      if (stats) {
        if (x0 + c0 + 16 <= W) {                            // warp-uniform: a full tile needs no per-pixel predicate
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            cs[nt][0] += c[nt][0] + c[nt][2]; cq[nt][0] = fmaf(c[nt][0], c[nt][0], fmaf(c[nt][2], c[nt][2], cq[nt][0]));
            cs[nt][1] += c[nt][1] + c[nt][3]; cq[nt][1] = fmaf(c[nt][1], c[nt][1], fmaf(c[nt][3], c[nt][3], cq[nt][1]));
          }
        } else {
          const bool v0 = x0 + c0 + g < W, v1 = x0 + c0 + g + 8 < W;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            if (v0) {
              cs[nt][0] += c[nt][0]; cq[nt][0] = fmaf(c[nt][0], c[nt][0], cq[nt][0]);
              cs[nt][1] += c[nt][1]; cq[nt][1] = fmaf(c[nt][1], c[nt][1], cq[nt][1]);
            }
            if (v1) {
              cs[nt][0] += c[nt][2]; cq[nt][0] = fmaf(c[nt][2], c[nt][2], cq[nt][0]);
              cs[nt][1] += c[nt][3]; cq[nt][1] = fmaf(c[nt][3], c[nt][3], cq[nt][1]);
            }
          }
        }
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        *reinterpret_cast<uint32_t*>(so + g * OSTR + nt * 16 + t * 4) = tc::pack2<T>(c[nt][0], c[nt][1]);
        *reinterpret_cast<uint32_t*>(so + (g + 8) * OSTR + nt * 16 + t * 4) = tc::pack2<T>(c[nt][2], c[nt][3]);
      }
      __syncwarp();
      // 16 pixels x COUT channels = 16 * NT chunks of 16 bytes.  Lane -> (pixel j % 16, chunk j / 16): a quarter-warp reads eight
      // consecutive pixels of one chunk column (stride 112 B = 7 x 16 B: eight distinct 16-byte bank groups), and the two half-warps
      // of one store instruction write chunks 2i and 2i+1 of sixteen pixels = sixteen whole 32-byte sectors
      T* yrow = y + (((int64_t)n * H + oy) * W + x0 + c0) * COUT;
#pragma unroll
      for (int j = lane; j < 16 * NT; j += 32) {
        const int px = j & 15, ch = j >> 4;
        if (x0 + c0 + px < W)
          *reinterpret_cast<uint4*>(yrow + px * COUT + ch * 8) = *reinterpret_cast<const uint4*>(so + px * OSTR + ch * 16);
      }
      __syncwarp();
    }
  }

  if (stats) {   // kernel argument: uniform
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float s = cs[nt][j], q = cq[nt][j];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
        if (g == 0) { atomicAdd(&s_sum[nt * 8 + 2 * t + j], s); atomicAdd(&s_sum[COUT + nt * 8 + 2 * t + j], q); }
      }
    }
    __syncthreads();
    if (tid < 16) {                                          // [image][group][sum, sum of squares]
      const int grp = tid >> 1, which = tid & 1;
      float v = 0.f;
#pragma unroll
      for (int i = 0; i < CPG; ++i) v += s_sum[which * COUT + grp * CPG + i];
      atomicAdd(stats + (int64_t)n * 16 + tid, (double)v);
    }
  }
}

}  // namespace

bool first_conv_mma_supported(const Tens& y, const ConvW& w) {
  static const int enabled = getenv("XRD_FIRST_MMA") ? atoi(getenv("XRD_FIRST_MMA")) : 1;
  return enabled && y.dt != DT_F32 && w.kh == 3 && w.kw == 3 && w.stride == 1 && w.pad == 1 && w.cin == 2 && w.cout == 48 && y.c == 48 && !w.d2s;
}

// y = conv3x3(cat[a, b]) + bias; a, b: (N,H,W) fp32 planes; w.w: [9][2][COUT] fp32 (ConvW's CUDA-core layout).
// stats (nullable, zero on entry): [N][8][2] += GroupNorm sums of y.
void first_conv_mma(Ctx& c, const float* a, const float* b, const ConvW& w, Tens& y, double* stats) {
  XRD_REQUIRE(first_conv_mma_supported(y, w), "first_conv_mma: unsupported configuration");
  dim3 grid(cdiv(y.w, kFcTW), cdiv(y.h, kFcTH), y.n);
  if (y.dt == DT_F16)
    XRD_LAUNCH(c, (k_first_conv_mma<__half, 48>), grid, 256, 0, a, b, w.w, w.bias, (__half*)y.p, stats, y.h, y.w);
  else
    XRD_LAUNCH(c, (k_first_conv_mma<__nv_bfloat16, 48>), grid, 256, 0, a, b, w.w, w.bias, (__nv_bfloat16*)y.p, stats, y.h, y.w);
}

}  // namespace xrd
