// kernels_edge.cu -- the HBM-bound kernels at the two ends of a UNet evaluation and at its three 2x upsample sites,
// rewritten so that each one moves its tensor exactly once at 16 bytes per access and leaves the CUDA-core work
// (a few hundred FMAs per pixel) far below the memory time:
//   * conv_smallcin2 : 3x3 conv from <=4 fp32 input planes to COUT channels (cat[x, condition] -> 48: HYB:362-363;
//                      NAFNet intro 1 -> 32; router 1 -> 32; fusion 3 -> 48).  thread = one pixel, all COUT outputs,
//                      weights broadcast from shared memory; optional GroupNorm sums of the output.
//   * conv_cout1_v2  : GroupNorm + SiLU + 3x3 conv to ONE channel (+ sampler update, HYB:353-357,410-416; NAFNet ending).
//                      Each input pixel is read and activated once and reduced to its nine per-tap dot products; the
//                      3x3 gather then runs on that 9-plane fp32 tile in shared memory.
//   * upsample2x_stats: bilinear 2x (align_corners=False, HYB:381-382) fused with the GroupNorm sums of its output.
#include "kernels.cuh"
#include <stdlib.h>

namespace xrd {

void conv_cout1_tiled(Ctx& c, const Cout1Args& a);   // kernels_fused.cu (previous version, generic C)

namespace {

template <typename T> __device__ __forceinline__ void ld8f(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void ld8f<__half>(const __half* p, float (&v)[8]) {
  uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
  const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <> __device__ __forceinline__ void ld8f<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <> __device__ __forceinline__ void ld8f<float>(const float* p, float (&v)[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <typename T> __device__ __forceinline__ void st8f(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void st8f<__half>(__half* p, const float (&v)[8]) {
  uint4 h;
  h.x = pack_h2_sat(v[0], v[1]); h.y = pack_h2_sat(v[2], v[3]); h.z = pack_h2_sat(v[4], v[5]); h.w = pack_h2_sat(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = h;
}
template <> __device__ __forceinline__ void st8f<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
  __nv_bfloat162 h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<uint4*>(h);
}
template <> __device__ __forceinline__ void st8f<float>(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *(reinterpret_cast<float4*>(p) + 1) = make_float4(v[4], v[5], v[6], v[7]);
}

// x*sigmoid(x) = h*tanh(h) + h with h = x/2: FMUL, MUFU.TANH, FFMA (the exp + divide form is 6 instructions)
__device__ __forceinline__ float silu_fast(float v) {
  const float h = 0.5f * v;
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
  return fmaf(h, th, h);
}

// ---------------------------------------------------------------------------------------------------------------
// conv_smallcin2
// ---------------------------------------------------------------------------------------------------------------
template <typename TO, int CIN, int COUT>
__global__ void __launch_bounds__(256) k_conv_smallcin2(const float* __restrict__ x1, const float* __restrict__ x2, int c1, int N, int H, int W,
                                                        const float* __restrict__ w, const float* __restrict__ bias, TO* __restrict__ y,
                                                        double* __restrict__ stats) {
  __shared__ __align__(16) float sw[9 * CIN * COUT + COUT];   // [tap][ci][COUT] + bias
  __shared__ float sred[8][16];
  for (int i = threadIdx.x; i < 9 * CIN * COUT; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) sw[9 * CIN * COUT + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int c2 = CIN - c1;
  const int64_t HW = (int64_t)H * W;
  const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = pix < (int64_t)N * HW;
  const int n = ok ? (int)(pix / HW) : 0;
  const int rem = ok ? (int)(pix - (int64_t)n * HW) : 0;
  const int oh = rem / W, ow = rem - oh * W;
  float acc[COUT];
#pragma unroll
  for (int j = 0; j < COUT; ++j) acc[j] = sw[9 * CIN * COUT + j];
  if (ok) {
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ih = oh + ky - 1;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int iw = ow + kx - 1;
        const bool in = ih >= 0 && ih < H && iw >= 0 && iw < W;
        const int64_t ip = ((int64_t)n * H + ih) * W + iw;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          float v = 0.f;
          if (in) v = ci < c1 ? __ldg(x1 + ip * c1 + ci) : __ldg(x2 + ip * c2 + (ci - c1));
          const float4* wp = reinterpret_cast<const float4*>(sw + ((ky * 3 + kx) * CIN + ci) * COUT);
#pragma unroll
          for (int j4 = 0; j4 < COUT / 4; ++j4) {
            const float4 ww = wp[j4];       // same address in every lane: one broadcast wavefront
            acc[4 * j4 + 0] = fmaf(v, ww.x, acc[4 * j4 + 0]); acc[4 * j4 + 1] = fmaf(v, ww.y, acc[4 * j4 + 1]);
            acc[4 * j4 + 2] = fmaf(v, ww.z, acc[4 * j4 + 2]); acc[4 * j4 + 3] = fmaf(v, ww.w, acc[4 * j4 + 3]);
          }
        }
      }
    }
    TO* yp = y + pix * COUT;
#pragma unroll
    for (int j8 = 0; j8 < COUT / 8; ++j8) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = acc[8 * j8 + j];
      st8f<TO>(yp + 8 * j8, o);
    }
  }
  if (stats) {   // host guarantees HW % 256 == 0 (a block never straddles two images) and COUT % 8 == 0
    constexpr int CPG = COUT / 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float gs[8], gq[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float s = 0.f, q = 0.f;
      if (ok) {
#pragma unroll
        for (int j = 0; j < CPG; ++j) { const float v = acc[g * CPG + j]; s += v; q = fmaf(v, v, q); }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
      gs[g] = s; gq[g] = q;
    }
    if (lane == 0) {
#pragma unroll
      for (int g = 0; g < 8; ++g) { sred[warp][2 * g] = gs[g]; sred[warp][2 * g + 1] = gq[g]; }
    }
    __syncthreads();
    if (threadIdx.x < 16) {
      float t = 0.f;
#pragma unroll
      for (int wi = 0; wi < 8; ++wi) t += sred[wi][threadIdx.x];
      const int nb = (int)(((int64_t)blockIdx.x * blockDim.x) / HW);
      atomicAdd(stats + (size_t)nb * 16 + threadIdx.x, (double)t);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// conv_cout1_v2
// ---------------------------------------------------------------------------------------------------------------
struct Cout1P {
  const void* x; int N, H, W;
  const float* w; const float* bias;
  const double* gn_sums; int groups; const float* gamma; const float* beta; float eps; int act_in;
  int mode, sanitize;
  const float* inp; float* y; const float* x_cur; float* x_next; float c1, c2;
};

constexpr int kO1TW = 32, kO1TH = 16;                       // output tile
constexpr int kO1HW = kO1TW + 2, kO1HH = kO1TH + 2;         // halo tile 34 x 18
constexpr int kO1HP = kO1HW * kO1HH;                        // 612 halo pixels
constexpr int kO1PS = 616;                                  // plane stride (floats)

template <typename T, int C>
__global__ void __launch_bounds__(256) k_conv_cout1_v2(Cout1P p) {
  __shared__ __align__(16) float s_w[9 * C];                // [tap][C]
  __shared__ __align__(16) float s_scale[C], s_shift[C];
  __shared__ float s_t[9 * kO1PS];                          // nine per-tap dot-product planes of the halo tile
  const int n = blockIdx.z;
  const int ox0 = blockIdx.x * kO1TW, oy0 = blockIdx.y * kO1TH;
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) s_w[i] = p.w[i];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float sc = 1.f, sh = 0.f;
    if (p.gn_sums) {
      const int cpg = C / p.groups;
      const int g = c / cpg;
      const double cnt = (double)cpg * p.H * p.W;
      const double m = p.gn_sums[((int64_t)n * p.groups + g) * 2] / cnt;
      double var = p.gn_sums[((int64_t)n * p.groups + g) * 2 + 1] / cnt - m * m;
      if (var < 0) var = 0;
      sc = (float)(1.0 / sqrt(var + (double)p.eps)) * p.gamma[c];
      sh = p.beta[c] - (float)m * sc;
    }
    s_scale[c] = sc; s_shift[c] = sh;
  }
  __syncthreads();
  // phase 1: one halo pixel per thread and iteration: load C channels (16-byte loads), normalise, activate, nine dot products
  const T* xb = (const T*)p.x + (int64_t)n * p.H * p.W * C;
  const bool silu = p.gn_sums != nullptr && p.act_in == ACT_SILU;
  for (int hp = threadIdx.x; hp < kO1HP; hp += blockDim.x) {
    const int hy = hp / kO1HW, hx = hp - hy * kO1HW;
    const int iy = oy0 + hy - 1, ix = ox0 + hx - 1;
    float t[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) t[k] = 0.f;
    if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {        // zero padding lives in the activated domain
      const T* src = xb + ((int64_t)iy * p.W + ix) * C;
#pragma unroll
      for (int c8 = 0; c8 < C / 8; ++c8) {
        float a[8];
        ld8f<T>(src + c8 * 8, a);
        const float4 sc0 = *reinterpret_cast<const float4*>(s_scale + c8 * 8), sc1 = *reinterpret_cast<const float4*>(s_scale + c8 * 8 + 4);
        const float4 sh0 = *reinterpret_cast<const float4*>(s_shift + c8 * 8), sh1 = *reinterpret_cast<const float4*>(s_shift + c8 * 8 + 4);
        a[0] = fmaf(a[0], sc0.x, sh0.x); a[1] = fmaf(a[1], sc0.y, sh0.y); a[2] = fmaf(a[2], sc0.z, sh0.z); a[3] = fmaf(a[3], sc0.w, sh0.w);
        a[4] = fmaf(a[4], sc1.x, sh1.x); a[5] = fmaf(a[5], sc1.y, sh1.y); a[6] = fmaf(a[6], sc1.z, sh1.z); a[7] = fmaf(a[7], sc1.w, sh1.w);
        if (silu) {
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] = silu_fast(a[j]);
        } else if (p.gn_sums && p.act_in != ACT_NONE) {
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] = act_apply(a[j], p.act_in);
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const float4 w0 = *reinterpret_cast<const float4*>(s_w + k * C + c8 * 8), w1 = *reinterpret_cast<const float4*>(s_w + k * C + c8 * 8 + 4);
          float d = t[k];
          d = fmaf(a[0], w0.x, d); d = fmaf(a[1], w0.y, d); d = fmaf(a[2], w0.z, d); d = fmaf(a[3], w0.w, d);
          d = fmaf(a[4], w1.x, d); d = fmaf(a[5], w1.y, d); d = fmaf(a[6], w1.z, d); d = fmaf(a[7], w1.w, d);
          t[k] = d;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) s_t[k * kO1PS + hp] = t[k];
  }
  __syncthreads();
  // phase 2: out(y,x) = sum over taps (ky,kx) of plane[ky*3+kx] at halo position (y+ky, x+kx)
  const float b0 = p.bias ? p.bias[0] : 0.f;
  for (int o = threadIdx.x; o < kO1TW * kO1TH; o += blockDim.x) {
    const int ty = o / kO1TW, tx = o - ty * kO1TW;
    const int ox = ox0 + tx, oy = oy0 + ty;
    if (ox >= p.W || oy >= p.H) continue;
    float v = b0;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) v += s_t[(ky * 3 + kx) * kO1PS + (ty + ky) * kO1HW + tx + kx];
    const int64_t oi = ((int64_t)n * p.H + oy) * p.W + ox;
    if (p.mode == 3) {
      if (p.y) p.y[oi] = v;
      const float e = clamp_nan(v, -5.f, 5.f);
      const float xn = p.c1 * (p.x_cur[oi] - p.c2 * e);
      p.x_next[oi] = clamp_nan(xn, 0.f, 1.f);
      continue;
    }
    if (p.mode == 1) v += p.inp[oi];
    if (p.mode == 2) v = 1.0f / (1.0f + expf(-v));
    if (p.sanitize) v = sanitize01(v);
    p.y[oi] = v;
  }
}

template <typename T, int C>
void launch_cout1_v2(Ctx& c, const Cout1Args& a) {
  Cout1P p;
  p.x = a.x.p; p.N = a.x.n; p.H = a.x.h; p.W = a.x.w;
  p.w = a.w; p.bias = a.bias;
  p.gn_sums = a.gn_sums; p.groups = a.groups; p.gamma = a.gamma; p.beta = a.beta; p.eps = a.eps; p.act_in = a.act_in;
  p.mode = a.mode; p.sanitize = a.sanitize; p.inp = a.inp; p.y = a.y; p.x_cur = a.x_cur; p.x_next = a.x_next; p.c1 = a.c1; p.c2 = a.c2;
  dim3 grid(cdiv(a.x.w, kO1TW), cdiv(a.x.h, kO1TH), a.x.n);
  XRD_LAUNCH(c, (k_conv_cout1_v2<T, C>), grid, 256, 0, p);
}

// ---------------------------------------------------------------------------------------------------------------
// upsample2x + GroupNorm sums of the result
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(288) k_upsample2x_stats(const T* __restrict__ x, T* __restrict__ y, int H, int W, int C, int pix_per_block,
                                                          double* __restrict__ stats) {
  extern __shared__ float sm[];   // [2][C]
  const int V = C >> 3;           // 16-byte vector slots per pixel
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int ppi = blockDim.x / V;
  const int slot = threadIdx.x % V, lane = threadIdx.x / V;
  const int Ho = 2 * H, Wo = 2 * W, HWo = Ho * Wo;
  const T* xb = x + (int64_t)n * H * W * C + slot * 8;
  T* yb = y + (int64_t)n * HWo * C + slot * 8;
  float s[8], ss[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; ss[i] = 0.f; }
  const int p0 = blockIdx.x * pix_per_block, p1 = min(p0 + pix_per_block, HWo);
  for (int pp = p0 + lane; pp < p1; pp += ppi) {
    const int oh = pp / Wo, ow = pp - oh * Wo;
    int h0, h1, w0, w1; float fh, fw;                      // weight of the *second* sample
    if (oh & 1) { h0 = oh >> 1; h1 = min(h0 + 1, H - 1); fh = 0.25f; } else { h1 = oh >> 1; h0 = max(h1 - 1, 0); fh = 0.75f; }
    if (ow & 1) { w0 = ow >> 1; w1 = min(w0 + 1, W - 1); fw = 0.25f; } else { w1 = ow >> 1; w0 = max(w1 - 1, 0); fw = 0.75f; }
    float a[8], b[8], cc[8], d[8], o[8];
    ld8f<T>(xb + ((int64_t)h0 * W + w0) * C, a);
    ld8f<T>(xb + ((int64_t)h0 * W + w1) * C, b);
    ld8f<T>(xb + ((int64_t)h1 * W + w0) * C, cc);
    ld8f<T>(xb + ((int64_t)h1 * W + w1) * C, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float top = a[j] + fw * (b[j] - a[j]);
      const float bot = cc[j] + fw * (d[j] - cc[j]);
      o[j] = top + fh * (bot - top);
      s[j] += o[j]; ss[j] = fmaf(o[j], o[j], ss[j]);
    }
    st8f<T>(yb + (int64_t)pp * C, o);
  }
  if (stats) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { atomicAdd(&sm[slot * 8 + i], s[i]); atomicAdd(&sm[C + slot * 8 + i], ss[i]); }
    __syncthreads();
    if (threadIdx.x < 8) {
      const int cpg = C / 8;
      double a = 0.0, b = 0.0;
      for (int i = 0; i < cpg; ++i) { a += (double)sm[threadIdx.x * cpg + i]; b += (double)sm[C + threadIdx.x * cpg + i]; }
      atomicAdd(&stats[((int64_t)n * 8 + threadIdx.x) * 2 + 0], a);
      atomicAdd(&stats[((int64_t)n * 8 + threadIdx.x) * 2 + 1], b);
    }
  }
}

// Same arithmetic, one thread per INPUT pixel and 8 channels: the 2x2 output block of input pixel (i, j) needs the 3x3 input
// neighbourhood, i.e. 9 loads and conversions for four outputs instead of 16, and the three rows' horizontal interpolations are
// shared by the two output rows (the per-output kernel above was issue-bound: 2 TB/s).
template <typename T>
__global__ void __launch_bounds__(288) k_upsample2x_stats_v2(const T* __restrict__ x, T* __restrict__ y, int H, int W, int C, int pix_per_block,
                                                             double* __restrict__ stats) {
  extern __shared__ float sm[];   // [2][C]
  const int V = C >> 3;
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int ppi = blockDim.x / V;
  const int slot = threadIdx.x % V, lane = threadIdx.x / V;
  const int Wo = 2 * W, HW = H * W;
  const T* xb = x + (int64_t)n * HW * C + slot * 8;
  T* yb = y + (int64_t)n * 4 * HW * C + slot * 8;
  float s[8], ss[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; ss[i] = 0.f; }
  const int p0 = blockIdx.x * pix_per_block, p1 = min(p0 + pix_per_block, HW);
  for (int pp = p0 + lane; pp < p1; pp += ppi) {
    const int i = pp / W, j = pp - i * W;
    const int rr[3] = {max(i - 1, 0), i, min(i + 1, H - 1)};
    const int cm = max(j - 1, 0), cp = min(j + 1, W - 1);
    float hl[3][8], hr[3][8];              // per input row: value at output column 2j (weights .25/.75) and 2j+1 (.75/.25)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float a[8], b[8], d[8];
      const T* row = xb + (int64_t)rr[r] * W * C;
      ld8f<T>(row + (int64_t)cm * C, a);
      ld8f<T>(row + (int64_t)j * C, b);
      ld8f<T>(row + (int64_t)cp * C, d);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        hl[r][k] = a[k] + 0.75f * (b[k] - a[k]);
        hr[r][k] = b[k] + 0.25f * (d[k] - b[k]);
      }
    }
    float o[8];
    T* o00 = yb + ((int64_t)(2 * i) * Wo + 2 * j) * C;
#pragma unroll
    for (int q = 0; q < 4; ++q) {          // (output row parity, output column parity)
      const int dy = q >> 1, dxp = q & 1;
      const float fh = dy ? 0.25f : 0.75f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float top = dxp ? hr[dy][k] : hl[dy][k];
        const float bot = dxp ? hr[dy + 1][k] : hl[dy + 1][k];
        o[k] = top + fh * (bot - top);
        s[k] += o[k]; ss[k] = fmaf(o[k], o[k], ss[k]);
      }
      st8f<T>(o00 + ((int64_t)dy * Wo + dxp) * C, o);
    }
  }
  if (stats) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { atomicAdd(&sm[slot * 8 + i], s[i]); atomicAdd(&sm[C + slot * 8 + i], ss[i]); }
    __syncthreads();
    if (threadIdx.x < 8) {
      const int cpg = C / 8;
      double a = 0.0, b = 0.0;
      for (int i = 0; i < cpg; ++i) { a += (double)sm[threadIdx.x * cpg + i]; b += (double)sm[C + threadIdx.x * cpg + i]; }
      atomicAdd(&stats[((int64_t)n * 8 + threadIdx.x) * 2 + 0], a);
      atomicAdd(&stats[((int64_t)n * 8 + threadIdx.x) * 2 + 1], b);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_im2col_3x3_2ch(const float* __restrict__ a, const float* __restrict__ b, T* __restrict__ col, int N, int H,
                                                        int W) {
  const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (int64_t)N * H * W) return;
  const int ow = (int)(pix % W), oh = (int)((pix / W) % H);
  const int64_t base = pix - (int64_t)oh * W - ow;        // first pixel of this image
  float v[32];
#pragma unroll
  for (int i = 18; i < 32; ++i) v[i] = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ih = oh + ky - 1, iw = ow + kx - 1;
      const bool in = ih >= 0 && ih < H && iw >= 0 && iw < W;
      const int64_t ip = base + (int64_t)ih * W + iw;
      v[(ky * 3 + kx) * 2 + 0] = in ? __ldg(a + ip) : 0.f;
      v[(ky * 3 + kx) * 2 + 1] = in ? __ldg(b + ip) : 0.f;
    }
  }
  T* dst = col + pix * 32;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = v[j * 8 + i];
    st8f<T>(dst + j * 8, o);
  }
}

}  // namespace

void im2col_3x3_2ch(Ctx& c, const float* a, const float* b, Tens& col) {
  XRD_REQUIRE(col.c == 32 && col.dt != DT_F32, "im2col_3x3_2ch: 32 16-bit columns expected");
  const int64_t total = (int64_t)col.n * col.h * col.w;
  if (col.dt == DT_BF16)
    XRD_LAUNCH(c, (k_im2col_3x3_2ch<__nv_bfloat16>), (int)cdiv64(total, 256), 256, 0, a, b, (__nv_bfloat16*)col.p, col.n, col.h, col.w);
  else
    XRD_LAUNCH(c, (k_im2col_3x3_2ch<__half>), (int)cdiv64(total, 256), 256, 0, a, b, (__half*)col.p, col.n, col.h, col.w);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
bool conv_smallcin2_supported(const Tens& x1, const Tens* x2, const ConvW& w) {
  const int cin = x1.c + (x2 ? x2->c : 0);
  return x1.dt == DT_F32 && (!x2 || x2->dt == DT_F32) && cin >= 1 && cin <= 3 && w.kh == 3 && w.kw == 3 && w.stride == 1 && w.pad == 1 &&
         (w.cout == 32 || w.cout == 48) && !w.d2s;
}

bool conv_smallcin2_stats_supported(const Tens& x1) { return ((int64_t)x1.h * x1.w) % 256 == 0; }

template <typename TO, int CIN>
static void smallcin2_launch(Ctx& c, const Tens& x1, const Tens* x2, const ConvW& w, Tens& y, double* stats) {
  const int64_t total = (int64_t)x1.n * x1.h * x1.w;
  const int blocks = (int)cdiv64(total, 256);
  const float* p2 = (const float*)(x2 ? x2->p : nullptr);
  if (w.cout == 48)
    XRD_LAUNCH(c, (k_conv_smallcin2<TO, CIN, 48>), blocks, 256, 0, (const float*)x1.p, p2, x1.c, x1.n, x1.h, x1.w, w.w, w.bias, (TO*)y.p, stats);
  else
    XRD_LAUNCH(c, (k_conv_smallcin2<TO, CIN, 32>), blocks, 256, 0, (const float*)x1.p, p2, x1.c, x1.n, x1.h, x1.w, w.w, w.bias, (TO*)y.p, stats);
}

void conv_smallcin2(Ctx& c, const Tens& x1, const Tens* x2, const ConvW& w, Tens& y, double* stats) {
  XRD_REQUIRE(conv_smallcin2_supported(x1, x2, w), "conv_smallcin2: unsupported");
  const int cin = x1.c + (x2 ? x2->c : 0);
  XRD_REQUIRE(cin == w.cin && y.n == x1.n && y.h == x1.h && y.w == x1.w && y.c == w.cout, "conv_smallcin2: shape");
  if (stats) XRD_REQUIRE(conv_smallcin2_stats_supported(x1), "conv_smallcin2: statistics need H*W %% 256 == 0");
  XRD_DISPATCH(y.dt, TO, {
    if (cin == 1) smallcin2_launch<TO, 1>(c, x1, x2, w, y, stats);
    else if (cin == 2) smallcin2_launch<TO, 2>(c, x1, x2, w, y, stats);
    else smallcin2_launch<TO, 3>(c, x1, x2, w, y, stats);
  });
}

bool conv_cout1_v2_supported(const Cout1Args& a) { return a.k == 3 && (a.x.c == 48 || a.x.c == 32); }

void conv_cout1_v2(Ctx& c, const Cout1Args& a) {
  XRD_REQUIRE(conv_cout1_v2_supported(a), "conv_cout1_v2: unsupported");
  XRD_DISPATCH(a.x.dt, T, {
    if (a.x.c == 48) launch_cout1_v2<T, 48>(c, a);
    else launch_cout1_v2<T, 32>(c, a);
  });
}

// y = bilinear 2x of x; stats (nullable, [N][8][2], zeroed by the caller) += GroupNorm sums of y over 8 channel groups
void upsample2x_stats(Ctx& c, const Tens& x, Tens& y, double* stats) {
  XRD_REQUIRE(y.n == x.n && y.h == 2 * x.h && y.w == 2 * x.w && y.c == x.c && y.dt == x.dt, "upsample2x: shape");
  const int V = x.c / 8;
  if (x.dt == DT_F32 || x.c % 8 != 0 || V > 288) {      // check mode / odd channel counts: plain kernel + statistics pass
    upsample2x(c, x, y);
    if (stats) gn_stats(c, y, nullptr, 8, stats);
    return;
  }
  const int threads = (288 / V) * V;
  static const int v2 = getenv("XRD_UPS2") ? atoi(getenv("XRD_UPS2")) : 1;
  if (v2) {                      // one thread per input pixel: 2x2 outputs from 9 loads
    const int HWi = x.h * x.w;
    const int ppi = threads / V;
    int64_t ppb = cdiv64((int64_t)HWi * x.n, 148 * 8);
    ppb = std::max<int64_t>(ppb, (int64_t)ppi * 4);
    ppb = std::min<int64_t>(cdiv64(ppb, ppi) * ppi, HWi);
    dim3 grid(cdiv(HWi, (int)ppb), x.n);
    if (x.dt == DT_BF16)
      XRD_LAUNCH(c, (k_upsample2x_stats_v2<__nv_bfloat16>), grid, threads, 2 * x.c * sizeof(float), (const __nv_bfloat16*)x.p,
                 (__nv_bfloat16*)y.p, x.h, x.w, x.c, (int)ppb, stats);
    else
      XRD_LAUNCH(c, (k_upsample2x_stats_v2<__half>), grid, threads, 2 * x.c * sizeof(float), (const __half*)x.p, (__half*)y.p, x.h, x.w,
                 x.c, (int)ppb, stats);
    return;
  }
  const int HWo = y.h * y.w;
  const int ppi = threads / V;
  int64_t ppb = cdiv64((int64_t)HWo * x.n, 148 * 8);
  ppb = std::max<int64_t>(ppb, (int64_t)ppi * 8);
  ppb = std::min<int64_t>(cdiv64(ppb, ppi) * ppi, HWo);
  dim3 grid(cdiv(HWo, (int)ppb), x.n);
  if (x.dt == DT_BF16)
    XRD_LAUNCH(c, (k_upsample2x_stats<__nv_bfloat16>), grid, threads, 2 * x.c * sizeof(float), (const __nv_bfloat16*)x.p, (__nv_bfloat16*)y.p, x.h,
               x.w, x.c, (int)ppb, stats);
  else
    XRD_LAUNCH(c, (k_upsample2x_stats<__half>), grid, threads, 2 * x.c * sizeof(float), (const __half*)x.p, (__half*)y.p, x.h, x.w, x.c,
               (int)ppb, stats);
}

}  // namespace xrd
