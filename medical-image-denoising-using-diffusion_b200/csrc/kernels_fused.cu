// kernels_fused.cu -- bandwidth-oriented versions of the HBM-bound kernels on the UNet loop:
//   * GroupNorm statistics / apply+activation with 16-byte accesses and a fixed channel slot per thread,
//   * direct convolution for tiny Cin (the first conv of every network: cat[x,cond]->48, 1->32, 3->48),
//   * the single-output-channel 3x3 conv (UNet out_conv + sampler update, NAFNet ending) with the
//     GroupNorm+SiLU'd input tile staged once in shared memory.
#include "kernels.cuh"
#include <type_traits>
#include <algorithm>
#include "tc_common.cuh"

namespace xrd {

void gn_stats_v1(Ctx& c, const Tens& x1, const Tens* x2, int groups, double* sums);
void gn_act_v1(Ctx& c, const Tens& x1, const Tens* x2, int groups, const double* sums, const float* gamma, const float* beta,
               float eps, int act, Tens& y);
void conv_cout1_v1(Ctx& c, const Cout1Args& a);

template <typename T> struct V16 { static constexpr int N = 16 / sizeof(T); };

template <typename T> __device__ __forceinline__ void ldv(const T* p, float* v);
template <> __device__ __forceinline__ void ldv<float>(const float* p, float* v) {
  float4 t = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void ldv<__half>(const __half* p, float* v) {
  uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
  const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <> __device__ __forceinline__ void ldv<__nv_bfloat16>(const __nv_bfloat16* p, float* v) {
  uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

__device__ __forceinline__ float act_fast(float v, int act) {
  switch (act) {
    case ACT_SILU: return __fdividef(v, 1.0f + __expf(-v));
    case ACT_GELU: return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
    case ACT_SIGMOID: return __fdividef(1.0f, 1.0f + __expf(-v));
    default: return v;
  }
}

// -------------------------------------------------------------------------------------------------
// GroupNorm statistics: thread t owns vector slot (t % V) of every pixel it visits, V = C / (16 B worth of T)
// -------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(288) k_gn_stats2(const T* __restrict__ x1, const T* __restrict__ x2, int c1, int c2, int HW,
                                                   int groups, int pix_per_block, double* __restrict__ sums) {
  constexpr int VN = V16<T>::N;
  extern __shared__ float sm[];  // [2][ctot]
  const int ctot = c1 + c2, V = ctot / VN;
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * ctot; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int ppi = blockDim.x / V;
  const int slot = threadIdx.x % V, lane = threadIdx.x / V;
  const int c = slot * VN;
  const T* src; int cs, cc;
  if (c < c1) { src = x1; cs = c1; cc = c; } else { src = x2; cs = c2; cc = c - c1; }
  src += (int64_t)n * HW * cs + cc;
  float s[VN], ss[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) { s[i] = 0.f; ss[i] = 0.f; }
  const int p0 = blockIdx.x * pix_per_block, p1 = min(p0 + pix_per_block, HW);
  int pp = p0 + lane;
  for (; pp + 3 * ppi < p1; pp += 4 * ppi) {       // four independent 16 B loads in flight per thread
    float a[VN], b[VN], d[VN], e[VN];
    ldv<T>(src + (int64_t)pp * cs, a);
    ldv<T>(src + (int64_t)(pp + ppi) * cs, b);
    ldv<T>(src + (int64_t)(pp + 2 * ppi) * cs, d);
    ldv<T>(src + (int64_t)(pp + 3 * ppi) * cs, e);
#pragma unroll
    for (int i = 0; i < VN; ++i) {
      s[i] += (a[i] + b[i]) + (d[i] + e[i]);
      ss[i] = fmaf(a[i], a[i], ss[i]); ss[i] = fmaf(b[i], b[i], ss[i]);
      ss[i] = fmaf(d[i], d[i], ss[i]); ss[i] = fmaf(e[i], e[i], ss[i]);
    }
  }
  for (; pp < p1; pp += ppi) {
    float a[VN];
    ldv<T>(src + (int64_t)pp * cs, a);
#pragma unroll
    for (int i = 0; i < VN; ++i) { s[i] += a[i]; ss[i] = fmaf(a[i], a[i], ss[i]); }
  }
#pragma unroll
  for (int i = 0; i < VN; ++i) { atomicAdd(&sm[c + i], s[i]); atomicAdd(&sm[ctot + c + i], ss[i]); }
  __syncthreads();
  if (threadIdx.x < groups) {
    const int cpg = ctot / groups;
    double a = 0.0, b = 0.0;
    for (int i = 0; i < cpg; ++i) { a += (double)sm[threadIdx.x * cpg + i]; b += (double)sm[ctot + threadIdx.x * cpg + i]; }
    atomicAdd(&sums[((int64_t)n * groups + threadIdx.x) * 2 + 0], a);
    atomicAdd(&sums[((int64_t)n * groups + threadIdx.x) * 2 + 1], b);
  }
}

static int pick_ppb(int HW, int N, int ppi) {
  // aim for ~6 blocks per SM in total, at least 4 iterations of 4 loads each per block -- unless that leaves the machine
  // under-filled (one image at the deep levels: 22 blocks of 16 items per thread, 8.7 us for a 1.5 MB tensor): then one
  // iteration of 4 loads per block is the floor, so that the served shape spreads over the SMs
  int64_t target_blocks = 148 * 6;
  int64_t ppb = cdiv64((int64_t)HW * N, target_blocks);
  const int64_t deep = (int64_t)ppi * 16, shallow = (int64_t)ppi * 4;
  const bool filled = cdiv64(HW, deep) * N >= 2 * 148;
  ppb = std::max<int64_t>(ppb, filled ? deep : shallow);
  ppb = cdiv64(ppb, ppi) * ppi;
  return (int)std::min<int64_t>(ppb, std::max(HW, 1));
}

__global__ void k_gn_merge_stats(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;     // (n, g, which)
  if (i >= N * 16) return;
  const int n = i >> 4, g = (i >> 1) & 7, k = i & 1;
  const double* src = (g < 4 ? a : b) + (size_t)n * 16 + (size_t)(g & 3) * 4 + k;
  out[i] = src[0] + src[2];
}

void gn_merge_stats(Ctx& c, const double* a, const double* b, double* out, int N) {
  XRD_LAUNCH(c, k_gn_merge_stats, cdiv(N * 16, 128), 128, 0, a, b, out, N);
}

__global__ void k_gn_coef(const double* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int N,
                          int C, int groups, int HW, float2* __restrict__ coef) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  const int n = i / C, ch = i - n * C, cpg = C / groups, g = ch / cpg;
  const double cnt = (double)cpg * HW;
  const double m = sums[((int64_t)n * groups + g) * 2] / cnt;
  double var = sums[((int64_t)n * groups + g) * 2 + 1] / cnt - m * m;
  if (var < 0) var = 0;
  const float a = (float)(1.0 / sqrt(var + (double)eps)) * gamma[ch];
  coef[i] = make_float2(0.5f * a, 0.5f * (beta[ch] - (float)m * a));
}

void gn_coef(Ctx& c, const double* sums, const float* gamma, const float* beta, float eps, int N, int C, int groups, int HW, float2* coef) {
  XRD_LAUNCH(c, k_gn_coef, cdiv(N * C, 128), 128, 0, sums, gamma, beta, eps, N, C, groups, HW, coef);
}

void gn_stats(Ctx& c, const Tens& x1, const Tens* x2, int groups, double* sums) {
  const int c1 = x1.c, c2 = x2 ? x2->c : 0, ctot = c1 + c2;
  const int VN = (int)(16 / dsize(x1.dt));
  const bool ok = (c1 % VN) == 0 && (c2 % VN) == 0 && ctot % groups == 0 && groups <= 32 && ctot / VN <= 288 &&
                  (!x2 || x2->dt == x1.dt);
  if (!ok) { gn_stats_v1(c, x1, x2, groups, sums); return; }
  if (x2) XRD_REQUIRE(x2->n == x1.n && x2->h == x1.h && x2->w == x1.w, "gn_stats: source mismatch");
  const int V = ctot / VN;
  const int threads = (288 / V) * V;
  const int HW = x1.h * x1.w;
  const int ppb = pick_ppb(HW, x1.n, threads / V);
  dim3 grid(cdiv(HW, ppb), x1.n);
  XRD_DISPATCH(x1.dt, T, XRD_LAUNCH(c, (k_gn_stats2<T>), grid, threads, 2 * ctot * sizeof(float), (const T*)x1.p,
                                    (const T*)(x2 ? x2->p : nullptr), c1, c2, HW, groups, ppb, sums));
}

// -------------------------------------------------------------------------------------------------
// GroupNorm apply + activation (+ virtual concat): per-thread scale/shift for its fixed channel slot
// -------------------------------------------------------------------------------------------------
template <typename TI, typename TO, bool FAST>
__global__ void __launch_bounds__(288) k_gn_act2(const TI* __restrict__ x1, const TI* __restrict__ x2, int c1, int c2, int HW, int groups,
                                                 const double* __restrict__ sums, const float* __restrict__ gamma,
                                                 const float* __restrict__ beta, float eps, int act, TO* __restrict__ y,
                                                 int pix_per_block) {
  constexpr int VN = V16<TI>::N;
  __shared__ float s_mean[32], s_rstd[32];
  const int ctot = c1 + c2, cpg = ctot / groups, n = blockIdx.y, V = ctot / VN;
  if (threadIdx.x < groups) {
    double cnt = (double)cpg * HW;
    double m = sums[((int64_t)n * groups + threadIdx.x) * 2] / cnt;
    double var = sums[((int64_t)n * groups + threadIdx.x) * 2 + 1] / cnt - m * m;
    if (var < 0) var = 0;
    s_mean[threadIdx.x] = (float)m;
    s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const int ppi = blockDim.x / V;
  const int slot = threadIdx.x % V, lane = threadIdx.x / V;
  const int c = slot * VN;
  float sc[VN], sh[VN];
#pragma unroll
  for (int i = 0; i < VN; ++i) {
    const int g = (c + i) / cpg;
    const float a = s_rstd[g] * gamma[c + i];
    sc[i] = a;
    sh[i] = beta[c + i] - s_mean[g] * a;
  }
  const TI* src; int cs, cc;
  if (c < c1) { src = x1; cs = c1; cc = c; } else { src = x2; cs = c2; cc = c - c1; }
  src += (int64_t)n * HW * cs + cc;
  TO* dst = y + (int64_t)n * HW * ctot + c;
  const int p0 = blockIdx.x * pix_per_block, p1 = min(p0 + pix_per_block, HW);
  int pp = p0 + lane;
  if constexpr (FAST && sizeof(TI) == 2 && sizeof(TO) == 2) {
    // 16-bit fast path: raw 16-byte loads, one 16-byte store per pixel slot, and for SiLU the tanh form
    //   x*sigmoid(x) = h*tanh(h) + h with h = x/2  (FMA, MUFU.TANH, FMA instead of EX2 + RCP + two multiplies; the 1/2 is folded
    // into the GroupNorm scale/shift) -- the same form conv3r applies in shared memory.  The 8-byte stores and the
    // exp/divide sequence of the generic loop made this kernel issue-bound at 68 % of the HBM copy rate.
    const bool silu = act == ACT_SILU;
    if (silu) {
#pragma unroll
      for (int i = 0; i < VN; ++i) { sc[i] *= 0.5f; sh[i] *= 0.5f; }
    }
    auto apply = [&](const uint4& q) -> uint4 {
      float v[8];
      tc::unpack8<TI>(q, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float t = fmaf(v[i], sc[i], sh[i]);
        if (silu) {
          float th;
          asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(t));
          v[i] = fmaf(t, th, t);
        } else {
          v[i] = act_fast(t, act);
        }
      }
      uint4 o;
      o.x = tc::pack2<TO>(v[0], v[1]); o.y = tc::pack2<TO>(v[2], v[3]); o.z = tc::pack2<TO>(v[4], v[5]); o.w = tc::pack2<TO>(v[6], v[7]);
      return o;
    };
    for (; pp + 3 * ppi < p1; pp += 4 * ppi) {       // four independent 16-byte loads in flight per thread
      uint4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = __ldg(reinterpret_cast<const uint4*>(src + (int64_t)(pp + u * ppi) * cs));
#pragma unroll
      for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(dst + (int64_t)(pp + u * ppi) * ctot) = apply(q[u]);
    }
    for (; pp < p1; pp += ppi) *reinterpret_cast<uint4*>(dst + (int64_t)pp * ctot) = apply(__ldg(reinterpret_cast<const uint4*>(src + (int64_t)pp * cs)));
    return;
  }
  for (; pp + 3 * ppi < p1; pp += 4 * ppi) {       // four independent 16-byte loads in flight per thread
    float a[4][VN];
#pragma unroll
    for (int u = 0; u < 4; ++u) ldv<TI>(src + (int64_t)(pp + u * ppi) * cs, a[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < VN; ++i) {
        const float t = fmaf(a[u][i], sc[i], sh[i]);
        a[u][i] = FAST ? act_fast(t, act) : act_apply(t, act);
      }
#pragma unroll
      for (int i = 0; i < VN; i += 4) {
        float o4[4] = {a[u][i], a[u][i + 1], a[u][i + 2], a[u][i + 3]};
        st4<TO>(dst + (int64_t)(pp + u * ppi) * ctot + i, o4);
      }
    }
  }
  for (; pp < p1; pp += ppi) {
    float a[VN];
    ldv<TI>(src + (int64_t)pp * cs, a);
#pragma unroll
    for (int i = 0; i < VN; ++i) { const float ta = fmaf(a[i], sc[i], sh[i]); a[i] = FAST ? act_fast(ta, act) : act_apply(ta, act); }
#pragma unroll
    for (int i = 0; i < VN; i += 4) {
      float o4[4] = {a[i], a[i + 1], a[i + 2], a[i + 3]};
      st4<TO>(dst + (int64_t)pp * ctot + i, o4);
    }
  }
}

void gn_act(Ctx& c, const Tens& x1, const Tens* x2, int groups, const double* sums, const float* gamma, const float* beta, float eps,
            int act, Tens& y) {
  const int c1 = x1.c, c2 = x2 ? x2->c : 0, ctot = c1 + c2;
  const int VN = (int)(16 / dsize(x1.dt));
  const bool ok = (c1 % VN) == 0 && (c2 % VN) == 0 && ctot % groups == 0 && groups <= 32 && ctot / VN <= 288 &&
                  (!x2 || x2->dt == x1.dt) && y.dt == x1.dt;
  if (!ok) { gn_act_v1(c, x1, x2, groups, sums, gamma, beta, eps, act, y); return; }
  XRD_REQUIRE(y.c == ctot && y.n == x1.n && y.h == x1.h && y.w == x1.w, "gn_act: output shape mismatch");
  const int V = ctot / VN;
  const int threads = (288 / V) * V;
  const int HW = x1.h * x1.w;
  const int ppb = pick_ppb(HW, x1.n, threads / V);
  dim3 grid(cdiv(HW, ppb), x1.n);
  switch (x1.dt) {
    case DT_F32:
      XRD_LAUNCH(c, (k_gn_act2<float, float, false>), grid, threads, 0, (const float*)x1.p, (const float*)(x2 ? x2->p : nullptr), c1, c2,
                 HW, groups, sums, gamma, beta, eps, act, (float*)y.p, ppb);
      break;
    case DT_BF16:
      XRD_LAUNCH(c, (k_gn_act2<__nv_bfloat16, __nv_bfloat16, true>), grid, threads, 0, (const __nv_bfloat16*)x1.p,
                 (const __nv_bfloat16*)(x2 ? x2->p : nullptr), c1, c2, HW, groups, sums, gamma, beta, eps, act, (__nv_bfloat16*)y.p, ppb);
      break;
    case DT_F16:
      XRD_LAUNCH(c, (k_gn_act2<__half, __half, true>), grid, threads, 0, (const __half*)x1.p, (const __half*)(x2 ? x2->p : nullptr), c1,
                 c2, HW, groups, sums, gamma, beta, eps, act, (__half*)y.p, ppb);
      break;
  }
}

// -------------------------------------------------------------------------------------------------
// 3x3 (stride 1, pad 1) convolution with a tiny input-channel count read from fp32 planes / fp32 NHWC:
//   thread = (output pixel, 8 consecutive output channels); weights [9][cin][cout] staged in smem
// -------------------------------------------------------------------------------------------------
template <typename TO>
__global__ void __launch_bounds__(256) k_conv_smallcin(const float* __restrict__ x1, const float* __restrict__ x2, int c1, int c2, int N,
                                                       int H, int W, int Cout, const float* __restrict__ w, const float* __restrict__ bias,
                                                       TO* __restrict__ y) {
  extern __shared__ float sw[];  // [9*cin][Cout] + bias[Cout]
  const int cin = c1 + c2;
  for (int i = threadIdx.x; i < 9 * cin * Cout; i += blockDim.x) sw[i] = w[i];
  float* sb = sw + 9 * cin * Cout;
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sb[i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int O = Cout >> 3;
  const int64_t total = (int64_t)N * H * W * O;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int o8 = (int)(i % O) * 8;
    int64_t pix = i / O;
    const int ow = (int)(pix % W);
    const int oh = (int)((pix / W) % H);
    const int n = (int)(pix / ((int64_t)W * H));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = sb[o8 + j];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ih = oh + ky - 1;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int iw = ow + kx - 1;
        if (iw < 0 || iw >= W) continue;
        const int64_t ip = ((int64_t)n * H + ih) * W + iw;
        for (int ci = 0; ci < cin; ++ci) {
          const float v = ci < c1 ? __ldg(x1 + ip * c1 + ci) : __ldg(x2 + ip * c2 + (ci - c1));
          const float* wp = sw + ((ky * 3 + kx) * cin + ci) * Cout + o8;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, wp[j], acc[j]);
        }
      }
    }
    float lo[4] = {acc[0], acc[1], acc[2], acc[3]}, hi[4] = {acc[4], acc[5], acc[6], acc[7]};
    st4<TO>(y + pix * Cout + o8, lo);
    st4<TO>(y + pix * Cout + o8 + 4, hi);
  }
}

bool conv_smallcin_supported(const Tens& x1, const Tens* x2, const ConvW& w) {
  const int cin = x1.c + (x2 ? x2->c : 0);
  return x1.dt == DT_F32 && (!x2 || x2->dt == DT_F32) && cin <= 4 && w.kh == 3 && w.kw == 3 && w.stride == 1 && w.pad == 1 &&
         w.cout % 8 == 0 && !w.d2s && (size_t)(9 * cin + 1) * w.cout * 4 <= 48 * 1024;
}

void conv_smallcin(Ctx& c, const Tens& x1, const Tens* x2, const ConvW& w, Tens& y) {
  XRD_REQUIRE(conv_smallcin_supported(x1, x2, w), "conv_smallcin: unsupported");
  const int c1 = x1.c, c2 = x2 ? x2->c : 0;
  XRD_REQUIRE(c1 + c2 == w.cin && y.n == x1.n && y.h == x1.h && y.w == x1.w && y.c == w.cout, "conv_smallcin: shape");
  const int64_t total = (int64_t)x1.n * x1.h * x1.w * (w.cout / 8);
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv64(total, 256), 148 * 16));
  const size_t smem = (size_t)(9 * (c1 + c2) + 1) * w.cout * sizeof(float);
  XRD_DISPATCH(y.dt, TO, XRD_LAUNCH(c, (k_conv_smallcin<TO>), blocks, 256, smem, (const float*)x1.p, (const float*)(x2 ? x2->p : nullptr),
                                    c1, c2, x1.n, x1.h, x1.w, w.cout, w.w, w.bias, (TO*)y.p));
}

// -------------------------------------------------------------------------------------------------
// 3x3 conv to ONE output channel with the (GroupNorm + activation)'d input tile staged once in smem:
//   block = 16x16 output pixels; smem tile [C][18][18] fp32 (channel-major: conflict-free for both phases)
// -------------------------------------------------------------------------------------------------
struct Cout1T {
  const void* x; int N, H, W, C;
  const float* w; const float* bias;
  const double* gn_sums; int groups; const float* gamma; const float* beta; float eps; int act_in;
  int mode, sanitize;
  const float* inp; float* y; const float* x_cur; float* x_next; float c1, c2;
};

template <typename T, bool FAST>
__global__ void __launch_bounds__(256) k_conv_cout1_tiled(Cout1T p) {
  constexpr int TS = 16, HS = TS + 2, HP = HS * HS;     // 18x18 halo tile, 324 pixels
  constexpr int VN = V16<T>::N;
  extern __shared__ float sm[];                          // tile[C][HP] | w[9][C] | scale[C] | shift[C]
  const int C = p.C;
  float* tile = sm;
  float* s_w = tile + C * HP;
  float* s_scale = s_w + 9 * C;
  float* s_shift = s_scale + C;
  const int n = blockIdx.z;
  const int ox0 = blockIdx.x * TS, oy0 = blockIdx.y * TS;
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) s_w[i] = p.w[i];
  if (p.gn_sums) {
    const int cpg = C / p.groups;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const int g = c / cpg;
      const double cnt = (double)cpg * p.H * p.W;
      const double m = p.gn_sums[((int64_t)n * p.groups + g) * 2] / cnt;
      double var = p.gn_sums[((int64_t)n * p.groups + g) * 2 + 1] / cnt - m * m;
      if (var < 0) var = 0;
      const float sc = (float)(1.0 / sqrt(var + (double)p.eps)) * p.gamma[c];
      s_scale[c] = sc;
      s_shift[c] = p.beta[c] - (float)m * sc;
    }
  }
  __syncthreads();
  // phase 1: load + normalise + activate each halo pixel once.  item = (channel vector, halo pixel): consecutive threads
  // take consecutive pixels of one channel vector, so the transposed smem writes are conflict free.
  const int V = C / VN;
  const T* xb = (const T*)p.x + (int64_t)n * p.H * p.W * C;
  for (int i = threadIdx.x; i < V * HP; i += blockDim.x) {
    const int v = i / HP, hp = i - v * HP;
    const int hy = hp / HS, hx = hp - hy * HS;
    const int iy = oy0 + hy - 1, ix = ox0 + hx - 1;
    float a[VN];
    const bool ok = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
    if (ok) {
      ldv<T>(xb + ((int64_t)iy * p.W + ix) * C + v * VN, a);
      if (p.gn_sums) {
#pragma unroll
        for (int j = 0; j < VN; ++j) {
          const float t = fmaf(a[j], s_scale[v * VN + j], s_shift[v * VN + j]);
          a[j] = FAST ? act_fast(t, p.act_in) : act_apply(t, p.act_in);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < VN; ++j) a[j] = 0.f;          // zero padding lives in the activated domain
    }
#pragma unroll
    for (int j = 0; j < VN; ++j) tile[(v * VN + j) * HP + hp] = a[j];
  }
  __syncthreads();
  // phase 2: one output pixel per thread
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int ox = ox0 + tx, oy = oy0 + ty;
  float acc = 0.f;
  for (int c = 0; c < C; ++c) {
    const float* tp = tile + c * HP + ty * HS + tx;
    const float* wp = s_w + c;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) acc = fmaf(tp[ky * HS + kx], wp[(ky * 3 + kx) * C], acc);
  }
  if (ox >= p.W || oy >= p.H) return;
  float v = acc + (p.bias ? p.bias[0] : 0.f);
  const int64_t o = ((int64_t)n * p.H + oy) * p.W + ox;
  if (p.mode == 3) {
    if (p.y) p.y[o] = v;
    const float e = clamp_nan(v, -5.f, 5.f);
    const float xn = p.c1 * (p.x_cur[o] - p.c2 * e);
    p.x_next[o] = clamp_nan(xn, 0.f, 1.f);
    return;
  }
  if (p.mode == 1) v += p.inp[o];
  if (p.mode == 2) v = 1.0f / (1.0f + expf(-v));
  if (p.sanitize) v = sanitize01(v);
  p.y[o] = v;
}

void conv_cout1(Ctx& c, const Cout1Args& a) {
  if (conv_cout1_v2_supported(a)) { conv_cout1_v2(c, a); return; }
  const int VN = (int)(16 / dsize(a.x.dt));
  const size_t smem = ((size_t)a.x.c * 324 + 11 * (size_t)a.x.c) * sizeof(float);
  if (a.k != 3 || a.x.c % VN != 0 || smem > 200 * 1024) { conv_cout1_v1(c, a); return; }
  Cout1T p;
  p.x = a.x.p; p.N = a.x.n; p.H = a.x.h; p.W = a.x.w; p.C = a.x.c;
  p.w = a.w; p.bias = a.bias;
  p.gn_sums = a.gn_sums; p.groups = a.groups; p.gamma = a.gamma; p.beta = a.beta; p.eps = a.eps; p.act_in = a.act_in;
  p.mode = a.mode; p.sanitize = a.sanitize; p.inp = a.inp; p.y = a.y; p.x_cur = a.x_cur; p.x_next = a.x_next; p.c1 = a.c1; p.c2 = a.c2;
  dim3 grid(cdiv(a.x.w, 16), cdiv(a.x.h, 16), a.x.n);
  if (c.dry) return;
  switch (a.x.dt) {
    case DT_F32: {
      ensure_dyn_smem(k_conv_cout1_tiled<float, false>, 200 * 1024);
      XRD_LAUNCH(c, (k_conv_cout1_tiled<float, false>), grid, 256, smem, p);
    } break;
    case DT_BF16: {
      ensure_dyn_smem(k_conv_cout1_tiled<__nv_bfloat16, true>, 200 * 1024);
      XRD_LAUNCH(c, (k_conv_cout1_tiled<__nv_bfloat16, true>), grid, 256, smem, p);
    } break;
    case DT_F16: {
      ensure_dyn_smem(k_conv_cout1_tiled<__half, true>, 200 * 1024);
      XRD_LAUNCH(c, (k_conv_cout1_tiled<__half, true>), grid, 256, smem, p);
    } break;
  }
}

// ---------------------------------------------------------------- range audit of 16-bit activations
// The f16 mode stores activations with saturation (common.cuh sat_h): one large value cannot become inf/NaN downstream, but a
// clipped value is silently wrong.  The audit scans a tensor after its producer and counts elements sitting exactly at the
// saturation value, non-finite elements, and the largest finite magnitude, so that a caller (tests, a periodic production
// check) can prove the activations of a checkpoint stay inside the f16 range -- or see exactly how often they do not.
template <typename T>
__global__ void k_range_audit(const uint4* __restrict__ x, size_t nvec, RangeAudit* a) {
  unsigned int sat = 0, bad = 0;
  float amax = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 t = x[i];
    const T* h = reinterpret_cast<const T*>(&t);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = fabsf(ldf<T>(h + j));
      if (!(v <= 3.0e38f)) { ++bad; continue; }          // NaN or inf
      if (sizeof(T) == 2 && std::is_same<T, __half>::value && v >= 65504.f) ++sat;
      amax = fmaxf(amax, v);
    }
  }
#pragma unroll
  for (int of = 16; of > 0; of >>= 1) {
    sat += __shfl_xor_sync(0xffffffffu, sat, of);
    bad += __shfl_xor_sync(0xffffffffu, bad, of);
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, of));
  }
  if ((threadIdx.x & 31) == 0) {
    if (sat) atomicAdd(&a->saturated, (unsigned long long)sat);
    if (bad) atomicAdd(&a->nonfinite, (unsigned long long)bad);
    atomicMax(&a->absmax_bits, __float_as_uint(amax));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { atomicAdd(&a->elements, (unsigned long long)nvec * 8ull); atomicAdd(&a->tensors, 1u); }
}

void range_audit(Ctx& c, const Tens& t) {
  if (!c.audit || c.dry || t.dt == DT_F32 || !t.p) return;
  const size_t n = t.numel();
  if (n == 0 || n % 8 != 0) return;                       // every internal tensor has a multiple of 8 channels
  const size_t nvec = n / 8;
  const int grid = (int)std::min<size_t>((nvec + 255) / 256, 148 * 8);
  if (t.dt == DT_F16) XRD_LAUNCH(c, k_range_audit<__half>, grid, 256, 0, (const uint4*)t.p, nvec, c.audit);
  else XRD_LAUNCH(c, k_range_audit<__nv_bfloat16>, grid, 256, 0, (const uint4*)t.p, nvec, c.audit);
}

}  // namespace xrd
