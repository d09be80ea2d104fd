// kernels.cuh -- host-side launchers of every kernel in libxrd.
#pragma once
#include "common.cuh"

namespace xrd {

// Packed convolution weights (device memory owned by the handle).
struct ConvW {
  int kh = 0, kw = 0, stride = 1, pad = 0;
  int cin = 0;             // total input channels (both sources of a virtual concat)
  int cout = 0;            // output channels as computed by the GEMM (4*cf for depth-to-space)
  int d2s = 0;             // 1: GEMM column q=(i*2+j)*cf+c is stored at pixel (2h+i,2w+j), channel c
  float* w = nullptr;      // [kh*kw][cin][cout] fp32 (CUDA-core kernel + fp32 check mode)
  float* bias = nullptr;   // [cout] fp32 or null
  // tcgen05 packing: [n_kblocks][cout_pad][64] bf16/f16, K-major, built per operand dtype
  void* wtc[3] = {nullptr, nullptr, nullptr};
  int tc_c1 = -1;          // channel split the tc packing was built for (source-1 channels)
  int tc_nkb = 0;          // number of 64-wide K blocks
  int tc_npad = 0;         // cout rounded up to a multiple of 16
  // pre-swizzled second copy of the packing (same shape), for 1-D bulk loads of whole K blocks
  const void* wtc_swz(DType dt) const { return (const char*)wtc[dt] + (size_t)tc_nkb * tc_npad * 128; }
};

struct ConvEpi {
  const float* chan_add = nullptr;  // per-(n,cout) additive term (time embedding), row n at n*chan_add_bstride
  int chan_add_bstride = 0;
  const float* in_scale = nullptr;  // per-(n,cin) multiplicative scale on the input (NAFNet SCA), [N][cin]
  const float* out_scale = nullptr; // per-cout scale applied to (acc+bias) (NAFNet beta/gamma)
  Tens resid;                       // optional residual with the output's layout
  int act = ACT_NONE;
  double* stats_out = nullptr;      // optional [N][8][2] += (sum, sum of squares) of the stored output over 8 channel groups
  bool gate = false;                // SimpleGate on the output row (HYB:119-121): y has cout/2 channels (conv1 only)
  // GroupNorm + activation of the INPUT applied inside the kernel (conv3 only): [N][cin] (0.5*scale, 0.5*shift) from gn_coef()
  const float2* in_coef = nullptr;
  int in_act = ACT_NONE;
};

// --- contractions ---------------------------------------------------------
// CUDA-core implicit GEMM, fp32 accumulate; x2 (nullable) is the second half of a virtual channel concat.
void conv_simt(Ctx& c, const Tens& x1, const Tens* x2, const ConvW& w, const ConvEpi& e, Tens& y);
// tcgen05 implicit GEMM (conv_tc.cu); same contract, bf16/f16 operands.
void conv_tc(Ctx& c, const Tens& x1, const Tens* x2, ConvW& w, const ConvEpi& e, Tens& y);
bool conv_tc_supported(const Tens& x1, const Tens* x2, const ConvW& w, const ConvEpi& e);
bool conv_tc_stats_supported(const ConvW& w);   // can the epilogue emit GroupNorm sums (ConvEpi::stats_out) for this layer
// persistent halo-reusing variant for 3x3/s1/p1 on maps whose width is a multiple of 128 (conv3.cu)
bool conv3_supported(const Tens& x1, const Tens* x2, const ConvW& w, const ConvEpi& e);
void conv3(Ctx& c, const Tens& x1, const Tens* x2, ConvW& w, const ConvEpi& e, Tens& y);
// row-ring 3x3/s1/p1 for Cin <= 64, Cout in {48,96}, W % 128 == 0 (conv3r.cu); honours ConvEpi::in_coef (GroupNorm+SiLU of the input)
bool conv3r_supported(const Tens& x1, const Tens* x2, const ConvW& w, const ConvEpi& e);
void conv3r(Ctx& c, const Tens& x, ConvW& w, const ConvEpi& e, Tens& y);
// N-stacked row-ring 3x3/s1/p1 (conv3s.cu): the three vertical taps share one N=144 MMA per input row; Cout in {48,96} as slices of
// 48, input = one tensor of <= 128 channels or a virtual concat of two <= 64-channel tensors; honours ConvEpi::in_coef
bool conv3s_supported(const Tens& x1, const Tens* x2, const ConvW& w, const ConvEpi& e);
void conv3s(Ctx& c, const Tens& x1, const Tens* x2, ConvW& w, const ConvEpi& e, Tens& y);
// 3x3/s1/p1 on maps exactly 64 pixels wide (conv3w.cu): three column-shifted copies, 4-row tiles, Cout in {144,192}
bool conv3w_supported(const Tens& x1, const Tens* x2, const ConvW& w, const ConvEpi& e);
void conv3w(Ctx& c, const Tens& x1, const Tens* x2, ConvW& w, const ConvEpi& e, Tens& y);
// persistent 1x1 GEMM with resident weights (conv1.cu): res_conv / qkv / proj of the UNet
bool conv1_supported(const Tens& x1, const Tens* x2, const ConvW& w, const ConvEpi& e);
void conv1(Ctx& c, const Tens& x1, const Tens* x2, ConvW& w, const ConvEpi& e, Tens& y);
// build w.wtc[dt] from w.w (device side); c1 = channels of source 1
void conv_tc_pack(cudaStream_t s, ConvW& w, DType dt, int c1);

// direct 3x3/s1/p1 conv for Cin <= 4 read from fp32 planes (first conv of every network)
bool conv_smallcin_supported(const Tens& x1, const Tens* x2, const ConvW& w);
void conv_smallcin(Ctx& c, const Tens& x1, const Tens* x2, const ConvW& w, Tens& y);
// one pixel per thread, all output channels (Cin <= 3, Cout in {32,48}); stats: optional GroupNorm sums [N][8][2] of y
bool conv_smallcin2_supported(const Tens& x1, const Tens* x2, const ConvW& w);
bool conv_smallcin2_stats_supported(const Tens& x1);
void conv_smallcin2(Ctx& c, const Tens& x1, const Tens* x2, const ConvW& w, Tens& y, double* stats);

// softmax(q^T k / sqrt(d)) v; qkv [N,HW,3*heads*d] (channel = s*heads*d + head*d + j), out [N,HW,heads*d]
void attention_simt(Ctx& c, const Tens& qkv, int heads, Tens& out);
// `scratch`: attention_tc_scratch_floats() floats of workspace for the split-key launch shape of single images (null or a zero count:
// every CTA walks the whole key range)
void attention_tc(Ctx& c, const Tens& qkv, int heads, Tens& out, float* scratch = nullptr);
size_t attention_tc_scratch_floats(const Tens& qkv, int heads);
bool attention_tc_supported(const Tens& qkv, int heads);

// --- normalisation / elementwise ---------------------------------------------
// sums[N][groups][2] (double) += (sum, sum of squares) over the virtual concat [x1 | x2]; caller zeroes sums.
void gn_stats(Ctx& c, const Tens& x1, const Tens* x2, int groups, double* sums);
void gn_act(Ctx& c, const Tens& x1, const Tens* x2, int groups, const double* sums, const float* gamma,
            const float* beta, float eps, int act, Tens& y);
void zero_async(Ctx& c, void* p, size_t bytes);
// range audit (no-op unless c.audit is set and t is a 16-bit tensor): counts saturated / non-finite elements, tracks max |v|
void range_audit(Ctx& c, const Tens& t);
// coef[n][c] = 0.5 * (rstd*gamma[c], beta[c] - mean*rstd*gamma[c]) for GroupNorm(groups, C) with the given sums ([N][groups][2])
void gn_coef(Ctx& c, const double* sums, const float* gamma, const float* beta, float eps, int N, int C, int groups, int HW, float2* coef);
// out[n][g] = a[n][2g] + a[n][2g+1] (g < 4), b[n][2(g-4)] + b[n][2(g-4)+1] (g >= 4): sums of GroupNorm(8, 2C) over [a | b]
void gn_merge_stats(Ctx& c, const double* a, const double* b, double* out, int N);
// col[n,h,w, 0..31] = 3x3 neighbourhood (zero padded) of the two fp32 planes a, b: column tap*2 + {0: a, 1: b}, columns 18..31 zero
void im2col_3x3_2ch(Ctx& c, const float* a, const float* b, Tens& col);
// first_conv.cu: in_conv on cat([x, condition]) in one pass (warp-level MMA on an in-smem 16-bit copy of the two planes)
bool first_conv_mma_supported(const Tens& y, const ConvW& w);
void first_conv_mma(Ctx& c, const float* a, const float* b, const ConvW& w, Tens& y, double* stats);
void upsample2x(Ctx& c, const Tens& x, Tens& y);       // bilinear, align_corners=False, exact 2x
// same, 16 bytes per access, fused with the GroupNorm sums of y (stats nullable, [N][8][2], zeroed by the caller)
void upsample2x_stats(Ctx& c, const Tens& x, Tens& y, double* stats);
void layernorm(Ctx& c, const Tens& x, const float* g, const float* b, float eps, Tens& y);
// depthwise 3x3 (pad 1) on u[...,2C] -> SimpleGate -> g[...,C]; pool[N][C] += spatial sums (caller zeroes)
void dwconv_gate_pool(Ctx& c, const Tens& u, const float* w9, const float* bias, Tens& g, float* pool);
// scale[n][c] = b[c] + sum_j W[c][j] * pool[n][j] / HW
void sca_scale(Ctx& c, const float* pool, int N, int C, int HW, const float* W, const float* b, float* scale);

void simple_gate(Ctx& c, const Tens& u, Tens& g);        // g = u[..., :C] * u[..., C:]
void maxpool2x2(Ctx& c, const Tens& x, Tens& y);         // nn.MaxPool2d(2): y[n,h,w,c] = max of the 2x2 block (H, W even; NaN propagates)
void scale_nc(Ctx& c, Tens& x, const float* scale);      // x[n,h,w,c] *= scale[n][c]   (in place)
void interleave3(Ctx& c, const float* a, const float* b, const float* m, float* y, int64_t n);
void pad_crop_plane(Ctx& c, const float* src, float* dst, int N, int Hs, int Ws, int Hd, int Wd);

// --- single-output-channel convolutions with fused prologue/epilogue -----------
struct Cout1Args {
  Tens x;                          // [N,H,W,C] features
  int k = 1;                       // 1 or 3 (stride 1, pad k/2)
  const float* w = nullptr;        // [k*k][C]
  const float* bias = nullptr;     // 1 element
  const double* gn_sums = nullptr; // optional fused GroupNorm + activation on the input
  int groups = 0;
  const float* gamma = nullptr;
  const float* beta = nullptr;
  float eps = 1e-5f;
  int act_in = ACT_NONE;
  int mode = 0;                    // 0 plain, 1 "+ inp", 2 sigmoid, 3 sampler update
  int sanitize = 0;                // nan_to_num + clamp(0,1) on the result (modes 0-2)
  const float* inp = nullptr;      // mode 1
  float* y = nullptr;              // modes 0-2: (N,H,W) fp32 result; mode 3: optional raw eps tap
  const float* x_cur = nullptr;    // mode 3: current sampler state
  float* x_next = nullptr;         // mode 3: new sampler state (may alias x_cur)
  float c1 = 0.f, c2 = 0.f;        // mode 3: 1/sqrt(alpha_t), (1-alpha_t)/sqrt(1-alpha_hat_t)
};
void conv_cout1(Ctx& c, const Cout1Args& a);
bool conv_cout1_v2_supported(const Cout1Args& a);     // C in {32,48}, 3x3: per-pixel tap dot products + smem gather (kernels_edge.cu)
void conv_cout1_v2(Ctx& c, const Cout1Args& a);

// --- time embedding -----------------------------------------------------------
struct TimeEmbW {
  int mc = 0, ted = 0, total = 0;   // model_channels, time_emb_dim, sum of out_c over all ResidualBlocks
  const float *w1 = nullptr, *b1 = nullptr, *w2 = nullptr, *b2 = nullptr;  // time_mlp.1, time_mlp.3
  const float *wall = nullptr, *ball = nullptr;                            // concatenated per-block Linear(ted,out_c)
};
// out[r][total] for r < rows: t value = t_i64 ? t_i64[r] : t_list[r]
void time_embed(Ctx& c, const TimeEmbW& w, const int64_t* t_i64, const int* t_list, int rows, float* out);

// --- layout helpers (op hooks, image planes) -------------------------------------
void nchw_to_nhwc(Ctx& c, const float* x, Tens& y);   // fp32 NCHW -> NHWC of y.dt
void nhwc_to_nchw(Ctx& c, const Tens& x, float* y);
void sanitize_plane(Ctx& c, const float* x, float* y, size_t n);
void copy_plane(Ctx& c, const float* x, float* y, size_t n);

// --- weight packing (device side, fp32) ----------------------------------------------
// (Cout,Cin,kh,kw) -> [kh*kw][Cin][Cout]
void pack_conv_weight(cudaStream_t s, const float* w, float* out, int cout, int cin, int kh, int kw);
// ConvTranspose2d(4,2,1) (Cin,Cout,4,4) followed by 2x2 mean  ->  3x3 conv [9][Cin][Cout]
void pack_convT4_avg_weight(cudaStream_t s, const float* w, float* out, int cin, int cout);
// ConvTranspose2d(2,2) (Cin,Cout,2,2) -> [1][Cin][4*Cout] with q=(i*2+j)*Cout+co ; bias -> [4*Cout]
void pack_convT2_weight(cudaStream_t s, const float* w, const float* b, float* out, float* bout, int cin, int cout);
// 1x1 conv feeding PixelShuffle(2): (4*Cf,Cin,1,1) with q_ref=c*4+i*2+j -> [1][Cin][4*Cf] with q=(i*2+j)*Cf+c
void pack_pixelshuffle_weight(cudaStream_t s, const float* w, float* out, int cin, int cf);
// eval-mode BatchNorm2d folded into the bias-free conv in front of it (DirectUNetModel.py:163-165 etc.):
// wf[co][...] = w[co][...] * g[co]/sqrt(var[co]+eps),  bf[co] = beta[co] - mean[co] * g[co]/sqrt(var[co]+eps)
void fold_bn_weight(cudaStream_t s, const float* w, const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                    float* wf, float* bf, int cout, int per_cout);
// depthwise (C2,1,3,3) -> [9][C2]
void pack_dw_weight(cudaStream_t s, const float* w, float* out, int c2);

}  // namespace xrd
