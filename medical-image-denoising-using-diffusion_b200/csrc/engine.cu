// engine.cu -- weight ingestion/packing and the four networks of the /denoise hot path,
// expressed as launch sequences over the kernels in kernels_simt.cu / conv_tc.cu / attn_tc.cu.
#include <functional>
#include "engine.cuh"

#include <stdlib.h>

#include <math.h>
#include <string.h>
#include <algorithm>

namespace xrd {

// ------------------------------------------------------------------------------------------------
// handle utilities
// ------------------------------------------------------------------------------------------------
void* Handle::dalloc(size_t bytes) {
  void* p = nullptr;
  XRD_CUDA(cudaMalloc(&p, std::max<size_t>(bytes, 16)));
  owned.push_back(p);
  return p;
}
float* Handle::dalloc_f(size_t n) { return (float*)dalloc(n * sizeof(float)); }

void Handle::drop_graphs() {
  for (auto& kv : graphs) {
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    if (kv.second.temb) cudaFree(kv.second.temb);
  }
  graphs.clear();
}

static void free_convw_tc(ConvW& w) {
  for (int i = 0; i < 3; ++i)
    if (w.wtc[i]) { cudaFree(w.wtc[i]); w.wtc[i] = nullptr; }
  w.tc_c1 = -1;
}

void Handle::free_owned() {
  drop_graphs();
  free_convw_tc(unet.in_conv_g);
  auto fr = [](ResW& r) { free_convw_tc(r.c1); free_convw_tc(r.c2); free_convw_tc(r.rc); free_convw_tc(r.c1s); };
  for (auto& r : unet.res) fr(r);
  for (auto& a : unet.attn) { free_convw_tc(a.qkv); free_convw_tc(a.proj); }
  for (auto& d : unet.down) free_convw_tc(d);
  for (auto& u : unet.up) free_convw_tc(u);
  auto fb = [](NafBlockW& b) { free_convw_tc(b.c1); free_convw_tc(b.c3); free_convw_tc(b.c4); free_convw_tc(b.c5); };
  for (auto& s : naf.enc) for (auto& b : s) fb(b);
  for (auto& s : naf.dec) for (auto& b : s) fb(b);
  for (auto& b : naf.mid) fb(b);
  for (ConvW* w : {&expert.inc[0], &expert.inc[1], &expert.down1[0], &expert.down1[1], &expert.down2[0], &expert.down2[1], &expert.bott[0],
                   &expert.bott[1], &expert.upc2[0], &expert.upc2[1], &expert.upc1[0], &expert.upc1[1], &expert.fin, &expert.up2, &expert.up1})
    free_convw_tc(*w);
  for (auto& w : naf.downs) free_convw_tc(w);
  for (auto& w : naf.ups) free_convw_tc(w);
  for (auto& w : naf.skips) free_convw_tc(w);
  for (void* p : owned) cudaFree(p);
  owned.clear();
  unet = UNetW(); naf = NafW(); router = RouterW(); fusion = FusionW();
}

const Param& Handle::P(const std::string& key) const {
  auto it = params.find(key);
  if (it == params.end()) fail(XRD_ERR_MISSING_PARAM, "missing state_dict tensor '%s'", key.c_str());
  return it->second;
}

static void expect_shape(const Param& p, const std::string& key, std::initializer_list<int64_t> shape) {
  std::vector<int64_t> s(shape);
  if (p.shape != s) {
    std::string got, want;
    for (auto v : p.shape) got += std::to_string(v) + ",";
    for (auto v : s) want += std::to_string(v) + ",";
    fail(XRD_ERR_INVALID, "tensor '%s' has shape (%s) but the configuration needs (%s)", key.c_str(), got.c_str(), want.c_str());
  }
}

// ------------------------------------------------------------------------------------------------
// weight packing
// ------------------------------------------------------------------------------------------------
static ConvW make_conv(Handle& h, const std::string& key, int cout, int cin, int k, int stride, int pad, bool bias = true) {
  const Param& w = h.P(key + ".weight");
  expect_shape(w, key + ".weight", {cout, cin, k, k});
  ConvW c;
  c.kh = c.kw = k; c.stride = stride; c.pad = pad; c.cin = cin; c.cout = cout;
  c.w = h.dalloc_f((size_t)cout * cin * k * k);
  pack_conv_weight(nullptr, w.d, c.w, cout, cin, k, k);
  if (bias) {
    const Param& b = h.P(key + ".bias");
    expect_shape(b, key + ".bias", {cout});
    c.bias = b.d;
  }
  return c;
}

static void finalize_unet(Handle& h) {
  const xrd_config& cf = h.cfg;
  const std::string p = cf.unet_prefix;
  UNetW& u = h.unet;
  u = UNetW();
  XRD_REQUIRE(cf.unet_in_channels == 1, "UNet: only in_channels=1 (grayscale) is implemented, got %d", cf.unet_in_channels);
  const int mc = cf.unet_model_channels, ted = cf.unet_time_emb_dim, levels = cf.unet_n_levels;
  XRD_REQUIRE(mc % 16 == 0 && mc >= 16, "UNet: model_channels must be a multiple of 16");
  u.mc = mc; u.heads = cf.unet_num_heads;
  auto in_attn = [&](int lvl) {
    for (int i = 0; i < cf.unet_n_attn; ++i) if (cf.unet_attention_resolutions[i] == lvl) return true;
    return false;
  };
  std::vector<std::pair<std::string, int>> temb_blocks;  // (prefix, out_c) in creation order
  int temb_total = 0;
  auto add_res = [&](const std::string& q, int cin, int cout) {
    ResW r;
    r.cin = cin; r.cout = cout; r.temb_off = temb_total; temb_total += cout;
    temb_blocks.push_back({q, cout});
    r.g1 = h.PD(q + "block1.0.weight"); r.b1 = h.PD(q + "block1.0.bias");
    expect_shape(h.P(q + "block1.0.weight"), q + "block1.0.weight", {cin});
    r.c1 = make_conv(h, q + "block1.2", cout, cin, 3, 1, 1);
    r.g2 = h.PD(q + "block2.0.weight"); r.b2 = h.PD(q + "block2.0.bias");
    r.c2 = make_conv(h, q + "block2.3", cout, cout, 3, 1, 1);
    r.has_rc = cin != cout;
    if (r.has_rc) r.rc = make_conv(h, q + "res_conv", cout, cin, 1, 1, 0);
    expect_shape(h.P(q + "time_mlp.1.weight"), q + "time_mlp.1.weight", {cout, ted});
    u.res.push_back(r);
    return (int)u.res.size() - 1;
  };
  auto add_attn = [&](const std::string& q, int c) {
    AttnW a;
    a.c = c;
    a.g = h.PD(q + "norm.weight"); a.b = h.PD(q + "norm.bias");
    a.qkv = make_conv(h, q + "qkv", 3 * c, c, 1, 1, 0);
    a.proj = make_conv(h, q + "proj", c, c, 1, 1, 0);
    u.attn.push_back(a);
    return (int)u.attn.size() - 1;
  };

  {  // in_conv: (mc, 2, 3, 3)
    u.in_conv = make_conv(h, p + "in_conv", mc, 2, 3, 1, 1);
    // the same weights as a [32 x mc] GEMM operand: row k = tap*2 + channel for k < 18 (exactly the [tap][cin][cout] layout
    // of in_conv.w), rows 18..31 zero
    u.in_conv_g = ConvW();
    u.in_conv_g.kh = u.in_conv_g.kw = 1; u.in_conv_g.stride = 1; u.in_conv_g.pad = 0; u.in_conv_g.cin = 32; u.in_conv_g.cout = mc;
    u.in_conv_g.w = h.dalloc_f((size_t)32 * mc);
    XRD_CUDA(cudaMemset(u.in_conv_g.w, 0, (size_t)32 * mc * sizeof(float)));
    XRD_CUDA(cudaMemcpy(u.in_conv_g.w, u.in_conv.w, (size_t)18 * mc * sizeof(float), cudaMemcpyDeviceToDevice));
    u.in_conv_g.bias = u.in_conv.bias;
  }
  int ch = mc, li = 0;
  for (int lvl = 0; lvl < levels; ++lvl) {
    const int oc = mc * cf.unet_channel_mult[lvl];
    for (int r = 0; r < cf.unet_num_res_blocks; ++r) {
      u.downs.push_back({U_RES, add_res(p + "downs." + std::to_string(li++) + ".", ch, oc)});
      ch = oc;
      if (in_attn(lvl)) u.downs.push_back({U_ATTN, add_attn(p + "downs." + std::to_string(li++) + ".", ch)});
    }
    if (lvl != levels - 1) {
      u.down.push_back(make_conv(h, p + "downs." + std::to_string(li++), ch, ch, 3, 2, 1));
      u.downs.push_back({U_DOWN, (int)u.down.size() - 1});
    }
  }
  u.mid1 = add_res(p + "mid_block1.", ch, ch);
  u.mid_attn = add_attn(p + "mid_attn.", ch);
  u.mid2 = add_res(p + "mid_block2.", ch, ch);
  li = 0;
  for (int lvl = levels - 1; lvl >= 0; --lvl) {
    const int oc = mc * cf.unet_channel_mult[lvl];
    for (int r = 0; r < cf.unet_num_res_blocks + 1; ++r) {
      u.ups.push_back({U_RES, add_res(p + "ups." + std::to_string(li++) + ".", 2 * ch, oc)});
      ch = oc;
      if (in_attn(lvl)) u.ups.push_back({U_ATTN, add_attn(p + "ups." + std::to_string(li++) + ".", ch)});
    }
    if (lvl != 0) {
      // ConvTranspose2d(ch,ch,4,2,1): always consumed through an exact 0.5x bilinear resize (HYB:381-382), which
      // equals a 2x2 mean; the pair is pre-combined into one 3x3 convolution on the input grid (fp32, before any
      // 16-bit rounding).
      const std::string q = p + "ups." + std::to_string(li++);
      const Param& w = h.P(q + ".weight");
      expect_shape(w, q + ".weight", {ch, ch, 4, 4});
      ConvW c;
      c.kh = c.kw = 3; c.stride = 1; c.pad = 1; c.cin = ch; c.cout = ch;
      c.w = h.dalloc_f((size_t)9 * ch * ch);
      pack_convT4_avg_weight(nullptr, w.d, c.w, ch, ch);
      c.bias = h.P(q + ".bias").d;
      u.up.push_back(c);
      u.ups.push_back({U_UP, (int)u.up.size() - 1});
    }
  }
  // out_conv: GN(8,ch) + SiLU + conv3x3 ch->1
  u.out_c = ch;
  u.og = h.PD(p + "out_conv.0.weight"); u.ob = h.PD(p + "out_conv.0.bias");
  {
    const Param& w = h.P(p + "out_conv.2.weight");
    expect_shape(w, p + "out_conv.2.weight", {1, ch, 3, 3});
    u.ow = h.dalloc_f((size_t)9 * ch);
    pack_conv_weight(nullptr, w.d, u.ow, 1, ch, 3, 3);   // -> [9][ch][1]
    u.obias = h.PD(p + "out_conv.2.bias");
  }
  // time embedding
  expect_shape(h.P(p + "time_mlp.1.weight"), p + "time_mlp.1.weight", {ted, mc});
  expect_shape(h.P(p + "time_mlp.3.weight"), p + "time_mlp.3.weight", {ted, ted});
  float* wall = h.dalloc_f((size_t)temb_total * ted);
  float* ball = h.dalloc_f(temb_total);
  int off = 0;
  for (auto& tb : temb_blocks) {
    XRD_CUDA(cudaMemcpy(wall + (size_t)off * ted, h.PD(tb.first + "time_mlp.1.weight"), (size_t)tb.second * ted * 4, cudaMemcpyDeviceToDevice));
    XRD_CUDA(cudaMemcpy(ball + off, h.PD(tb.first + "time_mlp.1.bias"), (size_t)tb.second * 4, cudaMemcpyDeviceToDevice));
    off += tb.second;
  }
  u.te.mc = mc; u.te.ted = ted; u.te.total = temb_total;
  u.te.w1 = h.PD(p + "time_mlp.1.weight"); u.te.b1 = h.PD(p + "time_mlp.1.bias");
  u.te.w2 = h.PD(p + "time_mlp.3.weight"); u.te.b2 = h.PD(p + "time_mlp.3.bias");
  u.te.wall = wall; u.te.ball = ball;
  u.ready = true;
}

static NafBlockW make_nafblock(Handle& h, const std::string& q, int c) {
  NafBlockW b;
  b.c = c;
  b.n1w = h.PD(q + "norm1.weight"); b.n1b = h.PD(q + "norm1.bias");
  b.n2w = h.PD(q + "norm2.weight"); b.n2b = h.PD(q + "norm2.bias");
  expect_shape(h.P(q + "beta"), q + "beta", {1, c, 1, 1});
  b.beta = h.PD(q + "beta"); b.gamma = h.PD(q + "gamma");
  b.c1 = make_conv(h, q + "conv1", 2 * c, c, 1, 1, 0);
  b.c3 = make_conv(h, q + "conv3", c, c, 1, 1, 0);
  b.c4 = make_conv(h, q + "conv4", 2 * c, c, 1, 1, 0);
  b.c5 = make_conv(h, q + "conv5", c, c, 1, 1, 0);
  expect_shape(h.P(q + "conv2.weight"), q + "conv2.weight", {2 * c, 1, 3, 3});
  b.dw = h.dalloc_f((size_t)9 * 2 * c);
  pack_dw_weight(nullptr, h.PD(q + "conv2.weight"), b.dw, 2 * c);
  b.dwb = h.PD(q + "conv2.bias");
  expect_shape(h.P(q + "sca.1.weight"), q + "sca.1.weight", {c, c, 1, 1});
  b.scaw = h.PD(q + "sca.1.weight"); b.scab = h.PD(q + "sca.1.bias");
  return b;
}

static void finalize_nafnet(Handle& h) {
  const xrd_config& cf = h.cfg;
  const std::string p = cf.naf_prefix;
  NafW& n = h.naf;
  n = NafW();
  XRD_REQUIRE(cf.naf_img_channel == 1, "NAFNet: only img_channel=1 is implemented");
  XRD_REQUIRE(cf.naf_n_enc == cf.naf_n_dec, "NAFNet: enc_blk_nums and dec_blk_nums must have equal length");
  const int width = cf.naf_width;
  XRD_REQUIRE(width >= 16 && (width & (width - 1)) == 0, "NAFNet: width must be a power of two >= 16 (got %d)", width);
  n.width = width;
  n.intro = make_conv(h, p + "intro", width, 1, 3, 1, 1);
  int ch = width;
  for (int s = 0; s < cf.naf_n_enc; ++s) {
    std::vector<NafBlockW> st;
    for (int b = 0; b < cf.naf_enc_blk_nums[s]; ++b) st.push_back(make_nafblock(h, p + "encoders." + std::to_string(s) + "." + std::to_string(b) + ".", ch));
    n.enc.push_back(st);
    n.downs.push_back(make_conv(h, p + "downs." + std::to_string(s), 2 * ch, ch, 2, 2, 0));
    ch *= 2;
  }
  for (int b = 0; b < cf.naf_middle_blk_num; ++b) n.mid.push_back(make_nafblock(h, p + "middle_blks." + std::to_string(b) + ".", ch));
  for (int s = 0; s < cf.naf_n_dec; ++s) {
    // ups: 1x1 ch -> 2ch (no bias) + PixelShuffle(2) -> ch/2 channels at 2x resolution
    const std::string q = p + "ups." + std::to_string(s) + ".0.weight";
    expect_shape(h.P(q), q, {2 * ch, ch, 1, 1});
    ConvW up;
    up.kh = up.kw = 1; up.stride = 1; up.pad = 0; up.cin = ch; up.cout = 2 * ch; up.d2s = 1;
    up.w = h.dalloc_f((size_t)ch * 2 * ch);
    pack_pixelshuffle_weight(nullptr, h.PD(q), up.w, ch, ch / 2);
    n.ups.push_back(up);
    ch /= 2;
    n.skips.push_back(make_conv(h, p + "skip_convs." + std::to_string(s), ch, 2 * ch, 1, 1, 0));
    std::vector<NafBlockW> st;
    for (int b = 0; b < cf.naf_dec_blk_nums[s]; ++b) st.push_back(make_nafblock(h, p + "decoders." + std::to_string(s) + "." + std::to_string(b) + ".", ch));
    n.dec.push_back(st);
  }
  XRD_REQUIRE(ch == width, "NAFNet: decoder does not return to width");
  expect_shape(h.P(p + "ending.weight"), p + "ending.weight", {1, width, 3, 3});
  n.ending_w = h.dalloc_f((size_t)9 * width);
  pack_conv_weight(nullptr, h.PD(p + "ending.weight"), n.ending_w, 1, width, 3, 3);
  n.ending_b = h.PD(p + "ending.bias");
  n.ready = true;
}

static CGG make_cgg(Handle& h, const std::string& q, int cout, int cin, int stride, int groups) {
  CGG c;
  c.conv = make_conv(h, q + "0", cout, cin, 3, stride, 1);
  c.g = h.PD(q + "1.weight"); c.b = h.PD(q + "1.bias");
  c.groups = groups;
  return c;
}
static ConvW make_convT2(Handle& h, const std::string& key, int cin, int cout) {
  expect_shape(h.P(key + ".weight"), key + ".weight", {cin, cout, 2, 2});
  ConvW c;
  c.kh = c.kw = 1; c.stride = 1; c.pad = 0; c.cin = cin; c.cout = 4 * cout; c.d2s = 1;
  c.w = h.dalloc_f((size_t)cin * 4 * cout);
  c.bias = h.dalloc_f((size_t)4 * cout);
  pack_convT2_weight(nullptr, h.PD(key + ".weight"), h.PD(key + ".bias"), c.w, c.bias, cin, cout);
  return c;
}

static void finalize_router(Handle& h) {
  const std::string p = h.cfg.router_prefix;
  const int b = h.cfg.router_base_c;
  XRD_REQUIRE(b % 8 == 0, "router: base_c must be a multiple of 8");
  RouterW& r = h.router;
  r = RouterW();
  r.enc1 = make_cgg(h, p + "enc1.", b, 1, 1, 8);
  r.enc2 = make_cgg(h, p + "enc2.", 2 * b, b, 2, 8);
  r.enc3 = make_cgg(h, p + "enc3.", 4 * b, 2 * b, 2, 8);
  r.mid = make_cgg(h, p + "mid.", 4 * b, 4 * b, 1, 8);
  r.up3 = make_convT2(h, p + "up3", 4 * b, 2 * b);
  r.dec3 = make_cgg(h, p + "dec3.", 2 * b, 4 * b, 1, 8);
  r.up2 = make_convT2(h, p + "up2", 2 * b, b);
  r.dec2 = make_cgg(h, p + "dec2.", b, 2 * b, 1, 8);
  expect_shape(h.P(p + "out_conv.weight"), p + "out_conv.weight", {1, b, 1, 1});
  r.out_w = h.PD(p + "out_conv.weight");
  r.out_b = h.PD(p + "out_conv.bias");
  r.ready = true;
}

static void finalize_fusion(Handle& h) {
  const std::string p = h.cfg.fusion_prefix;
  const int b = h.cfg.fusion_base_c;
  XRD_REQUIRE(b % 8 == 0, "fusion: base_c must be a multiple of 8");
  FusionW& f = h.fusion;
  f = FusionW();
  f.conv1 = make_cgg(h, p + "conv1.", b, 3, 1, 8);
  f.conv2 = make_cgg(h, p + "conv2.", b / 2, b, 1, 4);
  expect_shape(h.P(p + "out_conv.weight"), p + "out_conv.weight", {1, b / 2, 1, 1});
  f.out_w = h.PD(p + "out_conv.weight");
  f.out_b = h.PD(p + "out_conv.bias");
  if (b == 48 || b == 96) {
    const int hb = b / 2;
    auto padded = [&](const float* src, size_t n_src, size_t n_all) {
      float* d = h.dalloc_f(n_all);
      XRD_CUDA(cudaMemset(d, 0, n_all * sizeof(float)));
      XRD_CUDA(cudaMemcpy(d, src, n_src * sizeof(float), cudaMemcpyDeviceToDevice));
      return d;
    };
    const float* wpad = padded(h.PD(p + "conv2.0.weight"), (size_t)hb * b * 9, (size_t)b * b * 9);   // [cout][cin][3][3]: rows hb.. are zero
    ConvW cw;
    cw.kh = cw.kw = 3; cw.stride = 1; cw.pad = 1; cw.cin = b; cw.cout = b;
    cw.w = h.dalloc_f((size_t)b * b * 9);
    pack_conv_weight(nullptr, wpad, cw.w, b, b, 3, 3);
    cw.bias = padded(h.PD(p + "conv2.0.bias"), hb, b);
    f.conv2p.conv = cw;
    f.conv2p.g = padded(f.conv2.g, hb, b);
    f.conv2p.b = padded(f.conv2.b, hb, b);
    f.conv2p.groups = 8;
    f.out_wp = padded(f.out_w, hb, b);
    f.padded = true;
  }
  f.ready = true;
}

// Conv2d(cin, cout, 3, padding=1, bias=False) + BatchNorm2d(cout) in eval mode -> one conv with bias (keys q+"N", q+"N+1")
static ConvW make_conv_bn(Handle& h, const std::string& q, int i, int cout, int cin) {
  const std::string ck = q + std::to_string(i), bk = q + std::to_string(i + 1);
  const Param& w = h.P(ck + ".weight");
  expect_shape(w, ck + ".weight", {cout, cin, 3, 3});
  for (const char* sfx : {".weight", ".bias", ".running_mean", ".running_var"}) expect_shape(h.P(bk + sfx), bk + sfx, {cout});
  float* wf = h.dalloc_f((size_t)cout * cin * 9);
  float* bf = h.dalloc_f(cout);
  fold_bn_weight(nullptr, w.d, h.PD(bk + ".weight"), h.PD(bk + ".bias"), h.PD(bk + ".running_mean"), h.PD(bk + ".running_var"), 1e-5f, wf, bf,
                 cout, cin * 9);
  ConvW c;
  c.kh = c.kw = 3; c.stride = 1; c.pad = 1; c.cin = cin; c.cout = cout;
  c.w = h.dalloc_f((size_t)cout * cin * 9);
  pack_conv_weight(nullptr, wf, c.w, cout, cin, 3, 3);
  c.bias = bf;
  return c;
}

static void finalize_expert(Handle& h) {
  const std::string p = h.cfg.expert_prefix;
  const int b = h.cfg.expert_base_c;
  XRD_REQUIRE(b >= 16 && b % 16 == 0, "ExpertDenoiser: base_channels must be a multiple of 16 (got %d)", b);
  ExpertW& e = h.expert;
  e = ExpertW();
  e.base = b;
  auto pair = [&](ConvW (&dst)[2], const std::string& q, int cin, int cout) {
    dst[0] = make_conv_bn(h, q, 0, cout, cin);
    dst[1] = make_conv_bn(h, q, 3, cout, cout);
  };
  pair(e.inc, p + "inc.", 1, b);
  pair(e.down1, p + "down1.", b, 2 * b);
  pair(e.down2, p + "down2.", 2 * b, 4 * b);
  pair(e.bott, p + "bottleneck.", 4 * b, 8 * b);
  e.up2 = make_convT2(h, p + "up2", 8 * b, 4 * b);
  pair(e.upc2, p + "upconv2.", 8 * b, 4 * b);
  e.up1 = make_convT2(h, p + "up1", 4 * b, 2 * b);
  pair(e.upc1, p + "upconv1.", 4 * b, 2 * b);
  e.fin = make_conv_bn(h, p + "final.", 0, b, 2 * b);
  expect_shape(h.P(p + "outc.weight"), p + "outc.weight", {1, b, 1, 1});
  e.out_w = h.PD(p + "outc.weight");
  e.out_b = h.PD(p + "outc.bias");
  e.ready = true;
}

void finalize(Handle& h, int which) {
  XRD_CUDA(cudaSetDevice(h.device));
  XRD_CUDA(cudaDeviceSynchronize());
  h.free_owned();
  h.unet.ready = h.naf.ready = h.router.ready = h.fusion.ready = h.expert.ready = false;   // free_owned() released every packed plan
  if (which & XRD_PART_UNET) finalize_unet(h);
  if (which & XRD_PART_NAFNET) finalize_nafnet(h);
  if (which & XRD_PART_ROUTER) finalize_router(h);
  if (which & XRD_PART_FUSION) finalize_fusion(h);
  if (which & XRD_PART_EXPERT) finalize_expert(h);
  // sampler coefficient tables (float32 like HYB:396-398)
  const int T = h.cfg.noise_steps;
  h.coef1.assign(T, 0.f); h.coef2.assign(T, 0.f);
  float ah = 1.0f;
  for (int t = 0; t < T; ++t) {
    double bt = T > 1 ? (double)h.cfg.beta_start + ((double)h.cfg.beta_end - (double)h.cfg.beta_start) * t / (double)(T - 1)
                      : (double)h.cfg.beta_start;
    float beta = (float)bt;
    float alpha = 1.0f - beta;
    ah = ah * alpha;
    h.coef1[t] = 1.0f / sqrtf(alpha);
    h.coef2[t] = (1.0f - alpha) / sqrtf(1.0f - ah);
  }
  XRD_CUDA(cudaDeviceSynchronize());
}

// ------------------------------------------------------------------------------------------------
// network building blocks
// ------------------------------------------------------------------------------------------------
static const bool g_fuse_gn = !(getenv("XRD_FUSE_GN") && atoi(getenv("XRD_FUSE_GN")) == 0);
static const bool g_fuse_gn_conv3 = getenv("XRD_FUSE_GN_CONV3") && atoi(getenv("XRD_FUSE_GN_CONV3")) != 0;
// N-stacked row-ring kernel (conv3s.cu): narrowest map it takes over, and where GroupNorm+SiLU is applied inside it
// (0 never, 1 one-chunk inputs only, 2 also two-chunk inputs, whose row ring is only three rows deep)
static const int g_c3s_minw = getenv("XRD_C3S_MINW") ? atoi(getenv("XRD_C3S_MINW")) : 128;
static const int g_c3s_gn = getenv("XRD_C3S_GN") ? atoi(getenv("XRD_C3S_GN")) : 2;
static bool use_conv3s(const Tens& x1, const Tens* x2, const ConvW& w, const ConvEpi& e) {
  return x1.w >= g_c3s_minw && conv3s_supported(x1, x2, w, e);
}

// A contraction.  e.stats_out (optional, [N][8][2], zeroed here) receives the GroupNorm sums of the output: from the
// kernel's own epilogue where it has one, else from a statistics pass over the stored tensor.
static void conv_impl(Ctx& c, const Tens& x1, const Tens* x2, ConvW& w, const ConvEpi& e, Tens& y);
static void conv(Ctx& c, const Tens& x1, const Tens* x2, ConvW& w, const ConvEpi& e, Tens& y) {
  conv_impl(c, x1, x2, w, e, y);
  range_audit(c, y);
}
static void conv_impl(Ctx& c, const Tens& x1, const Tens* x2, ConvW& w, const ConvEpi& e, Tens& y) {
  // e.stats_out comes from stats16(): already zero
  // row-ring kernel: measured faster than the halo kernel from 256 columns up (1.06x activation fetch instead of 2x), slower
  // at 128 (too few 32-row work items per SM)
  if (c.tc && use_conv3s(x1, x2, w, e)) { conv3s(c, x1, x2, w, e, y); return; }
  if (c.tc && x1.w >= 256 && conv3r_supported(x1, x2, w, e)) { conv3r(c, x1, w, e, y); return; }
  if (c.tc && conv3_supported(x1, x2, w, e)) { conv3(c, x1, x2, w, e, y); return; }
  if (c.tc && conv3w_supported(x1, x2, w, e)) { conv3w(c, x1, x2, w, e, y); return; }
  if (c.tc && conv1_supported(x1, x2, w, e)) { conv1(c, x1, x2, w, e, y); return; }
  XRD_REQUIRE(!e.gate, "conv: the fused SimpleGate epilogue exists in conv1 only (caller must check conv1_supported)");
  if (c.tc && conv_tc_supported(x1, x2, w, e)) { conv_tc(c, x1, x2, w, e, y); return; }
  ConvEpi e2 = e;
  e2.stats_out = nullptr;
  if (c.tc && conv_tc_supported(x1, x2, w, e2)) conv_tc(c, x1, x2, w, e2, y);
  else conv_simt(c, x1, x2, w, e2, y);
  if (e.stats_out) gn_stats(c, y, nullptr, 8, e.stats_out);
}

// first conv of a network (1..3 fp32 input planes).  stats (nullable, [N][8][2], zero on entry): GroupNorm sums of y.
// Returns false when the statistics were not produced (the caller then runs a statistics pass).
static bool conv_first(Ctx& c, const Tens& x1, const Tens* x2, ConvW& w, Tens& y, double* stats = nullptr) {
  if (conv_smallcin2_supported(x1, x2, w)) {
    const bool st = stats && conv_smallcin2_stats_supported(x1);
    conv_smallcin2(c, x1, x2, w, y, st ? stats : nullptr);
    return st;
  }
  if (conv_smallcin_supported(x1, x2, w)) conv_smallcin(c, x1, x2, w, y);
  else conv_simt(c, x1, x2, w, ConvEpi(), y);
  return false;
}

static double* new_sums(Ctx& c, int n, int groups) {
  const size_t cnt = (size_t)n * groups * 2;
  if (double* z = c.alloc_zeroed(cnt)) return z;
  double* s = c.allocd(cnt);
  zero_async(c, s, cnt * sizeof(double));
  return s;
}
// zeroed [n][8][2] GroupNorm sums for a producer's epilogue
static double* stats16(Ctx& c, int n) { return new_sums(c, n, 8); }

// An activation tensor together with the GroupNorm sums of its 8 channel groups ([N][8][2] doubles; null = unknown).
// Producers fill them in their epilogues so that the consuming GroupNorm (HYB:264,269,288,354) needs no pass of its own.
struct TS {
  Tens t;
  double* st = nullptr;
};

// sums over the 8 groups of GroupNorm(8, C1 + C2) applied to the virtual concat [x1 | x2]
static double* concat_stats(Ctx& c, const TS& x1, const TS* x2, int groups) {
  const int B = x1.t.n;
  if (!x2) {
    if (x1.st && groups == 8) return x1.st;
    double* s = new_sums(c, B, groups);
    gn_stats(c, x1.t, nullptr, groups, s);
    return s;
  }
  if (x1.st && x2->st && groups == 8 && x1.t.c == x2->t.c) {
    double* s = c.allocd((size_t)B * 16);   // fully overwritten
    gn_merge_stats(c, x1.st, x2->st, s, B);      // groups 0..3 = pairs of x1's groups, 4..7 = pairs of x2's
    return s;
  }
  double* s = new_sums(c, B, groups);
  gn_stats(c, x1.t, &x2->t, groups, s);
  return s;
}

static void resblock(Ctx& c, UNetW& u, ResW& r, const TS& x1, const TS* x2, const float* temb, int temb_bstride, TS& out) {
  // `out.t` is allocated by the caller; out.st is allocated here (it must outlive this block's scratch)
  const int B = x1.t.n, H = x1.t.h, W = x1.t.w;
  out.st = stats16(c, B);
  const size_t mk = c.a->mark();
  const Tens* x2t = x2 ? &x2->t : nullptr;
  double* s1 = concat_stats(c, x1, x2, u.groups);
  Tens hm = c.alloc(B, H, W, r.cout);
  double* s2 = stats16(c, B);
  ConvEpi e1;
  e1.chan_add = temb + r.temb_off; e1.chan_add_bstride = temb_bstride;
  e1.stats_out = s2;
  // GroupNorm + SiLU of the input applied inside the conv (on the landed shared-memory stage) where the kernel can:
  // the activated tensor -- and for the up path the concatenated one -- is then never written to HBM.
  auto fused_in = [&](const Tens& a, const Tens* b, ConvW& w, const ConvEpi& e, const double* sums, const float* g, const float* bt,
                      Tens& y) -> bool {
    ConvEpi pe = e;
    pe.in_coef = reinterpret_cast<const float2*>(sums);      // non-null probe
    pe.in_act = ACT_SILU;
    if (!c.tc || !g_fuse_gn) return false;
    // the row-ring kernel transforms every row once under a deep ring; in the 2-stage halo kernel the transform sits on
    // the load -> MMA critical path (measured: slower than the stand-alone pass), so conv3 only fuses on request
    // measured (tools/gn_fuse_time.py, batch 16): 48->48 @512^2 fused 430 us vs 295 + 172 us separate; at 256^2 it is a wash
    const bool two_chunks = b != nullptr || a.c > 64;
    const bool stack = g_c3s_gn >= (two_chunks ? 2 : 1) && use_conv3s(a, b, w, pe) && !(b && (!w.wtc[a.dt] || w.tc_c1 != a.c));
    const bool ring = !stack && a.w >= 512 && conv3r_supported(a, b, w, pe);
    const bool halo = !stack && !ring && g_fuse_gn_conv3 && conv3_supported(a, b, w, pe) && !(b && (!w.wtc[a.dt] || w.tc_c1 != a.c));
    if (!stack && !ring && !halo) return false;
    const int ct = a.c + (b ? b->c : 0);
    float2* cf = (float2*)c.a->alloc((size_t)B * ct * sizeof(float2));
    gn_coef(c, sums, g, bt, 1e-5f, B, ct, u.groups, H * W, cf);
    pe.in_coef = cf;
    if (stack) conv3s(c, a, b, w, pe, y); else if (ring) conv3r(c, a, w, pe, y); else conv3(c, a, b, w, pe, y);
    range_audit(c, y);
    return true;
  };
  ConvW& w1 = x2 ? r.c1s : r.c1;
  if (!fused_in(x1.t, x2t, w1, e1, s1, r.g1, r.b1, hm)) {
    Tens a1 = c.alloc(B, H, W, r.cin);
    gn_act(c, x1.t, x2t, u.groups, s1, r.g1, r.b1, 1e-5f, ACT_SILU, a1);
    range_audit(c, a1);
    conv(c, a1, nullptr, r.c1, e1, hm);
  }
  ConvEpi e2;
  if (r.has_rc) {
    Tens rr = c.alloc(B, H, W, r.cout);
    conv(c, x1.t, x2t, r.rc, ConvEpi(), rr);
    e2.resid = rr;
  } else {
    XRD_REQUIRE(!x2, "resblock: identity skip with a concatenated input");
    e2.resid = x1.t;
  }
  e2.stats_out = out.st;
  if (!fused_in(hm, nullptr, r.c2, e2, s2, r.g2, r.b2, out.t)) {
    Tens a2 = c.alloc(B, H, W, r.cout);
    gn_act(c, hm, nullptr, u.groups, s2, r.g2, r.b2, 1e-5f, ACT_SILU, a2);
    range_audit(c, a2);
    conv(c, a2, nullptr, r.c2, e2, out.t);
  }
  c.a->release(mk);
}

static void attention(Ctx& c, const Tens& qkv, int heads, Tens& o) {
  if (c.tc && attention_tc_supported(qkv, heads)) {
    const size_t nf = attention_tc_scratch_floats(qkv, heads);       // released with the block's other temporaries
    attention_tc(c, qkv, heads, o, nf ? c.allocf(nf) : nullptr);
  } else {
    attention_simt(c, qkv, heads, o);
  }
}

static void attnblock(Ctx& c, UNetW& u, AttnW& a, const TS& x, TS& out) {
  const int B = x.t.n, H = x.t.h, W = x.t.w;
  out.st = stats16(c, B);
  const size_t mk = c.a->mark();
  double* s = concat_stats(c, x, nullptr, u.groups);
  Tens xn = c.alloc(B, H, W, a.c);
  gn_act(c, x.t, nullptr, u.groups, s, a.g, a.b, 1e-5f, ACT_NONE, xn);
  range_audit(c, xn);
  Tens qkv = c.alloc(B, H, W, 3 * a.c);
  conv(c, xn, nullptr, a.qkv, ConvEpi(), qkv);
  Tens o = c.alloc(B, H, W, a.c);
  attention(c, qkv, u.heads, o);
  range_audit(c, o);
  ConvEpi e;
  e.resid = x.t;
  e.stats_out = out.st;
  conv(c, o, nullptr, a.proj, e, out.t);
  c.a->release(mk);
}

struct UNetOut {           // what the fused out_conv kernel does with eps
  int mode = 0;            // 0: write eps to y; 3: sampler update
  float* y = nullptr;      // mode 0: eps; mode 3: optional raw-eps tap
  const float* x_cur = nullptr;
  float* x_next = nullptr;
  float c1 = 0, c2 = 0;
};

// One UNet evaluation (HYB:359-388).  x, cond: (B,H,W) fp32 planes.  temb: [B or 1][total] table row(s).
static void unet_eval(Ctx& c, UNetW& u, const float* x, const float* cond, const float* temb, int temb_bstride, int B, int H, int W,
                      const UNetOut& o) {
  const size_t mk0 = c.a->mark();
  {   // every GroupNorm-sum buffer of this evaluation comes out of one pre-zeroed pool (one memset node)
    const size_t cnt = (size_t)96 * B * 16;
    c.zpool = c.allocd(cnt);
    c.zpool_off = 0; c.zpool_cap = cnt;
    zero_async(c, c.zpool, cnt * sizeof(double));
  }
  Tens xin; xin.p = (void*)x; xin.n = B; xin.h = H; xin.w = W; xin.c = 1; xin.dt = DT_F32;
  Tens cin = xin; cin.p = (void*)cond;
  TS h;
  h.t = c.alloc(B, H, W, u.mc);
  {
    double* st = stats16(c, B);
    ConvEpi e0;
    e0.stats_out = st;
    Tens col;
    col.n = B; col.h = H; col.w = W; col.c = 32; col.dt = c.adt;
    if (c.tc && first_conv_mma_supported(h.t, u.in_conv)) {
      // one pass: fp32 planes in, 16-bit NHWC activation and its GroupNorm sums out (first_conv.cu)
      first_conv_mma(c, x, cond, u.in_conv, h.t, st);
      h.st = st;
      range_audit(c, h.t);
    } else if (c.tc && u.in_conv_g.w && conv1_supported(col, nullptr, u.in_conv_g, e0)) {
      // tensor-core first conv: 16-bit im2col rows [pixel][9 taps x (x, cond) | zero pad] (64 B each), then the persistent
      // 1x1 GEMM with GroupNorm sums in its epilogue.  cat([x, condition]) is still never materialised.
      col = c.alloc(B, H, W, 32);
      im2col_3x3_2ch(c, x, cond, col);
      conv1(c, col, nullptr, u.in_conv_g, e0, h.t);
      h.st = st;
      range_audit(c, h.t);
    } else {
      if (conv_first(c, xin, &cin, u.in_conv, h.t, st)) h.st = st;      // cat([x, condition]) never materialised
      range_audit(c, h.t);
    }
  }
  std::vector<TS> skips;
  for (const ULayer& L : u.downs) {
    if (L.kind == U_RES) {
      ResW& r = u.res[L.idx];
      TS out;
      out.t = c.alloc(B, h.t.h, h.t.w, r.cout);
      resblock(c, u, r, h, nullptr, temb, temb_bstride, out);
      h = out;
    } else if (L.kind == U_ATTN) {
      TS out;
      out.t = c.alloc(B, h.t.h, h.t.w, h.t.c);
      attnblock(c, u, u.attn[L.idx], h, out);
      h = out;
    } else {
      ConvW& d = u.down[L.idx];
      TS out;
      out.t = c.alloc(B, (h.t.h + 2 - 3) / 2 + 1, (h.t.w + 2 - 3) / 2 + 1, d.cout);
      out.st = stats16(c, B);
      ConvEpi e;
      e.stats_out = out.st;
      conv(c, h.t, nullptr, d, e, out.t);
      h = out;
    }
    skips.push_back(h);
  }
  {
    TS o1, o2, o3;
    o1.t = c.alloc(B, h.t.h, h.t.w, h.t.c);
    resblock(c, u, u.res[u.mid1], h, nullptr, temb, temb_bstride, o1);
    o2.t = c.alloc(B, h.t.h, h.t.w, h.t.c);
    attnblock(c, u, u.attn[u.mid_attn], o1, o2);
    o3.t = c.alloc(B, h.t.h, h.t.w, h.t.c);
    resblock(c, u, u.res[u.mid2], o2, nullptr, temb, temb_bstride, o3);
    h = o3;
  }
  for (const ULayer& L : u.ups) {
    if (L.kind == U_RES) {
      XRD_REQUIRE(!skips.empty(), "UNet: skip stack underflow");
      TS skip = skips.back(); skips.pop_back();
      ResW& r = u.res[L.idx];
      TS xs = h;
      if (h.t.h != skip.t.h || h.t.w != skip.t.w) {
        XRD_REQUIRE(skip.t.h == 2 * h.t.h && skip.t.w == 2 * h.t.w, "UNet: unsupported skip resize %dx%d -> %dx%d", h.t.h, h.t.w, skip.t.h,
                    skip.t.w);
        xs.t = c.alloc(B, skip.t.h, skip.t.w, h.t.c);
        xs.st = new_sums(c, B, 8);                        // sums of the resized tensor are not those of its source
        upsample2x_stats(c, h.t, xs.t, xs.st);            // F.interpolate(bilinear) 2x (HYB:381-382)
        range_audit(c, xs.t);
      }
      TS out;
      out.t = c.alloc(B, skip.t.h, skip.t.w, r.cout);
      resblock(c, u, r, xs, &skip, temb, temb_bstride, out);
      h = out;
    } else if (L.kind == U_ATTN) {
      TS out;
      out.t = c.alloc(B, h.t.h, h.t.w, h.t.c);
      attnblock(c, u, u.attn[L.idx], h, out);
      h = out;
    } else {
      // ConvTranspose2d + the 0.5x bilinear that always follows it == one pre-combined 3x3 conv at this resolution
      XRD_REQUIRE(!skips.empty() && skips.back().t.h == h.t.h && skips.back().t.w == h.t.w,
                  "UNet: this configuration consumes a ConvTranspose2d output at full size; only the reference topology is implemented");
      ConvW& up = u.up[L.idx];
      TS out;
      out.t = c.alloc(B, h.t.h, h.t.w, up.cout);
      out.st = stats16(c, B);
      ConvEpi e;
      e.stats_out = out.st;
      conv(c, h.t, nullptr, up, e, out.t);
      h = out;
    }
  }
  XRD_REQUIRE(h.t.h == H && h.t.w == W && h.t.c == u.out_c, "UNet: output resolution mismatch");
  double* s = concat_stats(c, h, nullptr, u.groups);
  Cout1Args a;
  a.x = h.t; a.k = 3; a.w = u.ow; a.bias = u.obias;
  a.gn_sums = s; a.groups = u.groups; a.gamma = u.og; a.beta = u.ob; a.eps = 1e-5f; a.act_in = ACT_SILU;
  a.mode = o.mode; a.y = o.y; a.x_cur = o.x_cur; a.x_next = o.x_next; a.c1 = o.c1; a.c2 = o.c2;
  conv_cout1(c, a);
  c.zpool = nullptr; c.zpool_cap = c.zpool_off = 0;
  c.a->release(mk0);
}

// ---------------------------------------------------------------- NAFNet
static void nafblock(Ctx& c, NafBlockW& b, const Tens& x, Tens& out) {
  const size_t mk = c.a->mark();
  const int B = x.n, H = x.h, W = x.w, C = b.c;
  Tens t = c.alloc(B, H, W, C);
  layernorm(c, x, b.n1w, b.n1b, 1e-6f, t);
  range_audit(c, t);
  Tens u = c.alloc(B, H, W, 2 * C);
  conv(c, t, nullptr, b.c1, ConvEpi(), u);
  Tens g = c.alloc(B, H, W, C);
  float* pool = c.allocf((size_t)B * C);
  zero_async(c, pool, (size_t)B * C * 4);
  dwconv_gate_pool(c, u, b.dw, b.dwb, g, pool);
  range_audit(c, g);
  float* scale = c.allocf((size_t)B * C);
  sca_scale(c, pool, B, C, H * W, b.scaw, b.scab, scale);
  Tens y = c.alloc(B, H, W, C);
  ConvEpi e3;
  e3.out_scale = b.beta; e3.resid = x;
  if (c.tc && x.dt != DT_F32) {
    scale_nc(c, g, scale);                      // x * sca(x) (HYB:157) ahead of the tensor-core GEMM
    range_audit(c, g);
  } else {
    e3.in_scale = scale;                        // folded into the A-operand load of the CUDA-core GEMM
  }
  conv(c, g, nullptr, b.c3, e3, y);
  layernorm(c, y, b.n2w, b.n2b, 1e-6f, t);
  range_audit(c, t);
  ConvEpi e4;
  e4.gate = true;                               // SimpleGate in the GEMM epilogue where the 2C-wide row fits one accumulator
  if (c.tc && conv1_supported(t, nullptr, b.c4, e4)) {
    conv(c, t, nullptr, b.c4, e4, g);
  } else {
    conv(c, t, nullptr, b.c4, ConvEpi(), u);
    simple_gate(c, u, g);
    range_audit(c, g);
  }
  ConvEpi e5;
  e5.out_scale = b.gamma; e5.resid = y;
  conv(c, g, nullptr, b.c5, e5, out);
  c.a->release(mk);
}

// EnhancedNAFNet.forward (HYB:206-238).  inp/out: (B,H,W) fp32 planes; `sanitize`: nan_to_num+clamp on the result.
static void nafnet_forward(Ctx& c, NafW& n, const float* inp, float* out, int B, int H, int W, int sanitize) {
  const size_t mk0 = c.a->mark();
  const int mult = 1 << (int)n.enc.size();
  const int Hp = cdiv(H, mult) * mult, Wp = cdiv(W, mult) * mult;
  const float* ip = inp;
  if (Hp != H || Wp != W) {
    float* padded = c.allocf((size_t)B * Hp * Wp);
    pad_crop_plane(c, inp, padded, B, H, W, Hp, Wp);
    ip = padded;
  }
  Tens xin; xin.p = (void*)ip; xin.n = B; xin.h = Hp; xin.w = Wp; xin.c = 1; xin.dt = DT_F32;
  Tens x = c.alloc(B, Hp, Wp, n.width);
  conv_first(c, xin, nullptr, n.intro, x);
  range_audit(c, x);
  std::vector<Tens> encs;
  for (size_t s = 0; s < n.enc.size(); ++s) {
    for (auto& b : n.enc[s]) {
      Tens o = c.alloc(B, x.h, x.w, x.c);
      nafblock(c, b, x, o);
      x = o;
    }
    encs.push_back(x);
    Tens d = c.alloc(B, x.h / 2, x.w / 2, 2 * x.c);
    conv(c, x, nullptr, n.downs[s], ConvEpi(), d);
    x = d;
  }
  for (auto& b : n.mid) {
    Tens o = c.alloc(B, x.h, x.w, x.c);
    nafblock(c, b, x, o);
    x = o;
  }
  for (size_t s = 0; s < n.dec.size(); ++s) {
    Tens upx = c.alloc(B, 2 * x.h, 2 * x.w, x.c / 2);
    conv(c, x, nullptr, n.ups[s], ConvEpi(), upx);          // 1x1 + PixelShuffle(2) store
    Tens& skip = encs[encs.size() - 1 - s];
    XRD_REQUIRE(skip.h == upx.h && skip.w == upx.w, "NAFNet: skip size mismatch");
    Tens m = c.alloc(B, upx.h, upx.w, upx.c);
    conv(c, upx, &skip, n.skips[s], ConvEpi(), m);          // cat + skip_conv as a two-source GEMM
    x = m;
    for (auto& b : n.dec[s]) {
      Tens o = c.alloc(B, x.h, x.w, x.c);
      nafblock(c, b, x, o);
      x = o;
    }
  }
  Cout1Args a;
  a.x = x; a.k = 3; a.w = n.ending_w; a.bias = n.ending_b; a.mode = 1; a.inp = ip; a.sanitize = sanitize;
  if (Hp != H || Wp != W) {
    float* full = c.allocf((size_t)B * Hp * Wp);
    a.y = full;
    conv_cout1(c, a);
    pad_crop_plane(c, full, out, B, Hp, Wp, H, W);          // x[:, :, :H, :W]
  } else {
    a.y = out;
    conv_cout1(c, a);
  }
  c.a->release(mk0);
}

// ---------------------------------------------------------------- router / fusion (always fp32)
static Tens cgg(Ctx& c, CGG& L, const Tens& x1, const Tens* x2) {
  const int Ho = (x1.h + 2 - 3) / L.conv.stride + 1, Wo = (x1.w + 2 - 3) / L.conv.stride + 1;
  Tens y = c.alloc(x1.n, Ho, Wo, L.conv.cout, DT_F32);
  conv_first(c, x1, x2, L.conv, y);
  double* s = new_sums(c, x1.n, L.groups);
  gn_stats(c, y, nullptr, L.groups, s);
  Tens a = c.alloc(x1.n, Ho, Wo, L.conv.cout, DT_F32);
  gn_act(c, y, nullptr, L.groups, s, L.g, L.b, 1e-5f, ACT_GELU, a);
  return a;
}

static void router_forward(Ctx& c, RouterW& r, const float* x, float* mask, int B, int H, int W, int sanitize) {
  XRD_REQUIRE(H % 4 == 0 && W % 4 == 0, "router: H and W must be multiples of 4 (got %dx%d)", H, W);
  const size_t mk0 = c.a->mark();
  {   // every GroupNorm-sum buffer of this evaluation comes out of one pre-zeroed pool (one memset node)
    const size_t cnt = (size_t)96 * B * 16;
    c.zpool = c.allocd(cnt);
    c.zpool_off = 0; c.zpool_cap = cnt;
    zero_async(c, c.zpool, cnt * sizeof(double));
  }
  Tens xin; xin.p = (void*)x; xin.n = B; xin.h = H; xin.w = W; xin.c = 1; xin.dt = DT_F32;
  Tens e1 = cgg(c, r.enc1, xin, nullptr);
  Tens e2 = cgg(c, r.enc2, e1, nullptr);
  Tens e3 = cgg(c, r.enc3, e2, nullptr);
  Tens m = cgg(c, r.mid, e3, nullptr);
  Tens d3u = c.alloc(B, e2.h, e2.w, r.up3.cout / 4, DT_F32);
  conv_simt(c, m, nullptr, r.up3, ConvEpi(), d3u);
  Tens d3 = cgg(c, r.dec3, d3u, &e2);
  Tens d2u = c.alloc(B, e1.h, e1.w, r.up2.cout / 4, DT_F32);
  conv_simt(c, d3, nullptr, r.up2, ConvEpi(), d2u);
  Tens d2 = cgg(c, r.dec2, d2u, &e1);
  Cout1Args a;
  a.x = d2; a.k = 1; a.w = r.out_w; a.bias = r.out_b; a.mode = 2; a.sanitize = sanitize; a.y = mask;
  conv_cout1(c, a);
  c.zpool = nullptr; c.zpool_cap = c.zpool_off = 0;
  c.a->release(mk0);
}

static void fusion_forward(Ctx& c, FusionW& f, const float* naf, const float* diff, const float* mask, float* out, int B, int H, int W) {
  const size_t mk0 = c.a->mark();
  Tens x3 = c.alloc(B, H, W, 3, DT_F32);
  interleave3(c, naf, diff, mask, (float*)x3.p, (int64_t)B * H * W);
  if (c.tc && c.adt != DT_F32 && f.padded && conv_smallcin2_supported(x3, nullptr, f.conv1.conv)) {
    // 16-bit modes (HYB:552-557 on the tensor cores): conv1 -> 16-bit, GN+GELU pass, conv2 as a padded cout = base_c tcgen05 conv
    // whose epilogue emits the GroupNorm sums, GroupNorm + GELU folded into the prologue of the final 1x1
    const int b = f.conv1.conv.cout;
    Tens y1 = c.alloc(B, H, W, b);
    double* s1 = new_sums(c, B, 8);
    if (!conv_first(c, x3, nullptr, f.conv1.conv, y1, s1)) gn_stats(c, y1, nullptr, 8, s1);
    range_audit(c, y1);
    Tens a1 = c.alloc(B, H, W, b);
    gn_act(c, y1, nullptr, 8, s1, f.conv1.g, f.conv1.b, 1e-5f, ACT_GELU, a1);
    range_audit(c, a1);
    Tens y2 = c.alloc(B, H, W, b);
    ConvEpi e2;
    e2.stats_out = new_sums(c, B, 8);
    conv(c, a1, nullptr, f.conv2p.conv, e2, y2);
    Cout1Args a;
    a.x = y2; a.k = 1; a.w = f.out_wp; a.bias = f.out_b; a.mode = 0; a.y = out;
    a.gn_sums = e2.stats_out; a.groups = 8; a.gamma = f.conv2p.g; a.beta = f.conv2p.b; a.eps = 1e-5f; a.act_in = ACT_GELU;
    conv_cout1(c, a);
    c.zpool = nullptr; c.zpool_cap = c.zpool_off = 0;
    c.a->release(mk0);
    return;
  }
  Tens a1 = cgg(c, f.conv1, x3, nullptr);
  Tens a2 = cgg(c, f.conv2, a1, nullptr);
  Cout1Args a;
  a.x = a2; a.k = 1; a.w = f.out_w; a.bias = f.out_b; a.mode = 0; a.y = out;
  conv_cout1(c, a);
  c.zpool = nullptr; c.zpool_cap = c.zpool_off = 0;
  c.a->release(mk0);
}

// ---------------------------------------------------------------- ExpertDenoiser (DirectUNetModel.py:232-255)
static void expert_forward(Ctx& c, ExpertW& e, const float* inp, float* out, int B, int H, int W) {
  XRD_REQUIRE(H % 4 == 0 && W % 4 == 0, "ExpertDenoiser: H and W must be multiples of 4 (got %dx%d)", H, W);
  const size_t mk0 = c.a->mark();
  ConvEpi relu;
  relu.act = ACT_RELU;
  auto cbr = [&](const Tens& a, const Tens* b, ConvW& w) {       // conv + folded BatchNorm + ReLU
    Tens y = c.alloc(a.n, a.h, a.w, w.cout);
    conv(c, a, b, w, relu, y);
    return y;
  };
  Tens xin; xin.p = (void*)inp; xin.n = B; xin.h = H; xin.w = W; xin.c = 1; xin.dt = DT_F32;
  Tens x1 = c.alloc(B, H, W, e.base);
  conv_simt(c, xin, nullptr, e.inc[0], relu, x1);                 // 1 -> base from the fp32 plane (CUDA cores: 9 MACs per output)
  range_audit(c, x1);
  x1 = cbr(x1, nullptr, e.inc[1]);
  Tens x2 = cbr(cbr(x1, nullptr, e.down1[0]), nullptr, e.down1[1]);
  Tens x2p = c.alloc(B, H / 2, W / 2, x2.c);
  maxpool2x2(c, x2, x2p);
  Tens x3 = cbr(cbr(x2p, nullptr, e.down2[0]), nullptr, e.down2[1]);
  Tens x3p = c.alloc(B, H / 4, W / 4, x3.c);
  maxpool2x2(c, x3, x3p);
  Tens x4 = cbr(cbr(x3p, nullptr, e.bott[0]), nullptr, e.bott[1]);
  Tens u2 = c.alloc(B, H / 2, W / 2, e.up2.cout / 4);
  conv(c, x4, nullptr, e.up2, ConvEpi(), u2);                     // ConvTranspose2d(2,2): GEMM + depth-to-space store
  Tens d2 = cbr(cbr(u2, &x3, e.upc2[0]), nullptr, e.upc2[1]);     // cat([xd2, x3]) as a two-source contraction
  Tens u1 = c.alloc(B, H, W, e.up1.cout / 4);
  conv(c, d2, nullptr, e.up1, ConvEpi(), u1);
  Tens d1 = cbr(cbr(u1, &x2, e.upc1[0]), nullptr, e.upc1[1]);
  Tens f = cbr(d1, nullptr, e.fin);
  Cout1Args a;
  a.x = f; a.k = 1; a.w = e.out_w; a.bias = e.out_b; a.mode = 0; a.y = out;
  conv_cout1(c, a);
  c.a->release(mk0);
}

// ------------------------------------------------------------------------------------------------
// exported to capi.cu
// ------------------------------------------------------------------------------------------------
void run_expert(Ctx& c, Handle& h, const float* inp, float* out, int B, int H, int W) { expert_forward(c, h.expert, inp, out, B, H, W); }

std::vector<int> ddim_timesteps(int noise_steps, int inference_steps) {
  std::vector<int> t;
  if (inference_steps < 1) inference_steps = 1;
  const int step = std::max(1, noise_steps / inference_steps);
  for (int i = 0; i < noise_steps; i += step) t.push_back(i);
  std::reverse(t.begin(), t.end());
  return t;
}

void run_unet_eps(Ctx& c, Handle& h, const float* x, const float* cond, const int64_t* t, float* eps, int B, int H, int W) {
  const size_t mk = c.a->mark();
  float* temb = c.allocf((size_t)B * h.unet.te.total);
  time_embed(c, h.unet.te, t, nullptr, B, temb);
  UNetOut o;
  o.mode = 0; o.y = eps;
  unet_eval(c, h.unet, x, cond, temb, h.unet.te.total, B, H, W, o);
  c.a->release(mk);
}

// the reverse loop (HYB:403-418) for one micro-batch; xcur is updated in place
void run_ddim_loop(Ctx& c, Handle& h, const float* noisy, float* xcur, const float* temb_table, const std::vector<int>& ts,
                   float* eps_trace, float* xin_trace, const float* teacher_x, size_t trace_stride, int B, int H, int W) {
  const size_t plane = (size_t)B * H * W;
  for (size_t e = 0; e < ts.size(); ++e) {
    if (teacher_x) copy_plane(c, teacher_x + e * trace_stride, xcur, plane);
    if (xin_trace) copy_plane(c, xcur, xin_trace + e * trace_stride, plane);
    UNetOut o;
    o.mode = 3;
    o.y = eps_trace ? eps_trace + e * trace_stride : nullptr;
    o.x_cur = xcur; o.x_next = xcur;
    o.c1 = h.coef1[ts[e]]; o.c2 = h.coef2[ts[e]];
    unet_eval(c, h.unet, xcur, noisy, temb_table + e * h.unet.te.total, 0, B, H, W, o);
  }
}

void run_nafnet(Ctx& c, Handle& h, const float* inp, float* out, int B, int H, int W, int sanitize) {
  nafnet_forward(c, h.naf, inp, out, B, H, W, sanitize);
}
void run_router(Ctx& c, Handle& h, const float* x, float* mask, int B, int H, int W, int sanitize) {
  router_forward(c, h.router, x, mask, B, H, W, sanitize);
}
void run_fusion(Ctx& c, Handle& h, const float* naf, const float* diff, const float* mask, float* out, int B, int H, int W) {
  fusion_forward(c, h.fusion, naf, diff, mask, out, B, H, W);
}

// pre-pack tensor-core weights for every conv the current mode will route to conv_tc (cudaMalloc is illegal
// during stream capture, so this runs before any graph is recorded)
void prepack_tc(Handle& h, DType dt) {
  if (dt == DT_F32) return;
  auto pk = [&](ConvW& w, int c1) {
    if (!w.w) return;
    if (w.cin % 16 != 0 || c1 % 16 != 0 || (w.cin - c1) % 16 != 0 || w.cout % 8 != 0) return;
    conv_tc_pack(nullptr, w, dt, c1);
  };
  UNetW& u = h.unet;
  if (u.ready) {
    // which ResidualBlocks see a concatenated input: exactly the ones in `ups`
    std::vector<char> cat(u.res.size(), 0);
    for (auto& L : u.ups) if (L.kind == U_RES) cat[L.idx] = 1;
    for (size_t i = 0; i < u.res.size(); ++i) {
      ResW& r = u.res[i];
      pk(r.c1, r.cin);                                   // conv1 reads the materialised GN output (one source) ...
      if (cat[i]) {                                      // ... or, GroupNorm fused into the conv, the two concat sources directly
        r.c1s = r.c1;
        for (int k = 0; k < 3; ++k) r.c1s.wtc[k] = nullptr;
        r.c1s.tc_c1 = -1;
        pk(r.c1s, r.cin / 2);
      }
      pk(r.c2, r.cout);
      if (r.has_rc) pk(r.rc, cat[i] ? r.cin / 2 : r.cin);
    }
    for (auto& a : u.attn) { pk(a.qkv, a.c); pk(a.proj, a.c); }
    for (auto& d : u.down) pk(d, d.cin);
    for (auto& w : u.up) pk(w, w.cin);
    pk(u.in_conv_g, 32);
  }
  NafW& n = h.naf;
  if (n.ready) {
    auto pb = [&](NafBlockW& b) { pk(b.c1, b.c); pk(b.c3, b.c); pk(b.c4, b.c); pk(b.c5, b.c); };
    for (auto& s : n.enc) for (auto& b : s) pb(b);
    for (auto& s : n.dec) for (auto& b : s) pb(b);
    for (auto& b : n.mid) pb(b);
    for (auto& w : n.downs) pk(w, w.cin);
    for (auto& w : n.ups) pk(w, w.cin);
    for (auto& w : n.skips) pk(w, w.cin / 2);
  }
  if (h.fusion.ready && h.fusion.padded) pk(h.fusion.conv2p.conv, h.fusion.conv2p.conv.cin);
  if (h.expert.ready) {
    ExpertW& e = h.expert;
    pk(e.inc[1], e.inc[1].cin);
    for (ConvW* w : {&e.down1[0], &e.down1[1], &e.down2[0], &e.down2[1], &e.bott[0], &e.bott[1], &e.upc2[1], &e.upc1[1], &e.fin, &e.up2, &e.up1}) pk(*w, w->cin);
    pk(e.upc2[0], e.upc2[0].cin / 2);                      // two sources: [up(x4) | x3]
    pk(e.upc1[0], e.upc1[0].cin / 2);
  }
  XRD_CUDA(cudaDeviceSynchronize());
}

}  // namespace xrd
