// engine.cuh -- the handle: weight store, packed layer plans, arena, graph cache.
#pragma once
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "kernels.cuh"

namespace xrd {

struct Param {
  std::vector<int64_t> shape;
  float* d = nullptr;
  size_t n = 0;
  bool in_slab = false;     // d points into Handle::slab (xrd_import_weights): not freed on its own
};

struct ResW {
  int cin = 0, cout = 0, temb_off = 0;
  bool has_rc = false;
  const float *g1 = nullptr, *b1 = nullptr, *g2 = nullptr, *b2 = nullptr;
  ConvW c1, c2, rc;
  ConvW c1s;                 // conv1 again, tensor-core packing split at cin/2: reads a concatenated input as two sources
                             // (same fp32 weights; used when GroupNorm+SiLU is applied inside the conv, so no concat tensor exists)
};
struct AttnW {
  int c = 0;
  const float *g = nullptr, *b = nullptr;
  ConvW qkv, proj;
};
enum UKind { U_RES = 0, U_ATTN = 1, U_DOWN = 2, U_UP = 3 };
struct ULayer { int kind; int idx; };

struct UNetW {
  bool ready = false;
  int mc = 0, groups = 8, heads = 2;
  ConvW in_conv;
  ConvW in_conv_g;           // in_conv as a 1x1 GEMM over the im2col matrix [pixels][32] (18 real columns): tensor-core first conv
  std::vector<ResW> res;     // downs..., mid1, mid2, ups... in creation order
  std::vector<AttnW> attn;
  std::vector<ConvW> down, up;
  std::vector<ULayer> downs, ups;
  int mid1 = -1, mid2 = -1, mid_attn = -1;
  const float *og = nullptr, *ob = nullptr;   // out_conv.0 GroupNorm
  float* ow = nullptr;                        // out_conv.2 weight packed [9][C]
  const float* obias = nullptr;
  int out_c = 0;
  TimeEmbW te;
};

struct NafBlockW {
  int c = 0;
  const float *n1w = nullptr, *n1b = nullptr, *n2w = nullptr, *n2b = nullptr, *beta = nullptr, *gamma = nullptr;
  ConvW c1, c3, c4, c5;
  float* dw = nullptr;        // [9][2c]
  const float* dwb = nullptr;
  const float *scaw = nullptr, *scab = nullptr;
};
struct NafW {
  bool ready = false;
  int width = 0;
  ConvW intro;
  std::vector<std::vector<NafBlockW>> enc, dec;
  std::vector<NafBlockW> mid;
  std::vector<ConvW> downs, ups, skips;
  float* ending_w = nullptr;  // [9][width]
  const float* ending_b = nullptr;
};
struct CGG {  // conv -> GroupNorm -> GELU
  ConvW conv;
  const float *g = nullptr, *b = nullptr;
  int groups = 8;
};
struct RouterW {
  bool ready = false;
  CGG enc1, enc2, enc3, mid, dec3, dec2;
  ConvW up3, up2;
  const float* out_w = nullptr;   // (1,base,1,1) == [1][base]
  const float* out_b = nullptr;
};
struct FusionW {
  bool ready = false;
  CGG conv1, conv2;
  const float* out_w = nullptr;
  const float* out_b = nullptr;
  // 16-bit modes: conv2 (base_c -> base_c/2) zero-padded to base_c outputs so that it runs on the tcgen05 3x3 kernels;
  // GroupNorm(4, base_c/2) is the first four groups of GroupNorm(8, base_c) over the padded tensor (same group size)
  bool padded = false;
  CGG conv2p;
  const float* out_wp = nullptr;
};

// ExpertDenoiser (DirectUNet/DirectUNetModel.py:160-255): Conv3x3(no bias) + BatchNorm2d(eval) + ReLU pairs, MaxPool2d(2),
// ConvTranspose2d(2,2) ups, 1x1 head.  BatchNorm is folded into the conv weights/bias at pack time.
struct ExpertW {
  bool ready = false;
  int base = 0;
  ConvW inc[2], down1[2], down2[2], bott[2], upc2[2], upc1[2], fin;
  ConvW up2, up1;
  const float* out_w = nullptr;   // outc (1,base,1,1) == [1][base]
  const float* out_b = nullptr;
};

struct GraphEntry {
  cudaGraphExec_t exec = nullptr;
  uint64_t kernels = 0;     // kernel nodes per replay (launch accounting)
  float* in = nullptr;      // staging planes inside the arena the graph reads/writes
  float* out = nullptr;
  float* temb = nullptr;    // time-embedding table owned by this entry (cudaMalloc)
};

struct Handle {
  int device = 0;
  xrd_config cfg;
  int mode = XRD_MODE_FP16;
  bool use_graph = true;
  RangeAudit* audit_dev = nullptr;   // device counters of the range audit (allocated on first use)
  bool audit = false;
  std::mutex mu;

  std::unordered_map<std::string, Param> params;
  float* slab = nullptr;          // one allocation holding every tensor of an imported weight blob
  std::vector<void*> owned;       // packed weights etc. freed on refinalize / destroy
  UNetW unet;
  NafW naf;
  RouterW router;
  FusionW fusion;
  ExpertW expert;

  Arena arena;
  std::map<std::string, size_t> plan_cache;      // call signature -> arena bytes
  std::map<std::string, GraphEntry> graphs;

  // sampler tables (host, float32 like the reference)
  std::vector<float> coef1, coef2;               // per timestep index t
  // last op hook for xrd_op_time_last
  std::function<void(Ctx&)> last_op;
  size_t last_op_bytes = 0;

  float* dalloc_f(size_t n);
  void* dalloc(size_t bytes);
  void free_owned();
  void drop_graphs();
  const Param& P(const std::string& key) const;
  const float* PD(const std::string& key) const { return P(key).d; }
  bool has(const std::string& key) const { return params.count(key) != 0; }
};

}  // namespace xrd
