// capi.cu -- the extern "C" surface declared in include/xrd.h.  Catches every internal
// exception at the boundary, plans/grows the arena, captures and replays the sampler graph,
// splits large batches into micro-batches.
#include <functional>
#include <cxxabi.h>
#include <stdlib.h>
#include "engine.cuh"

#include <string.h>
#include <algorithm>

#define XRD_EXPORT extern "C" __attribute__((visibility("default")))

namespace xrd {
void finalize(Handle& h, int which);
std::vector<int> ddim_timesteps(int noise_steps, int inference_steps);
void run_unet_eps(Ctx& c, Handle& h, const float* x, const float* cond, const int64_t* t, float* eps, int B, int H, int W);
void run_ddim_loop(Ctx& c, Handle& h, const float* noisy, float* xcur, const float* temb_table, const std::vector<int>& ts,
                   float* eps_trace, float* xin_trace, const float* teacher_x, size_t trace_stride, int B, int H, int W);
void run_nafnet(Ctx& c, Handle& h, const float* inp, float* out, int B, int H, int W, int sanitize);
void run_router(Ctx& c, Handle& h, const float* x, float* mask, int B, int H, int W, int sanitize);
void run_fusion(Ctx& c, Handle& h, const float* naf, const float* diff, const float* mask, float* out, int B, int H, int W);
void run_expert(Ctx& c, Handle& h, const float* inp, float* out, int B, int H, int W);
void prepack_tc(Handle& h, DType dt);
void tiles_check(int B, int H, int W, int T, int halo);
void tiles_extract(Ctx& c, const float* img, float* tiles, int B, int H, int W, int T, int halo);
void tiles_blend(Ctx& c, const float* tiles, float* img, int B, int H, int W, int T, int halo);
int tiles_count_host(int L, int T, int halo);
int tiles_origin_host(int k, int L, int T, int halo);
void resize_bicubic_u8(Ctx& c, const uint8_t* src, uint8_t* dst, uint8_t* tmp, int N, int Hin, int Win, int Hout, int Wout);
int resample_table_host(int in_size, int out_size, int* bounds, int* kk, int cap_k);
void u8_to_unit(Ctx& c, const uint8_t* in, float* out, int64_t n);
void unit_to_u8(Ctx& c, const float* in, uint8_t* out, int64_t n);
}  // namespace xrd

using namespace xrd;

static thread_local std::string g_last_error;

struct xrd_handle {
  Handle h;
  cudaStream_t cap_stream = nullptr;   // capture-only stream (the caller's stream may be the legacy default stream)
  cudaStream_t last_stream = nullptr;
  bool have_last_stream = false;
  DType packed_dt = DT_F32;
  float* scratch = nullptr;            // hybrid: sanitised NAFNet / sampler / mask planes of one micro-batch
  size_t scratch_n = 0;
  // hybrid side branches (HYB:612-624: the fast path, the quality path and the routing mask only share their input): NAFNet and
  // the router run on two private streams with private workspaces beside the sampler graph and join before the fusion stack
  cudaStream_t side[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
  Arena side_arena[2];
  bool overlap = true;
  int64_t overlap_max_px = INT64_MAX;  // side branches only for micro-batches of at most this many pixels (XRD_OVERLAP_MAXPIX)
  // op-hook state
  ConvW op_w;
  std::vector<void*> op_owned;
};

template <class F>
static int guarded(F&& f) {
  try {
    f();
    return XRD_OK;
  } catch (const Error& e) {
    g_last_error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return XRD_ERR_INVALID;
  } catch (...) {
    g_last_error = "unknown error";
    return XRD_ERR_INVALID;
  }
}

static DType mode_dtype(int mode) { return mode == XRD_MODE_FP32_CHECK ? DT_F32 : (mode == XRD_MODE_FP16 ? DT_F16 : DT_BF16); }

static Ctx make_ctx(xrd_handle* H, cudaStream_t s, bool dry, Arena* a) {
  Ctx c;
  c.s = s; c.a = a; c.dry = dry;
  c.adt = mode_dtype(H->h.mode);
  c.tc = H->h.mode != XRD_MODE_FP32_CHECK;
  c.audit = (H->h.audit && !dry) ? H->h.audit_dev : nullptr;
  return c;
}

static void bind_stream(xrd_handle* H, cudaStream_t s) {
  XRD_CUDA(cudaSetDevice(H->h.device));
  if (H->have_last_stream && H->last_stream != s) XRD_CUDA(cudaStreamSynchronize(H->last_stream));
  H->last_stream = s; H->have_last_stream = true;
  DType dt = mode_dtype(H->h.mode);
  if (dt != DT_F32 && H->packed_dt != dt) {
    prepack_tc(H->h, dt);
    H->packed_dt = dt;
  }
}

// Plan (dry run) + grow the arena, then run `fn` for real with the arena reset.  `A` is the handle's main arena (the one
// the captured graphs point into: growing it drops them) or one of the side-branch arenas.
static void with_arena_in(xrd_handle* H, cudaStream_t s, const std::string& key, Arena& A, const std::function<void(Ctx&)>& fn) {
  Handle& h = H->h;
  const bool main_arena = &A == &h.arena;
  size_t need;
  auto it = h.plan_cache.find(key);
  if (it != h.plan_cache.end()) {
    need = it->second;
  } else {
    Arena dry;
    dry.dry = true;
    Ctx c = make_ctx(H, s, true, &dry);
    fn(c);
    need = dry.peak + 4096;
    h.plan_cache[key] = need;
  }
  if (need > A.cap) {
    XRD_CUDA(cudaDeviceSynchronize());
    if (main_arena) h.drop_graphs();
    if (A.base) XRD_CUDA(cudaFree(A.base));
    A.base = nullptr; A.cap = 0;
    size_t cap = need + (need >> 4);
    cudaError_t e = cudaMalloc((void**)&A.base, cap);
    if (e != cudaSuccess) fail(XRD_ERR_CUDA, "cannot allocate a %zu MiB workspace: %s", cap >> 20, cudaGetErrorString(e));
    A.cap = cap;
  }
  A.off = 0; A.peak = 0; A.dry = false;
  Ctx c = make_ctx(H, s, false, &A);
  fn(c);
}
static void with_arena(xrd_handle* H, cudaStream_t s, const std::string& key, const std::function<void(Ctx&)>& fn) {
  with_arena_in(H, s, key, H->h.arena, fn);
}

// images per micro-batch: bounds the workspace (and keeps one graph shape for any batch size)
static int micro_batch(int B, int H, int W) {
  const int64_t budget = (int64_t)16 * 512 * 512;
  int mb = (int)std::max<int64_t>(1, budget / ((int64_t)H * W));
  return std::min(B, mb);
}

static std::string keyf(const char* tag, xrd_handle* H, int B, int Hh, int W, int e0 = 0, int e1 = 0, int e2 = 0, int e3 = 0,
                        int e4 = 0) {
  char buf[192];
  snprintf(buf, sizeof(buf), "%s|%d|%d|%d|%d.%d.%d.%d.%d|m%d", tag, B, Hh, W, e0, e1, e2, e3, e4, H->h.mode);
  return buf;
}

static void check_img(int B, int H, int W) {
  XRD_REQUIRE(B >= 1 && H >= 1 && W >= 1, "empty batch or image (B=%d,H=%d,W=%d)", B, H, W);
}
static void check_unet_shape(xrd_handle* H, int Hh, int W) {
  const int div = 1 << (H->h.cfg.unet_n_levels - 1);
  XRD_REQUIRE(Hh % div == 0 && W % div == 0, "UNet: H and W must be multiples of %d (got %dx%d)", div, Hh, W);
}

// ------------------------------------------------------------------------------------------------
XRD_EXPORT void xrd_default_config(xrd_config* c) {
  memset(c, 0, sizeof(*c));
  c->unet_in_channels = 1; c->unet_model_channels = 48; c->unet_n_levels = 4;
  const int cm[4] = {1, 2, 3, 4};
  for (int i = 0; i < 4; ++i) c->unet_channel_mult[i] = cm[i];
  c->unet_num_res_blocks = 2; c->unet_n_attn = 1; c->unet_attention_resolutions[0] = 3;
  c->unet_time_emb_dim = 192; c->unet_num_heads = 2;
  c->naf_img_channel = 1; c->naf_width = 32; c->naf_middle_blk_num = 8;
  c->naf_n_enc = 4; c->naf_n_dec = 4;
  const int eb[4] = {2, 2, 4, 6}, db[4] = {2, 2, 2, 2};
  for (int i = 0; i < 4; ++i) { c->naf_enc_blk_nums[i] = eb[i]; c->naf_dec_blk_nums[i] = db[i]; }
  c->router_base_c = 32; c->fusion_base_c = 48;
  c->noise_steps = 50; c->beta_start = 1e-4f; c->beta_end = 0.02f;
  c->expert_base_c = 64;
}

XRD_EXPORT int xrd_api_version(void) { return XRD_API_VERSION; }
XRD_EXPORT const char* xrd_last_error(void) { return g_last_error.c_str(); }
XRD_EXPORT uint64_t xrd_kernel_launch_count(void) { return g_launches.load(); }

// ---- in-situ launch profile (common.cuh: LaunchProf) ------------------------------------------------------------------------
XRD_EXPORT int xrd_profile_begin(void) {
  return guarded([&] {
    XRD_REQUIRE(!g_prof, "a launch profile is already running on this thread");
    g_prof = new LaunchProf();
    g_prof->recs.reserve(1 << 14);
  });
}
// Ends the profile of the calling thread and writes one line per kernel, "name<TAB>launches<TAB>total_ms\n" (kernel = the launch
// expression, template arguments as written in the source), in order of first launch.  buf = NULL / too small: *need only.
XRD_EXPORT int xrd_profile_end(char* buf, uint64_t cap, uint64_t* need) {
  return guarded([&] {
    XRD_REQUIRE(g_prof && need, "no launch profile is running on this thread");
    LaunchProf* pr = g_prof;
    g_prof = nullptr;
    std::vector<std::string> order;
    std::unordered_map<std::string, std::pair<uint64_t, double>> agg;
    cudaError_t first_err = cudaSuccess;
    for (auto& r : pr->recs) {
      float ms = 0.f;
      cudaError_t e = cudaEventSynchronize(r.e1);
      if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r.e0, r.e1);
      if (e != cudaSuccess && first_err == cudaSuccess) first_err = e;
      cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
      std::string name = r.name;
      if (name.rfind("_Z", 0) == 0) {
        int st = 0;
        char* dm = abi::__cxa_demangle(name.c_str(), nullptr, nullptr, &st);
        if (dm) { if (st == 0) name = dm; free(dm); }
      }
      // "void xrd::k_conv3s<__half, 3, 0, 8, 1, true>(CUtensorMap_st, ...)" (or the launch expression "(k_foo<T, 1>)") -> "k_conv3s"
      while (!name.empty() && (name[0] == '(' || name[0] == ' ')) name.erase(0, 1);
      if (name.compare(0, 5, "void ") == 0) name.erase(0, 5);
      for (size_t an; (an = name.find("(anonymous namespace)::")) != std::string::npos;) name.erase(an, 23);
      const size_t lt = name.find_first_of("<(");
      if (lt != std::string::npos) name.resize(lt);
      const size_t ns = name.rfind("::");
      if (ns != std::string::npos) name.erase(0, ns + 2);
      auto it = agg.find(name);
      if (it == agg.end()) { order.push_back(name); it = agg.emplace(name, std::make_pair(0ull, 0.0)).first; }
      it->second.first += 1; it->second.second += ms;
    }
    delete pr;
    if (first_err != cudaSuccess) fail(XRD_ERR_CUDA, "launch profile: %s (profiles need eager launches, not a graph capture)", cudaGetErrorString(first_err));
    std::string out;
    char line[256];
    for (auto& n : order) {
      snprintf(line, sizeof(line), "%s\t%llu\t%.6f\n", n.c_str(), (unsigned long long)agg[n].first, agg[n].second);
      out += line;
    }
    *need = out.size() + 1;
    if (buf && cap >= *need) memcpy(buf, out.c_str(), *need);
  });
}

XRD_EXPORT int xrd_ddim_num_evals(int noise_steps, int inference_steps) {
  if (noise_steps < 1) return 0;
  return (int)ddim_timesteps(noise_steps, inference_steps).size();
}

XRD_EXPORT int xrd_create(int device, const xrd_config* cfg, xrd_handle** out) {
  return guarded([&] {
    XRD_REQUIRE(cfg && out, "null argument");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      fail(XRD_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "count=0");
    XRD_REQUIRE(device >= 0 && device < ndev, "device %d out of range (have %d)", device, ndev);
    cudaDeviceProp prop;
    XRD_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
      fail(XRD_ERR_NO_DEVICE, "device %d is sm_%d%d; libxrd is built for sm_100a (B200) only", device, prop.major, prop.minor);
    XRD_REQUIRE(cfg->unet_n_levels >= 1 && cfg->unet_n_levels <= XRD_MAX_LEVELS && cfg->naf_n_enc <= XRD_MAX_LEVELS &&
                    cfg->naf_n_dec <= XRD_MAX_LEVELS && cfg->noise_steps >= 1,
                "invalid configuration");
    DeviceScope dscope(device);
    xrd_handle* H = new xrd_handle();
    H->h.device = device;
    H->h.cfg = *cfg;
    H->h.cfg.unet_prefix[63] = H->h.cfg.naf_prefix[63] = H->h.cfg.router_prefix[63] = H->h.cfg.fusion_prefix[63] = 0;
    H->h.cfg.expert_prefix[63] = 0;
    // kernel nodes inherit the priority of the stream they were captured on: the sampler graph (the critical path of the hybrid)
    // is captured at the highest priority, so the side branches below fill the SMs it leaves idle instead of competing with it
    int prio_least = 0, prio_greatest = 0;
    XRD_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    XRD_CUDA(cudaStreamCreateWithPriority(&H->cap_stream, cudaStreamNonBlocking, prio_greatest));
    for (int i = 0; i < 2; ++i) {
      XRD_CUDA(cudaStreamCreateWithFlags(&H->side[i], cudaStreamNonBlocking));
      XRD_CUDA(cudaEventCreateWithFlags(&H->ev_join[i], cudaEventDisableTiming));
    }
    XRD_CUDA(cudaEventCreateWithFlags(&H->ev_fork, cudaEventDisableTiming));
    H->overlap = !(getenv("XRD_OVERLAP") && atoi(getenv("XRD_OVERLAP")) == 0);
    if (getenv("XRD_OVERLAP_MAXPIX")) H->overlap_max_px = atoll(getenv("XRD_OVERLAP_MAXPIX"));
    *out = H;
  });
}

static void free_op_state(xrd_handle* H) {
  for (int i = 0; i < 3; ++i)
    if (H->op_w.wtc[i]) { cudaFree(H->op_w.wtc[i]); H->op_w.wtc[i] = nullptr; }
  H->op_w = ConvW();
  for (void* p : H->op_owned) cudaFree(p);
  H->op_owned.clear();
}

XRD_EXPORT void xrd_destroy(xrd_handle* H) {
  if (!H) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(H->h.device);
  cudaDeviceSynchronize();
  free_op_state(H);
  H->h.free_owned();
  for (auto& kv : H->h.params) if (!kv.second.in_slab) cudaFree(kv.second.d);
  if (H->h.slab) cudaFree(H->h.slab);
  if (H->h.arena.base) cudaFree(H->h.arena.base);
  if (H->scratch) cudaFree(H->scratch);
  if (H->h.audit_dev) cudaFree(H->h.audit_dev);
  if (H->cap_stream) cudaStreamDestroy(H->cap_stream);
  for (int i = 0; i < 2; ++i) {
    if (H->side_arena[i].base) cudaFree(H->side_arena[i].base);
    if (H->side[i]) cudaStreamDestroy(H->side[i]);
    if (H->ev_join[i]) cudaEventDestroy(H->ev_join[i]);
  }
  if (H->ev_fork) cudaEventDestroy(H->ev_fork);
  if (prev >= 0) cudaSetDevice(prev);
  delete H;
}

XRD_EXPORT int xrd_set_param(xrd_handle* H, const char* key, const void* data, const int64_t* shape, int ndim, int is_device) {
  return guarded([&] {
    XRD_REQUIRE(H && key && data && ndim >= 0 && ndim <= 8, "bad argument");
    std::lock_guard<std::mutex> lk(H->h.mu);
    DeviceScope dscope(H->h.device);
    // Kernels of this handle still in flight (on a non-blocking stream the blocking copy below does not wait for) may read the
    // tensor being replaced; and every packed plan / captured graph derived from the old value is now stale: quiesce, then
    // invalidate, so that a run without xrd_finalize_weights fails loudly instead of mixing new and old weights.
    if (H->have_last_stream) XRD_CUDA(cudaStreamSynchronize(H->last_stream));
    H->h.unet.ready = H->h.naf.ready = H->h.router.ready = H->h.fusion.ready = H->h.expert.ready = false;
    H->h.drop_graphs();
    size_t n = 1;
    std::vector<int64_t> shp;
    for (int i = 0; i < ndim; ++i) { XRD_REQUIRE(shape[i] >= 0, "negative dimension"); n *= (size_t)shape[i]; shp.push_back(shape[i]); }
    Param& p = H->h.params[key];
    if (p.n != n || !p.d) {
      // weights referenced by packed plans may be replaced: invalidate plans first
      if (p.d && !p.in_slab) { XRD_CUDA(cudaDeviceSynchronize()); cudaFree(p.d); }
      p.d = nullptr; p.in_slab = false;
      XRD_CUDA(cudaMalloc((void**)&p.d, std::max<size_t>(n * 4, 16)));
    }
    p.n = n; p.shape = shp;
    XRD_CUDA(cudaMemcpy(p.d, data, n * 4, is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
  });
}

// ---- weight blob: every state_dict tensor of a handle in one relocatable buffer ---------------------------
// layout: "XRDW0001" | u64 count | u64 payload_offset | u64 payload_floats | entries | pad | payload (float32)
//         entry = u32 key_len | key | u32 ndim | i64 dims[ndim] | u64 offset_in_floats      (tensors start at multiples of 64 floats)
namespace {
struct BlobWriter {
  char* p; uint64_t cap, off = 0;
  void put(const void* src, size_t n) { if (p && off + n <= cap) memcpy(p + off, src, n); off += n; }
  template <typename V> void val(V v) { put(&v, sizeof(V)); }
};
struct BlobReader {
  const char* p; uint64_t n, off = 0;
  void get(void* dst, size_t k) { XRD_REQUIRE(off + k <= n, "weight blob truncated"); memcpy(dst, p + off, k); off += k; }
  template <typename V> V val() { V v; get(&v, sizeof(V)); return v; }
};
}  // namespace

XRD_EXPORT int xrd_export_weights(xrd_handle* H, void* buf, uint64_t cap, uint64_t* need) {
  return guarded([&] {
    XRD_REQUIRE(H && need, "null argument");
    std::lock_guard<std::mutex> lk(H->h.mu);
    DeviceScope dscope(H->h.device);
    std::vector<std::string> keys;
    for (auto& kv : H->h.params) keys.push_back(kv.first);
    std::sort(keys.begin(), keys.end());
    uint64_t hdr = 8 + 3 * 8, floats = 0;
    for (auto& k : keys) {
      const Param& p = H->h.params[k];
      hdr += 4 + k.size() + 4 + 8 * p.shape.size() + 8;
      floats += (p.n + 63) & ~(uint64_t)63;
    }
    const uint64_t pay = (hdr + 255) & ~(uint64_t)255;
    *need = pay + floats * 4;
    if (!buf || cap < *need) return;                       // size query
    if (H->have_last_stream) XRD_CUDA(cudaStreamSynchronize(H->last_stream));
    BlobWriter w{(char*)buf, cap};
    w.put("XRDW0001", 8);
    w.val<uint64_t>(keys.size()); w.val<uint64_t>(pay); w.val<uint64_t>(floats);
    uint64_t o = 0;
    for (auto& k : keys) {
      const Param& p = H->h.params[k];
      w.val<uint32_t>((uint32_t)k.size()); w.put(k.data(), k.size());
      w.val<uint32_t>((uint32_t)p.shape.size());
      for (auto d : p.shape) w.val<int64_t>(d);
      w.val<uint64_t>(o);
      o += (p.n + 63) & ~(uint64_t)63;
    }
    memset((char*)buf + w.off, 0, pay - w.off);
    o = 0;
    for (auto& k : keys) {
      const Param& p = H->h.params[k];
      const uint64_t span = (p.n + 63) & ~(uint64_t)63;
      float* dst = (float*)((char*)buf + pay) + o;
      XRD_CUDA(cudaMemcpy(dst, p.d, p.n * 4, cudaMemcpyDeviceToHost));
      for (uint64_t i = p.n; i < span; ++i) dst[i] = 0.f;
      o += span;
    }
  });
}

XRD_EXPORT int xrd_import_weights(xrd_handle* H, const void* blob, uint64_t bytes) {
  return guarded([&] {
    XRD_REQUIRE(H && blob, "null argument");
    std::lock_guard<std::mutex> lk(H->h.mu);
    DeviceScope dscope(H->h.device);
    BlobReader r{(const char*)blob, bytes};
    char magic[8];
    r.get(magic, 8);
    XRD_REQUIRE(memcmp(magic, "XRDW0001", 8) == 0, "not a libxrd weight blob");
    const uint64_t count = r.val<uint64_t>(), pay = r.val<uint64_t>(), floats = r.val<uint64_t>();
    XRD_REQUIRE(pay <= bytes && floats <= (bytes - pay) / 4 && count < (1u << 20), "weight blob header out of range");
    struct Ent { std::string key; std::vector<int64_t> shape; uint64_t off, n; };
    std::vector<Ent> ents(count);
    for (auto& e : ents) {
      const uint32_t kl = r.val<uint32_t>();
      XRD_REQUIRE(kl < 4096, "weight blob: key too long");
      e.key.resize(kl);
      r.get(&e.key[0], kl);
      const uint32_t nd = r.val<uint32_t>();
      XRD_REQUIRE(nd <= 8, "weight blob: too many dimensions");
      e.n = 1;
      for (uint32_t i = 0; i < nd; ++i) { const int64_t d = r.val<int64_t>(); XRD_REQUIRE(d >= 0, "weight blob: negative dimension"); e.shape.push_back(d); e.n *= (uint64_t)d; }
      e.off = r.val<uint64_t>();
      XRD_REQUIRE(e.off <= floats && e.n <= floats - e.off, "weight blob: tensor '%s' outside the payload", e.key.c_str());
    }
    // quiesce and invalidate exactly as xrd_set_param does, then swap the whole parameter set
    XRD_CUDA(cudaDeviceSynchronize());
    H->h.unet.ready = H->h.naf.ready = H->h.router.ready = H->h.fusion.ready = H->h.expert.ready = false;
    H->h.drop_graphs();
    float* slab = nullptr;
    XRD_CUDA(cudaMalloc((void**)&slab, std::max<uint64_t>(floats * 4, 16)));
    cudaError_t ce = cudaMemcpy(slab, (const char*)blob + pay, floats * 4, cudaMemcpyHostToDevice);     // ONE copy for all tensors
    if (ce != cudaSuccess) { cudaFree(slab); fail(XRD_ERR_CUDA, "weight blob upload failed: %s", cudaGetErrorString(ce)); }
    for (auto& kv : H->h.params) if (!kv.second.in_slab && kv.second.d) cudaFree(kv.second.d);
    H->h.params.clear();
    if (H->h.slab) cudaFree(H->h.slab);
    H->h.slab = slab;
    for (auto& e : ents) {
      Param& p = H->h.params[e.key];
      p.shape = e.shape; p.n = e.n; p.d = slab + e.off; p.in_slab = true;
    }
  });
}

XRD_EXPORT int xrd_finalize_weights(xrd_handle* H, int which) {
  return guarded([&] {
    XRD_REQUIRE(H, "null handle");
    std::lock_guard<std::mutex> lk(H->h.mu);
    DeviceScope dscope(H->h.device);
    finalize(H->h, which);
    H->packed_dt = DT_F32;
  });
}

XRD_EXPORT int xrd_set_mode(xrd_handle* H, int mode) {
  return guarded([&] {
    XRD_REQUIRE(H && (mode == XRD_MODE_BF16 || mode == XRD_MODE_FP32_CHECK || mode == XRD_MODE_FP16), "bad mode %d", mode);
    std::lock_guard<std::mutex> lk(H->h.mu);
    H->h.mode = mode;
  });
}
XRD_EXPORT int xrd_get_mode(xrd_handle* H) { return H ? H->h.mode : XRD_ERR_INVALID; }
XRD_EXPORT int xrd_set_use_graph(xrd_handle* H, int enable) {
  if (!H) return XRD_ERR_INVALID;
  H->h.use_graph = enable != 0;
  return XRD_OK;
}

XRD_EXPORT int xrd_set_side_branches(xrd_handle* H, int enable) {
  if (!H) return XRD_ERR_INVALID;
  std::lock_guard<std::mutex> lk(H->h.mu);
  H->overlap = enable != 0;
  return XRD_OK;
}

// ------------------------------------------------------------------------------------------------
XRD_EXPORT int xrd_unet_eps(xrd_handle* H, const float* x, const float* cond, const int64_t* t, float* eps, int B, int Hh, int W,
                            void* stream) {
  return guarded([&] {
    XRD_REQUIRE(H && x && cond && t && eps, "null argument");
    check_img(B, Hh, W);
    std::lock_guard<std::mutex> lk(H->h.mu);
    XRD_REQUIRE(H->h.unet.ready, "UNet weights are not finalised");
    check_unet_shape(H, Hh, W);
    cudaStream_t s = (cudaStream_t)stream;
    DeviceScope dscope(H->h.device);
    bind_stream(H, s);
    const int mb = micro_batch(B, Hh, W);
    const size_t plane = (size_t)Hh * W;
    for (int b0 = 0; b0 < B; b0 += mb) {
      const int nb = std::min(mb, B - b0);
      with_arena(H, s, keyf("unet", H, nb, Hh, W), [&](Ctx& c) {
        run_unet_eps(c, H->h, x + b0 * plane, cond + b0 * plane, t + b0, eps + b0 * plane, nb, Hh, W);
      });
    }
  });
}

// temb table for a list of timesteps, computed once per graph entry / eager call
static void compute_temb_table(xrd_handle* H, cudaStream_t s, const std::vector<int>& ts, float* table_dev) {
  int* tl = nullptr;
  XRD_CUDA(cudaMalloc((void**)&tl, ts.size() * sizeof(int)));
  XRD_CUDA(cudaMemcpy(tl, ts.data(), ts.size() * sizeof(int), cudaMemcpyHostToDevice));
  Arena none;
  Ctx c = make_ctx(H, s, false, &none);
  time_embed(c, H->h.unet.te, nullptr, tl, (int)ts.size(), table_dev);
  XRD_CUDA(cudaStreamSynchronize(s));
  cudaFree(tl);
}

static uint64_t count_kernel_nodes(cudaGraph_t g) {
  size_t n = 0;
  if (cudaGraphGetNodes(g, nullptr, &n) != cudaSuccess || n == 0) return 0;
  std::vector<cudaGraphNode_t> nodes(n);
  if (cudaGraphGetNodes(g, nodes.data(), &n) != cudaSuccess) return 0;
  uint64_t k = 0;
  for (auto nd : nodes) {
    cudaGraphNodeType ty;
    if (cudaGraphNodeGetType(nd, &ty) == cudaSuccess && ty == cudaGraphNodeTypeKernel) ++k;
  }
  return k;
}

// one micro-batch of the sampler; noisy/out are user device pointers
static void ddim_chunk(xrd_handle* H, cudaStream_t s, const float* noisy, int inference_steps, float* out, float* eps_trace,
                       float* xin_trace, const float* teacher_x, size_t trace_stride, int nb, int Hh, int W) {
  Handle& h = H->h;
  const std::vector<int> ts = ddim_timesteps(h.cfg.noise_steps, inference_steps);
  const size_t plane = (size_t)nb * Hh * W;
  const bool traced = eps_trace || xin_trace || teacher_x;
  const std::string key = keyf("ddim", H, nb, Hh, W, inference_steps);

  if (!h.use_graph || traced || h.audit) {      // the audit kernels are not part of the captured graph
    with_arena(H, s, key + (traced ? "|eager-trace" : "|eager"), [&](Ctx& c) {
      float* xcur = c.allocf(plane);
      float* temb = c.allocf(ts.size() * (size_t)h.unet.te.total);
      int* tl = (int*)c.a->alloc(ts.size() * sizeof(int));
      if (!c.dry) {
        XRD_CUDA(cudaMemcpyAsync(tl, ts.data(), ts.size() * sizeof(int), cudaMemcpyHostToDevice, c.s));
        XRD_CUDA(cudaStreamSynchronize(c.s));    // ts is a stack object
      }
      time_embed(c, h.unet.te, nullptr, tl, (int)ts.size(), temb);
      copy_plane(c, noisy, xcur, plane);
      run_ddim_loop(c, h, noisy, xcur, temb, ts, eps_trace, xin_trace, teacher_x, trace_stride, nb, Hh, W);
      copy_plane(c, xcur, out, plane);
    });
    return;
  }

  // ---- graph path: plan, (re)capture if needed, replay ----
  auto body = [&](Ctx& c, float** in_p, float** x_p, const float* temb) {
    float* in = c.allocf(plane);
    float* xcur = c.allocf(plane);
    if (in_p) *in_p = in;
    if (x_p) *x_p = xcur;
    copy_plane(c, in, xcur, plane);
    run_ddim_loop(c, h, in, xcur, temb, ts, nullptr, nullptr, nullptr, 0, nb, Hh, W);
  };
  auto git = h.graphs.find(key);
  if (git == h.graphs.end()) {
    // plan + make sure the arena is large enough (this may drop other graphs)
    with_arena(H, s, key + "|plan", [&](Ctx& c) {
      if (c.dry) body(c, nullptr, nullptr, nullptr);
    });
    GraphEntry ge;
    XRD_CUDA(cudaMalloc((void**)&ge.temb, ts.size() * (size_t)h.unet.te.total * sizeof(float)));
    compute_temb_table(H, s, ts, ge.temb);
    // warm-up pass outside capture (sets function attributes, validates every launch configuration)
    {
      h.arena.off = 0;
      Ctx c = make_ctx(H, s, false, &h.arena);
      body(c, &ge.in, &ge.out, ge.temb);
      XRD_CUDA(cudaStreamSynchronize(s));
    }
    cudaGraph_t graph = nullptr;
    XRD_CUDA(cudaStreamBeginCapture(H->cap_stream, cudaStreamCaptureModeThreadLocal));
    try {
      h.arena.off = 0;
      Ctx c = make_ctx(H, H->cap_stream, false, &h.arena);
      const uint64_t before = g_launches.load();
      body(c, &ge.in, &ge.out, ge.temb);
      g_launches.store(before);                  // capture launches nothing; replays are counted below
    } catch (...) {
      cudaStreamEndCapture(H->cap_stream, &graph);
      if (graph) cudaGraphDestroy(graph);
      cudaFree(ge.temb);
      throw;
    }
    XRD_CUDA(cudaStreamEndCapture(H->cap_stream, &graph));
    ge.kernels = count_kernel_nodes(graph);
    cudaError_t e = cudaGraphInstantiate(&ge.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { cudaFree(ge.temb); fail(XRD_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e)); }
    git = h.graphs.emplace(key, ge).first;
  }
  GraphEntry& ge = git->second;
  XRD_CUDA(cudaMemcpyAsync(ge.in, noisy, plane * 4, cudaMemcpyDeviceToDevice, s));
  XRD_CUDA(cudaGraphLaunch(ge.exec, s));
  g_launches.fetch_add(ge.kernels);
  XRD_CUDA(cudaMemcpyAsync(out, ge.out, plane * 4, cudaMemcpyDeviceToDevice, s));
}

XRD_EXPORT int xrd_ddim_denoise(xrd_handle* H, const float* noisy, int inference_steps, float* out, float* eps_trace, float* xin_trace,
                                const float* teacher_x, int B, int Hh, int W, void* stream) {
  return guarded([&] {
    XRD_REQUIRE(H && noisy && out, "null argument");
    XRD_REQUIRE(inference_steps >= 1, "inference_steps must be >= 1");
    check_img(B, Hh, W);
    std::lock_guard<std::mutex> lk(H->h.mu);
    XRD_REQUIRE(H->h.unet.ready, "UNet weights are not finalised");
    check_unet_shape(H, Hh, W);
    cudaStream_t s = (cudaStream_t)stream;
    DeviceScope dscope(H->h.device);
    bind_stream(H, s);
    const int mb = micro_batch(B, Hh, W);
    const size_t img = (size_t)Hh * W;
    const size_t tstride = (size_t)B * img;   // traces are (n_evals, B, H, W)
    for (int b0 = 0; b0 < B; b0 += mb) {
      const int nb = std::min(mb, B - b0);
      ddim_chunk(H, s, noisy + b0 * img, inference_steps, out + b0 * img, eps_trace ? eps_trace + b0 * img : nullptr,
                 xin_trace ? xin_trace + b0 * img : nullptr, teacher_x ? teacher_x + b0 * img : nullptr, tstride, nb, Hh, W);
    }
  });
}

XRD_EXPORT int xrd_nafnet(xrd_handle* H, const float* inp, float* out, int B, int Hh, int W, void* stream) {
  return guarded([&] {
    XRD_REQUIRE(H && inp && out, "null argument");
    check_img(B, Hh, W);
    std::lock_guard<std::mutex> lk(H->h.mu);
    XRD_REQUIRE(H->h.naf.ready, "NAFNet weights are not finalised");
    cudaStream_t s = (cudaStream_t)stream;
    DeviceScope dscope(H->h.device);
    bind_stream(H, s);
    const int mb = micro_batch(B, Hh, W);
    const size_t img = (size_t)Hh * W;
    for (int b0 = 0; b0 < B; b0 += mb) {
      const int nb = std::min(mb, B - b0);
      with_arena(H, s, keyf("naf", H, nb, Hh, W), [&](Ctx& c) { run_nafnet(c, H->h, inp + b0 * img, out + b0 * img, nb, Hh, W, 0); });
    }
  });
}

XRD_EXPORT int xrd_router(xrd_handle* H, const float* x, float* mask, int B, int Hh, int W, void* stream) {
  return guarded([&] {
    XRD_REQUIRE(H && x && mask, "null argument");
    check_img(B, Hh, W);
    std::lock_guard<std::mutex> lk(H->h.mu);
    XRD_REQUIRE(H->h.router.ready, "router weights are not finalised");
    cudaStream_t s = (cudaStream_t)stream;
    DeviceScope dscope(H->h.device);
    bind_stream(H, s);
    const int mb = micro_batch(B, Hh, W);
    const size_t img = (size_t)Hh * W;
    for (int b0 = 0; b0 < B; b0 += mb) {
      const int nb = std::min(mb, B - b0);
      with_arena(H, s, keyf("router", H, nb, Hh, W), [&](Ctx& c) { run_router(c, H->h, x + b0 * img, mask + b0 * img, nb, Hh, W, 0); });
    }
  });
}

XRD_EXPORT int xrd_fusion(xrd_handle* H, const float* naf, const float* diff, const float* mask, float* out, int B, int Hh, int W,
                          void* stream) {
  return guarded([&] {
    XRD_REQUIRE(H && naf && diff && mask && out, "null argument");
    check_img(B, Hh, W);
    std::lock_guard<std::mutex> lk(H->h.mu);
    XRD_REQUIRE(H->h.fusion.ready, "fusion weights are not finalised");
    cudaStream_t s = (cudaStream_t)stream;
    DeviceScope dscope(H->h.device);
    bind_stream(H, s);
    const int mb = micro_batch(B, Hh, W);
    const size_t img = (size_t)Hh * W;
    for (int b0 = 0; b0 < B; b0 += mb) {
      const int nb = std::min(mb, B - b0);
      with_arena(H, s, keyf("fusion", H, nb, Hh, W),
                 [&](Ctx& c) { run_fusion(c, H->h, naf + b0 * img, diff + b0 * img, mask + b0 * img, out + b0 * img, nb, Hh, W); });
    }
  });
}

XRD_EXPORT int xrd_expert(xrd_handle* H, const float* inp, float* out, int B, int Hh, int W, void* stream) {
  return guarded([&] {
    XRD_REQUIRE(H && inp && out, "null argument");
    check_img(B, Hh, W);
    std::lock_guard<std::mutex> lk(H->h.mu);
    XRD_REQUIRE(H->h.expert.ready, "ExpertDenoiser weights are not finalised");
    cudaStream_t s = (cudaStream_t)stream;
    DeviceScope dscope(H->h.device);
    bind_stream(H, s);
    // 128 + 128 channels live at full resolution: a quarter of the sampler's micro-batch bounds the workspace
    const int mb = std::max(1, micro_batch(B, Hh, W) / 4);
    const size_t img = (size_t)Hh * W;
    for (int b0 = 0; b0 < B; b0 += mb) {
      const int nb = std::min(mb, B - b0);
      with_arena(H, s, keyf("expert", H, nb, Hh, W), [&](Ctx& c) { run_expert(c, H->h, inp + b0 * img, out + b0 * img, nb, Hh, W); });
    }
  });
}

XRD_EXPORT int xrd_set_range_audit(xrd_handle* H, int enable) {
  return guarded([&] {
    XRD_REQUIRE(H, "null handle");
    std::lock_guard<std::mutex> lk(H->h.mu);
    DeviceScope dscope(H->h.device);
    if (enable && !H->h.audit_dev) {
      XRD_CUDA(cudaMalloc((void**)&H->h.audit_dev, sizeof(RangeAudit)));
      XRD_CUDA(cudaMemset(H->h.audit_dev, 0, sizeof(RangeAudit)));
    }
    H->h.audit = enable != 0;
  });
}

XRD_EXPORT int xrd_get_range_report(xrd_handle* H, uint64_t* saturated, uint64_t* nonfinite, uint64_t* elements, float* absmax,
                                    uint32_t* tensors, int reset) {
  return guarded([&] {
    XRD_REQUIRE(H, "null handle");
    std::lock_guard<std::mutex> lk(H->h.mu);
    DeviceScope dscope(H->h.device);
    RangeAudit r;
    memset(&r, 0, sizeof(r));
    if (H->h.audit_dev) {
      if (H->have_last_stream) XRD_CUDA(cudaStreamSynchronize(H->last_stream));
      XRD_CUDA(cudaMemcpy(&r, H->h.audit_dev, sizeof(r), cudaMemcpyDeviceToHost));
      if (reset) XRD_CUDA(cudaMemset(H->h.audit_dev, 0, sizeof(RangeAudit)));
    }
    if (saturated) *saturated = r.saturated;
    if (nonfinite) *nonfinite = r.nonfinite;
    if (elements) *elements = r.elements;
    if (tensors) *tensors = r.tensors;
    if (absmax) { float f; memcpy(&f, &r.absmax_bits, 4); *absmax = f; }
  });
}

XRD_EXPORT int xrd_hybrid(xrd_handle* H, const float* noisy, int inference_steps, float* out, float* naf_out, float* diff_out,
                          float* mask_out, int B, int Hh, int W, void* stream) {
  return guarded([&] {
    XRD_REQUIRE(H && noisy && out, "null argument");
    XRD_REQUIRE(inference_steps >= 1, "inference_steps must be >= 1");
    check_img(B, Hh, W);
    std::lock_guard<std::mutex> lk(H->h.mu);
    Handle& h = H->h;
    XRD_REQUIRE(h.unet.ready && h.naf.ready && h.router.ready && h.fusion.ready, "hybrid weights are not finalised");
    check_unet_shape(H, Hh, W);
    cudaStream_t s = (cudaStream_t)stream;
    DeviceScope dscope(H->h.device);
    bind_stream(H, s);
    const int mb = micro_batch(B, Hh, W);
    const size_t img = (size_t)Hh * W;
    // persistent per-call planes (outside the arena, which every stage resets)
    const size_t need = 3 * (size_t)mb * img;
    if (H->scratch_n < need) {
      if (H->scratch) { XRD_CUDA(cudaDeviceSynchronize()); cudaFree(H->scratch); H->scratch = nullptr; H->scratch_n = 0; }
      XRD_CUDA(cudaMalloc((void**)&H->scratch, need * 4));
      H->scratch_n = need;
    }
    float* scratch = H->scratch;
    try {
    for (int b0 = 0; b0 < B; b0 += mb) {
      const int nb = std::min(mb, B - b0);
      const size_t plane = (size_t)nb * img;
      float* nafp = naf_out ? naf_out + b0 * img : scratch;
      float* difp = diff_out ? diff_out + b0 * img : scratch + (size_t)mb * img;
      float* mskp = mask_out ? mask_out + b0 * img : scratch + 2 * (size_t)mb * img;
      const float* in = noisy + b0 * img;
      // The three branches read only `in` (HYB:612-624).  With side branches on, NAFNet and the router are enqueued on the two
      // private streams behind a fork event and work out of their own arenas; the sampler graph stays on the caller's stream; the
      // fusion stack waits for both joins.  The next micro-batch forks after this one's fusion, so the scratch planes are never
      // overwritten while they are read.  The range audit and the in-situ launch profile want one ordered stream: serial then.
      const bool par = H->overlap && !h.audit && !g_prof && (int64_t)nb * Hh * W <= H->overlap_max_px;
      cudaStream_t s_naf = par ? H->side[0] : s, s_rt = par ? H->side[1] : s;
      if (par) {
        XRD_CUDA(cudaEventRecord(H->ev_fork, s));
        XRD_CUDA(cudaStreamWaitEvent(s_naf, H->ev_fork, 0));
        XRD_CUDA(cudaStreamWaitEvent(s_rt, H->ev_fork, 0));
      }
      // fast path: NAFNet -> nan_to_num + clamp (HYB:614-616)
      with_arena_in(H, s_naf, keyf("naf", H, nb, Hh, W, 1), par ? H->side_arena[0] : h.arena,
                    [&](Ctx& c) { run_nafnet(c, h, in, nafp, nb, Hh, W, 1); });
      if (par) {
        XRD_CUDA(cudaEventRecord(H->ev_join[0], s_naf));
        // routing mask (HYB:622-624)
        with_arena_in(H, s_rt, keyf("router", H, nb, Hh, W, 1), H->side_arena[1], [&](Ctx& c) { run_router(c, h, in, mskp, nb, Hh, W, 1); });
        XRD_CUDA(cudaEventRecord(H->ev_join[1], s_rt));
      }
      // quality path: sampler -> nan_to_num + clamp (HYB:618-620); x is already clamped to [0,1] by the last update
      ddim_chunk(H, s, in, inference_steps, difp, nullptr, nullptr, nullptr, 0, nb, Hh, W);
      {
        Arena none;
        Ctx c = make_ctx(H, s, false, &none);
        sanitize_plane(c, difp, difp, plane);
      }
      if (par) {
        XRD_CUDA(cudaStreamWaitEvent(s, H->ev_join[0], 0));
        XRD_CUDA(cudaStreamWaitEvent(s, H->ev_join[1], 0));
      } else {
        with_arena(H, s, keyf("router", H, nb, Hh, W, 1), [&](Ctx& c) { run_router(c, h, in, mskp, nb, Hh, W, 1); });
      }
      // fusion (HYB:626)
      with_arena(H, s, keyf("fusion", H, nb, Hh, W), [&](Ctx& c) { run_fusion(c, h, nafp, difp, mskp, out + b0 * img, nb, Hh, W); });
    }
    } catch (...) {
      // a failed stage must not leave a side branch running behind the caller's back (it reads the caller's input)
      for (int i = 0; i < 2; ++i) cudaStreamSynchronize(H->side[i]);
      throw;
    }
  });
}

// ------------------------------------------------------------------------------------------------
// kernel-level hooks
// ------------------------------------------------------------------------------------------------
XRD_EXPORT int xrd_op_conv2d(xrd_handle* H, int impl, const float* x, const float* weight, const float* bias, float* y, int B, int Cin,
                             int Hh, int W, int Cout, int k, int stride, int pad, void* stream) {
  return xrd_op_conv2d_stats(H, impl, x, weight, bias, y, nullptr, B, Cin, Hh, W, Cout, k, stride, pad, stream);
}

XRD_EXPORT int xrd_op_conv2d_stats(xrd_handle* H, int impl, const float* x, const float* weight, const float* bias, float* y, double* stats,
                                   int B, int Cin, int Hh, int W, int Cout, int k, int stride, int pad, void* stream) {
  return guarded([&] {
    XRD_REQUIRE(H && x && weight && y, "null argument");
    std::lock_guard<std::mutex> lk(H->h.mu);
    cudaStream_t s = (cudaStream_t)stream;
    DeviceScope dscope(H->h.device);
    bind_stream(H, s);
    XRD_CUDA(cudaStreamSynchronize(s));
    free_op_state(H);
    ConvW& w = H->op_w;
    w.kh = w.kw = k; w.stride = stride; w.pad = pad; w.cin = Cin; w.cout = Cout;
    void* wp = nullptr;
    XRD_CUDA(cudaMalloc(&wp, (size_t)Cout * Cin * k * k * 4));
    H->op_owned.push_back(wp);
    w.w = (float*)wp;
    pack_conv_weight(s, weight, w.w, Cout, Cin, k, k);
    w.bias = (float*)bias;
    const DType dt = mode_dtype(H->h.mode);
    // impl: 0 CUDA cores, 1 tcgen05 per-tap, 2 tcgen05 persistent halo (conv3), 3 = conv3 over a virtual concat of the two
    // channel halves of x (B must be 1 so that each half is a contiguous NCHW block), 4 = per-tap kernel over the same concat,
    // 5 = persistent 1x1 GEMM (conv1), 6 = conv1 over the concat, 7 = 64-wide 3x3 kernel (conv3w), 8 = conv3w over the concat,
    // 11 = row-ring kernel (conv3r), 12 = conv3r with GroupNorm(8) + SiLU of the input applied inside the kernel,
    // 9 / 10 = conv3 (single source / concat) with GroupNorm(8) + SiLU of the input applied inside the kernel; affine parameters
    // gamma[c] = 1 + 0.01*(c % 7), beta[c] = 0.02*(c % 5) - 0.03
    // 13 = conv1 with the NAFBlock FFN epilogue (HYB:165-169): y = SimpleGate(conv(x) + bias) * bias[:Cout/2] + x  (Cin == Cout/2),
    // 14 = conv1 with the scaled-residual epilogue (HYB:161): y = (conv(x) + bias) * bias + x  (Cin == Cout); bias doubles as the scale
    // 15 = N-stacked row-ring kernel (conv3s), 16 = conv3s with GroupNorm(8) + SiLU of the input applied inside the kernel,
    // 17 / 18 = the same two over the virtual concat of the channel halves
    // 19 = the UNet's first conv (first_conv.cu): Cin == 2, Cout == 48, the two input channels read as fp32 planes (B must be 1)
    if (impl == 19) {
      XRD_REQUIRE(dt != DT_F32 && B == 1 && Cin == 2 && k == 3 && stride == 1 && pad == 1, "hook 19: B == 1, Cin == 2, 3x3/s1/p1, 16-bit mode");
      H->h.last_op = nullptr;
      with_arena(H, s, keyf("opfirst", H, B, Hh, W, Cin, Cout), [&](Ctx& c) {
        Tens yo = c.alloc(B, Hh, W, Cout);
        double* st = c.allocd((size_t)B * 16);
        XRD_REQUIRE(first_conv_mma_supported(yo, H->op_w), "hook 19: unsupported shape");
        auto run = [H, x, yo, st, Hh, W](Ctx& cc) mutable {
          Tens yy = yo;
          zero_async(cc, st, (size_t)yo.n * 16 * sizeof(double));
          first_conv_mma(cc, x, x + (size_t)Hh * W, H->op_w, yy, st);
        };
        run(c);
        nhwc_to_nchw(c, yo, y);
        if (stats && !c.dry) XRD_CUDA(cudaMemcpyAsync(stats, st, (size_t)B * 16 * sizeof(double), cudaMemcpyDeviceToDevice, c.s));
        H->h.last_op = run;
        H->h.last_op_bytes = yo.bytes() + (size_t)2 * Hh * W * 4;
      });
      return;
    }
    const bool split = impl == 3 || impl == 4 || impl == 6 || impl == 8 || impl == 10 || impl == 17 || impl == 18;
    const bool fuse_gn = impl == 9 || impl == 10 || impl == 12 || impl == 16 || impl == 18;
    if (split) XRD_REQUIRE(Cin % 32 == 0, "split-input conv hook needs Cin %% 32 == 0");
    if (impl >= 1) {
      XRD_REQUIRE(dt != DT_F32, "the tcgen05 kernels need a 16-bit mode");
      conv_tc_pack(s, w, dt, split ? Cin / 2 : Cin);
    }
    const int Ho = (Hh + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    H->h.last_op = nullptr;
    with_arena(H, s, keyf("opconv", H, B, Hh, W, Cin, Cout, k, stride * 8 + pad, impl), [&](Ctx& c) {
      Tens xi = c.alloc(B, Hh, W, split ? Cin / 2 : Cin);
      Tens xj = split ? c.alloc(B, Hh, W, Cin / 2) : Tens();
      if (impl == 13) XRD_REQUIRE(bias && Cin * 2 == Cout && k == 1, "hook 13 needs a bias, k == 1 and Cin == Cout/2");
      if (impl == 14) XRD_REQUIRE(bias && Cin == Cout && k == 1, "hook 14 needs a bias, k == 1 and Cin == Cout");
      Tens yo = c.alloc(B, Ho, Wo, impl == 13 ? Cout / 2 : Cout);
      double* st = c.allocd((size_t)B * 16);
      double* gsum = fuse_gn ? c.allocd((size_t)B * 16) : nullptr;
      float2* gcoef = fuse_gn ? (float2*)c.a->alloc((size_t)B * Cin * sizeof(float2)) : nullptr;
      float* gaff = fuse_gn ? c.allocf((size_t)2 * Cin) : nullptr;
      if (!split) {
        nchw_to_nhwc(c, x, xi);
      } else {              // the two channel halves of every image become the two sources of the virtual concat
        for (int b = 0; b < B; ++b) {
          for (int half = 0; half < 2; ++half) {
            Tens v = half ? xj : xi;
            v.n = 1;
            v.p = (char*)v.p + (size_t)b * Hh * W * (Cin / 2) * dsize(v.dt);
            nchw_to_nhwc(c, x + ((size_t)b * Cin + (size_t)half * (Cin / 2)) * Hh * W, v);
          }
        }
      }
      if (fuse_gn && !c.dry) {
        std::vector<float> aff(2 * (size_t)Cin);
        for (int ch = 0; ch < Cin; ++ch) { aff[ch] = 1.0f + 0.01f * (float)(ch % 7); aff[Cin + ch] = 0.02f * (float)(ch % 5) - 0.03f; }
        XRD_CUDA(cudaMemcpyAsync(gaff, aff.data(), aff.size() * sizeof(float), cudaMemcpyHostToDevice, s));
        XRD_CUDA(cudaStreamSynchronize(s));
      }
      if (fuse_gn) {
        // the GroupNorm coefficients of the input: computed once here, outside the timed closure -- in the networks the sums
        // come out of the producing kernel's epilogue, the conv launch is all that the fused layer costs
        zero_async(c, gsum, (size_t)yo.n * 16 * sizeof(double));
        gn_stats(c, xi, split ? &xj : nullptr, 8, gsum);
        gn_coef(c, gsum, gaff, gaff + Cin, 1e-5f, yo.n, Cin, 8, xi.h * xi.w, gcoef);
      }
      auto run = [H, xi, xj, yo, st, impl, split, fuse_gn, gcoef](Ctx& cc) mutable {
        Tens yy = yo;
        ConvEpi e;
        if (fuse_gn) {
          e.in_coef = gcoef; e.in_act = ACT_SILU;
          e.stats_out = st;
          zero_async(cc, st, (size_t)yo.n * 16 * sizeof(double));
          if (impl == 12) conv3r(cc, xi, H->op_w, e, yy);
          else if (impl == 16 || impl == 18) conv3s(cc, xi, split ? &xj : nullptr, H->op_w, e, yy);
          else conv3(cc, xi, split ? &xj : nullptr, H->op_w, e, yy);
        } else if (impl == 15 || impl == 17) {
          e.stats_out = st;
          zero_async(cc, st, (size_t)yo.n * 16 * sizeof(double));
          conv3s(cc, xi, split ? &xj : nullptr, H->op_w, e, yy);
        } else if (impl == 11) {
          e.stats_out = st;
          zero_async(cc, st, (size_t)yo.n * 16 * sizeof(double));
          conv3r(cc, xi, H->op_w, e, yy);
        } else if (impl == 2 || impl == 3) {
          e.stats_out = st;
          zero_async(cc, st, (size_t)yo.n * 16 * sizeof(double));
          conv3(cc, xi, split ? &xj : nullptr, H->op_w, e, yy);
        } else if (impl == 7 || impl == 8) {
          e.stats_out = st;
          zero_async(cc, st, (size_t)yo.n * 16 * sizeof(double));
          conv3w(cc, xi, split ? &xj : nullptr, H->op_w, e, yy);
        } else if (impl == 5 || impl == 6) {
          if (conv1_supported(xi, split ? &xj : nullptr, H->op_w, [&] { ConvEpi t; t.stats_out = st; return t; }())) {
            e.stats_out = st;
            zero_async(cc, st, (size_t)yo.n * 16 * sizeof(double));
          }
          conv1(cc, xi, split ? &xj : nullptr, H->op_w, e, yy);
        }
        else if (impl == 13 || impl == 14) {
          e.gate = impl == 13;
          e.out_scale = H->op_w.bias;
          e.resid = xi;
          conv1(cc, xi, nullptr, H->op_w, e, yy);
        }
        else if (impl == 1 || impl == 4) conv_tc(cc, xi, split ? &xj : nullptr, H->op_w, e, yy);
        else conv_simt(cc, xi, nullptr, H->op_w, e, yy);
      };
      run(c);
      if (!c.dry && stats && (impl == 2 || impl == 3 || impl >= 5)) XRD_CUDA(cudaMemcpyAsync(stats, st, (size_t)B * 16 * sizeof(double), cudaMemcpyDeviceToDevice, s));
      nhwc_to_nchw(c, yo, y);
      if (!c.dry) {
        H->h.last_op = run;
        H->h.last_op_bytes = xi.bytes() + yo.bytes();
      }
    });
  });
}

XRD_EXPORT int xrd_op_groupnorm_act(xrd_handle* H, const float* x, const float* gamma, const float* beta, float* y, int B, int C, int Hh,
                                    int W, int groups, int act, void* stream) {
  return guarded([&] {
    XRD_REQUIRE(H && x && gamma && beta && y, "null argument");
    std::lock_guard<std::mutex> lk(H->h.mu);
    cudaStream_t s = (cudaStream_t)stream;
    DeviceScope dscope(H->h.device);
    bind_stream(H, s);
    H->h.last_op = nullptr;
    with_arena(H, s, keyf("opgn", H, B, Hh, W, C, groups), [&](Ctx& c) {
      Tens xi = c.alloc(B, Hh, W, C);
      Tens yo = c.alloc(B, Hh, W, C);
      double* sums = c.allocd((size_t)B * groups * 2);
      nchw_to_nhwc(c, x, xi);
      auto run = [xi, yo, sums, gamma, beta, groups, act, B](Ctx& cc) mutable {
        Tens yy = yo;
        zero_async(cc, sums, (size_t)B * groups * 2 * sizeof(double));
        gn_stats(cc, xi, nullptr, groups, sums);
        gn_act(cc, xi, nullptr, groups, sums, gamma, beta, 1e-5f, act, yy);
      };
      run(c);
      nhwc_to_nchw(c, yo, y);
      if (!c.dry) { H->h.last_op = run; H->h.last_op_bytes = 2 * xi.bytes() + yo.bytes(); }
    });
  });
}

XRD_EXPORT int xrd_op_attention(xrd_handle* H, int impl, const float* qkv, float* out, int B, int heads, int d, int Hh, int W,
                                void* stream) {
  return guarded([&] {
    XRD_REQUIRE(H && qkv && out, "null argument");
    std::lock_guard<std::mutex> lk(H->h.mu);
    cudaStream_t s = (cudaStream_t)stream;
    DeviceScope dscope(H->h.device);
    bind_stream(H, s);
    H->h.last_op = nullptr;
    with_arena(H, s, keyf("opattn", H, B, Hh, W, heads, d, impl), [&](Ctx& c) {
      Tens q = c.alloc(B, Hh, W, 3 * heads * d);
      Tens o = c.alloc(B, Hh, W, heads * d);
      float* ws = nullptr;                       // workspace of the split-key launch shape (single images)
      if (impl == 1 && attention_tc_supported(q, heads)) {
        const size_t nf = attention_tc_scratch_floats(q, heads);
        if (nf) ws = c.allocf(nf);
      }
      nchw_to_nhwc(c, qkv, q);
      auto run = [q, o, heads, impl, ws](Ctx& cc) mutable {
        Tens oo = o;
        if (impl == 1) attention_tc(cc, q, heads, oo, ws);
        else attention_simt(cc, q, heads, oo);
      };
      run(c);
      nhwc_to_nchw(c, o, out);
      if (!c.dry) { H->h.last_op = run; H->h.last_op_bytes = q.bytes() + o.bytes(); }
    });
  });
}

XRD_EXPORT int xrd_op_time_last(xrd_handle* H, int iters, float* ms_per_launch, void* stream) {
  return guarded([&] {
    XRD_REQUIRE(H && ms_per_launch && iters >= 1, "bad argument");
    std::lock_guard<std::mutex> lk(H->h.mu);
    XRD_REQUIRE((bool)H->h.last_op, "no op to time: call an xrd_op_* function first");
    cudaStream_t s = (cudaStream_t)stream;
    DeviceScope dscope(H->h.device);
    bind_stream(H, s);
    Arena none;
    Ctx c = make_ctx(H, s, false, &none);
    cudaEvent_t e0, e1;
    XRD_CUDA(cudaEventCreate(&e0));
    XRD_CUDA(cudaEventCreate(&e1));
    // warm up for ~0.3 s of GPU time so the SM clock has ramped before the timed launches
    {
      XRD_CUDA(cudaEventRecord(e0, s));
      for (int i = 0; i < 3; ++i) H->h.last_op(c);
      XRD_CUDA(cudaEventRecord(e1, s));
      XRD_CUDA(cudaEventSynchronize(e1));
      float w3 = 0.f;
      XRD_CUDA(cudaEventElapsedTime(&w3, e0, e1));
      int extra = (int)std::min(2000.0f, 300.0f / std::max(w3 / 3.0f, 0.001f));
      for (int i = 0; i < extra; ++i) H->h.last_op(c);
    }
    XRD_CUDA(cudaEventRecord(e0, s));
    for (int i = 0; i < iters; ++i) H->h.last_op(c);
    XRD_CUDA(cudaEventRecord(e1, s));
    XRD_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    XRD_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_per_launch = ms / (float)iters;
  });
}

// ------------------------------------------------------------------------------------------------
// overlap tiling (BASELINE configs[4]); no handle: the kernels need a stream only
XRD_EXPORT int xrd_tiles_plan(int H, int W, int tile, int halo, int* ny, int* nx, int* oy, int* ox, int cap) {
  return guarded([&] {
    tiles_check(1, H, W, tile, halo);
    const int cy = tiles_count_host(H, tile, halo), cx = tiles_count_host(W, tile, halo);
    if (ny) *ny = cy;
    if (nx) *nx = cx;
    if (oy) { XRD_REQUIRE(cap >= cy, "origin buffer too small (%d < %d)", cap, cy); for (int k = 0; k < cy; ++k) oy[k] = tiles_origin_host(k, H, tile, halo); }
    if (ox) { XRD_REQUIRE(cap >= cx, "origin buffer too small (%d < %d)", cap, cx); for (int k = 0; k < cx; ++k) ox[k] = tiles_origin_host(k, W, tile, halo); }
  });
}

XRD_EXPORT int xrd_tiles_extract(const float* img, float* tiles, int B, int H, int W, int tile, int halo, void* stream) {
  return guarded([&] {
    XRD_REQUIRE(img && tiles, "null argument");
    Ctx c;
    c.s = (cudaStream_t)stream;
    DeviceScope dscope(device_of(img));
    tiles_extract(c, img, tiles, B, H, W, tile, halo);
  });
}

XRD_EXPORT int xrd_tiles_blend(const float* tiles, float* img, int B, int H, int W, int tile, int halo, void* stream) {
  return guarded([&] {
    XRD_REQUIRE(img && tiles, "null argument");
    Ctx c;
    c.s = (cudaStream_t)stream;
    DeviceScope dscope(device_of(tiles));
    tiles_blend(c, tiles, img, B, H, W, tile, halo);
  });
}

// ------------------------------------------------------------------------------------------------
// pre/post-processing of `/denoise` on the GPU (Pillow-exact 8-bit bicubic resampling); no handle
XRD_EXPORT int xrd_resize_bicubic_u8(const uint8_t* src, uint8_t* dst, uint8_t* tmp, int N, int Hin, int Win, int Hout, int Wout, void* stream) {
  return guarded([&] {
    XRD_REQUIRE(src && dst, "null argument");
    Ctx c;
    c.s = (cudaStream_t)stream;
    DeviceScope dscope(device_of(src));
    resize_bicubic_u8(c, src, dst, tmp, N, Hin, Win, Hout, Wout);
  });
}

XRD_EXPORT int xrd_resample_table(int in_size, int out_size, int* ksize, int* bounds, int* kk, int cap_k) {
  return guarded([&] {
    XRD_REQUIRE(in_size >= 1 && out_size >= 1 && ksize, "bad argument");
    *ksize = resample_table_host(in_size, out_size, bounds, kk, cap_k);
  });
}

XRD_EXPORT int xrd_u8_to_unit(const uint8_t* src, float* dst, int64_t n, void* stream) {
  return guarded([&] {
    XRD_REQUIRE(src && dst && n >= 0, "bad argument");
    Ctx c;
    c.s = (cudaStream_t)stream;
    DeviceScope dscope(device_of(src));
    if (n) u8_to_unit(c, src, dst, n);
  });
}

XRD_EXPORT int xrd_unit_to_u8(const float* src, uint8_t* dst, int64_t n, void* stream) {
  return guarded([&] {
    XRD_REQUIRE(src && dst && n >= 0, "bad argument");
    Ctx c;
    c.s = (cudaStream_t)stream;
    DeviceScope dscope(device_of(src));
    if (n) unit_to_u8(c, src, dst, n);
  });
}
