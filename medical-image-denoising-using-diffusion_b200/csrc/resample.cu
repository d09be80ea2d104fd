// resample.cu -- the pixel work either side of the networks in `/denoise`, on the GPU and bit-exact:
//   pre  (RUN:191-201): PIL 'L' image -> transforms.Resize((512,512), BICUBIC) -> ToTensor()            = uint8 resample, then v / 255
//   post (RUN:143-149): clamp(0,1) -> (v * 255).astype('uint8') (truncation) -> Image.resize(size, BICUBIC) = truncate, uint8 resample
// The arithmetic lives in the reference's third-party dependency Pillow (Backend/requirements.txt; libImaging/Resample.c), whose
// published algorithm for 8-bit images is restated here: per output coordinate a window [xmin, xmin+n) of the Keys cubic
// (a = -0.5) stretched by max(1, in/out) (antialiasing when shrinking), computed in double, normalised, converted to fixed
// point with 22 fractional bits (round half away from zero); a horizontal pass then a vertical pass, each accumulating in
// int32 from 1 << 21 and clipping (sum >> 22) to [0, 255] -- the intermediate image is 8-bit, as in Pillow.
// The coefficient tables are built on the host in double exactly as Pillow's C does and cached per (device, in, out).
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

namespace xrd {

namespace {
constexpr int kPrecisionBits = 32 - 8 - 2;

inline double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

struct Table {
  int ksize = 0;
  int* bounds = nullptr;   // device [out][2]: first input index, count
  int* kk = nullptr;       // device [out][ksize] fixed-point weights
};

// Pillow precompute_coeffs + normalize_coeffs_8bpc for a full-image box
void build_table(int in_size, int out_size, std::vector<int>& bounds, std::vector<int>& kk, int& ksize) {
  const double scale = (double)in_size / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  ksize = (int)std::ceil(support) * 2 + 1;
  bounds.assign((size_t)out_size * 2, 0);
  kk.assign((size_t)out_size * ksize, 0);
  std::vector<double> k((size_t)ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0.0 + (xx + 0.5) * scale;
    double ww = 0.0;
    const double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    int x;
    for (x = 0; x < xmax; ++x) {
      const double w = bicubic_filter((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (x = 0; x < xmax; ++x)
      if (ww != 0.0) k[x] /= ww;
    for (; x < ksize; ++x) k[x] = 0;
    bounds[(size_t)xx * 2 + 0] = xmin;
    bounds[(size_t)xx * 2 + 1] = xmax;
    for (x = 0; x < ksize; ++x) {
      const double v = k[x];
      kk[(size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << kPrecisionBits)) : (int)(0.5 + v * (1 << kPrecisionBits));
    }
  }
}

// Device tables are cached per (device, in, out): a table lives on the device it was built on.  The cache is bounded (every
// upload size of `/denoise` adds a (512 -> W0/H0) pair): least recently used entries are freed -- cudaFree waits for the
// kernels that may still read them.
std::mutex g_mu;
struct Entry { Table t; uint64_t stamp = 0; };
std::map<std::tuple<int, int, int>, Entry> g_tables;
uint64_t g_stamp = 0;
constexpr size_t kMaxTables = 64;

Table table_for(int in_size, int out_size) {
  int dev = 0;
  XRD_CUDA(cudaGetDevice(&dev));           // the entry point made the buffers' device current
  std::lock_guard<std::mutex> lk(g_mu);
  const auto key = std::make_tuple(dev, in_size, out_size);
  auto it = g_tables.find(key);
  if (it != g_tables.end()) { it->second.stamp = ++g_stamp; return it->second.t; }
  if (g_tables.size() >= kMaxTables) {
    auto old = g_tables.begin();
    for (auto j = g_tables.begin(); j != g_tables.end(); ++j) if (j->second.stamp < old->second.stamp) old = j;
    {
      DeviceScope ds(std::get<0>(old->first));
      cudaFree(old->second.t.bounds);
      cudaFree(old->second.t.kk);
    }
    g_tables.erase(old);
  }
  std::vector<int> b, k;
  Entry e;
  build_table(in_size, out_size, b, k, e.t.ksize);
  XRD_CUDA(cudaMalloc((void**)&e.t.bounds, b.size() * sizeof(int)));
  XRD_CUDA(cudaMalloc((void**)&e.t.kk, k.size() * sizeof(int)));
  XRD_CUDA(cudaMemcpy(e.t.bounds, b.data(), b.size() * sizeof(int), cudaMemcpyHostToDevice));
  XRD_CUDA(cudaMemcpy(e.t.kk, k.data(), k.size() * sizeof(int), cudaMemcpyHostToDevice));
  e.stamp = ++g_stamp;
  g_tables.emplace(key, e);
  return e.t;
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}
}  // namespace

// out[n][y][xx] = clip8(2^21 + sum_x in[n][y][xmin + x] * k[xx][x]);  thread = one output pixel
__global__ void k_resample_h_u8(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int N, int H, int Win, int Wout,
                                const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
  const int64_t total = (int64_t)N * H * Wout;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xx = (int)(i % Wout);
    const int64_t row = i / Wout;
    const int xmin = __ldg(bounds + 2 * xx), cnt = __ldg(bounds + 2 * xx + 1);
    const uint8_t* src = in + row * Win + xmin;
    const int* k = kk + (int64_t)xx * ksize;
    int ss = 1 << (kPrecisionBits - 1);
    for (int x = 0; x < cnt; ++x) ss += (int)__ldg(src + x) * __ldg(k + x);
    out[i] = clip8(ss);
  }
}

__global__ void k_resample_v_u8(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int N, int Hin, int Hout, int W,
                                const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
  const int64_t total = (int64_t)N * Hout * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int yy = (int)((i / W) % Hout);
    const int n = (int)(i / ((int64_t)W * Hout));
    const int ymin = __ldg(bounds + 2 * yy), cnt = __ldg(bounds + 2 * yy + 1);
    const uint8_t* src = in + ((int64_t)n * Hin + ymin) * W + x;
    const int* k = kk + (int64_t)yy * ksize;
    int ss = 1 << (kPrecisionBits - 1);
    for (int y = 0; y < cnt; ++y) ss += (int)__ldg(src + (int64_t)y * W) * __ldg(k + y);
    out[i] = clip8(ss);
  }
}

// ToTensor (torchvision): uint8 -> float32 / 255
__global__ void k_u8_to_unit(const uint8_t* __restrict__ in, float* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __fdiv_rn((float)in[i], 255.0f);
}
// RUN:115,145: clamp(0,1) then (v * 255).astype('uint8'): float32 product, truncation toward zero
__global__ void k_unit_to_u8(const float* __restrict__ in, uint8_t* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = in[i];
    v = fminf(fmaxf(v, 0.0f), 1.0f);              // torch.clamp: NaN stays NaN; (NaN * 255).astype(uint8) is undefined upstream, 0 here
    const float s = __fmul_rn(v, 255.0f);
    out[i] = (uint8_t)(int)s;
  }
}

static int ew_grid(int64_t n) { return (int)std::min<int64_t>((n + 255) / 256, 148 * 32); }

// (N, Hin, Win) uint8 -> (N, Hout, Wout) uint8; tmp must hold N * Hin * Wout bytes (unused when Win == Wout)
void resize_bicubic_u8(Ctx& c, const uint8_t* src, uint8_t* dst, uint8_t* tmp, int N, int Hin, int Win, int Hout, int Wout) {
  XRD_REQUIRE(N >= 1 && Hin >= 1 && Win >= 1 && Hout >= 1 && Wout >= 1, "resize: empty image");
  const bool need_h = Win != Wout, need_v = Hin != Hout;
  if (!need_h && !need_v) {
    XRD_CUDA(cudaMemcpyAsync(dst, src, (size_t)N * Hin * Win, cudaMemcpyDeviceToDevice, c.s));
    return;
  }
  const uint8_t* cur = src;
  if (need_h) {
    const Table t = table_for(Win, Wout);
    uint8_t* o = need_v ? tmp : dst;
    XRD_REQUIRE(o != nullptr, "resize: temporary buffer required");
    XRD_LAUNCH(c, k_resample_h_u8, ew_grid((int64_t)N * Hin * Wout), 256, 0, cur, o, N, Hin, Win, Wout, t.bounds, t.kk, t.ksize);
    cur = o;
  }
  if (need_v) {
    const Table t = table_for(Hin, Hout);
    XRD_LAUNCH(c, k_resample_v_u8, ew_grid((int64_t)N * Hout * Wout), 256, 0, cur, dst, N, Hin, Hout, Wout, t.bounds, t.kk, t.ksize);
  }
}

// host only: the fixed-point table Pillow would build for this size pair (tests pin it against Pillow on the CPU)
int resample_table_host(int in_size, int out_size, int* bounds, int* kk, int cap_k) {
  std::vector<int> b, k;
  int ksize = 0;
  build_table(in_size, out_size, b, k, ksize);
  if (bounds) std::copy(b.begin(), b.end(), bounds);
  if (kk) {
    XRD_REQUIRE((size_t)cap_k >= k.size(), "resample table buffer too small (%d < %zu)", cap_k, k.size());
    std::copy(k.begin(), k.end(), kk);
  }
  return ksize;
}

void u8_to_unit(Ctx& c, const uint8_t* in, float* out, int64_t n) { XRD_LAUNCH(c, k_u8_to_unit, ew_grid(n), 256, 0, in, out, n); }
void unit_to_u8(Ctx& c, const float* in, uint8_t* out, int64_t n) { XRD_LAUNCH(c, k_unit_to_u8, ew_grid(n), 256, 0, in, out, n); }

}  // namespace xrd
