// tiles.cu -- overlap-tiled evaluation of images larger than the network's native field (BASELINE configs[4]:
// 1024x1024 through 512x512 tiles with halos).  The reference has no tiling code; the decomposition is defined here
// and restated in oracle/xrd_oracle.py (tile_origins / tile_weights / extract_tiles / blend_tiles), DESIGN.md section 9:
//
//   origins along an axis of length L, tile T, halo h:  stride S = T - 2h,  n = L == T ? 1 : ceil((L - T) / S) + 1,
//                                                       o_k = min(k * S, L - T)
//   weight of local coordinate i in tile k:             min(k == 0 ? R : i + 1,  k == n-1 ? R : T - i,  R) / R,  R = max(1, 2h)
//   blended pixel:                                      sum_t (wy*wx) * v_t  /  sum_t (wy*wx)   over covering tiles, ty then tx ascending
//
// All arithmetic is fp32 with explicit round-to-nearest mul/add/div (no FMA contraction) so the result is bit-identical
// to the numpy restatement.  Both kernels are one 16-byte access per thread: HBM-bound copies.
#include "common.cuh"

namespace xrd {

__host__ __device__ inline int tile_count(int L, int T, int halo) {
  if (L == T) return 1;
  const int S = T - 2 * halo;
  return (L - T + S - 1) / S + 1;
}
__host__ __device__ inline int tile_origin(int k, int L, int T, int halo) {
  const int o = k * (T - 2 * halo);
  return o < L - T ? o : L - T;
}
__device__ __forceinline__ float tile_weight(int i, int k, int n, int T, int R) {
  int m = R;
  if (k != 0 && i + 1 < m) m = i + 1;
  if (k != n - 1 && T - i < m) m = T - i;
  return __fdiv_rn((float)m, (float)R);
}

// tiles[((b*ny + ty)*nx + tx)][T][T] = img[b][oy+..][ox+..]; thread = 4 consecutive pixels of one tile row
__global__ void k_tiles_extract(const float* __restrict__ img, float* __restrict__ tiles, int B, int H, int W, int T, int halo,
                                int ny, int nx) {
  const int64_t total = (int64_t)B * ny * nx * T * (T / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int x4 = (int)(r % (T / 4)); r /= (T / 4);
    const int y = (int)(r % T); r /= T;
    const int tx = (int)(r % nx); r /= nx;
    const int ty = (int)(r % ny);
    const int b = (int)(r / ny);
    const int oy = tile_origin(ty, H, T, halo), ox = tile_origin(tx, W, T, halo);
    const float* src = img + ((int64_t)b * H + oy + y) * W + ox + x4 * 4;
    float4 v;
    if (((ox | W) & 3) == 0) v = __ldg(reinterpret_cast<const float4*>(src));
    else v = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3));
    reinterpret_cast<float4*>(tiles)[i] = v;
  }
}

// thread = one output pixel; gathers the <= (2 or 3)^2 covering tiles in fixed order
__global__ void k_tiles_blend(const float* __restrict__ tiles, float* __restrict__ img, int B, int H, int W, int T, int halo, int ny,
                              int nx) {
  const int R = halo > 0 ? 2 * halo : 1;
  const int64_t total = (int64_t)B * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int b = (int)(i / ((int64_t)W * H));
    float num = 0.f, den = 0.f;
    for (int ty = 0; ty < ny; ++ty) {
      const int ly = y - tile_origin(ty, H, T, halo);
      if (ly < 0 || ly >= T) continue;
      const float wy = tile_weight(ly, ty, ny, T, R);
      for (int tx = 0; tx < nx; ++tx) {
        const int lx = x - tile_origin(tx, W, T, halo);
        if (lx < 0 || lx >= T) continue;
        const float w = __fmul_rn(wy, tile_weight(lx, tx, nx, T, R));
        const float v = __ldg(tiles + (((int64_t)(b * ny + ty) * nx + tx) * T + ly) * T + lx);
        num = __fadd_rn(num, __fmul_rn(w, v));
        den = __fadd_rn(den, w);
      }
    }
    img[i] = __fdiv_rn(num, den);
  }
}

void tiles_check(int B, int H, int W, int T, int halo) {
  XRD_REQUIRE(B >= 1 && T >= 4 && T % 4 == 0, "tiles: tile size must be a positive multiple of 4 (got %d)", T);
  XRD_REQUIRE(halo >= 0 && 2 * halo < T, "tiles: halo %d does not leave a positive stride for tile %d", halo, T);
  XRD_REQUIRE(H >= T && W >= T, "tiles: image %dx%d is smaller than the tile %d", H, W, T);
}

void tiles_extract(Ctx& c, const float* img, float* tiles, int B, int H, int W, int T, int halo) {
  tiles_check(B, H, W, T, halo);
  const int ny = tile_count(H, T, halo), nx = tile_count(W, T, halo);
  const int64_t total = (int64_t)B * ny * nx * T * (T / 4);
  const int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
  XRD_LAUNCH(c, k_tiles_extract, grid, 256, 0, img, tiles, B, H, W, T, halo, ny, nx);
}

void tiles_blend(Ctx& c, const float* tiles, float* img, int B, int H, int W, int T, int halo) {
  tiles_check(B, H, W, T, halo);
  const int ny = tile_count(H, T, halo), nx = tile_count(W, T, halo);
  const int64_t total = (int64_t)B * H * W;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 16);
  XRD_LAUNCH(c, k_tiles_blend, grid, 256, 0, tiles, img, B, H, W, T, halo, ny, nx);
}

int tiles_count_host(int L, int T, int halo) { return tile_count(L, T, halo); }
int tiles_origin_host(int k, int L, int T, int halo) { return tile_origin(k, L, T, halo); }

}  // namespace xrd
