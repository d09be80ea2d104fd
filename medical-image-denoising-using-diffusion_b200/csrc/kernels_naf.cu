// kernels_naf.cu -- bandwidth-oriented versions of the NAFBlock's HBM-bound passes for 16-bit NHWC storage
// (LayerNorm2d HYB:108-115, depthwise 3x3 + SimpleGate + pool HYB:155-157).  Same arithmetic as the generic kernels in
// kernels_simt.cu (fp32 math, two-pass variance, value-as-stored pooling); what changes is the memory access: 16-byte
// loads/stores and several independent pixels in flight per thread, because the generic versions (8-byte accesses, one
// pixel in flight) were latency-bound at 10-48 % of the HBM copy rate (profiles/README.md section 6).
#include "kernels.cuh"
#include "tc_common.cuh"

namespace xrd {

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm over the channels of each pixel.  lpp lanes share a pixel, each lane owns NV 16-byte chunks (8 channels each):
// chunk j of lane `sub` covers channels (j*lpp + sub)*8..+8.  U pixels per thread are loaded before any is reduced.
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int NV, int U>
__global__ void __launch_bounds__(256) k_layernorm16(const T* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                                                     float eps, T* __restrict__ y, int64_t npix, int C, int lpp) {
  const int lane = threadIdx.x & 31;
  const int sub = lane % lpp, psub = lane / lpp;
  const int ppw = 32 / lpp;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float invC = 1.0f / (float)C;
  float gg[NV][8], bb[NV][8];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c0 = (j * lpp + sub) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) { gg[j][i] = __ldg(g + c0 + i); bb[j][i] = __ldg(b + c0 + i); }
  }
  for (int64_t p0 = warp * (ppw * U); p0 < npix; p0 += nwarps * (ppw * U)) {
    uint4 q[U][NV];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t pix = p0 + u * ppw + psub;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        if (pix < npix) q[u][j] = __ldg(reinterpret_cast<const uint4*>(x + pix * C + (j * lpp + sub) * 8));
        else q[u][j] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t pix = p0 + u * ppw + psub;
      float v[NV][8];
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        tc::unpack8<T>(q[u][j], v[j]);
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[j][i];
      }
      for (int o = lpp >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s * invC;
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[j][i] - mean; ss = fmaf(d, d, ss); }
      for (int o = lpp >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float rstd = 1.0f / sqrtf(ss * invC + eps);
      if (pix < npix) {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          float o8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o8[i] = (v[j][i] - mean) * rstd * gg[j][i] + bb[j][i];
          uint4 pk;
          pk.x = tc::pack2<T>(o8[0], o8[1]); pk.y = tc::pack2<T>(o8[2], o8[3]);
          pk.z = tc::pack2<T>(o8[4], o8[5]); pk.w = tc::pack2<T>(o8[6], o8[7]);
          *reinterpret_cast<uint4*>(y + pix * C + (j * lpp + sub) * 8) = pk;
        }
      }
    }
  }
}

bool layernorm16_supported(const Tens& x, const Tens& y) {
  static const int enabled = getenv("XRD_LN16") ? atoi(getenv("XRD_LN16")) : 1;
  if (!enabled || x.dt == DT_F32 || y.dt != x.dt) return false;
  const int chunks = x.c / 8;
  if (x.c % 8 != 0 || chunks < 1) return false;
  if (chunks <= 32) return (chunks & (chunks - 1)) == 0;          // 8..256 channels: one chunk per lane, lpp = chunks
  return chunks == 64;                                             // 512 channels: two chunks per lane
}

void layernorm16(Ctx& c, const Tens& x, const float* g, const float* b, float eps, Tens& y) {
  XRD_REQUIRE(layernorm16_supported(x, y) && y.numel() == x.numel() && y.c == x.c, "layernorm16: unsupported configuration");
  const int chunks = x.c / 8;
  const int nv = chunks > 32 ? 2 : 1;
  const int lpp = chunks / nv;
  const int64_t npix = (int64_t)x.n * x.h * x.w;
  const int U = nv == 2 ? 2 : 4;
  const int64_t per_block = (int64_t)(32 / lpp) * U * 8;
  const int bx = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv64(npix, per_block), 148 * 8));
  if (x.dt == DT_F16) {
    using T = __half;
    if (nv == 2) XRD_LAUNCH(c, (k_layernorm16<T, 2, 2>), bx, 256, 0, (const T*)x.p, g, b, eps, (T*)y.p, npix, x.c, lpp);
    else XRD_LAUNCH(c, (k_layernorm16<T, 1, 4>), bx, 256, 0, (const T*)x.p, g, b, eps, (T*)y.p, npix, x.c, lpp);
  } else {
    using T = __nv_bfloat16;
    if (nv == 2) XRD_LAUNCH(c, (k_layernorm16<T, 2, 2>), bx, 256, 0, (const T*)x.p, g, b, eps, (T*)y.p, npix, x.c, lpp);
    else XRD_LAUNCH(c, (k_layernorm16<T, 1, 4>), bx, 256, 0, (const T*)x.p, g, b, eps, (T*)y.p, npix, x.c, lpp);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// depthwise 3x3 (2C channels) + SimpleGate + global-average-pool partial sums, C in {32, 64, 128}.
//   thread = (pixel column x, 8 consecutive channels of the 2C-channel tensor); it walks down a strip of rows.
//   Each input row (pixels x-1, x, x+1: three 16-byte chunks) is unpacked once and feeds the three output rows it
//   touches (ky = 2, 1, 0) held as running accumulators, so every input element is converted once.  The 72 weights of the
//   thread's 8 channels live in registers.  Lanes l and l ^ nq hold channel c and channel C + c of the same pixel: one
//   shuffle exchange forms the gate product; the lower lane stores 16 bytes.
//   Round 1 read the rows straight from global memory with one row prefetched per thread in registers (72 weights + 24
//   accumulators leave room for no more): ~8 KB of distinct bytes in flight per SM, and the 1.36 TB/s (21 % of the copy
//   rate) it measured is what Little's law gives for that at ~800 ns of HBM latency.  Here (round 2) one thread of the block
//   keeps a ring of S row segments ((PX + 2) pixels x 2C channels, 4.3-5.1 KB each, one cp.async.bulk request per row) in
//   flight on mbarriers, independent of anyone's registers: 2 blocks x 4 slots = 35-40 KB in flight per SM.  The compute reads
//   its three 16-byte neighbours from the ring (a warp reads 512 contiguous bytes: conflict-free).  Halo pixels outside the
//   image are zeroed once per slot and never overwritten; rows outside the image are never fetched (their barrier phase is
//   completed by a plain arrive).  432 us at C = 32 @512x512, batch 16 (1.9 TB/s; 592 us before).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kDwSlots = 4;

template <typename T>
__global__ void __launch_bounds__(256, 2) k_dwconv_gate_pool16s(const T* __restrict__ u, const float* __restrict__ w9,
                                                                const float* __restrict__ bias, T* __restrict__ g, float* __restrict__ pool,
                                                                int H, int W, int C, int rows_per_block) {
  extern __shared__ __align__(128) uint8_t s_raw[];
  const int nq = C >> 3;                       // 8-channel chunks per half: 4, 8 or 16
  const int ppw = 32 / (2 * nq);               // pixels per warp
  const int PX = 8 * ppw;                      // pixel columns per block
  const int C2 = 2 * C;
  const uint32_t pix_bytes = (uint32_t)C2 * sizeof(T);
  const uint32_t slot_bytes = (uint32_t)(PX + 2) * pix_bytes;
  uint8_t* ring = s_raw;                                                      // [kDwSlots][(PX + 2) * 2C] T
  uint64_t* full = reinterpret_cast<uint64_t*>(s_raw + kDwSlots * slot_bytes);
  float* s_pool = reinterpret_cast<float*>(full + kDwSlots);                  // [C]

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane & (nq - 1);
  const int half = (lane / nq) & 1;
  const int px = warp * ppw + lane / (2 * nq); // pixel column within the block
  const int x0 = blockIdx.x * PX;
  const int x = x0 + px;
  const int n = blockIdx.z;
  const int y0 = blockIdx.y * rows_per_block;
  const int y1 = min(y0 + rows_per_block, H);
  const int n_in = y1 - y0 + 2;                // input rows y0-1 .. y1
  const int cb = half * C + q * 8;             // first of this thread's 8 channels in the 2C-channel tensor

  // the part of a row segment that exists in the image: pixels [xs, xe) land at slot pixel xs - (x0 - 1)
  const int xs = max(x0 - 1, 0), xe = min(x0 + PX + 1, W);
  const uint32_t dst_off = (uint32_t)(xs - (x0 - 1)) * pix_bytes;
  const uint32_t row_bytes = (uint32_t)(xe - xs) * pix_bytes;
  const T* row0 = u + ((int64_t)n * H * W + xs) * C2;                          // + r * W * C2 per image row

  for (int i = threadIdx.x; i < C; i += blockDim.x) s_pool[i] = 0.f;
  // zero the slot pixels no copy will ever write (image borders, ragged right edge)
  for (uint32_t i = threadIdx.x * 16u; i < kDwSlots * slot_bytes; i += blockDim.x * 16u) {
    const uint32_t o = i % slot_bytes;
    if (o < dst_off || o >= dst_off + row_bytes) *reinterpret_cast<uint4*>(ring + i) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kDwSlots; ++s) tc::mbar_init(&full[s], 1);
    tc::fence_barrier_init();
  }
  __syncthreads();

  auto issue = [&](int i) {                    // thread 0: input row i of this block into slot i % kDwSlots
    const int r = y0 - 1 + i;
    uint64_t* bar = &full[i % kDwSlots];
    if (r >= 0 && r < H) {
      tc::mbar_expect_tx(bar, row_bytes);
      tc::bulk_load_1d(ring + (size_t)(i % kDwSlots) * slot_bytes + dst_off, row0 + (int64_t)r * W * C2, row_bytes, bar);
    } else {
      tc::mbar_arrive(bar);                    // nothing to fetch: the row is zero padding
    }
  };
  if (threadIdx.x == 0)
    for (int i = 0; i < kDwSlots && i < n_in; ++i) issue(i);

  float w[9][8], bs[8];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(w9 + t * C2 + cb)), b = __ldg(reinterpret_cast<const float4*>(w9 + t * C2 + cb + 4));
    w[t][0] = a.x; w[t][1] = a.y; w[t][2] = a.z; w[t][3] = a.w; w[t][4] = b.x; w[t][5] = b.y; w[t][6] = b.z; w[t][7] = b.w;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) bs[i] = __ldg(bias + cb + i);

  const bool xin = x < W;
  float a0[8], a1[8], a2[8], ps[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a0[i] = bs[i]; a1[i] = bs[i]; a2[i] = bs[i]; ps[i] = 0.f; }
  const uint32_t my_off = (uint32_t)px * pix_bytes + (uint32_t)cb * sizeof(T);   // pixel x-1 of this thread's chunk inside a slot
  for (int i = 0; i < n_in; ++i) {
    const int r = y0 - 1 + i;
    const int slot = i % kDwSlots;
    tc::mbar_wait_relaxed(&full[slot], (uint32_t)(i / kDwSlots) & 1u, 70);
    if (r >= 0 && r < H) {                      // block-uniform
      const uint8_t* sp = ring + (size_t)slot * slot_bytes + my_off;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const uint4 raw = *reinterpret_cast<const uint4*>(sp + (uint32_t)k * pix_bytes);
        float v[8];
        tc::unpack8<T>(raw, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a0[j] = fmaf(v[j], w[6 + k][j], a0[j]);   // output row r-1 sees this row through ky = 2
          a1[j] = fmaf(v[j], w[3 + k][j], a1[j]);   // output row r   through ky = 1
          a2[j] = fmaf(v[j], w[k][j], a2[j]);       // output row r+1 through ky = 0
        }
      }
    }
    // output row r-1 is complete
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = a0[j] * __shfl_xor_sync(0xffffffffu, a0[j], nq);
    if (r - 1 >= y0 && xin && half == 0) {
      uint4 pk;
      pk.x = tc::pack2<T>(o[0], o[1]); pk.y = tc::pack2<T>(o[2], o[3]); pk.z = tc::pack2<T>(o[4], o[5]); pk.w = tc::pack2<T>(o[6], o[7]);
      *reinterpret_cast<uint4*>(g + (((int64_t)n * H + (r - 1)) * W + x) * C + q * 8) = pk;
      float st[8];
      tc::unpack8<T>(pk, st);                    // the pool must see what the next layer sees: the value as stored
#pragma unroll
      for (int j = 0; j < 8; ++j) ps[j] += st[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { a0[j] = a1[j]; a1[j] = a2[j]; a2[j] = bs[j]; }
    __syncthreads();                             // every thread has read slot `slot`: it may be refilled
    if (threadIdx.x == 0 && i + kDwSlots < n_in) issue(i + kDwSlots);
  }
  if (half == 0 && xin) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&s_pool[q * 8 + j], ps[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&pool[(int64_t)n * C + i], s_pool[i]);
}

// ---------------------------------------------------------------------------------------------------------------
// Variant for C = 256 / 512 (round 2): the staged kernel above synchronises the whole block once per row (the thread that refills
// the ring is one of the computing threads), which left it at 1.9 TB/s.  Here the ring has a warp of its own -- a ninth warp
// whose lane 0 waits on per-slot `empty` barriers and issues the bulk copies -- so the eight computing warps only ever wait
// for data.  The thread mapping changes so that the gate needs no shuffle and any C in {32 .. 512} fits: a thread owns four
// channels c..c+3 of the first half AND the four partner channels C+c..C+c+3 of one pixel (two 8-byte shared loads per
// neighbour), multiplies them in registers and stores 8 bytes; a pixel is C/4 consecutive threads, so warp-wide accesses
// stay contiguous.  C = 256 and 512 (the NAFNet levels the 16-byte kernels did not cover: 0.7 TB/s in the generic kernel)
// run here too.  Arithmetic and summation order per output are those of the kernels above.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kDw3Slots = 4;
constexpr int kDw3Threads = 288;               // 8 computing warps + the ring warp

template <typename T>
__global__ void __maxnreg__(112) k_dwconv_gate_pool16v3(const T* __restrict__ u, const float* __restrict__ w9,
                                                                         const float* __restrict__ bias, T* __restrict__ g,
                                                                         float* __restrict__ pool, int H, int W, int C, int rows_per_block) {
  extern __shared__ __align__(128) uint8_t s_raw[];
  const int tpp = C >> 2;                      // threads per pixel: 8 .. 128
  const int PX = 256 / tpp;                    // pixel columns per block: 32 .. 2
  const int C2 = 2 * C;
  const uint32_t pix_bytes = (uint32_t)C2 * sizeof(T);
  const uint32_t slot_bytes = (uint32_t)(PX + 2) * pix_bytes;
  uint8_t* ring = s_raw;                                                      // [kDw3Slots][(PX + 2) * 2C] T
  uint64_t* full = reinterpret_cast<uint64_t*>(s_raw + kDw3Slots * slot_bytes);
  uint64_t* empty = full + kDw3Slots;
  float* s_pool = reinterpret_cast<float*>(empty + kDw3Slots);                // [C]
  float* s_bias = s_pool + C;                                                 // [2C]: the accumulators restart from it every row

  const int warp = threadIdx.x >> 5;
  const int x0 = blockIdx.x * PX;
  const int n = blockIdx.z;
  const int y0 = blockIdx.y * rows_per_block;
  const int y1 = min(y0 + rows_per_block, H);
  const int n_in = y1 - y0 + 2;                // input rows y0-1 .. y1

  // the part of a row segment that exists in the image: pixels [xs, xe) land at slot pixel xs - (x0 - 1)
  const int xs = max(x0 - 1, 0), xe = min(x0 + PX + 1, W);
  const uint32_t dst_off = (uint32_t)(xs - (x0 - 1)) * pix_bytes;
  const uint32_t row_bytes = (uint32_t)(xe - xs) * pix_bytes;

  for (int i = threadIdx.x; i < C; i += blockDim.x) s_pool[i] = 0.f;
  for (int i = threadIdx.x; i < C2; i += blockDim.x) s_bias[i] = __ldg(bias + i);
  // zero the slot pixels no copy will ever write (image borders, ragged right edge)
  for (uint32_t i = threadIdx.x * 16u; i < kDw3Slots * slot_bytes; i += blockDim.x * 16u) {
    const uint32_t o = i % slot_bytes;
    if (o < dst_off || o >= dst_off + row_bytes) *reinterpret_cast<uint4*>(ring + i) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kDw3Slots; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 256); }
    tc::fence_barrier_init();
  }
  __syncthreads();

  if (warp == 8) {
    // ===================== ring warp =====================
    if ((threadIdx.x & 31) == 0) {
      const T* row0 = u + ((int64_t)n * H * W + xs) * C2;                       // + r * W * C2 per image row
      for (int i = 0; i < n_in; ++i) {
        const int slot = i % kDw3Slots;
        if (i >= kDw3Slots) tc::mbar_wait_relaxed(&empty[slot], (uint32_t)(i / kDw3Slots - 1) & 1u, 71);
        const int r = y0 - 1 + i;
        if (r >= 0 && r < H) {
          tc::mbar_expect_tx(&full[slot], row_bytes);
          tc::bulk_load_1d(ring + (size_t)slot * slot_bytes + dst_off, row0 + (int64_t)r * W * C2, row_bytes, &full[slot]);
        } else {
          tc::mbar_arrive(&full[slot]);        // nothing to fetch: the row is zero padding
        }
      }
    }
  } else {
    // ===================== computing warps =====================
    const int px = threadIdx.x / tpp;            // pixel column within the block
    const int q4 = (threadIdx.x - px * tpp) * 4; // first of this thread's four channels in each half
    const int x = x0 + px;
    const bool xin = x < W;
    float w[9][8];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(w9 + t * C2 + q4)), b = __ldg(reinterpret_cast<const float4*>(w9 + t * C2 + C + q4));
      w[t][0] = a.x; w[t][1] = a.y; w[t][2] = a.z; w[t][3] = a.w; w[t][4] = b.x; w[t][5] = b.y; w[t][6] = b.z; w[t][7] = b.w;
    }
    float a0[8], a1[8], a2[8], ps[4];
    {
      const float4 a = *reinterpret_cast<const float4*>(s_bias + q4), b = *reinterpret_cast<const float4*>(s_bias + C + q4);
      a0[0] = a1[0] = a2[0] = a.x; a0[1] = a1[1] = a2[1] = a.y; a0[2] = a1[2] = a2[2] = a.z; a0[3] = a1[3] = a2[3] = a.w;
      a0[4] = a1[4] = a2[4] = b.x; a0[5] = a1[5] = a2[5] = b.y; a0[6] = a1[6] = a2[6] = b.z; a0[7] = a1[7] = a2[7] = b.w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) ps[i] = 0.f;
    const uint32_t off_lo = (uint32_t)px * pix_bytes + (uint32_t)q4 * sizeof(T);   // pixel x-1, first half
    const uint32_t off_hi = off_lo + (uint32_t)C * sizeof(T);                      // same pixel, partner channels
    for (int i = 0; i < n_in; ++i) {
      const int r = y0 - 1 + i;
      const int slot = i % kDw3Slots;
      tc::mbar_wait(&full[slot], (uint32_t)(i / kDw3Slots) & 1u, 72);
      if (r >= 0 && r < H) {                      // block-uniform
        const uint8_t* sp = ring + (size_t)slot * slot_bytes;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const uint2 lo = *reinterpret_cast<const uint2*>(sp + off_lo + (uint32_t)k * pix_bytes);
          const uint2 hi = *reinterpret_cast<const uint2*>(sp + off_hi + (uint32_t)k * pix_bytes);
          float v[8];
          { const float2 f0 = tc::unpack2<T>(lo.x), f1 = tc::unpack2<T>(lo.y), f2 = tc::unpack2<T>(hi.x), f3 = tc::unpack2<T>(hi.y);
            v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y; v[4] = f2.x; v[5] = f2.y; v[6] = f3.x; v[7] = f3.y; }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            a0[j] = fmaf(v[j], w[6 + k][j], a0[j]);   // output row r-1 sees this row through ky = 2
            a1[j] = fmaf(v[j], w[3 + k][j], a1[j]);   // output row r   through ky = 1
            a2[j] = fmaf(v[j], w[k][j], a2[j]);       // output row r+1 through ky = 0
          }
        }
      }
      tc::mbar_arrive(&empty[slot]);              // release semantics: this thread's reads of the slot are ordered before it
      // output row r-1 is complete: SimpleGate in registers
      if (r - 1 >= y0 && xin) {
        uint2 pk;
        pk.x = tc::pack2<T>(a0[0] * a0[4], a0[1] * a0[5]); pk.y = tc::pack2<T>(a0[2] * a0[6], a0[3] * a0[7]);
        *reinterpret_cast<uint2*>(g + (((int64_t)n * H + (r - 1)) * W + x) * C + q4) = pk;
        const float2 s0 = tc::unpack2<T>(pk.x), s1 = tc::unpack2<T>(pk.y);   // the pool must see the value as stored
        ps[0] += s0.x; ps[1] += s0.y; ps[2] += s1.x; ps[3] += s1.y;
      }
      {
        const float4 a = *reinterpret_cast<const float4*>(s_bias + q4), b = *reinterpret_cast<const float4*>(s_bias + C + q4);
#pragma unroll
        for (int j = 0; j < 8; ++j) { a0[j] = a1[j]; a1[j] = a2[j]; }
        a2[0] = a.x; a2[1] = a.y; a2[2] = a.z; a2[3] = a.w; a2[4] = b.x; a2[5] = b.y; a2[6] = b.z; a2[7] = b.w;
      }
    }
    if (xin) {
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(&s_pool[q4 + j], ps[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&pool[(int64_t)n * C + i], s_pool[i]);
}

bool dwconv_gate_pool16_supported(const Tens& u, const Tens& g) {
  static const int enabled = getenv("XRD_DW16") ? atoi(getenv("XRD_DW16")) : 1;
  if (!enabled || u.dt == DT_F32 || g.dt != u.dt) return false;
  static const int v3 = getenv("XRD_DW16_V3") ? atoi(getenv("XRD_DW16_V3")) : 1;
  if (v3 && (g.c == 256 || g.c == 512)) return true;
  return g.c == 32 || g.c == 64 || g.c == 128;
}

void dwconv_gate_pool16(Ctx& c, const Tens& u, const float* w9, const float* bias, Tens& g, float* pool) {
  const int C = g.c;
  XRD_REQUIRE(dwconv_gate_pool16_supported(u, g) && u.c == 2 * C && g.n == u.n && g.h == u.h && g.w == u.w, "dwconv_gate_pool16: shape");
  // measured per launch at batch 16 (gpurun_out r4d): C = 256 @64x64 91 us here against 150 us in the generic kernel, C = 512 @32x32
  // 54 against 69; at C <= 128 the 16-byte staged kernel wins (432 against 674 us at C = 32 @512x512): 8-byte accesses and
  // the ninth warp's register squeeze (112 per thread, 28 spilled) cost more than the per-row block barrier
  static const int v3 = getenv("XRD_DW16_V3") ? atoi(getenv("XRD_DW16_V3")) : 1;
  if ((v3 == 1 && C >= 256) || v3 == 2) {
    const int px = 256 / (C / 4);
    const int rows3 = u.h >= 256 ? 32 : 16;
    dim3 grid3(cdiv(u.w, px), cdiv(u.h, rows3), u.n);
    const size_t smem = (size_t)kDw3Slots * (px + 2) * 2 * C * dsize(u.dt) + 2 * kDw3Slots * sizeof(uint64_t) + 3 * C * sizeof(float);
    if (u.dt == DT_F16) XRD_LAUNCH(c, (k_dwconv_gate_pool16v3<__half>), grid3, kDw3Threads, smem, (const __half*)u.p, w9, bias, (__half*)g.p, pool, u.h, u.w, C, rows3);
    else XRD_LAUNCH(c, (k_dwconv_gate_pool16v3<__nv_bfloat16>), grid3, kDw3Threads, smem, (const __nv_bfloat16*)u.p, w9, bias, (__nv_bfloat16*)g.p, pool, u.h, u.w, C, rows3);
    return;
  }
  const int ppb = 8 * (32 / (2 * (C / 8)));          // pixel columns per block
  const int rows = u.h >= 256 ? 32 : 16;
  dim3 grid(cdiv(u.w, ppb), cdiv(u.h, rows), u.n);
  const size_t smem = (size_t)kDwSlots * (ppb + 2) * 2 * C * dsize(u.dt) + kDwSlots * sizeof(uint64_t) + C * sizeof(float);
  if (u.dt == DT_F16) XRD_LAUNCH(c, (k_dwconv_gate_pool16s<__half>), grid, 256, smem, (const __half*)u.p, w9, bias, (__half*)g.p, pool, u.h, u.w, C, rows);
  else XRD_LAUNCH(c, (k_dwconv_gate_pool16s<__nv_bfloat16>), grid, 256, smem, (const __nv_bfloat16*)u.p, w9, bias, (__nv_bfloat16*)g.p, pool, u.h, u.w, C, rows);
}

}  // namespace xrd
