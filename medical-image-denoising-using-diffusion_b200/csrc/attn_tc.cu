// attn_tc.cu -- flash-style self-attention on tcgen05 (AttentionBlock, HYB:292-305).
//
//   per CTA: 128 queries of one (image, head); loop over 128-key tiles j:
//     S_j = Q K_j^T        tcgen05.mma, A = Q smem (K-major), B = K smem (K-major), D = TMEM S[j&1] (double buffered)
//     P_j = exp2((S_j - m) * scale * log2e)   softmax warps: ONE tcgen05.ld pass -> registers -> f16/bf16 -> smem P[j&1]
//     O  += P_j V_j        tcgen05.mma, A = P smem, B = V smem (MN-major: keys are the K dimension), D = TMEM O, accumulating
//   The running output never leaves TMEM: the reference maximum m is only raised when a row's new maximum exceeds
//   it by more than 2^8 (P stays below 256, exact after the final division by the row sum accumulated with the same
//   m); only then is O rescaled in place (tcgen05.ld / tcgen05.st), which after the first tiles almost never happens.
//   So the per-tile critical path is the softmax warpgroup alone; Q K^T of tile j+1 and P V of tile j-1 run under it.
//   warp 0 = TMA producer (Q once, K/V double buffered), warp 1 = MMA issuer + TMEM owner,
//   warps 2..5 = softmax / correction / epilogue.  The (n, N, N) score matrix never exists in HBM.
#include "kernels.cuh"
#include "tc_common.cuh"

namespace xrd {

struct AttnTcP {
  int HW, heads, d;
  int nkv;                 // number of 128-key tiles
  int nchunk;              // 64-channel chunks of the head dim (1 or 2)
  float scale_log2e;       // d^-0.5 * log2(e)
  uint32_t idesc_qk, idesc_pv;
  void* out;
};

constexpr int kAttThreads = 192;
constexpr int kTile = 128 * 128;       // one [128 rows x 64 x 16-bit] swizzled tile = 16 KB

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// MN-major B operand (V): 64-element d blocks `lbo` bytes apart, 8-key groups 1024 B apart, 128B swizzle
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t dsc = 0;
  dsc |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  dsc |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  dsc |= (uint64_t)(1024 >> 4) << 32;
  dsc |= (uint64_t)1 << 46;
  dsc |= (uint64_t)2 << 61;
  return dsc;
}

template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <typename T> __device__ __forceinline__ float round16(float a);
template <> __device__ __forceinline__ float round16<__nv_bfloat16>(float a) { return __bfloat162float(__float2bfloat16_rn(a)); }
template <> __device__ __forceinline__ float round16<__half>(float a) { return __half2float(__float2half_rn(a)); }

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
        "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr float kRescaleLog2 = 8.0f;   // raise the reference maximum only when P would exceed 2^8

template <typename T, int DC>   // DC = ceil(d / 32): 32-column chunks of the output row
__global__ void __launch_bounds__(kAttThreads, 1) k_attn_tc(const __grid_constant__ CUtensorMap tmQKV, const AttnTcP p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // layout: Q [2 tiles] | KV stage 0: K [2 tiles] V [2 tiles] | KV stage 1 | P buffer 0 [2 tiles] | P buffer 1 | barriers
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + 2 * kTile;
  uint8_t* sP = sKV + 2 * 4 * kTile;
  uint64_t* bars = (uint64_t*)(sP + 4 * kTile);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;     // [2]
  uint64_t* s_free = bars + 7;     // [2]
  uint64_t* p_full = bars + 9;     // [2]
  uint64_t* p_free = bars + 11;    // [2]
  uint64_t* o_done = bars + 13;
  uint32_t* tmem_slot = (uint32_t*)(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, head = blockIdx.y, img = blockIdx.z;
  const int qoff = head * p.d, koff = p.heads * p.d + head * p.d, voff = 2 * p.heads * p.d + head * p.d;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmQKV);
    tc::mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&kv_full[s], 1); tc::mbar_init(&kv_empty[s], 1);
      tc::mbar_init(&s_full[s], 1); tc::mbar_init(&s_free[s], 128);
      tc::mbar_init(&p_full[s], 128); tc::mbar_init(&p_free[s], 1);
    }
    tc::mbar_init(o_done, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, 512);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 256;      // S buffers at +0 and +128

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (tc::elect_one()) {
      tc::mbar_expect_tx(q_full, (uint32_t)p.nchunk * kTile);
      for (int c = 0; c < p.nchunk; ++c) tc::tma_load_3d(sQ + c * kTile, &tmQKV, q_full, qoff + 64 * c, q0, img);
      for (int j = 0; j < p.nkv; ++j) {
        const int st = j & 1;
        tc::mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        uint8_t* sK = sKV + st * 4 * kTile;
        uint8_t* sV = sK + 2 * kTile;
        tc::mbar_expect_tx(&kv_full[st], (uint32_t)(2 * p.nchunk) * kTile);
        for (int c = 0; c < p.nchunk; ++c) {
          tc::tma_load_3d(sK + c * kTile, &tmQKV, &kv_full[st], koff + 64 * c, j * 128, img);
          tc::tma_load_3d(sV + c * kTile, &tmQKV, &kv_full[st], voff + 64 * c, j * 128, img);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (one elected lane; the warp stays converged) =====================
    const uint32_t aQ = tc::smem_u32(sQ);
    const int ks0 = min(64, p.d) >> 4, ks1 = p.nchunk > 1 ? (min(64, p.d - 64) >> 4) : 0;
    auto issue_pv = [&](int j) {      // O (+)= P_j V_j
      const int st = j & 1, pb = j & 1;
      tc::mbar_wait(&p_full[pb], (j >> 1) & 1);
      tc::tc_fence_after();
      if (tc::elect_one()) {
        const uint32_t aP = tc::smem_u32(sP + pb * 2 * kTile);
        const uint32_t bV = tc::smem_u32(sKV + st * 4 * kTile + 2 * kTile);
#pragma unroll
        for (int k = 0; k < 8; ++k) {   // 8 x 16 keys
          const uint64_t ad = tc::umma_desc_sw128(aP + (k >> 2) * kTile + (k & 3) * 32);
          const uint64_t bd = umma_desc_mn_sw128(bV + k * 2048, kTile);
          tc::umma_f16(tmem_O, ad, bd, p.idesc_pv, (j | k) ? 1u : 0u);
        }
        tc::umma_commit(&p_free[pb]);
        tc::umma_commit(&kv_empty[st]);
        tc::umma_commit(o_done);
      }
      __syncwarp();
    };
    tc::mbar_wait(q_full, 0);
    for (int j = 0; j < p.nkv; ++j) {
      const int st = j & 1, sb = j & 1;
      tc::mbar_wait(&kv_full[st], (j >> 1) & 1);
      tc::mbar_wait(&s_free[sb], ((j >> 1) & 1) ^ 1);       // softmax has drained S_{j-2} from this TMEM buffer
      tc::tc_fence_after();
      if (tc::elect_one()) {
        const uint32_t bK = tc::smem_u32(sKV + st * 4 * kTile);
        const uint32_t tS = tmem_base + (uint32_t)sb * 128u;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < ks0) tc::umma_f16(tS, tc::umma_desc_sw128(aQ + k * 32), tc::umma_desc_sw128(bK + k * 32), p.idesc_qk, k ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < ks1) tc::umma_f16(tS, tc::umma_desc_sw128(aQ + kTile + k * 32), tc::umma_desc_sw128(bK + kTile + k * 32), p.idesc_qk, 1u);
        tc::umma_commit(&s_full[sb]);
      }
      __syncwarp();
      if (j > 0) issue_pv(j - 1);
    }
    issue_pv(p.nkv - 1);
  } else {
    // ===================== softmax / correction / epilogue (warps 2..5, one query row per thread) =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int q = q0 + row;
    float m_run = -INFINITY, l_run = 0.f;
    const int sw = row & 7;
    for (int j = 0; j < p.nkv; ++j) {
      const int sb = j & 1;
      tc::mbar_wait(&s_full[sb], (j >> 1) & 1);
      tc::tc_fence_after();
      float v[128];
      const uint32_t tS = tmem_base + (uint32_t)sb * 128u + lane_addr;
      tmem_ld32_nowait(tS, v);
      tmem_ld32_nowait(tS + 32, v + 32);
      tmem_ld32_nowait(tS + 64, v + 64);
      tmem_ld32_nowait(tS + 96, v + 96);
      tmem_ld_wait();
      tc::tc_fence_before();
      tc::mbar_arrive(&s_free[sb]);                  // S_j is in registers: Q K^T of tile j+2 may overwrite this buffer
      const int kvalid = min(128, p.HW - j * 128);
      if (kvalid < 128) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= kvalid) v[i] = -INFINITY;
      }
      float mx0 = v[0], mx1 = v[1], mx2 = v[2], mx3 = v[3];
#pragma unroll
      for (int i = 4; i < 128; i += 4) {
        mx0 = fmaxf(mx0, v[i]); mx1 = fmaxf(mx1, v[i + 1]); mx2 = fmaxf(mx2, v[i + 2]); mx3 = fmaxf(mx3, v[i + 3]);
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      // lazy reference maximum: keep m_run unless this tile would push P above 2^kRescaleLog2
      const bool need = (mx - m_run) * p.scale_log2e > kRescaleLog2;      // true on the first tile (m_run = -inf)
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = need ? mx : m_run;
        if (j > 0) {
          const float alpha = need ? ex2((m_run - m_new) * p.scale_log2e) : 1.0f;
          tc::mbar_wait(o_done, (j - 1) & 1);          // P V of tile j-1 has landed in O
          tc::tc_fence_after();
#pragma unroll
          for (int c = 0; c < DC; ++c) {
            float o[32];
            tmem_ld32_nowait(tmem_O + lane_addr + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] *= alpha;
            tmem_st32(tmem_O + lane_addr + c * 32, o);
          }
          tmem_st_wait();
          tc::tc_fence_before();
          l_run *= alpha;
        }
        m_run = m_new;
      }
      const float mb = m_run * p.scale_log2e;
      // P = exp2(S*c - m*c), rounded to the MMA operand format, written K-major / 128B-swizzled into P[j&1]
      tc::mbar_wait(&p_free[sb], ((j >> 1) & 1) ^ 1);   // P V of tile j-2 has consumed this buffer
      uint8_t* prow = sP + sb * 2 * kTile + row * 128;
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float a = ex2(fmaf(v[c * 32 + i], p.scale_log2e, -mb));
          const float b = ex2(fmaf(v[c * 32 + i + 1], p.scale_log2e, -mb));
          rs0 += a; rs1 += b;
          pk[i >> 1] = pack2<T>(a, b);
        }
        // 32 keys = 64 B = 4 x 16 B chunks of this row; keys [64*t, 64*t+64) live in tile t
        uint8_t* trow = prow + (c >> 1) * kTile;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int chunk = (c & 1) * 4 + u;
          *reinterpret_cast<uint4*>(trow + ((chunk ^ sw) << 4)) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
        }
      }
      tc::fence_async_smem();                       // generic-proxy stores -> visible to the tensor core (async proxy)
      tc::mbar_arrive(&p_full[sb]);
      l_run += rs0 + rs1;
    }
    // epilogue: O / l -> 16-bit -> out[img, q, head*d + :]
    tc::mbar_wait(o_done, (p.nkv - 1) & 1);
    tc::tc_fence_after();
    const float inv = 1.0f / l_run;
    T* dst = (T*)p.out + ((int64_t)img * p.HW + q) * (p.heads * p.d) + head * p.d;
#pragma unroll
    for (int c = 0; c < DC; ++c) {
      float o[32];
      tmem_ld32_nowait(tmem_O + lane_addr + c * 32, o);
      tmem_ld_wait();
      if (q < p.HW) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          if (c * 32 + i < p.d) {
            uint4 w;
            w.x = pack2<T>(o[i] * inv, o[i + 1] * inv); w.y = pack2<T>(o[i + 2] * inv, o[i + 3] * inv);
            w.z = pack2<T>(o[i + 4] * inv, o[i + 5] * inv); w.w = pack2<T>(o[i + 6] * inv, o[i + 7] * inv);
            *reinterpret_cast<uint4*>(dst + c * 32 + i) = w;
          }
        }
      }
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

bool attention_tc_supported(const Tens& qkv, int heads) {
  if (qkv.dt == DT_F32) return false;
  if (qkv.c % (3 * heads) != 0) return false;
  const int d = qkv.c / (3 * heads);
  if (d % 16 != 0 || d < 16 || d > 128) return false;
  if ((heads * d) % 8 != 0) return false;
  return true;
}

void attention_tc(Ctx& c, const Tens& qkv, int heads, Tens& out) {
  XRD_REQUIRE(attention_tc_supported(qkv, heads), "attention_tc: unsupported shape (C=%d heads=%d)", qkv.c, heads);
  const int d = qkv.c / (3 * heads);
  XRD_REQUIRE(out.c == heads * d && out.n == qkv.n && out.h == qkv.h && out.w == qkv.w && out.dt == qkv.dt, "attention_tc: output shape");
  if (c.dry) return;
  const int HW = qkv.h * qkv.w;
  AttnTcP p;
  p.HW = HW; p.heads = heads; p.d = d;
  p.nkv = cdiv(HW, 128);
  p.nchunk = cdiv(d, 64);
  p.scale_log2e = (float)((1.0 / sqrt((double)d)) * 1.4426950408889634);
  const int fmt = qkv.dt == DT_BF16 ? 1 : 0;
  p.idesc_qk = tc::umma_idesc(128, 128, fmt);
  p.idesc_pv = tc::umma_idesc(128, d, fmt) | (1u << 16);   // B (V) is MN-major
  p.out = out.p;
  alignas(64) CUtensorMap tm;
  {
    const cuuint64_t dims[3] = {(cuuint64_t)qkv.c, (cuuint64_t)HW, (cuuint64_t)qkv.n};
    const cuuint64_t strides[2] = {(cuuint64_t)qkv.c * 2, (cuuint64_t)HW * qkv.c * 2};
    const cuuint32_t box[3] = {64, 128, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = get_encode_tiled()(&tm, tmap_dtype(qkv.dt), 3, qkv.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail(XRD_ERR_CUDA, "cuTensorMapEncodeTiled(qkv) failed: %d", (int)r);
  }
  const size_t smem = 1024 + (size_t)(2 + 8 + 4) * kTile + 16 * 8;
  dim3 grid(cdiv(HW, 128), heads, qkv.n);
  const int dc = cdiv(d, 32);
#define XRD_ATT_CASE(TT, DCV)                                                                                             \
  case DCV: {                                                                                                             \
    static bool attr = false;                                                                                             \
    if (!attr) { XRD_CUDA(cudaFuncSetAttribute(k_attn_tc<TT, DCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; } \
    XRD_LAUNCH(c, (k_attn_tc<TT, DCV>), grid, kAttThreads, smem, tm, p);                                                  \
  } break;
  if (qkv.dt == DT_BF16) {
    switch (dc) { XRD_ATT_CASE(__nv_bfloat16, 1) XRD_ATT_CASE(__nv_bfloat16, 2) XRD_ATT_CASE(__nv_bfloat16, 3) XRD_ATT_CASE(__nv_bfloat16, 4) }
  } else {
    switch (dc) { XRD_ATT_CASE(__half, 1) XRD_ATT_CASE(__half, 2) XRD_ATT_CASE(__half, 3) XRD_ATT_CASE(__half, 4) }
  }
#undef XRD_ATT_CASE
}

}  // namespace xrd
