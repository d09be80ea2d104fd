// attn_tc.cu -- flash-style self-attention on tcgen05 (AttentionBlock, HYB:292-305).
//
//   per CTA: 128 queries of one (image, head); loop over 128-key tiles:
//     S  = Q K^T          tcgen05.mma, A = Q smem (K-major), B = K smem (K-major), D = TMEM[0,128)
//     P  = exp2((S - m) * scale * log2e)   softmax warps: tcgen05.ld -> registers -> f16/bf16 -> smem (K-major, 128B swizzle)
//     Ot = P V            tcgen05.mma, A = P smem, B = V smem (MN-major: keys are the K dimension), D = TMEM[128,128+d)
//     O  = O * alpha + Ot  running output kept in registers (one query row per thread), rescaled online
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

template <typename T, int DC>   // DC = ceil(d / 32): 32-column chunks of the output row kept in registers
__global__ void __launch_bounds__(kAttThreads, 1) k_attn_tc(const __grid_constant__ CUtensorMap tmQKV, const AttnTcP p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // layout: Q [2 tiles] | KV stage 0: K [2 tiles] V [2 tiles] | KV stage 1 | P [2 tiles] | barriers
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + 2 * kTile;
  uint8_t* sP = sKV + 2 * 4 * kTile;
  uint64_t* bars = (uint64_t*)(sP + 2 * kTile);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_free = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* o_full = bars + 8;
  uint32_t* tmem_slot = (uint32_t*)(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, head = blockIdx.y, img = blockIdx.z;
  const int qoff = head * p.d, koff = p.heads * p.d + head * p.d, voff = 2 * p.heads * p.d + head * p.d;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmQKV);
    tc::mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&kv_full[s], 1); tc::mbar_init(&kv_empty[s], 1); }
    tc::mbar_init(s_full, 1);
    tc::mbar_init(s_free, 128);
    tc::mbar_init(p_full, 128);
    tc::mbar_init(o_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, 256);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 128;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
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
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      auto issue_pv = [&](int j) {
        const int st = j & 1;
        tc::mbar_wait(p_full, j & 1);
        tc::tc_fence_after();
        const uint32_t aP = tc::smem_u32(sP);
        const uint32_t bV = tc::smem_u32(sKV + st * 4 * kTile + 2 * kTile);
#pragma unroll
        for (int k = 0; k < 8; ++k) {   // 8 x 16 keys
          const uint64_t ad = tc::umma_desc_sw128(aP + (k >> 2) * kTile + (k & 3) * 32);
          const uint64_t bd = umma_desc_mn_sw128(bV + k * 2048, kTile);
          tc::umma_f16(tmem_O, ad, bd, p.idesc_pv, k ? 1u : 0u);
        }
        tc::umma_commit(o_full);
        tc::umma_commit(&kv_empty[st]);
      };
      tc::mbar_wait(q_full, 0);
      for (int j = 0; j < p.nkv; ++j) {
        const int st = j & 1;
        tc::mbar_wait(&kv_full[st], (j >> 1) & 1);
        if (j > 0) tc::mbar_wait(s_free, (j - 1) & 1);      // softmax has drained S_{j-1} from TMEM
        tc::tc_fence_after();
        const uint32_t aQ = tc::smem_u32(sQ);
        const uint32_t bK = tc::smem_u32(sKV + st * 4 * kTile);
        int first = 1;
        for (int c = 0; c < p.nchunk; ++c) {
          const int ks = min(64, p.d - 64 * c) >> 4;
          for (int k = 0; k < ks; ++k) {
            tc::umma_f16(tmem_S, tc::umma_desc_sw128(aQ + c * kTile + k * 32), tc::umma_desc_sw128(bK + c * kTile + k * 32),
                         p.idesc_qk, first ? 0u : 1u);
            first = 0;
          }
        }
        tc::umma_commit(s_full);
        if (j > 0) issue_pv(j - 1);
      }
      issue_pv(p.nkv - 1);
    }
  } else {
    // ===================== softmax / correction / epilogue (warps 2..5, one query row per thread) =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int q = q0 + row;
    float m_run = -INFINITY, l_run = 0.f;
    float o[DC * 32];
#pragma unroll
    for (int i = 0; i < DC * 32; ++i) o[i] = 0.f;
    uint8_t* prow = sP + row * 128;
    const int sw = row & 7;
    for (int j = 0; j < p.nkv; ++j) {
      tc::mbar_wait(s_full, j & 1);
      tc::tc_fence_after();
      const int kbase = j * 128;
      const int kvalid = min(128, p.HW - kbase);
      // pass 1: row max
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v[32];
        tc::tmem_ld32(tmem_S + lane_addr + c * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c * 32 + i < kvalid) mx = fmaxf(mx, v[i]);
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = (m_run == -INFINITY) ? 0.f : ex2((m_run - m_new) * p.scale_log2e);
      const float mb = m_new * p.scale_log2e;
      // pass 2: P = exp2(S*c - m*c), rounded to the MMA operand format, written K-major / 128B-swizzled
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v[32];
        tc::tmem_ld32(tmem_S + lane_addr + c * 32, v);
        if (c == 3) {
          tc::tc_fence_before();
          tc::mbar_arrive(s_free);                 // S_j fully in registers: the next Q K^T may overwrite TMEM
        }
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float a = (c * 32 + i < kvalid) ? ex2(fmaf(v[i], p.scale_log2e, -mb)) : 0.f;
          float b = (c * 32 + i + 1 < kvalid) ? ex2(fmaf(v[i + 1], p.scale_log2e, -mb)) : 0.f;
          a = round16<T>(a); b = round16<T>(b);
          rs += a + b;
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
      tc::mbar_arrive(p_full);
      l_run = l_run * alpha + rs;
      m_run = m_new;
      // correction + accumulate: O = O*alpha + P V
      tc::mbar_wait(o_full, j & 1);
      tc::tc_fence_after();
#pragma unroll
      for (int c = 0; c < DC; ++c) {
        float v[32];
        tc::tmem_ld32(tmem_O + lane_addr + c * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], alpha, v[i]);
      }
      tc::tc_fence_before();
    }
    // epilogue: O / l -> 16-bit -> out[img, q, head*d + :]
    if (q < p.HW) {
      const float inv = 1.0f / l_run;
      T* dst = (T*)p.out + ((int64_t)img * p.HW + q) * (p.heads * p.d) + head * p.d;
#pragma unroll
      for (int i = 0; i < DC * 32; i += 8) {
        if (i < p.d) {
          uint4 w;
          w.x = pack2<T>(o[i] * inv, o[i + 1] * inv); w.y = pack2<T>(o[i + 2] * inv, o[i + 3] * inv);
          w.z = pack2<T>(o[i + 4] * inv, o[i + 5] * inv); w.w = pack2<T>(o[i + 6] * inv, o[i + 7] * inv);
          *reinterpret_cast<uint4*>(dst + i) = w;
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 256);
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
  const size_t smem = 1024 + (size_t)(2 + 8 + 2) * kTile + 16 * 8;
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
