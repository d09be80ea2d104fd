// attn_tc.cu -- flash-style self-attention on tcgen05 (AttentionBlock, HYB:292-305).
//
//   per CTA: 256 queries (two 128-row Q tiles) of one (image, head); loop over 64-key tiles j; for each Q tile t:
//     S_t,j = Q_t K_j^T     tcgen05.mma, A = Q smem (K-major), B = K smem (K-major), D = TMEM S[t][j&1]
//     P_t,j = exp2((S - m) * scale * log2e)   softmax warps: one tcgen05.ld pass -> registers -> f16/bf16 -> smem P[t][j&1]
//     O_t  += P_t,j V_j     tcgen05.mma, A = P smem, B = V smem (MN-major: keys are the K dimension), D = TMEM O[t]
//   * K/V tiles are what every CTA re-streams from L2 (measured: that traffic, not the tensor pipe, bounded the
//     one-Q-tile version); two Q tiles per CTA halve it, and two softmax warpgroups keep the MUFU pipe busy.
//   * The running output never leaves TMEM: the reference maximum m is only raised when a row's new maximum exceeds
//     it by more than 2^8 (P stays below 256, exact after the final division by the row sum accumulated with the
//     same m); only then is O rescaled in place (tcgen05.ld / tcgen05.st) -- after the first tiles almost never.
//   * 64-key tiles keep S in 64 registers per thread and leave room for a 6-deep K/V ring in shared memory.
//   warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..5 / 6..9 = softmax of Q tile 0 / 1.
//   The (n, N, N) score matrix never exists in HBM.
#include "kernels.cuh"
#include "tc_common.cuh"

#include <stdlib.h>
#include <type_traits>

namespace xrd {

struct AttnTcP {
  int HW, heads, d;
  int nkv;                 // number of 64-key tiles
  int nchunk;              // 64-channel chunks of the head dim (1 or 2)
  float scale_log2e;       // d^-0.5 * log2(e)
  int ahead;               // key tiles Q K^T runs ahead of P V: 1 or 2
  int probe;               // probe the next Q K^T block's barriers under the P V MMAs
  void* out;
};

constexpr int kAttThreads = 320;       // warp 0 TMA, warp 1 MMA issuer, warps 2..9 softmax
constexpr int kTile = 128 * 128;       // Q / P tile: [128 rows x 64 x 16-bit], 128B-swizzled = 16 KB
constexpr int kKvTile = 64 * 128;      // K / V chunk tile: [64 keys x 64 ch] = 8 KB
constexpr int kRing = 3;               // K ring slots and V ring slots (2 chunk tiles each)

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

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

template <typename T, int D16>   // head dim d = 16 * D16 (compile time: the MMA issue loop must not carry runtime predicates)
__global__ void __launch_bounds__(kAttThreads, 1)
k_attn_tc(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AttnTcP p) {
  constexpr int D = 16 * D16;
  constexpr int DC = (D + 31) / 32;                          // 32-column chunks of the output row
  constexpr int NCH = (D + 63) / 64;                         // 64-channel chunks of the head dim
  constexpr int KS0 = (D < 64 ? D : 64) / 16, KS1 = D > 64 ? (D - 64) / 16 : 0;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // layout: Q [2 q-tiles][2 chunks] | P [2 q-tiles][2 buffers] | K ring [kRing][2 chunks] | V ring [kRing][2 chunks] | barriers
  uint8_t* sQ = smem;
  uint8_t* sP = sQ + 4 * kTile;
  uint8_t* sK = sP + 4 * kTile;
  uint8_t* sV = sK + kRing * 2 * kKvTile;
  uint64_t* bars = (uint64_t*)(sV + kRing * 2 * kKvTile);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;              // [kRing]
  uint64_t* k_empty = k_full + kRing;       // [kRing]
  uint64_t* v_full = k_empty + kRing;       // [kRing]
  uint64_t* v_empty = v_full + kRing;       // [kRing]
  uint64_t* s_full = v_empty + kRing;       // [2 q-tiles][2]
  uint64_t* s_free = s_full + 4;
  uint64_t* p_full = s_free + 4;
  uint64_t* p_free = p_full + 4;
  uint64_t* o_final = p_free + 4;           // [2] completes once, after the last P V
  uint32_t* tmem_slot = (uint32_t*)(o_final + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256, head = blockIdx.y, img = blockIdx.z;
  const int qoff = head * p.d, koff = p.heads * p.d + head * p.d, voff = 2 * p.heads * p.d + head * p.d;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmQ);
    tc::tma_prefetch_desc(&tmKV);
    tc::mbar_init(q_full, 1);
    for (int s = 0; s < kRing; ++s) {
      tc::mbar_init(&k_full[s], 1); tc::mbar_init(&k_empty[s], 1); tc::mbar_init(&v_full[s], 1); tc::mbar_init(&v_empty[s], 1);
    }
    for (int s = 0; s < 4; ++s) {
      tc::mbar_init(&s_full[s], 1); tc::mbar_init(&s_free[s], 128);
      tc::mbar_init(&p_full[s], 128); tc::mbar_init(&p_free[s], 1);
    }
    tc::mbar_init(&o_final[0], 1); tc::mbar_init(&o_final[1], 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, 512);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  // One CTA per SM (shared memory) and this is its only allocation: the allocator returns column 0.  Treating the base as the
  // constant 0 keeps every tcgen05.mma operand in uniform registers (no R2UR per instruction in the issue loop).
  if (*tmem_slot != 0u) {
    if (threadIdx.x == 0) printf("libxrd: attn_tc expects TMEM base 0, got %u\n", *tmem_slot);
    __trap();
  }
  constexpr uint32_t tmem_base = 0u;
  // TMEM columns: S[t][b] at t*128 + b*64 (64 wide), O[t] at 256 + t*128 (d <= 128 wide)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (tc::elect_one()) {
      tc::mbar_expect_tx(q_full, (uint32_t)(2 * NCH) * kTile);
      for (int t = 0; t < 2; ++t)
        for (int c = 0; c < NCH; ++c) tc::tma_load_3d(sQ + (t * 2 + c) * kTile, &tmQ, q_full, qoff + 64 * c, q0 + t * 128, img);
      uint32_t slot = 0, phase = 0;
      for (int j = 0; j < p.nkv; ++j) {
        tc::mbar_wait(&k_empty[slot], phase ^ 1);
        tc::mbar_expect_tx(&k_full[slot], (uint32_t)NCH * kKvTile);
        for (int c = 0; c < NCH; ++c) tc::tma_load_3d(sK + (slot * 2 + c) * kKvTile, &tmKV, &k_full[slot], koff + 64 * c, j * 64, img);
        tc::mbar_wait(&v_empty[slot], phase ^ 1);
        tc::mbar_expect_tx(&v_full[slot], (uint32_t)NCH * kKvTile);
        for (int c = 0; c < NCH; ++c) tc::tma_load_3d(sV + (slot * 2 + c) * kKvTile, &tmKV, &v_full[slot], voff + 64 * c, j * 64, img);
        if (++slot == kRing) { slot = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (one elected lane; the warp stays converged) =====================
    // per step j: P V of tile j for both Q tiles as soon as its P tiles are published, then Q K^T of tile j+2.
    // The ncu source view of the previous version showed this warp, not the softmax warps, on the critical path: ~135 cycles per
    // MMA, i.e. ~14 instructions (R2UR moves, runtime k-step predicates, per-MMA constant loads) retired at the single-thread rate
    // of one per ~4-8 cycles, while the softmax warps sat in the S-ready wait for 47 % of their time.  Now: everything that
    // varies is a 64-bit descriptor base computed once per block of 8-12 MMAs, every per-MMA offset is a compile-time constant,
    // the head dim and both instruction descriptors are template constants, and one elect block covers both Q tiles.
    constexpr uint32_t IDESC_QK = tc::umma_idesc(128, 64, tc::umma_fmt<T>());
    constexpr uint32_t IDESC_PV = tc::umma_idesc(128, D, tc::umma_fmt<T>()) | (1u << 16);     // B (V) is MN-major
    const uint64_t dQ = tc::umma_desc_sw128(tc::smem_u32(sQ));           // + t * (2 * kTile >> 4)
    const uint64_t dK = tc::umma_desc_sw128(tc::smem_u32(sK));           // + slot * (2 * kKvTile >> 4)
    const uint64_t dP = tc::umma_desc_sw128(tc::smem_u32(sP));           // + (t * 2 + b) * (kTile >> 4)
    const uint64_t dV = umma_desc_mn_sw128(tc::smem_u32(sV), kKvTile);   // + slot * (2 * kKvTile >> 4)
    uint32_t kslot = 0, kphase = 0, vslot = 0, vphase = 0;
    bool qk_probed = false;           // this lane's barrier of the next Q K^T block was already seen complete
    auto issue_pv = [&](int j, int qk_next) {      // O_t (+)= P_t,j V_j for both Q tiles; qk_next: key tile of the Q K^T block that follows (-1: none)
      const int b = j & 1, u = j >> 1;
      {   // the three waits as ONE instruction on three lanes (each wait in front of the MMAs idles the tensor pipe for its latency)
        uint64_t* const bar = lane == 0 ? &v_full[vslot] : &p_full[(lane == 1 ? 0 : 2) + b];
        if (lane < 3) tc::mbar_wait(bar, lane == 0 ? vphase : (uint32_t)(u & 1));
        __syncwarp();
      }
      tc::tc_fence_after();
      if (p.probe && qk_next >= 0 && lane < 3) {   // probe the barriers of the next Q K^T block now: the round trip runs under the MMAs below
        const int bn = qk_next & 1, un = qk_next >> 1;
        uint64_t* const bar = lane == 0 ? &k_full[kslot] : &s_free[(lane == 1 ? 0 : 2) + bn];
        qk_probed = tc::mbar_test(bar, lane == 0 ? kphase : (uint32_t)((un & 1) ^ 1));
      }
      if (tc::elect_one()) {
        const uint64_t bD = dV + (uint64_t)(vslot * (2 * kKvTile >> 4));
        const uint64_t aD0 = dP + (uint64_t)(b * (kTile >> 4));
        const uint32_t acc0 = j ? 1u : 0u;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const uint64_t aD = aD0 + (uint64_t)(t * 2 * (kTile >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k)     // 4 x 16 keys: +32 B in the K-major P tile, +2048 B (16 key rows) in the MN-major V tile
            tc::umma_f16(256u + (uint32_t)t * 128u, aD + (uint64_t)(k * 2), bD + (uint64_t)(k * (2048 >> 4)), IDESC_PV, k ? 1u : acc0);
        }
        tc::umma_commit(&p_free[b]);
        tc::umma_commit(&p_free[2 + b]);
        if (j == p.nkv - 1) { tc::umma_commit(&o_final[0]); tc::umma_commit(&o_final[1]); }
        tc::umma_commit(&v_empty[vslot]);
      }
      __syncwarp();
      if (++vslot == kRing) { vslot = 0; vphase ^= 1; }
    };
    auto issue_qk = [&](int j) {      // S_t,j = Q_t K_j^T for both Q tiles
      const int b = j & 1, u = j >> 1;
      {   // K_j landed; softmax has drained S_t,j-2 from these TMEM buffers -- one wait instruction on three lanes
        uint64_t* const bar = lane == 0 ? &k_full[kslot] : &s_free[(lane == 1 ? 0 : 2) + b];
        if (lane < 3) tc::mbar_wait_probed(qk_probed, bar, lane == 0 ? kphase : (uint32_t)((u & 1) ^ 1));
        qk_probed = false;
        __syncwarp();
      }
      tc::tc_fence_after();
      if (tc::elect_one()) {
        const uint64_t bD = dK + (uint64_t)(kslot * (2 * kKvTile >> 4));
        const uint32_t tS0 = (uint32_t)b * 64u;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const uint64_t aD = dQ + (uint64_t)(t * (2 * kTile >> 4));
          const uint32_t tS = tS0 + (uint32_t)t * 128u;
#pragma unroll
          for (int k = 0; k < KS0; ++k) tc::umma_f16(tS, aD + (uint64_t)(k * 2), bD + (uint64_t)(k * 2), IDESC_QK, k ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < KS1; ++k)
            tc::umma_f16(tS, aD + (uint64_t)((kTile >> 4) + k * 2), bD + (uint64_t)((kKvTile >> 4) + k * 2), IDESC_QK, 1u);
          tc::umma_commit(&s_full[t * 2 + b]);
        }
        tc::umma_commit(&k_empty[kslot]);
      }
      __syncwarp();
      if (++kslot == kRing) { kslot = 0; kphase ^= 1; }
    };
    // Q K^T runs one key tile ahead of P V (S is double buffered).  Two ahead (p.ahead == 2: S_j+2 reuses the TMEM buffer the
    // softmax of tile j drained at its start) measured 4 % slower: the kernel is bound by the MIO queue the MUFU.EX2 stream,
    // the P stores and the UTCHMMA issue share, not by S arriving late.
    tc::mbar_wait(q_full, 0);
    issue_qk(0);
    if (p.ahead == 2) {
      if (p.nkv > 1) issue_qk(1);
      for (int j = 0; j < p.nkv; ++j) {
        issue_pv(j, j + 2 < p.nkv ? j + 2 : -1);
        if (j + 2 < p.nkv) issue_qk(j + 2);
      }
    } else {
      for (int j = 0; j < p.nkv; ++j) {
        if (j + 1 < p.nkv) issue_qk(j + 1);
        issue_pv(j, j + 2 < p.nkv ? j + 2 : -1);
      }
    }
  } else {
    // ===================== softmax / correction / epilogue (warps 2..9, one query row per thread) =====================
    const int t = (warp - 2) >> 2;                 // Q tile
    const int quad = warp & 3;                     // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int q = q0 + t * 128 + row;
    const uint32_t tO = tmem_base + 256u + (uint32_t)t * 128u + lane_addr;
    float m_run = -INFINITY, l_run = 0.f;
    const int sw = row & 7;
    const uint32_t prow0 = tc::smem_u32(sP) + (uint32_t)(t * 2) * kTile + (uint32_t)row * 128u;
    auto step = [&](const int j, auto ragged) {
      const int b = j & 1, u = j >> 1;
      tc::mbar_wait(&s_full[t * 2 + b], u & 1);
      tc::tc_fence_after();
      float v[64];
      const uint32_t tS = tmem_base + (uint32_t)t * 128u + (uint32_t)b * 64u + lane_addr;
      tmem_ld32_nowait(tS, v);
      tmem_ld32_nowait(tS + 32, v + 32);
      tmem_ld_wait();
      tc::tc_fence_before();
      tc::mbar_arrive(&s_free[t * 2 + b]);           // S is in registers: Q K^T of tile j+2 may overwrite this buffer
      if (decltype(ragged)::value) {               // only the last, partial key tile pays for the masking (64 compare + select)
        const int kvalid = p.HW - j * 64;
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= kvalid) v[i] = -INFINITY;
      }
      float mx0 = v[0], mx1 = v[1], mx2 = v[2], mx3 = v[3];
#pragma unroll
      for (int i = 4; i < 64; i += 4) {
        mx0 = fmaxf(mx0, v[i]); mx1 = fmaxf(mx1, v[i + 1]); mx2 = fmaxf(mx2, v[i + 2]); mx3 = fmaxf(mx3, v[i + 3]);
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      // P V of tile j-2 has consumed P[t][b]
      tc::mbar_wait(&p_free[t * 2 + b], (u & 1) ^ 1);
      // lazy reference maximum: keep m_run unless this tile would push P above 2^kRescaleLog2
      const bool need = (mx - m_run) * p.scale_log2e > kRescaleLog2;      // true on the first tile (m_run = -inf)
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = need ? mx : m_run;
        if (j > 0) {
          const float alpha = need ? ex2((m_run - m_new) * p.scale_log2e) : 1.0f;
          // P V of tile j-1 has landed in O: it is the event that frees P[t][(j-1)&1], phase (j-1)>>1 of that barrier.  At this
          // point the barrier has completed either that phase or only the one before, so the parity wait is unambiguous.
          tc::mbar_wait(&p_free[t * 2 + ((j - 1) & 1)], ((j - 1) >> 1) & 1);
          tc::tc_fence_after();
#pragma unroll
          for (int c = 0; c < DC; ++c) {
            float o[32];
            tmem_ld32_nowait(tO + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] *= alpha;
            tmem_st32(tO + c * 32, o);
          }
          tmem_st_wait();
          tc::tc_fence_before();
          l_run *= alpha;
        }
        m_run = m_new;
      }
      const float sc = p.scale_log2e;
      const float mb = m_run * sc;
      // P = exp2(S*c - m*c), rounded to the MMA operand format, written K-major / 128B-swizzled into P[t][b]
      const uint32_t prow = prow0 + (uint32_t)b * kTile;
      float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {                 // 8 keys = one 16-byte chunk of this row
        const float e0 = ex2(fmaf(v[c * 8 + 0], sc, -mb)), e1 = ex2(fmaf(v[c * 8 + 1], sc, -mb));
        const float e2 = ex2(fmaf(v[c * 8 + 2], sc, -mb)), e3 = ex2(fmaf(v[c * 8 + 3], sc, -mb));
        const float e4 = ex2(fmaf(v[c * 8 + 4], sc, -mb)), e5 = ex2(fmaf(v[c * 8 + 5], sc, -mb));
        const float e6 = ex2(fmaf(v[c * 8 + 6], sc, -mb)), e7 = ex2(fmaf(v[c * 8 + 7], sc, -mb));
        rs0 += e0; rs1 += e1; rs2 += e2; rs3 += e3; rs0 += e4; rs1 += e5; rs2 += e6; rs3 += e7;
        st_shared_v4(prow + (uint32_t)((c ^ sw) << 4), pack2<T>(e0, e1), pack2<T>(e2, e3), pack2<T>(e4, e5), pack2<T>(e6, e7));
      }
      tc::fence_async_smem();                       // generic-proxy stores -> visible to the tensor core (async proxy)
      tc::mbar_arrive(&p_full[t * 2 + b]);
      l_run += (rs0 + rs1) + (rs2 + rs3);
    };
    const int nfull = p.HW / 64;                   // full 64-key tiles; at most one partial tile follows
    for (int j = 0; j < nfull; ++j) step(j, std::false_type());
    if (nfull < p.nkv) step(nfull, std::true_type());
    // epilogue: O / l -> 16-bit -> out[img, q, head*d + :]
    tc::mbar_wait(&o_final[t], 0);
    tc::tc_fence_after();
    const float inv = 1.0f / l_run;
    T* dst = (T*)p.out + ((int64_t)img * p.HW + q) * (p.heads * p.d) + head * p.d;
#pragma unroll
    for (int c = 0; c < DC; ++c) {
      float o[32];
      tmem_ld32_nowait(tO + c * 32, o);
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
  p.nkv = cdiv(HW, 64);
  p.nchunk = cdiv(d, 64);
  p.scale_log2e = (float)((1.0 / sqrt((double)d)) * 1.4426950408889634);
  static const int ahead = getenv("XRD_ATT_AHEAD") ? atoi(getenv("XRD_ATT_AHEAD")) : 1;   // measured: 0.327 ms (1) vs 0.340 ms (2) at B=16
  p.ahead = ahead == 1 ? 1 : 2;
  static const int probe = getenv("XRD_ATT_PROBE") ? atoi(getenv("XRD_ATT_PROBE")) : 1;
  p.probe = probe;
  p.out = out.p;
  alignas(64) CUtensorMap tmQ, tmKV;
  for (int which = 0; which < 2; ++which) {
    const cuuint64_t dims[3] = {(cuuint64_t)qkv.c, (cuuint64_t)HW, (cuuint64_t)qkv.n};
    const cuuint64_t strides[2] = {(cuuint64_t)qkv.c * 2, (cuuint64_t)HW * qkv.c * 2};
    const cuuint32_t box[3] = {64, which == 0 ? 128u : 64u, 1};     // Q tiles: 128 queries; K/V tiles: 64 keys
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = get_encode_tiled()(which == 0 ? &tmQ : &tmKV, tmap_dtype(qkv.dt), 3, qkv.p, dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail(XRD_ERR_CUDA, "cuTensorMapEncodeTiled(qkv) failed: %d", (int)r);
  }
  const size_t smem = 1024 + (size_t)8 * kTile + (size_t)2 * kRing * 2 * kKvTile + 64 * 8;
  dim3 grid(cdiv(HW, 256), heads, qkv.n);
  const int dc = d / 16;
#define XRD_ATT_CASE(TT, DCV)                                                                                             \
  case DCV: {                                                                                                             \
    static bool attr = false;                                                                                             \
    if (!attr) { XRD_CUDA(cudaFuncSetAttribute(k_attn_tc<TT, DCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; } \
    XRD_LAUNCH(c, (k_attn_tc<TT, DCV>), grid, kAttThreads, smem, tmQ, tmKV, p);                                                  \
  } break;
  if (qkv.dt == DT_BF16) {
    switch (dc) { XRD_ATT_CASE(__nv_bfloat16, 1) XRD_ATT_CASE(__nv_bfloat16, 2) XRD_ATT_CASE(__nv_bfloat16, 3) XRD_ATT_CASE(__nv_bfloat16, 4)
                  XRD_ATT_CASE(__nv_bfloat16, 5) XRD_ATT_CASE(__nv_bfloat16, 6) XRD_ATT_CASE(__nv_bfloat16, 7) XRD_ATT_CASE(__nv_bfloat16, 8) }
  } else {
    switch (dc) { XRD_ATT_CASE(__half, 1) XRD_ATT_CASE(__half, 2) XRD_ATT_CASE(__half, 3) XRD_ATT_CASE(__half, 4)
                  XRD_ATT_CASE(__half, 5) XRD_ATT_CASE(__half, 6) XRD_ATT_CASE(__half, 7) XRD_ATT_CASE(__half, 8) }
  }
#undef XRD_ATT_CASE
}

}  // namespace xrd
