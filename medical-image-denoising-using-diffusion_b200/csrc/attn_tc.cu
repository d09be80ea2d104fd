// attn_tc.cu -- flash-style self-attention on tcgen05 (AttentionBlock, HYB:292-305).
//
//   per CTA: 256 queries (two 128-row Q tiles) of one (image, head); loop over 64-key tiles j; for each Q tile t:
//     S_t,j = Q_t K_j^T     tcgen05.mma, A = Q smem (K-major), B = K smem (K-major), D = TMEM S[t][j&1]
//     P_t,j = exp2((S - m) * scale * log2e)   softmax warps: one tcgen05.ld pass -> registers -> f16/bf16 -> tcgen05.st back into
//                                             the first 32 columns of S[t][j&1] (two keys per 32-bit column)
//     O_t  += P_t,j V_j     tcgen05.mma, A = P in TENSOR MEMORY, B = V smem (MN-major: keys are the K dimension), D = TMEM O[t]
//   * P never touches shared memory: the first version stored it there (32 KB of stores + 32 KB of operand reads per key tile on
//     top of the 130 KB the Q/K/V operands cost: ~1470 cycles of the 128 B/clk shared-memory port per step against 1024 cycles of
//     MMA work, plus a fence.proxy.async per step); with the A operand in TMEM the port carries ~960 cycles per step.
//   * K/V tiles are what every CTA re-streams from L2 (measured: that traffic, not the tensor pipe, bounded the
//     one-Q-tile version); two Q tiles per CTA halve it, and two softmax warpgroups keep the MUFU pipe busy.
//   * The running output never leaves TMEM: the reference maximum m is only raised when a row's new maximum exceeds
//     it by more than 2^8 (P stays below 256, exact after the final division by the row sum accumulated with the
//     same m); only then is O rescaled in place (tcgen05.ld / tcgen05.st) -- after the first tiles almost never.
//   * 64-key tiles keep S in 64 registers per thread and leave room for a 6-deep K/V ring in shared memory.
//   * Launch shapes for small batches (attention_tc_shape): one Q tile per CTA while twice the CTAs still fit the SMs; for a single
//     image two CTAs share a Q tile's key range and k_attn_combine merges their partial results (the per-CTA chain over the key
//     tiles is latency-bound: 64 one-tile CTAs took 66 us, hardly less than 32 two-tile ones).
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
  int ntq;                 // Q tiles per CTA: 2 (throughput shape: the two tiles share every K/V tile), or 1 when the whole launch
                           // is so small (one or two images) that twice the CTAs of half the work fill more of the machine
  int nsplit;              // key-range splits per Q tile: 1, or 2 for single-image launches (one-Q-tile CTAs would fill under half of
                           // the SMs: 64 of 148 at 64x64).  With 2, CTA (x = 2*qtile + s) walks keys [s * nkv * 64, (s+1) * nkv * 64)
                           // (nkv is then the tile count of ONE range) and writes its un-normalised fp32 output, reference maximum
                           // and row sum; k_attn_combine merges the two ranges.
  float* part;             // nsplit == 2: [split][image][query][heads*d] fp32 partial outputs
  float2* ml;              // nsplit == 2: [split][image][head][query] (reference maximum * scale * log2 e, row sum)
  void* out;
};

constexpr int kAttThreads = 352;       // warp 0 TMA, warps 1 / 10 MMA issuers of Q tile 0 / 1, warps 2..9 softmax
constexpr int kAttWarpB = 10;
constexpr int kTile = 128 * 128;       // Q / P tile: [128 rows x 64 x 16-bit], 128B-swizzled = 16 KB
constexpr int kKvTile = 64 * 128;      // K / V chunk tile: [64 keys x 64 ch] = 8 KB
constexpr int kRing = 4;               // K ring slots and V ring slots (2 chunk tiles each)

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

// tcgen05.mma with the A operand in tensor memory (lanes = rows, one 32-bit column = two consecutive 16-bit K elements)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// which of every 8 key pairs take the polynomial exp2 (bit i = pair i): 3 of 8, spread so both paths stay interleaved
constexpr uint32_t kEmuMask = 0x00;     // measured: 0x92 (3 of 8) 0.307 ms vs 0.297 ms with none -- the MUFU pipe is not what paces the kernel

// (measured and dropped: ex2.approx.f16x2 -- one MUFU per two keys -- 0.389 ms against 0.381 ms, errors 1.5-2x: the MUFU pipe does not
// pace this kernel)
constexpr float kRescaleLog2 = 8.0f;   // raise the reference maximum only when P would exceed 2^8

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}


// ---- MMA-issuing warp: barrier block layout (byte offsets from `bars`) and the per-tile issue routines --------------------------
constexpr uint32_t kOffKFull = 8, kOffKEmpty = kOffKFull + 8 * kRing, kOffVFull = kOffKEmpty + 8 * kRing, kOffVEmpty = kOffVFull + 8 * kRing;
constexpr uint32_t kOffSFull = kOffVEmpty + 8 * kRing, kOffPFull = kOffSFull + 32, kOffPFree = kOffPFull + 32, kOffOFinal = kOffPFree + 32;
constexpr uint32_t kOffSFree = kOffOFinal + 16;

// TMEM map.  Two layouts exist.  Default (SEP = false): S[t][b] at t*128 + b*64, P[t][b] aliased onto the first 32 columns of S[t][b],
// O[t] at 256 + t*128; Q K^T of tile j+2 then queues behind the P V of tile j on the tensor pipe (same issuing thread).
// SEP (head dims <= 96 only): O[t] at 256 + t*96 and P[t] in its own 32 columns at 448 + t*32, so that Q K^T of tile j+2 may
// overwrite S[t][b] as soon as the softmax has READ S_t,j.  Measured at B=16, 4096 tokens, d=96: 0.401 ms against 0.279 ms for the
// aliased layout -- the extra "S read" barrier (128 arrivals per tile and step) and the second wait point in each issuing warp
// cost more than the earlier Q K^T buys.  Kept for experiments, off.
template <int D> struct AttMap {
  static constexpr bool SEP = false && D <= 96;
  static constexpr uint32_t OW = SEP ? 96u : 128u;
  __host__ __device__ static constexpr uint32_t S(int t, int b) { return (uint32_t)t * 128u + (uint32_t)b * 64u; }
  __host__ __device__ static constexpr uint32_t O(int t) { return 256u + (uint32_t)t * OW; }
  __host__ __device__ static constexpr uint32_t P(int t, int b) { return SEP ? 448u + (uint32_t)t * 32u : S(t, b); }
};

__device__ __forceinline__ void mbar_wait_addr(uint32_t addr, uint32_t parity) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) break;
    if (++spins > (1u << 22)) { printf("libxrd: attn_tc MMA warp: barrier %u timed out (block %d,%d,%d)\n", addr, blockIdx.x, blockIdx.y, blockIdx.z); __trap(); }
  }
}
__device__ __forceinline__ void umma_commit_addr(uint32_t addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(addr) : "memory");
}

// S_t = Q_t K^T for Q tile MT, K from ring slot SLOT, into S buffer B.  One thread.
template <typename T, int D16, int MT, int SLOT, int B>
__device__ __forceinline__ void att_issue_qk(uint32_t bar0, uint64_t dQ, uint64_t dK) {
  constexpr int D = 16 * D16;
  constexpr int KS0 = (D < 64 ? D : 64) / 16, KS1 = D > 64 ? (D - 64) / 16 : 0;
  constexpr uint32_t IDESC_QK = tc::umma_idesc(128, 64, tc::umma_fmt<T>());
  asm volatile("" : "+l"(dQ), "+l"(dK));        // keep `base + constant` in the uniform datapath (no hoisting into vector registers)
  const uint64_t bD = dK + (uint64_t)(SLOT * (2 * kKvTile >> 4));
  const uint64_t aD = dQ + (uint64_t)(MT * (2 * kTile >> 4));
  constexpr uint32_t tS = AttMap<D>::S(MT, B);
#pragma unroll
  for (int k = 0; k < KS0; ++k) tc::umma_f16(tS, aD + (uint64_t)(k * 2), bD + (uint64_t)(k * 2), IDESC_QK, k ? 1u : 0u);
#pragma unroll
  for (int k = 0; k < KS1; ++k) tc::umma_f16(tS, aD + (uint64_t)((kTile >> 4) + k * 2), bD + (uint64_t)((kKvTile >> 4) + k * 2), IDESC_QK, 1u);
  umma_commit_addr(bar0 + kOffSFull + 8u * (uint32_t)(MT * 2 + B));
  umma_commit_addr(bar0 + kOffKEmpty + 8u * SLOT);       // one of two arrivals (one per issuing warp)
}

// Key tile j = 4*m + S (ring slot S, S buffer S & 1) of Q tile MT.  `ph`: parity of the ring barriers for this round of four tiles.
//   separate P columns:  wait K_j+2 landed + S_MT,j read  ->  Q K^T of tile j+2;   wait V_j landed + P_MT,j stored  ->  P V of tile j
//   P aliased onto S:    wait V_j, P_MT,j, K_j+2 at once  ->  P V of tile j, then Q K^T of tile j+2 (ordered by the tensor pipe)
template <typename T, int D16, int MT, int S>
__device__ __forceinline__ void att_mma_step(int j, int nkv, uint32_t ph, uint32_t bar0, uint64_t dQ, uint64_t dK, uint64_t dV, int lane) {
  constexpr int D = 16 * D16;
  using Map = AttMap<D>;
  constexpr uint32_t IDESC_PV = tc::umma_idesc(128, D, tc::umma_fmt<T>()) | (1u << 16);     // B (V) is MN-major
  constexpr int B = S & 1, U1 = (S >> 1) & 1, S2 = (S + 2) & 3;
  const bool more = j + 2 < nkv;
  const uint32_t kpar = ph ^ (S >= 2 ? 1u : 0u);
  if (Map::SEP) {
    if (more) {
      if (lane < 2) mbar_wait_addr(bar0 + (lane == 0 ? kOffKFull + 8u * S2 : kOffSFree + 8u * (MT * 2 + B)), lane == 0 ? kpar : (uint32_t)U1);
      __syncwarp();
      tc::tc_fence_after();
      if (tc::elect_one()) att_issue_qk<T, D16, MT, S2, B>(bar0, dQ, dK);
      __syncwarp();
    }
    if (lane < 2) mbar_wait_addr(bar0 + (lane == 0 ? kOffVFull + 8u * S : kOffPFull + 8u * (MT * 2 + B)), lane == 0 ? ph : (uint32_t)U1);
  } else {
    if (lane < 3) {
      const uint32_t addr = bar0 + (lane == 0 ? kOffVFull + 8u * S : lane == 1 ? kOffPFull + 8u * (MT * 2 + B) : kOffKFull + 8u * S2);
      const uint32_t par = lane == 0 ? ph : lane == 2 ? kpar : (uint32_t)U1;
      if (lane < 2 || more) mbar_wait_addr(addr, par);
    }
  }
  __syncwarp();
  tc::tc_fence_after();
  if (tc::elect_one()) {
    uint64_t v_ = dV;
    asm volatile("" : "+l"(v_));
    const uint64_t bD = v_ + (uint64_t)(S * (2 * kKvTile >> 4));
    const uint32_t acc0 = j ? 1u : 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k)     // 4 x 16 keys: +8 columns (two keys each) of P in TMEM, +2048 B (16 key rows) in the MN-major V tile
      umma_f16_ts(Map::O(MT), Map::P(MT, B) + (uint32_t)(k * 8), bD + (uint64_t)(k * (2048 >> 4)), IDESC_PV, k ? 1u : acc0);
    umma_commit_addr(bar0 + kOffPFree + 8u * (MT * 2 + (Map::SEP ? 0 : B)));
    if (j == nkv - 1) umma_commit_addr(bar0 + kOffOFinal + 8u * MT);
    umma_commit_addr(bar0 + kOffVEmpty + 8u * S);         // one of two arrivals
    if (!Map::SEP && more) att_issue_qk<T, D16, MT, S2, B>(bar0, dQ, dK);
  }
  __syncwarp();
}

// The whole job of one MMA-issuing warp: every tensor-core instruction of Q tile MT.  Issue order: QK(0), QK(1), then per key
// tile j:  P V of tile j,  Q K^T of tile j+2  (S is double buffered; S[t][j&1] is rewritten by Q K^T of tile j+2 right behind the
// P V that read P_t,j out of it -- same thread, so the tensor pipe orders them).
// Measured on the one-warp version (ncu source view + a run with the softmax arithmetic removed: 0.371 ms against 0.384 ms with
// it): the kernel is paced by THIS instruction stream, not by the softmax -- 20 MMAs + 8 commits per key tile of 48 cycles each
// leave the single issuing thread ~4 cycles per instruction for its ~230 instructions.  Two issuing warps (one per Q tile) halve the
// stream each has to retire and decouple the two tiles' softmax -> P V -> Q K^T chains.
template <typename T, int D16, int MT>
__device__ __forceinline__ void att_mma_warp(int nkv, uint32_t bar0, uint64_t dQ, uint64_t dK, uint64_t dV, int lane) {
  if (lane == 0) mbar_wait_addr(bar0, 0);                              // q_full
  if (lane == 1) mbar_wait_addr(bar0 + kOffKFull, 0);
  if (lane == 2 && nkv > 1) mbar_wait_addr(bar0 + kOffKFull + 8u, 0);
  __syncwarp();
  tc::tc_fence_after();
  if (tc::elect_one()) {
    att_issue_qk<T, D16, MT, 0, 0>(bar0, dQ, dK);
    if (nkv > 1) att_issue_qk<T, D16, MT, 1, 1>(bar0, dQ, dK);
  }
  __syncwarp();
  uint32_t ph = 0;                  // parity of the K / V ring barriers for the four tiles of this round
  for (int j0 = 0; j0 < nkv; j0 += 4, ph ^= 1u) {
    att_mma_step<T, D16, MT, 0>(j0, nkv, ph, bar0, dQ, dK, dV, lane);
    if (j0 + 1 < nkv) att_mma_step<T, D16, MT, 1>(j0 + 1, nkv, ph, bar0, dQ, dK, dV, lane);
    if (j0 + 2 < nkv) att_mma_step<T, D16, MT, 2>(j0 + 2, nkv, ph, bar0, dQ, dK, dV, lane);
    if (j0 + 3 < nkv) att_mma_step<T, D16, MT, 3>(j0 + 3, nkv, ph, bar0, dQ, dK, dV, lane);
  }
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
  // layout: Q [2 q-tiles][2 chunks] | K ring [kRing][2 chunks] | V ring [kRing][2 chunks] | barriers
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + 4 * kTile;
  uint8_t* sV = sK + kRing * 2 * kKvTile;
  uint64_t* bars = (uint64_t*)(sV + kRing * 2 * kKvTile);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;              // [kRing]
  uint64_t* k_empty = k_full + kRing;       // [kRing]
  uint64_t* v_full = k_empty + kRing;       // [kRing]
  uint64_t* v_empty = v_full + kRing;       // [kRing]
  uint64_t* s_full = v_empty + kRing;       // [2 q-tiles][2]
  uint64_t* p_full = s_full + 4;            // P_t,j stored to TMEM (and, implied, S_t,j read)
  uint64_t* p_free = p_full + 4;            // P V of that tile completed (waited for only by the lazy rescale)
  uint64_t* o_final = p_free + 4;           // [2] completes once, after the last P V
  uint64_t* s_free = o_final + 2;           // [2 q-tiles][2] S_t,j is in the softmax warps' registers (separate-P map only)
  uint32_t* tmem_slot = (uint32_t*)(s_free + 4);
  using Map = AttMap<D>;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // split launches: blockIdx.x = nsplit * (Q tile) + key range (recomputed where it is used: nothing extra stays live in the loops)
  const int q0 = (p.nsplit > 1 ? (int)blockIdx.x / p.nsplit : (int)blockIdx.x) * 128 * p.ntq, head = blockIdx.y, img = blockIdx.z;
  const int qoff = head * p.d, koff = p.heads * p.d + head * p.d, voff = 2 * p.heads * p.d + head * p.d;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmQ);
    tc::tma_prefetch_desc(&tmKV);
    tc::mbar_init(q_full, 1);
    for (int s = 0; s < kRing; ++s) {
      // K / V slots are released by every issuing warp at work: one per Q tile
      tc::mbar_init(&k_full[s], 1); tc::mbar_init(&k_empty[s], (uint32_t)p.ntq); tc::mbar_init(&v_full[s], 1); tc::mbar_init(&v_empty[s], (uint32_t)p.ntq);
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
  // TMEM columns: see AttMap.

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (tc::elect_one()) {
      tc::mbar_expect_tx(q_full, (uint32_t)(p.ntq * NCH) * kTile);
      for (int t = 0; t < p.ntq; ++t)
        for (int c = 0; c < NCH; ++c) tc::tma_load_3d(sQ + (t * 2 + c) * kTile, &tmQ, q_full, qoff + 64 * c, q0 + t * 128, img);
      const int kbase = p.nsplit > 1 ? ((int)blockIdx.x % p.nsplit) * p.nkv * 64 : 0;      // first key of this CTA's range
      uint32_t slot = 0, phase = 0;
      for (int j = 0; j < p.nkv; ++j) {
        tc::mbar_wait(&k_empty[slot], phase ^ 1);
        tc::mbar_expect_tx(&k_full[slot], (uint32_t)NCH * kKvTile);
        for (int c = 0; c < NCH; ++c) tc::tma_load_3d(sK + (slot * 2 + c) * kKvTile, &tmKV, &k_full[slot], koff + 64 * c, kbase + j * 64, img);
        tc::mbar_wait(&v_empty[slot], phase ^ 1);
        tc::mbar_expect_tx(&v_full[slot], (uint32_t)NCH * kKvTile);
        for (int c = 0; c < NCH; ++c) tc::tma_load_3d(sV + (slot * 2 + c) * kKvTile, &tmKV, &v_full[slot], voff + 64 * c, kbase + j * 64, img);
        if (++slot == kRing) { slot = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == kAttWarpB) {
    // ===================== MMA issuers: warp 1 = Q tile 0, warp 10 = Q tile 1 (see att_mma_warp) =====================
    const uint32_t bar0 = tc::smem_u32(bars);
    const uint64_t dQ = tc::umma_desc_sw128(tc::smem_u32(sQ));           // + t * (2 * kTile >> 4)
    const uint64_t dK = tc::umma_desc_sw128(tc::smem_u32(sK));           // + slot * (2 * kKvTile >> 4)
    const uint64_t dV = umma_desc_mn_sw128(tc::smem_u32(sV), kKvTile);   // + slot * (2 * kKvTile >> 4)
    if (warp == 1) att_mma_warp<T, D16, 0>(p.nkv, bar0, dQ, dK, dV, lane);
    else if (p.ntq == 2) att_mma_warp<T, D16, 1>(p.nkv, bar0, dQ, dK, dV, lane);
  } else if (((warp - 2) >> 2) < p.ntq) {
    // ===================== softmax / correction / epilogue (warps 2..9, one query row per thread) =====================
    const int t = (warp - 2) >> 2;                 // Q tile
    const int quad = warp & 3;                     // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int q = q0 + t * 128 + row;
    const uint32_t tO = tmem_base + Map::O(t) + lane_addr;
    float m_run = -INFINITY, l_run = 0.f;
    auto step = [&](const int j, auto ragged) {
      const int b = j & 1, u = j >> 1;
      tc::mbar_wait(&s_full[t * 2 + b], u & 1);
      tc::tc_fence_after();
      float v[64];
      const uint32_t tS = tmem_base + Map::S(t, b) + lane_addr;
      tmem_ld32_nowait(tS, v);
      tmem_ld32_nowait(tS + 32, v + 32);
      tmem_ld_wait();
      if (Map::SEP) {                                // S is in registers: Q K^T of tile j+2 may overwrite this buffer
        tc::tc_fence_before();
        tc::mbar_arrive(&s_free[t * 2 + b]);
      }
      if (decltype(ragged)::value) {               // only the last, partial key tile pays for the masking (64 compare + select)
        const int kvalid = p.HW - j * 64;            // (never a split launch: those have no partial tile)
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= kvalid) v[i] = -INFINITY;
      }
      if (p.probe == 2) {        // timing experiment only (wrong results): no max / exp, P = packed raw S
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) pk[i] = pack2<T>(v[2 * i], v[2 * i + 1]);
        if (Map::SEP && j > 0) { tc::mbar_wait(&p_free[t * 2], (j - 1) & 1); tc::tc_fence_after(); }
        tmem_st32(tmem_base + Map::P(t, b) + lane_addr, reinterpret_cast<const float*>(pk));
        tmem_st_wait();
        tc::tc_fence_before();
        tc::mbar_arrive(&p_full[t * 2 + b]);
        l_run = 1.f;
        return;
      }
      // row maximum: three-input max (FMNMX3), four independent chains
      float mx0 = max3(v[0], v[1], v[2]), mx1 = max3(v[3], v[4], v[5]), mx2 = max3(v[6], v[7], v[8]), mx3 = max3(v[9], v[10], v[11]);
#pragma unroll
      for (int i = 12; i < 60; i += 8) {
        mx0 = max3(mx0, v[i], v[i + 1]); mx1 = max3(mx1, v[i + 2], v[i + 3]);
        mx2 = max3(mx2, v[i + 4], v[i + 5]); mx3 = max3(mx3, v[i + 6], v[i + 7]);
      }
      mx0 = max3(mx0, v[60], v[61]); mx1 = max3(mx1, v[62], v[63]);
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      // lazy reference maximum: keep m_run unless this tile would push P above 2^kRescaleLog2
      const bool need = (mx - m_run) * p.scale_log2e > kRescaleLog2;      // true on the first tile (m_run = -inf)
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = need ? mx : m_run;
        if (j > 0) {
          const float alpha = need ? ex2((m_run - m_new) * p.scale_log2e) : 1.0f;
          // P V of tile j-1 has landed in O: it is the event that frees P[t][(j-1)&1], phase (j-1)>>1 of that barrier.  At this
          // point the barrier has completed either that phase or only the one before, so the parity wait is unambiguous.
          if (Map::SEP) tc::mbar_wait(&p_free[t * 2], (j - 1) & 1);       // one barrier per Q tile, one phase per key tile
          else tc::mbar_wait(&p_free[t * 2 + ((j - 1) & 1)], ((j - 1) >> 1) & 1);
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
      // P = exp2(S*c - m*c), rounded to the MMA operand format, stored over the first 32 columns of S[t][b] (two keys per column).
      // The MUFU pipe (4 lanes per scheduler: 8 cycles per warp instruction) is the busiest unit of this kernel -- 64 exponentials
      // per thread and key tile, two softmax warps per scheduler = 1024 cycles per step against ~960 of MMA work.  kEmu of every 8
      // key pairs therefore take exp2 on the FMA / ALU pipes instead (packed fp32 pairs): x = n + f with n = round(x) from the
      // magic-number add, 2^f by a degree-3 polynomial on [-1/2, 1/2] (max relative error 7.5e-5, a third of the 16-bit rounding
      // of P that follows), 2^n by an integer add into the exponent field.  Scale, subtraction and row sums are packed too.
      const float2 sc2 = make_float2(sc, sc), nmb2 = make_float2(-mb, -mb);
      float2 rsa = make_float2(0.f, 0.f), rsb = make_float2(0.f, 0.f);
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {                 // pair i = keys 2i, 2i+1
        float2 x = tc::ffma2(make_float2(v[2 * i], v[2 * i + 1]), sc2, nmb2);
        float2 e;
        if (kEmuMask >> (i & 7) & 1) {
          x.x = fmaxf(x.x, -126.f); x.y = fmaxf(x.y, -126.f);
          const float2 tt = tc::fadd2(x, make_float2(12582912.f, 12582912.f));            // 1.5 * 2^23: the low mantissa bits are round(x)
          const float2 rr = tc::fadd2(tt, make_float2(-12582912.f, -12582912.f));
          const float2 f = tc::ffma2(rr, make_float2(-1.f, -1.f), x);                        // x - round(x)
          float2 q = tc::ffma2(f, make_float2(0.05517165f, 0.05517165f), make_float2(0.24261113f, 0.24261113f));
          q = tc::ffma2(q, f, make_float2(0.69326097f, 0.69326097f));
          q = tc::ffma2(q, f, make_float2(0.99992806f, 0.99992806f));
          e.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(tt.x) << 23));
          e.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(tt.y) << 23));
        } else {
          e.x = ex2(x.x); e.y = ex2(x.y);
        }
        if (i & 1) rsb = tc::fadd2(rsb, e); else rsa = tc::fadd2(rsa, e);
        pk[i] = pack2<T>(e.x, e.y);
      }
      const float rs0 = rsa.x + rsa.y, rs1 = rsb.x + rsb.y, rs2 = 0.f, rs3 = 0.f;
      if (Map::SEP && j > 0) {                       // P[t] is single buffered: P V of tile j-1 must have read it
        tc::mbar_wait(&p_free[t * 2], (j - 1) & 1);
        tc::tc_fence_after();
      }
      tmem_st32(tmem_base + Map::P(t, b) + lane_addr, reinterpret_cast<const float*>(pk));
      tmem_st_wait();
      tc::tc_fence_before();
      tc::mbar_arrive(&p_full[t * 2 + b]);
      l_run += (rs0 + rs1) + (rs2 + rs3);
    };
    // full 64-key tiles; at most one partial tile follows (split launches need HW % 128 == 0: every tile of a range is full)
    const int nfull = p.nsplit > 1 ? p.nkv : p.HW / 64;
    for (int j = 0; j < nfull; ++j) step(j, std::false_type());
    if (nfull < p.nkv) step(nfull, std::true_type());
    // epilogue: O / l -> 16-bit -> out[img, q, head*d + :]
    tc::mbar_wait(&o_final[t], 0);
    tc::tc_fence_after();
    if (p.nsplit > 1) {
      const int split = (int)blockIdx.x % p.nsplit;
      // partial result of this key range: O (un-normalised, relative to the reference maximum), the maximum in exp2 units, the sum
      float* dp = p.part + (((int64_t)split * gridDim.z + img) * p.HW + q) * (p.heads * p.d) + head * p.d;
#pragma unroll
      for (int c = 0; c < DC; ++c) {
        float o[32];
        tmem_ld32_nowait(tO + c * 32, o);
        tmem_ld_wait();
        if (q < p.HW) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            if (c * 32 + i < p.d) *reinterpret_cast<float4*>(dp + c * 32 + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
        }
      }
      if (q < p.HW) p.ml[(((int64_t)split * gridDim.z + img) * p.heads + head) * p.HW + q] = make_float2(m_run * p.scale_log2e, l_run);
    } else {
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
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

// Merge of the two key ranges of a split launch: out = (w0 O0 + w1 O1) / (w0 l0 + w1 l1), w_s = 2^(m_s - max(m0, m1)).
// One thread per (query, 8 channels): 2 x 32 B of partials in, 16 B out.
template <typename T>
__global__ void __launch_bounds__(256) k_attn_combine(const float* __restrict__ part, const float2* __restrict__ ml, T* __restrict__ out,
                                                      int n, int HW, int heads, int d) {
  const int C = heads * d, g8 = C / 8;
  const int64_t rows = (int64_t)n * HW;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * g8) return;
  const int64_t row = idx / g8;
  const int ch = (int)(idx - row * g8) * 8;
  const int head = ch / d;
  const int64_t img = row / HW, q = row - img * HW;
  const float2 a = ml[(img * heads + head) * HW + q];
  const float2 b = ml[(((int64_t)n + img) * heads + head) * HW + q];
  const float M = fmaxf(a.x, b.x);
  float w0 = ex2(a.x - M), w1 = ex2(b.x - M);
  const float inv = 1.0f / (w0 * a.y + w1 * b.y);
  w0 *= inv; w1 *= inv;
  const float4* p0 = reinterpret_cast<const float4*>(part + row * C + ch);
  const float4* p1 = reinterpret_cast<const float4*>(part + (rows + row) * C + ch);
  const float4 x0 = p0[0], x1 = p0[1], y0 = p1[0], y1 = p1[1];
  uint4 w;
  w.x = pack2<T>(w0 * x0.x + w1 * y0.x, w0 * x0.y + w1 * y0.y); w.y = pack2<T>(w0 * x0.z + w1 * y0.z, w0 * x0.w + w1 * y0.w);
  w.z = pack2<T>(w0 * x1.x + w1 * y1.x, w0 * x1.y + w1 * y1.y); w.w = pack2<T>(w0 * x1.z + w1 * y1.z, w0 * x1.w + w1 * y1.w);
  *reinterpret_cast<uint4*>(out + row * C + ch) = w;
}

// Launch shape.  ntq: one Q tile per CTA while twice the CTAs still fit the SMs (one or two images at 64x64: the served shape ran 32
// CTAs of 72 us each, profiles/r02_launches_ddim_b1.csv).  nsplit: when even the one-tile CTAs fill under half of the machine (one
// image: 64 CTAs of 66 us, each a serial chain over 64 key tiles), two CTAs share a Q tile's key range.
static void attention_tc_shape(const Tens& qkv, int heads, int& ntq, int& nsplit) {
  static const int ntq_env = getenv("XRD_ATT_NTQ") ? atoi(getenv("XRD_ATT_NTQ")) : 0;
  static const int split_env = getenv("XRD_ATT_SPLIT") ? atoi(getenv("XRD_ATT_SPLIT")) : 1;
  const int HW = qkv.h * qkv.w, nsm = sm_count();
  ntq = (ntq_env == 1 || ntq_env == 2) ? ntq_env : (2 * cdiv(HW, 256) * heads * qkv.n <= nsm ? 1 : 2);
  nsplit = 1;
  if (split_env && ntq == 1 && HW % 128 == 0 && HW / 64 >= 8 && 2 * (HW / 128) * heads * qkv.n <= nsm) nsplit = 2;
}

size_t attention_tc_scratch_floats(const Tens& qkv, int heads) {
  if (!attention_tc_supported(qkv, heads)) return 0;
  int ntq, nsplit;
  attention_tc_shape(qkv, heads, ntq, nsplit);
  if (nsplit == 1) return 0;
  const size_t rows = (size_t)qkv.n * qkv.h * qkv.w;
  return (size_t)nsplit * rows * (size_t)(qkv.c / 3) + (size_t)nsplit * rows * heads * 2;
}

bool attention_tc_supported(const Tens& qkv, int heads) {
  if (qkv.dt == DT_F32) return false;
  if (qkv.c % (3 * heads) != 0) return false;
  const int d = qkv.c / (3 * heads);
  if (d % 16 != 0 || d < 16 || d > 128) return false;
  if ((heads * d) % 8 != 0) return false;
  return true;
}

void attention_tc(Ctx& c, const Tens& qkv, int heads, Tens& out, float* scratch) {
  XRD_REQUIRE(attention_tc_supported(qkv, heads), "attention_tc: unsupported shape (C=%d heads=%d)", qkv.c, heads);
  const int d = qkv.c / (3 * heads);
  XRD_REQUIRE(out.c == heads * d && out.n == qkv.n && out.h == qkv.h && out.w == qkv.w && out.dt == qkv.dt, "attention_tc: output shape");
  if (c.dry) return;
  const int HW = qkv.h * qkv.w;
  AttnTcP p;
  p.HW = HW; p.heads = heads; p.d = d;
  attention_tc_shape(qkv, heads, p.ntq, p.nsplit);
  if (!scratch) p.nsplit = 1;                       // no workspace for partial results: one CTA walks the whole key range
  p.nkv = cdiv(HW, 64) / p.nsplit;                  // split launches: HW % 128 == 0
  p.part = scratch;
  p.ml = scratch ? reinterpret_cast<float2*>(scratch + (size_t)p.nsplit * qkv.n * HW * (heads * d)) : nullptr;
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
  const size_t smem = 1024 + (size_t)4 * kTile + (size_t)2 * kRing * 2 * kKvTile + 64 * 8;
  dim3 grid(cdiv(HW, 128 * p.ntq) * p.nsplit, heads, qkv.n);
  const int dc = d / 16;
#define XRD_ATT_CASE(TT, DCV)                                                                                             \
  case DCV: {                                                                                                             \
    ensure_dyn_smem(k_attn_tc<TT, DCV>, (int)smem);                                                                       \
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
  if (p.nsplit > 1) {
    const int64_t items = (int64_t)qkv.n * HW * (heads * d / 8);
    const int blocks = (int)cdiv64(items, 256);
    if (qkv.dt == DT_BF16) XRD_LAUNCH(c, k_attn_combine<__nv_bfloat16>, blocks, 256, 0, p.part, p.ml, (__nv_bfloat16*)out.p, qkv.n, HW, heads, d);
    else XRD_LAUNCH(c, k_attn_combine<__half>, blocks, 256, 0, p.part, p.ml, (__half*)out.p, qkv.n, HW, heads, d);
  }
}

}  // namespace xrd
