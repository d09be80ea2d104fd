// conv3s.cu -- row-ring tcgen05 3x3 convolution (stride 1, pad 1) with the three VERTICAL taps stacked along N.
//
// Why: tcgen05.mma M=128, K=16 from shared memory costs max(N/2, 32 + N/4) cycles (tools/probes/mma_probe.cu): an N=48 MMA takes
// 44.5 cycles (54 % of the tensor peak -- it is bound by the 128 B/clk shared-memory read of A), an N=144 MMA 72.5 cycles (100 %).
// A 3x3 convolution with 48 output channels therefore wastes half the tensor pipe when every tap is its own N=48 MMA
// (conv3r.cu / conv3.cu).  Here each landed INPUT row r is multiplied once per horizontal tap dx by the stacked weight block
//        B(dx) = [ W(dy=+1,dx) | W(dy=0,dx) | W(dy=-1,dx) ]          (N = 3 x 48 = 144)
// and the three 48-column results are the contributions of row r to the output rows r-1, r, r+1.  The accumulators of
// consecutive output rows are consecutive 48-column blocks of a TMEM ring (10 blocks), so the N=144 result lands directly on
// the three accumulators it belongs to: no epilogue-side summation, 3 MMAs per k-step instead of 9 (217 vs 400 cycles).
//   * output row o is first touched by input row o-1... in ring order: by the LAST block of the MMA of input row (o-1)+... see
//     c3s_issue_row: the first k-step of every input row is issued block by block so that the freshly recycled accumulator block
//     is overwritten (accumulate = 0) while its two neighbours accumulate; a block run that wraps around the ring end is split
//     into two MMAs;
//   * work item = (image, 128-column block, segment of up to `seg` image rows); every input row of the segment (+ one halo row
//     above and below) is fetched ONCE by TMA into a ring of row slots (conv3r.cu's ring), one 17 KB slot per 64-channel chunk;
//   * up to two chunks per row: a 65..128-channel input (64 + tail) or a virtual concat of two <= 64-channel tensors
//     (torch.cat([x, skip]) of the UNet's up path, HYB:383) -- the concat is never materialised;
//   * output channels in slices of 48: a CTA owns one slice (its 9 x chunks weight blocks stay resident), so 96 outputs are two
//     slices over the same input rows (the activations are read twice through L2; the N=144 MMAs more than pay for it);
//   * GroupNorm + SiLU of the input (HYB:264-265, 269-270) applied in place to each landed row by four extra warps (GN variant);
//   * 8-warp register epilogue as in conv3.cu: + bias + time-embedding row + residual, GroupNorm sums of the output.
#include "kernels.cuh"
#include "tc_common.cuh"

#include <algorithm>
#include <mutex>
#include <type_traits>
#include <vector>
#include <stdlib.h>

#ifndef XRD_C3S_XF_BOTH
#define XRD_C3S_XF_BOTH 1
#endif

namespace xrd {

struct Conv3SP {
  int H, W, nimg;
  int ncb, dpc, ndec;           // column blocks per row, decades (kSNB output rows) per image column, decades in total (per slice)
  int cout_total;               // 48 * slices
  int c0, c1;                   // channels of chunk 0 / chunk 1 (0: single chunk)
  int two_src;                  // chunk 1 is the second tensor (tmA1, channel 0) instead of channels 64.. of the first
  const float* bias;
  const float* chan_add; int chan_add_bstride;
  const void* resid;
  void* y;
  double* stats;
  const float2* in_coef;        // GN variant: [nimg][c0 + c1] (0.5*scale, 0.5*shift)
};

// Warp roles, aligned to warpgroups so that each role gets its own register budget (setmaxnreg):
//   warpgroup 0: warp 0 TMA producer, warps 1 + 2 MMA issuers (ping-pong), warp 3 idle        -> 56 registers
//   warpgroups 1, 2: the two epilogue groups (warps 4..7, 8..11; warp % 4 = its TMEM lane quarter) -> 184 (GN) / 216
//   warpgroups 3, 4 (GN variant only): input transform, warps 12..15 take the even landed rows, 16..19 the odd ones -> 80
//     (ncu of the version with ONE transform warpgroup: its warps were busy 90 % of the time, 2150 cycles per row against 1160
//     of the plain kernel -- the transform, not the tensor pipe, paced the kernel)
// (the first version had 11 / 15 warps under one budget: the GN variant's 480 threads left 128 registers per thread and its
// epilogue spilled; ptxas -v of that version: 88 bytes of spill stores, 320 of spill loads.)
constexpr int kSThreads = 384, kSThreadsGN = 640;
constexpr bool kSXfBoth = XRD_C3S_XF_BOTH;         // both transform warpgroups work on every landed row (half the pixels each)
constexpr int kSWarpB = 2;                          // the second MMA-issuing warp
constexpr int kSRegLow = 56, kSRegLowGN = 40, kSRegXf = 80, kSRegEpiGN = 136, kSRegEpi = 216;
// The pool setmaxnreg draws from is what the CTA was LAUNCHED with (threads x registers per thread), not the whole register file:
// GN variant 640 x 96 = 480 x 128 >= (40 + 2 * 136 + 2 * 80) x 128 (measured: 80 / 136 beats 64 / 152 -- GN / plain time 1.13 against 1.37); plain 384 x 168 = 504 x 128 >= (56 + 2 * 216) x 128.  An inc beyond the
// pool never returns (first 640-thread version: 48 + 2*160 + 2*72 = 512 units against 480 -- the second epilogue group hung).
static_assert(kSRegLowGN + 2 * kSRegEpiGN + 2 * kSRegXf <= (kSThreadsGN * 96) / 128, "GN variant: setmaxnreg budget exceeds the launch allocation");
static_assert(kSRegLow + 2 * kSRegEpi <= (kSThreads * 168) / 128, "plain variant: setmaxnreg budget exceeds the launch allocation");
constexpr int kSBox = 130;                   // pixels fetched per row (128 + halo column each side)
constexpr uint32_t kSSlot = 136 * 128;       // one chunk of one row: 17 KB keeps every row 1024-byte aligned
constexpr int kSNB = 10;                     // accumulator blocks (48 columns each) in the TMEM ring
constexpr uint32_t kSBlk = 48 * 128;         // one weight block: 48 output channels x 64 input channels, 16 bit
constexpr uint32_t kSBlk16 = kSBlk >> 4;

__device__ __forceinline__ void s3_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// All MMAs of one landed input row of a SEG-row work item.  J (the row's index inside the item, 0 .. SEG+1) and SEG are literal
// at every call site, so after inlining every quantity below is a compile-time constant: which outputs the row feeds
// (J-2 .. J clipped to the item), which weight blocks of the stack those are, which ring blocks of TMEM (items always start at
// block 0: SEG is a multiple of the ring length), where a run of blocks wraps around the ring end and has to be split into two
// MMAs, and which block is touched for the first time.  The issue thread executes nothing but descriptor additions and MMAs.
// (ncu source view of the first, runtime-parameterised version: ~250 scalar instructions per row in the single issuing
// thread, 1335 cycles per row against 700 cycles of MMA work -- the issue thread, not the tensor pipe or HBM, was the bound.)
// `baton` (nullable): arrived on just before the last horizontal tap of the row is issued -- the other issuing warp may then start
// its burst; its MMAs interleave with the last few of this row, which is harmless: every block this row touches has had its
// first (overwriting) MMA long before, and accumulation order does not matter.
template <int KS0, int KS1, int J, int SEG>
__device__ __forceinline__ void c3s_issue_row(uint64_t a_base, uint64_t b_base, uint32_t id48, uint64_t* baton) {
  constexpr int LO = J - 2 < 0 ? 0 : J - 2, HI = J > SEG - 1 ? SEG - 1 : J, N = HI - LO + 1;
  constexpr int FB = 2 - (J - LO);             // first weight block of the stack: 0 = dy +1 ... 2 = dy -1
  constexpr bool FRESH = J <= SEG - 1;         // output J exists: its block is overwritten by this row's first MMA
  constexpr int T0 = LO % kSNB;
  auto idn = [&](int nn) { return id48 + ((uint32_t)((nn - 1) * 6) << 17); };      // N field = N >> 3 at bit 17
  // first k-step: the accumulating blocks (one or two pieces), then the fresh block on its own
  {
    constexpr int NA = FRESH ? N - 1 : N;
    constexpr int A1 = NA < kSNB - T0 ? NA : kSNB - T0;
    if (A1 > 0) tc::umma_f16((uint32_t)(T0 * 48), a_base, b_base + (uint64_t)(FB * kSBlk16), idn(A1), 1u);
    if (NA - A1 > 0) tc::umma_f16(0u, a_base, b_base + (uint64_t)((FB + A1) * kSBlk16), idn(NA - A1), 1u);
    if (FRESH) tc::umma_f16((uint32_t)(((T0 + N - 1) % kSNB) * 48), a_base, b_base + (uint64_t)((FB + N - 1) * kSBlk16), id48, 0u);
  }
  constexpr int P1 = N < kSNB - T0 ? N : kSNB - T0;
#pragma unroll
  for (int c = 0; c < (KS1 ? 2 : 1); ++c) {
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
      for (int k = 0; k < (c ? KS1 : KS0); ++k) {
        if (c == 0 && dx == 0 && k == 0) continue;
        if (c == (KS1 ? 1 : 0) && dx == 2 && k == 0 && baton) { tc::tc_fence_before(); tc::mbar_arrive(baton); }
        const uint64_t ao = a_base + (uint64_t)(c * (kSSlot >> 4) + dx * 8 + k * 2);
        const uint64_t bo = b_base + (uint64_t)((c * 9 + dx * 3 + FB) * kSBlk16 + k * 2);
        tc::umma_f16((uint32_t)(T0 * 48), ao, bo, idn(P1), 1u);
        if (N - P1 > 0) tc::umma_f16(0u, ao, bo + (uint64_t)(P1 * kSBlk16), idn(N - P1), 1u);
      }
    }
  }
}

// MMA-warp state that survives from row to row: everything else about a row is a compile-time constant
struct C3SState {
  uint32_t sl;     // ring slot of the next input row
  uint32_t rph;    // bit s: phase of r_full / r_ready[s] the next use of slot s waits for
  uint32_t accb;   // how many times every accumulator block was used before this item (only the parity matters)
  uint32_t k;      // running iteration number: iteration k belongs to issuing warp k & 1
};
struct C3SBars { uint64_t *rbar, *r_empty, *acc_full, *acc_empty, *baton; };

// One issue iteration = the two landed input rows J, J+1 (pattern indices; see c3s_issue_row) of a work item; `dec` = use index
// of the accumulator blocks of outputs J, J+1 within the item (the decade of the output row).
// Two warps issue alternately (`my` = 0 / 1): while one is blocked feeding its burst into the (shallow) MMA queue, the other runs
// the barrier waits and the address arithmetic of the next iteration, so the tensor pipe does not idle through them (ncu of
// the single-issuer version: ~1800 cycles of burst and ~2000 cycles of waits / warp synchronisation / set-up per iteration).
// Both warps walk every iteration and keep the same state; only the owner waits, issues and commits.
template <int KS0, int KS1, int R, int J, int SEG, bool SPLIT>
__device__ __forceinline__ void c3s_iteration(C3SState& st, const C3SBars& b, uint32_t dec, uint64_t adesc0, uint64_t bdesc0, uint32_t id48,
                                              uint32_t row16, int lane, uint32_t my) {
  const uint32_t s0 = st.sl, s1 = (st.sl + 1 == (uint32_t)R) ? 0u : st.sl + 1;
  if ((st.k & 1u) == my) {
    // lanes 0 / 1: the two input rows have landed (and are transformed); lanes 8 / 9: the accumulator blocks the rows touch for the
    // first time (outputs J, J+1, when they exist) have been drained by the epilogue; lane 2: the other warp has issued its burst
    // A parity wait can only tell "the current phase" from "the one before": with a ring of fewer than four rows this iteration
    // shares a slot with the previous one (the other warp's), and probing that slot's barrier before the other warp's wait has
    // been satisfied would succeed on the WRONG phase.  There the baton (passed after the other warp's waits) is taken first.
    if (R < 4) {
      if (lane == 2 && st.k > 0u) tc::mbar_wait(&b.baton[my], ((st.k >> 1) & 1u) ^ (my ? 0u : 1u), 34);
      __syncwarp();
    }
    // SPLIT (shallow ring + input transform: the two-chunk GN variants): the ring cannot hold the two rows of the NEXT iteration
    // while this one runs, so the second row is waited for only after the first has been issued and its load + transform overlap
    // the first row's MMAs (both waits up front: 3700 cycles per row measured, 96->48 @512 574 us; split: 489 us).  Without a
    // transform the split costs more than it hides (plain 96->48 @512: 370 us up front, 426 us split).
    if (lane < (SPLIT ? 1 : 2)) {
      const uint32_t sx = lane ? s1 : s0;
      tc::mbar_wait(&b.rbar[sx], (st.rph >> sx) & 1u, 32);
    } else if (lane >= 8 && lane < 10 && J + (lane - 8) < SEG) {
      const int o = J + (lane - 8);
      tc::mbar_wait(&b.acc_empty[o % kSNB], ((st.accb + dec) & 1u) ^ 1u, 33);
    } else if (R >= 4 && lane == 2 && st.k > 0u) {
      tc::mbar_wait(&b.baton[my], ((st.k >> 1) & 1u) ^ (my ? 0u : 1u), 34);
    }
    __syncwarp();
    tc::tc_fence_after();
    // opaque copy of the weight descriptor: without it the compiler hoists every `bdesc0 + constant` of the unrolled bursts out
    // of the item loop into VECTOR registers and moves them back (R2UR) in front of each MMA
    uint64_t bd = bdesc0;
    asm volatile("" : "+l"(bd));
    const uint64_t a0 = adesc0 + (uint64_t)(s0 * row16), a1 = adesc0 + (uint64_t)(s1 * row16);
    const bool leader = tc::elect_one();       // ONE thread issues and commits both rows (a commit covers the issuing thread's MMAs)
    if (leader) {
      c3s_issue_row<KS0, KS1, J, SEG>(a0, bd, id48, nullptr);
      if (J < SEG) tc::umma_commit(&b.acc_full[J % kSNB]);               // output J: its first row is in (arrival 1 of 2)
      if (J >= 2) tc::umma_commit(&b.acc_full[(J - 2) % kSNB]);          // output J-2: its last row is in (arrival 2 of 2)
      tc::umma_commit(&b.r_empty[s0]);                                   // the input row is not needed again
    }
    if (SPLIT) {
      __syncwarp();
      if (lane == 1) tc::mbar_wait(&b.rbar[s1], (st.rph >> s1) & 1u, 32);
      __syncwarp();
      tc::tc_fence_after();
    }
    if (leader) {
      c3s_issue_row<KS0, KS1, J + 1, SEG>(a1, bd, id48, &b.baton[my ^ 1u]);
      if (J + 1 < SEG) tc::umma_commit(&b.acc_full[(J + 1) % kSNB]);
      if (J >= 1) tc::umma_commit(&b.acc_full[(J - 1) % kSNB]);
      tc::umma_commit(&b.r_empty[s1]);
    }
    __syncwarp();
  }
  st.rph ^= (1u << s0) | (1u << s1);
  st.sl = (s1 + 1 == (uint32_t)R) ? 0u : s1 + 1;
  st.k += 1u;
}

// A work item of `seg` output rows (any multiple of the accumulator ring length kSNB): rows 0, 1 (top edge), the interior rows
// decade by decade (the five iterations of a decade have the same patterns in every decade), rows seg, seg+1 (bottom edge: they
// feed the outputs seg-2, seg-1 = ring blocks 8, 9 only, whatever the length of the item -- the pattern of rows 10, 11 of a
// 10-row item).
template <int KS0, int KS1, int R, bool SPLIT>
__device__ __forceinline__ void c3s_run_item(C3SState& st, const C3SBars& b, int seg, uint64_t adesc0, uint64_t bdesc0, uint32_t id48,
                                             uint32_t row16, int lane, uint32_t my) {
  constexpr int INT = 1000;      // "long enough": rows 2 .. 11 of such an item are interior rows
  c3s_iteration<KS0, KS1, R, 0, INT, SPLIT>(st, b, 0u, adesc0, bdesc0, id48, row16, lane, my);
  for (int m = 0; m * kSNB + 2 < seg; ++m) {
    const uint32_t um = (uint32_t)m;
    c3s_iteration<KS0, KS1, R, 2, INT, SPLIT>(st, b, um, adesc0, bdesc0, id48, row16, lane, my);
    c3s_iteration<KS0, KS1, R, 4, INT, SPLIT>(st, b, um, adesc0, bdesc0, id48, row16, lane, my);
    c3s_iteration<KS0, KS1, R, 6, INT, SPLIT>(st, b, um, adesc0, bdesc0, id48, row16, lane, my);
    c3s_iteration<KS0, KS1, R, 8, INT, SPLIT>(st, b, um, adesc0, bdesc0, id48, row16, lane, my);
    if (m * kSNB + 10 < seg) c3s_iteration<KS0, KS1, R, 10, INT, SPLIT>(st, b, um + 1u, adesc0, bdesc0, id48, row16, lane, my);
  }
  c3s_iteration<KS0, KS1, R, kSNB, kSNB, SPLIT>(st, b, 0u, adesc0, bdesc0, id48, row16, lane, my);
  st.accb += (uint32_t)(seg / kSNB);
}

// Work distribution.  The output is cut into DECADES (kSNB = 10 consecutive rows of one 128-column block of one image); the
// decades of a slice are numbered image by image, column block by column block, top to bottom, and worker w of nw takes the
// contiguous range [w*ndec/nw, (w+1)*ndec/nw): every CTA gets the same number of decades +-1 (the first version dealt 40- and
// 10-row items round-robin: 12 % idle tail at 512x512, batch 16).  A work item is a maximal run of a worker's decades inside one
// image column, so a worker has one or two items per column it touches and fetches two halo rows per item, not per decade.
// Every role walks the same sequence with this iterator.
struct C3SItems {
  int d, d1;
  __device__ __forceinline__ C3SItems(const Conv3SP& p, int wi, int nw) {
    d = (int)(((int64_t)wi * p.ndec) / nw);
    d1 = (int)(((int64_t)(wi + 1) * p.ndec) / nw);
  }
  // rows past the image end (last decade of a column) are zero-filled by TMA, multiplied like any other row and dropped by the
  // epilogue, so that ring slots, accumulator blocks and barrier phases stay in step with the compile-time pattern of the MMA warp
  __device__ __forceinline__ bool next(const Conv3SP& p, int& img, int& cb, int& r0, int& rows) {
    if (d >= d1) return false;
    const int col = d / p.dpc, dd = d - col * p.dpc;
    const int n = min(p.dpc - dd, d1 - d);
    cb = col % p.ncb; img = col / p.ncb;
    r0 = dd * kSNB; rows = n * kSNB;
    d += n;
    return true;
  }
};

template <uint32_t N> __device__ __forceinline__ void c3s_reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <uint32_t N> __device__ __forceinline__ void c3s_reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

// Epilogue of 16 consecutive output channels (CH0 .. CH0+15 of the slice) of one pixel: + (bias + time embedding) + residual,
// GroupNorm partial sums, one 32-byte store.  bb / rc: this thread's 48 additive constants / 48 residual values (16-bit pairs).
template <typename T, int CH0, int CPG, int GPS>
__device__ __forceinline__ void c3s_epi16(const uint32_t (&v)[16], const float2 (&bb)[24], const uint4 (&rc)[6], bool has_res, bool has_stats,
                                          float2 (&gs)[GPS], float2 (&gq)[GPS], T* dst) {
  uint32_t pk[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = CH0 + 2 * j;
    float2 r = tc::fadd2(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), bb[ch / 2]);
    if (has_res) {
      const uint4& q = rc[ch / 8];
      const uint32_t w = (ch % 8) / 2 == 0 ? q.x : (ch % 8) / 2 == 1 ? q.y : (ch % 8) / 2 == 2 ? q.z : q.w;
      r = tc::fadd2(r, tc::unpack2<T>(w));
    }
    if (has_stats) {
      const int g = ch / CPG;                         // both channels of a pair lie in the same group (CPG is even)
      gs[g] = tc::fadd2(gs[g], r);
      gq[g] = tc::ffma2(r, r, gq[g]);
    }
    pk[j] = tc::pack2<T>(r.x, r.y);
  }
  tc::st_global_v8(dst + CH0, make_uint4(pk[0], pk[1], pk[2], pk[3]), make_uint4(pk[4], pk[5], pk[6], pk[7]));
}

// KS0 / KS1: 16-channel k-steps of chunk 0 / chunk 1 (KS1 = 0: one chunk); R: ring rows; NSL: output slices of 48 (1 or 2);
// GN: GroupNorm + SiLU applied to the landed rows.
template <typename T, int KS0, int KS1, int R, int NSL, bool GN>
__global__ void __launch_bounds__(GN ? kSThreadsGN : kSThreads, 1)
k_conv3s(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB,
         const Conv3SP p) {
  constexpr int NCH = KS1 ? 2 : 1;
  constexpr int CPG = (48 * NSL) / 8;                       // channels per GroupNorm group of the WHOLE output (8 groups)
  constexpr int GPS = 8 / NSL;                              // groups per slice
  constexpr uint32_t ROW_BYTES = NCH * kSSlot;
  static_assert(R >= 3, "two rows per iteration must leave the producer a free slot");
  static_assert(NSL == 1 || NSL == 2, "48 or 96 output channels");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                        // [R][NCH][17 KB]
  uint8_t* sB = sA + (size_t)R * ROW_BYTES;                  // [NCH][3 dx][3 blocks][6 KB]
  uint64_t* bars = (uint64_t*)(sB + (size_t)NCH * 9 * kSBlk);
  uint64_t* r_full = bars;                 // [R]  TMA landed
  uint64_t* r_ready = bars + R;            // [R]  GN variant: transformed
  uint64_t* r_empty = bars + 2 * R;        // [R]  the MMAs that read the row have completed
  uint64_t* acc_full = bars + 3 * R;       // [NB]
  uint64_t* acc_empty = acc_full + kSNB;   // [NB]
  uint64_t* w_full = acc_empty + kSNB;
  uint64_t* baton = w_full + 1;            // [2]  baton[w]: the other issuing warp has issued (almost all of) its burst, warp w may issue
  uint32_t* tmem_slot = (uint32_t*)(baton + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = NSL == 1 ? 0 : (int)(blockIdx.x % NSL);
  const int wi = (int)blockIdx.x / NSL, nw = (int)gridDim.x / NSL;      // this CTA's worker index among the CTAs of its slice

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA0);
    tc::tma_prefetch_desc(&tmA1);
    tc::tma_prefetch_desc(&tmB);
    for (int s = 0; s < R; ++s) { tc::mbar_init(&r_full[s], 1); tc::mbar_init(&r_ready[s], kSXfBoth ? 256 : 128); tc::mbar_init(&r_empty[s], 1); }
    // acc_full: two arrivals per use -- a commit of the warp that issued the output's first row and one of the warp that issued its last
    for (int s = 0; s < kSNB; ++s) { tc::mbar_init(&acc_full[s], 2); tc::mbar_init(&acc_empty[s], 128); }
    tc::mbar_init(&baton[0], 1); tc::mbar_init(&baton[1], 1);
    tc::mbar_init(w_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, 512);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (*tmem_slot != 0u) {      // one CTA per SM (shared memory) and its only allocation: base 0 keeps MMA operands uniform
    if (threadIdx.x == 0) printf("libxrd: conv3s expects TMEM base 0, got %u\n", *tmem_slot);
    __trap();
  }

  if (warp < 4) {
    c3s_reg_dec<GN ? kSRegLowGN : kSRegLow>();
    if (warp == 0) {
      // ===================== TMA producer: one box per (input row, chunk) =====================
      if (tc::elect_one()) {
        // resident weights of this slice, restacked: slot (chunk, dx, block b) <- tap (dy = 2 - b, dx) of the chunk
        tc::mbar_expect_tx(w_full, (uint32_t)(NCH * 9) * kSBlk);
        for (int c = 0; c < NCH; ++c)
          for (int dx = 0; dx < 3; ++dx)
            for (int b = 0; b < 3; ++b)
              tc::tma_load_3d(sB + (size_t)((c * 3 + dx) * 3 + b) * kSBlk, &tmB, w_full, 0, slice * 48, c * 9 + (2 - b) * 3 + dx);
        uint32_t slot = 0, phase = 0;
        C3SItems it(p, wi, nw);
        int img, cb, r0, rows;
        while (it.next(p, img, cb, r0, rows)) {
          auto fetch = [&](int j, uint32_t sl, uint32_t ph) {
            tc::mbar_wait_relaxed(&r_empty[sl], ph ^ 1, 31);
            tc::mbar_expect_tx(&r_full[sl], (uint32_t)NCH * kSBox * 128u);
            uint8_t* dst = sA + (size_t)sl * ROW_BYTES;
            tc::tma_load_4d(dst, &tmA0, &r_full[sl], 0, cb * 128 - 1, r0 - 1 + j, img);      // rows / columns outside the image: zero fill
            if (NCH == 2) {
              if (p.two_src) tc::tma_load_4d(dst + kSSlot, &tmA1, &r_full[sl], 0, cb * 128 - 1, r0 - 1 + j, img);
              else tc::tma_load_4d(dst + kSSlot, &tmA0, &r_full[sl], 64, cb * 128 - 1, r0 - 1 + j, img);
            }
          };
          for (int j = 0; j < rows + 2; ++j) {
#ifdef XRD_C3S_RACE_TEST
            // Fault injection (tools/race_c3s.sh): every few rows the NEXT row is fetched first and this one several microseconds
            // later -- an exaggerated out-of-order completion of two TMA loads.  The row protocol must not care in which order rows land.
            if ((j % 5) == 1 && j + 1 < rows + 2) {
              uint32_t s2 = slot + 1, p2 = phase;
              if (s2 == R) { s2 = 0; p2 ^= 1; }
              fetch(j + 1, s2, p2);
              __nanosleep(6000);
              fetch(j, slot, phase);
              slot = s2; phase = p2; ++j;
              if (++slot == R) { slot = 0; phase ^= 1; }
              continue;
            }
#endif
            fetch(j, slot, phase);
            if (++slot == R) { slot = 0; phase ^= 1; }
          }
        }
      }
      __syncwarp();
    } else if (warp == 1 || warp == kSWarpB) {
      // ===================== MMA issuers (two warps, alternating iterations) =====================
      const uint32_t my = warp == 1 ? 0u : 1u;
      tc::mbar_wait(w_full, 0, 37);
      const uint32_t id48 = tc::umma_idesc(128, 48, tc::umma_fmt<T>());
      const uint64_t adesc0 = tc::umma_desc_sw128(tc::smem_u32(sA));
      const uint64_t bdesc0 = tc::umma_desc_sw128(tc::smem_u32(sB));
      C3SState st{0u, 0u, 0u, 0u};
      const C3SBars bs{GN ? r_ready : r_full, r_empty, acc_full, acc_empty, baton};
      C3SItems it(p, wi, nw);
      int img, cb, r0, rows;
      while (it.next(p, img, cb, r0, rows)) c3s_run_item<KS0, KS1, R, (GN && R < 4)>(st, bs, rows, adesc0, bdesc0, id48, ROW_BYTES >> 4, lane, my);
    }
  } else if (GN && warp >= 12) {
    // ===================== input transform (warps 12..19): a = SiLU(GroupNorm(x)) in place, once per landed row =====================
    // two warpgroups, each taking every other landed row
    // thread = (16-byte chunk j of the valid channels, pixel lane); logical chunk j of ring pixel sp sits at physical
    // chunk j ^ (sp & 7) (128B swizzle on absolute addresses; slots are 1024-aligned).  Out-of-image pixels stay zero.
    // Arithmetic on packed fp32 pairs (FFMA2): x*sigmoid(x) = h*tanh(h) + h with h = x/2 folded into the coefficients.
    c3s_reg_dec<kSRegXf>();
    const uint32_t xw = kSXfBoth ? 0u : (uint32_t)(warp - 12) >> 2;
    const int tt = threadIdx.x - 384 - (int)xw * 128;
    // a thread works on ONE 16-byte channel chunk (its 8 coefficient pairs stay in registers) of every pixel it visits.
    // Quarter-warp q (8 consecutive threads: one shared-memory wavefront of a 16-byte access) = chunk v of EIGHT CONSECUTIVE
    // pixels: their physical chunks v ^ (pixel & 7) are the eight different 16-byte bank groups, so every LDS.128 / STS.128 is one
    // wavefront.  (The first mapping put the chunks of one pixel on consecutive threads; with 6 of 8 chunks valid a quarter-warp
    // straddled two pixels whose swizzled chunks collided: ncu counted 510 shared-memory wavefronts per row for 196 ideal, on a
    // pipe the N=144 MMAs alone keep ~95 % busy.)  With NVT chunks per pixel the 16 quarter-warps of the warpgroup form
    // OL = 16 / NVT octet lanes (48 channels: 2 lanes, the fourth warp idles; two chunks of 48: one lane).
    constexpr int NV0 = KS0 * 2, NV1 = KS1 * 2, NVT = NV0 + NV1, OL = (kSXfBoth ? 32 : 16) / NVT;
    constexpr int NOCT = (kSBox + 7) / 8, SLOTS = (NOCT + OL - 1) / OL;       // pixel octets per row; octets a thread visits
    constexpr int U = SLOTS <= 5 ? SLOTS : 3;                                  // pixels in flight per thread
    const int q = tt >> 3, l8 = tt & 7;
    const int v = q % NVT, ol = q / NVT;
    const bool active = q < OL * NVT;
    const int cm = (NCH == 2 && v >= NV0) ? 1 : 0;
    const int j8 = cm ? v - NV0 : v;
    // Everything that does not depend on the row is computed once: the thread's byte offset inside a row slot (pixel ol*8 + l8 of
    // the first octet it visits, its physical 16-byte chunk j8 ^ l8 -- every octet starts at a multiple of 8 pixels, so the swizzle
    // term is the same for all of them; later octets are +OL KB, an immediate of the load), the coefficient pairs as 64-bit
    // operands of the packed FMAs, and -- per work item -- one bit per visited octet saying whether the pixel is inside the image
    // and the box.  (SASS of the version that recomputed addresses and predicates per pixel and passed the coefficients as scalar
    // floats: 230 instructions per body of which 120 were loads, conversions, FMAs, MUFU and stores -- 68 MOVs re-pairing the
    // coefficient registers for every pixel, 40 address / predicate instructions; ncu: IPC 0.63 per scheduler, the transform's
    // 64 M warp instructions against 45 M of the whole plain kernel.)
    uint64_t sc[4], sh[4];
    const uint32_t toff = (uint32_t)((ol * 8 + l8) * 128) + (uint32_t)((j8 ^ l8) << 4) + (uint32_t)cm * kSSlot;
    int cur_img = -1;
    uint32_t slot = 0, phase = 0, nrow = 0;
    C3SItems it(p, wi, nw);
    int img, cb, r0, rows;
    while (it.next(p, img, cb, r0, rows)) {
      if (img != cur_img) {
        const float2* cf = p.in_coef + (size_t)img * (p.c0 + p.c1) + (cm ? p.c0 : 0) + j8 * 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 a = __ldg(cf + 2 * i), b = __ldg(cf + 2 * i + 1);
          // "+ 0" makes each pair the RESULT of a packed instruction, i.e. a value living in its own aligned register pair; a bare
          // mov.b64 {a.x, b.x} is rematerialised by ptxas in front of every use (two MOVs per packed operand per pixel)
          sc[i] = tc::fadd2_64(tc::pack_f32x2(a.x, b.x), 0ull); sh[i] = tc::fadd2_64(tc::pack_f32x2(a.y, b.y), 0ull);
        }
      }
      cur_img = img;
      uint32_t okmask = 0;                  // bit s: the pixel of octet ol + OL*s this thread visits is inside the box and the image
      if (active) {
#pragma unroll
        for (int sx = 0; sx < SLOTS; ++sx) {
          const int cc = (ol + OL * sx) * 8 + l8, iw = cb * 128 - 1 + cc;
          if (cc < kSBox && iw >= 0 && iw < p.W) okmask |= 1u << sx;
        }
      }
      for (int j = 0; j < rows + 2; ++j, ++nrow) {
        if (!kSXfBoth && (nrow & 1u) != xw) {
          // A parity wait can only tell "the current phase" from "the one before".  With an ODD ring two warpgroups taking alternate
          // rows alternate on every slot too, so a group that simply skipped the other group's rows would see only every other
          // phase of r_full[slot]: its next wait on the slot passes on the STALE phase if the other group's row has not landed
          // yet (TMA loads complete out of order under load), the group then runs ahead of the producer through every later
          // wait and its arrivals scramble r_ready (seen once in ~1e5 launches as an MMA-warp time-out on r_ready).  It therefore
          // observes the skipped row's landing too.  (Default build: both groups work on every row and nothing is skipped.)
#ifndef XRD_C3S_RACE_NOFIX
          if (R & 1) tc::mbar_wait_relaxed(&r_full[slot], phase, 38);
#endif
          if (++slot == R) { slot = 0; phase ^= 1; }
          continue;
        }
        tc::mbar_wait_relaxed(&r_full[slot], phase, 35);
        const int ih = r0 - 1 + j;
        if (ih >= 0 && ih < p.H && okmask) {
          const uint32_t sbase = tc::smem_u32(sA + (size_t)slot * ROW_BYTES) + toff;
#pragma unroll
          for (int s0 = 0; s0 < SLOTS; s0 += U) {
            // straight-line code for the U pixels in flight (loads and arithmetic unconditional, only the store is predicated; an
            // octet past the box reads whatever follows in shared memory and drops it): with a branch per pixel the compiler ran
            // the dependent chains LDS -> cvt -> FFMA2 -> MUFU -> FFMA2 -> cvt -> STS one after the other
            uint32_t w[U][4];
#pragma unroll
            for (int u = 0; u < U; ++u)
              if (s0 + u < SLOTS)
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[u][0]), "=r"(w[u][1]), "=r"(w[u][2]), "=r"(w[u][3])
                             : "r"(sbase + (uint32_t)((s0 + u) * OL * 1024)));      // base + constant: folded into the instruction's immediate
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
              for (int u = 0; u < U; ++u) {
                if (s0 + u >= SLOTS) continue;
                const uint64_t h = tc::ffma2_64(tc::pack_f32x2(tc::unpack2<T>(w[u][i])), sc[i], sh[i]);    // 0.5 * GroupNorm(x)
                const float2 hf = tc::unpack_f32x2(h);
                float2 th;
                asm("tanh.approx.f32 %0, %1;" : "=f"(th.x) : "f"(hf.x));
                asm("tanh.approx.f32 %0, %1;" : "=f"(th.y) : "f"(hf.y));
                const float2 o = tc::unpack_f32x2(tc::ffma2_64(h, tc::pack_f32x2(th), h));
                w[u][i] = tc::pack2<T>(o.x, o.y);
              }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
              if (s0 + u < SLOTS && (okmask >> (s0 + u) & 1u))
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbase + (uint32_t)((s0 + u) * OL * 1024)), "r"(w[u][0]), "r"(w[u][1]),
                             "r"(w[u][2]), "r"(w[u][3]) : "memory");
          }
        }
        tc::fence_async_smem();
        tc::mbar_arrive(&r_ready[slot]);
        if (++slot == R) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp < 12) {
    // ===================== epilogue (warps 4..11): group g drains the output rows whose running index is == g mod 2 =====================
    c3s_reg_inc<GN ? kSRegEpiGN : kSRegEpi>();
    const int quad = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int CT = p.cout_total;
    T* yp = (T*)p.y + slice * 48;
    const T* rp = p.resid ? (const T*)p.resid + slice * 48 : nullptr;
    const bool has_res = rp != nullptr, has_stats = p.stats != nullptr;
    // this thread's pixel gets the same 48 additive constants (bias + time-embedding row of the image) in every row: registers
    float2 bb[24];
    // GroupNorm partial sums of this thread's pixels, two lanes per group (even / odd channel of each pair: CPG is even)
    float2 gs[GPS], gq[GPS];
#pragma unroll
    for (int g = 0; g < GPS; ++g) { gs[g] = make_float2(0.f, 0.f); gq[g] = make_float2(0.f, 0.f); }
    auto flush_stats = [&](int img) {
      if (!has_stats || img < 0) return;
      float s1[GPS], s2[GPS];
#pragma unroll
      for (int g = 0; g < GPS; ++g) {
        s1[g] = gs[g].x + gs[g].y; s2[g] = gq[g].x + gq[g].y;
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) {
          s1[g] += __shfl_xor_sync(0xffffffffu, s1[g], of);
          s2[g] += __shfl_xor_sync(0xffffffffu, s2[g], of);
        }
      }
      if (lane < 2 * GPS) {
        float v = 0.f;
#pragma unroll
        for (int g = 0; g < GPS; ++g) { if (lane == 2 * g) v = s1[g]; if (lane == 2 * g + 1) v = s2[g]; }
        atomicAdd(p.stats + (size_t)img * 16 + slice * 2 * GPS + lane, (double)v);
      }
#pragma unroll
      for (int g = 0; g < GPS; ++g) { gs[g] = make_float2(0.f, 0.f); gq[g] = make_float2(0.f, 0.f); }
    };
    uint32_t o = 0;
    int cur_img = -1;
    C3SItems it(p, wi, nw);
    int img, cb, r0, rows;
    while (it.next(p, img, cb, r0, rows)) {
      if (img != cur_img) {
        flush_stats(cur_img);
#pragma unroll
        for (int j = 0; j < 24; ++j) {
          float2 b = make_float2(0.f, 0.f);
          if (p.bias) { b.x = __ldg(p.bias + slice * 48 + 2 * j); b.y = __ldg(p.bias + slice * 48 + 2 * j + 1); }
          if (p.chan_add) {
            const float* ca = p.chan_add + (int64_t)img * p.chan_add_bstride + slice * 48 + 2 * j;
            b.x += __ldg(ca); b.y += __ldg(ca + 1);
          }
          bb[j] = b;
        }
        cur_img = img;
      }
      for (int r = 0; r < rows; ++r, ++o) {
        if ((int)(o & 1) != grp) continue;
        const uint32_t a = o % kSNB, use = o / kSNB;
        const bool row_ok = r0 + r < p.H;                    // rows past the image end (last decade) are computed and dropped
        const int64_t pix = ((int64_t)img * p.H + r0 + r) * p.W + cb * 128 + quad * 32 + lane;
        uint4 rc[6];
        if (has_res && row_ok) {
#pragma unroll
          for (int j = 0; j < 6; ++j) rc[j] = __ldg(reinterpret_cast<const uint4*>(rp + pix * CT) + j);
        }
        tc::mbar_wait_relaxed(&acc_full[a], use & 1, 36);
        tc::tc_fence_after();
        if (!row_ok) {                                       // nothing to read: hand the block straight back
          tc::tc_fence_before();
          tc::mbar_arrive(&acc_empty[a]);
          continue;
        }
        const uint32_t tacc = a * 48u + ((uint32_t)(quad * 32) << 16);
        T* dst = yp + pix * CT;
        // 16 columns at a time, the next load in flight under the arithmetic of the previous one
        uint32_t va[16], vb[16];
        s3_tmem_ld16(tacc, va);
        s3_tmem_ld16(tacc + 16u, vb);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        c3s_epi16<T, 0, CPG, GPS>(va, bb, rc, has_res, has_stats, gs, gq, dst);
        s3_tmem_ld16(tacc + 32u, va);
        c3s_epi16<T, 16, CPG, GPS>(vb, bb, rc, has_res, has_stats, gs, gq, dst);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        // the block is in registers: hand it back to the MMA warps before the last third of the arithmetic and its store
        tc::tc_fence_before();
        tc::mbar_arrive(&acc_empty[a]);
        c3s_epi16<T, 32, CPG, GPS>(va, bb, rc, has_res, has_stats, gs, gq, dst);
      }
    }
    flush_stats(cur_img);
  }
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(0u, 512);
  }
}

static int c3s_env(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

bool conv3s_supported(const Tens& x1, const Tens* x2, const ConvW& w, const ConvEpi& e) {
  static const int enabled = c3s_env("XRD_CONV3S", 1);
  if (!enabled) return false;
  if (x1.dt == DT_F32) return false;
  if (!(w.kh == 3 && w.kw == 3 && w.stride == 1 && w.pad == 1) || w.d2s) return false;
  if (!(w.cout == 48 || w.cout == 96)) return false;
  if (x1.w % 128 != 0) return false;
  if (e.in_scale || e.out_scale || e.act != ACT_NONE || e.gate) return false;
  if (e.in_coef && e.in_act != ACT_SILU) return false;
  // (k-steps of chunk 0, of chunk 1) the kernel is instantiated for
  int k0, k1;
  if (x2) {
    if (x1.c % 16 || x2->c % 16 || x1.c > 64 || x2->c > 64) return false;
    k0 = x1.c / 16; k1 = x2->c / 16;
  } else {
    if (x1.c % 16 || x1.c > 128) return false;
    k0 = std::min(x1.c, 64) / 16; k1 = (x1.c - std::min(x1.c, 64)) / 16;
  }
  const bool one = k1 == 0 && k0 >= 1 && k0 <= 4;
  const bool two = (k0 == 3 && k1 == 3) || (k0 == 4 && k1 == 2);
  if (!one && !two) return false;
  if (w.cout == 96 && !((k0 == 3 && k1 == 0) || (k0 == 4 && k1 == 2) || (k0 == 3 && k1 == 3))) return false;
  return true;
}

void conv3s(Ctx& c, const Tens& x1, const Tens* x2, ConvW& w, const ConvEpi& e, Tens& y) {
  XRD_REQUIRE(conv3s_supported(x1, x2, w, e), "conv3s: unsupported configuration");
  const int cin = x1.c + (x2 ? x2->c : 0);
  XRD_REQUIRE(cin == w.cin && y.n == x1.n && y.h == x1.h && y.w == x1.w && y.c == w.cout && y.dt == x1.dt, "conv3s: shape mismatch");
  if (x2) XRD_REQUIRE(x2->n == x1.n && x2->h == x1.h && x2->w == x1.w && x2->dt == x1.dt, "conv3s: source mismatch");
  if (e.resid.p) XRD_REQUIRE(e.resid.dt == y.dt && e.resid.numel() == y.numel(), "conv3s: residual mismatch");
  if (c.dry) return;
  // packed weights [kblock = chunk*9 + tap][cout][64]: the chunk boundary is x1's channel count for a concat, 64 otherwise
  const int c1pack = x2 ? x1.c : x1.c;
  if (!w.wtc[x1.dt] || w.tc_c1 != c1pack) conv_tc_pack(c.s, w, x1.dt, c1pack);
  Conv3SP p;
  p.H = x1.h; p.W = x1.w; p.nimg = x1.n;
  p.cout_total = w.cout;
  const int nsl = w.cout / 48;
  if (x2) { p.c0 = x1.c; p.c1 = x2->c; p.two_src = 1; }
  else { p.c0 = std::min(x1.c, 64); p.c1 = x1.c - p.c0; p.two_src = 0; }
  const int nch = p.c1 ? 2 : 1;
  XRD_REQUIRE(w.tc_nkb == nch * 9 && w.tc_npad == w.cout, "conv3s: packed weights out of date (nkb %d, npad %d)", w.tc_nkb, w.tc_npad);
  const int nsm = sm_count();
  p.ncb = x1.w / 128;
  p.dpc = cdiv(x1.h, kSNB);
  p.ndec = x1.n * p.ncb * p.dpc;
  p.bias = w.bias;
  p.chan_add = e.chan_add; p.chan_add_bstride = e.chan_add_bstride;
  p.resid = e.resid.p; p.y = y.p;
  p.stats = e.stats_out;
  p.in_coef = e.in_coef;

  alignas(64) CUtensorMap tmA0, tmA1, tmB;
  auto enc_act = [&](CUtensorMap* m, const Tens& x) {
    const cuuint64_t dims[4] = {(cuuint64_t)x.c, (cuuint64_t)x.w, (cuuint64_t)x.h, (cuuint64_t)x.n};
    const cuuint64_t strides[3] = {(cuuint64_t)x.c * 2, (cuuint64_t)x.w * x.c * 2, (cuuint64_t)x.h * x.w * x.c * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)kSBox, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = get_encode_tiled()(m, tmap_dtype(x.dt), 4, x.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail(XRD_ERR_CUDA, "cuTensorMapEncodeTiled(conv3s activations) failed: %d", (int)r);
  };
  enc_act(&tmA0, x1);
  if (x2) enc_act(&tmA1, *x2); else tmA1 = tmA0;
  {
    const cuuint64_t dims[3] = {64, (cuuint64_t)w.cout, (cuuint64_t)(nch * 9)};
    const cuuint64_t strides[2] = {128, (cuuint64_t)w.cout * 128};
    const cuuint32_t box[3] = {64, 48, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = get_encode_tiled()(&tmB, tmap_dtype(x1.dt), 3, w.wtc[x1.dt], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail(XRD_ERR_CUDA, "cuTensorMapEncodeTiled(conv3s weights) failed: %d", (int)r);
  }
  const int R = nch == 1 ? 8 : 3;
  const size_t smem = 1024 + (size_t)R * nch * kSSlot + (size_t)nch * 9 * kSBlk + (3 * R + 2 * kSNB + 3) * 8 + 64;
  int grid = std::min(p.ndec * nsl, (nsm / nsl) * nsl);
  grid = std::max(nsl, (grid / nsl) * nsl);
  const bool gn = e.in_coef != nullptr;
  const int k0 = p.c0 / 16, k1 = p.c1 / 16;
  auto launch = [&](auto kern) {
    ensure_dyn_smem(kern, 227 * 1024);
    XRD_LAUNCH(c, kern, grid, gn ? kSThreadsGN : kSThreads, smem, tmA0, tmA1, tmB, p);
  };
  auto pick = [&](auto tag) {
    using T = decltype(tag);
#define C3S_CASE(K0, K1, RR, NS)                                                                                       \
    if (k0 == K0 && k1 == K1 && nsl == NS) {                                                                           \
      if (gn) launch(k_conv3s<T, K0, K1, RR, NS, true>); else launch(k_conv3s<T, K0, K1, RR, NS, false>);               \
      return;                                                                                                          \
    }
    C3S_CASE(1, 0, 8, 1) C3S_CASE(2, 0, 8, 1) C3S_CASE(3, 0, 8, 1) C3S_CASE(4, 0, 8, 1)
    C3S_CASE(3, 3, 3, 1) C3S_CASE(4, 2, 3, 1)
    C3S_CASE(3, 0, 8, 2) C3S_CASE(3, 3, 3, 2) C3S_CASE(4, 2, 3, 2)
#undef C3S_CASE
    fail(XRD_ERR_INVALID, "conv3s: no kernel for k-steps (%d,%d), %d slices", k0, k1, nsl);
  };
  if (x1.dt == DT_BF16) pick(__nv_bfloat16()); else pick(__half());
}

}  // namespace xrd
