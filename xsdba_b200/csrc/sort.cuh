// Register-blocked column sorter for the fast train / rank kernels (float32, 32 columns per CTA).
//
// Layout: buf[row][32] floats in shared memory, column = lane, so every access of a warp hits 32
// different banks whatever the rows are (rows are warp-uniform).  The 2^(NB+1) rows are two halves of
// 2^NB rows; each half is sorted ascending by a bitonic network, and the caller then selects order
// statistics from the two sorted runs by a merge-path binary search (two_run_pair below) -- this
// skips the last, most expensive merge phase (NB+1 stages over every element) for the ~100 order
// statistics the quantile step needs.
//
// The network runs in "passes".  In one pass a thread owns the 32 elements of its column whose
// in-half index differs only in 5 chosen bits S = {s0..s4}: it loads them into registers, applies
// every consecutive stage of the network whose exchange bit lies in S, and stores them back.  With
// 1024 threads = 32 warps x 32 lanes a CTA covers 2 halves x 16 blocks x 32 columns per pass, i.e.
// one block per thread, and the 45 stages of a 512-row half take 7 passes (7 LDS + 7 STS + 45 FMNMX
// per element instead of 2 LDS + 2 STS + 2 FMNMX per element per stage).
//
// Direction of a stage of phase p (merging runs of 2^p) is given by bit p of the element index:
// if that bit is in S it is a compile-time function of the register index, otherwise it is
// warp-uniform and selected by one branch per pass (template parameter DESC).  Phase NB is always
// ascending.
#pragma once
#include <cuda_runtime.h>

#ifndef XS_CE_IMAD
#define XS_CE_IMAD 1
#endif

namespace xsdba {

__constant__ unsigned xs_ce_consts[2] = {1u, 0xffffffffu};

template <int B0, int B1, int B2, int B3, int B4>
struct BitSet {
  static constexpr int b[5] = {B0, B1, B2, B3, B4};
  static constexpr int mask = (1 << B0) | (1 << B1) | (1 << B2) | (1 << B3) | (1 << B4);
  // register-index bit that controls element bit `bit`, or -1
  static __host__ __device__ constexpr int pos(int bit) {
    return bit == B0 ? 0 : bit == B1 ? 1 : bit == B2 ? 2 : bit == B3 ? 3 : bit == B4 ? 4 : -1;
  }
  // element-index offset of register i
  static __host__ __device__ constexpr int off(int i) {
    return ((i & 1) << B0) | (((i >> 1) & 1) << B1) | (((i >> 2) & 1) << B2) | (((i >> 3) & 1) << B3) |
           (((i >> 4) & 1) << B4);
  }
  // spread the low bits of v over the element bits NOT in S (ascending), NB bits in total
  template <int NB>
  static __device__ __forceinline__ int deposit(int v) {
    int e = 0, k = 0;
#pragma unroll
    for (int bit = 0; bit < NB; ++bit) {
      if (!((mask >> bit) & 1)) { e |= ((v >> k) & 1) << bit; ++k; }
    }
    return e;
  }
};

// one compare-exchange; `imad` selects the FMA-pipe reconstruction of the max (see sort_stage)
template <bool DESC>
__device__ __forceinline__ void constexpr_ce(float& a, float& b, bool imad, unsigned ce_one, unsigned ce_mone) {
  const float lo = fminf(a, b);
  float hi;
#if XS_CE_IMAD
  if (imad) {
    unsigned s_, h_;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(s_) : "r"(__float_as_uint(a)), "r"(ce_one), "r"(__float_as_uint(b)));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(h_) : "r"(__float_as_uint(lo)), "r"(ce_mone), "r"(s_));
    hi = __uint_as_float(h_);
  } else
#endif
  hi = fmaxf(a, b);
  a = DESC ? hi : lo;
  b = DESC ? lo : hi;
}

// Batcher's odd-even merge sort of 32 inputs: 191 compare-exchanges (depth 15) against 240 for the
// bitonic phases 1..5 -- used for the first pass, where a thread sorts its whole 32-element block.
// (generated and 0-1 verified offline; pairs are (lo index, hi index), applied in this order)
__device__ constexpr unsigned char kOem32A[191] = {0,2,0,1,1,4,6,4,5,5,0,2,2,1,3,3,1,3,5,8,10,8,9,9,12,14,12,13,13,8,10,10,9,11,11,9,11,13,0,4,4,2,6,6,2,6,10,1,5,5,3,7,7,3,7,11,1,3,5,7,9,11,13,16,18,16,17,17,20,22,20,21,21,16,18,18,17,19,19,17,19,21,24,26,24,25,25,28,30,28,29,29,24,26,26,25,27,27,25,27,29,16,20,20,18,22,22,18,22,26,17,21,21,19,23,23,19,23,27,17,19,21,23,25,27,29,0,8,8,4,12,12,4,12,20,2,10,10,6,14,14,6,14,22,2,6,10,14,18,22,26,1,9,9,5,13,13,5,13,21,3,11,11,7,15,15,7,15,23,3,7,11,15,19,23,27,1,3,5,7,9,11,13,15,17,19,21,23,25,27,29};
__device__ constexpr unsigned char kOem32B[191] = {1,3,2,3,2,5,7,6,7,6,4,6,4,5,7,5,2,4,6,9,11,10,11,10,13,15,14,15,14,12,14,12,13,15,13,10,12,14,8,12,8,10,14,10,4,8,12,9,13,9,11,15,11,5,9,13,2,4,6,8,10,12,14,17,19,18,19,18,21,23,22,23,22,20,22,20,21,23,21,18,20,22,25,27,26,27,26,29,31,30,31,30,28,30,28,29,31,29,26,28,30,24,28,24,26,30,26,20,24,28,25,29,25,27,31,27,21,25,29,18,20,22,24,26,28,30,16,24,16,20,28,20,8,16,24,18,26,18,22,30,22,10,18,26,4,8,12,16,20,24,28,17,25,17,21,29,21,9,17,25,19,27,19,23,31,23,11,19,27,5,9,13,17,21,25,29,2,4,6,8,10,12,14,16,18,20,22,24,26,28,30};

template <bool DESC>
__device__ __forceinline__ void sort32_oem(float (&r)[32], unsigned ce_one, unsigned ce_mone) {
#pragma unroll
  for (int k = 0; k < 191; ++k) {
    constexpr_ce<DESC>(r[kOem32A[k]], r[kOem32B[k]], (k % 3) != 0, ce_one, ce_mone);
  }
}

// one stage (phase, exchange bit) on the 32 registers of a thread
template <class S, bool DESC, int TOP, int PHASE, int BIT>
__device__ __forceinline__ void sort_stage(float (&r)[32], unsigned ce_one, unsigned ce_mone) {
  constexpr int pos = S::pos(BIT);
  constexpr int ppos = S::pos(PHASE);
  static_assert(pos >= 0, "exchange bit must be in the pass's bit set");
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    if (i & (1 << pos)) continue;
    const int j = i | (1 << pos);
    const bool desc = (PHASE >= TOP) ? false : (ppos >= 0 ? (((i >> ppos) & 1) != 0) : DESC);
    const float lo = fminf(r[i], r[j]);
    float hi;
#if XS_CE_IMAD
    // Pipe balancing (profiles/microbench_ce.cu): FMNMX runs on the half-rate ALU pipe, which bounds the
    // sorter.  For two CEs out of three the max is rebuilt on the FMA pipe instead: lo is bit-identical
    // to one of the inputs, so hi = a + b - lo in integer arithmetic is the other one, exactly.  The
    // multipliers (1, -1) come from constant memory so that ptxas keeps two IMADs (an IADD3 would go
    // back to the ALU pipe).
    if ((i + (i >> pos)) % 3 != 0) {
      unsigned s_, h_;
      asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(s_) : "r"(__float_as_uint(r[i])), "r"(ce_one), "r"(__float_as_uint(r[j])));
      asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(h_) : "r"(__float_as_uint(lo)), "r"(ce_mone), "r"(s_));
      hi = __uint_as_float(h_);
    } else
#endif
    hi = fmaxf(r[i], r[j]);
    r[i] = desc ? hi : lo;
    r[j] = desc ? lo : hi;
  }
}

// stage lists are encoded as ints PHASE*16 + BIT
template <class S, bool DESC, int TOP, int... ST>
__device__ __forceinline__ void sort_stages(float (&r)[32], unsigned ce_one, unsigned ce_mone) {
  (sort_stage<S, DESC, TOP, ST / 16, ST % 16>(r, ce_one, ce_mone), ...);
}

// One pass.  NB = log2(rows per half); RT_PHASE = the phase whose direction bit is warp-uniform in this
// pass (-1 if none).  half_rows_base points at buf[half * 2^NB][lane].
template <class S, int NB, int RT_PHASE, int... ST>
__device__ __forceinline__ void sort_pass(float* __restrict__ half_col, int block_idx) {
  const int ebase = S::template deposit<NB>(block_idx);
  float* p = half_col + ebase * 32;
  float r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = p[S::off(i) * 32];
  bool desc = false;
  if (RT_PHASE >= 0 && RT_PHASE < NB) desc = ((ebase >> (RT_PHASE < 0 ? 0 : RT_PHASE)) & 1) != 0;
  const unsigned ce_one = xs_ce_consts[0], ce_mone = xs_ce_consts[1];
  if (desc) sort_stages<S, true, NB, ST...>(r, ce_one, ce_mone);
  else sort_stages<S, false, NB, ST...>(r, ce_one, ce_mone);
#pragma unroll
  for (int i = 0; i < 32; ++i) p[S::off(i) * 32] = r[i];
}

#define XS_ST(p, b) ((p) * 16 + (b))

__device__ __forceinline__ void group_barrier(int id, int n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

// Sort both 512-row halves of buf[1024][32] ascending.  blockDim.x == 1024; ends with __syncthreads.
// Synchronisation is as fine as the data flow allows: up to and including phase 8 every pass stays
// inside a 256-row quarter (bit 8 of the in-half index is never an exchange bit), so the 8 warps of a
// quarter only wait for each other (named barriers 1..4); the two passes that touch phase 9 need the
// 16 warps of the half (barriers 5, 6).  Four independent groups per CTA drift apart, so the
// shared-memory phase of one overlaps the FMNMX / IMAD phase of another.  stagger_ns > 0 delays group
// k by k*stagger_ns at the start to force the phase offset.
__device__ __forceinline__ void sort_halves_512(float* buf, int stagger_ns = 0) {
  constexpr int NB = 9;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hb = warp >> 4;          // half (0, 1)
  const int qb = warp >> 3;          // quarter group (0..3): half*2 + bit 8 of the block base
  float* hc = buf + (size_t)hb * (1 << NB) * 32 + lane;
  const int blk = warp & 15;
  if (stagger_ns > 0 && qb > 0) __nanosleep((unsigned)(stagger_ns * qb));
  // pass 1: every thread sorts its 32-row block (bits {0..4}); ascending or descending by bit 5 of the block
  // (= what bitonic phases 1..5 would leave), with the 191-CE odd-even merge network
  {
    const int ebase = BitSet<0, 1, 2, 3, 4>::template deposit<NB>(blk);
    float* p = hc + ebase * 32;
    float r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = p[i * 32];
    const unsigned ce_one = xs_ce_consts[0], ce_mone = xs_ce_consts[1];
    if ((ebase >> 5) & 1) sort32_oem<true>(r, ce_one, ce_mone);
    else sort32_oem<false>(r, ce_one, ce_mone);
#pragma unroll
    for (int i = 0; i < 32; ++i) p[i * 32] = r[i];
  }
  group_barrier(1 + qb, 256);
  // pass 2: phase 6 bits 5..1 (direction bit 6 warp-uniform)
  sort_pass<BitSet<1, 2, 3, 4, 5>, NB, 6, XS_ST(6, 5), XS_ST(6, 4), XS_ST(6, 3), XS_ST(6, 2), XS_ST(6, 1)>(hc, blk);
  group_barrier(1 + qb, 256);
  // pass 3: phase 6 bit 0 (direction bit 6 in S), phase 7 bits 6..3 (direction bit 7 warp-uniform)
  sort_pass<BitSet<0, 3, 4, 5, 6>, NB, 7, XS_ST(6, 0), XS_ST(7, 6), XS_ST(7, 5), XS_ST(7, 4), XS_ST(7, 3)>(hc, blk);
  group_barrier(1 + qb, 256);
  // pass 4: phase 7 bits 2..0 (direction bit 7 in S), phase 8 bits 7,6 (direction bit 8 warp-uniform)
  sort_pass<BitSet<0, 1, 2, 6, 7>, NB, 8, XS_ST(7, 2), XS_ST(7, 1), XS_ST(7, 0), XS_ST(8, 7), XS_ST(8, 6)>(hc, blk);
  group_barrier(1 + qb, 256);
  // pass 5: phase 8 bits 5..1 (direction bit 8 warp-uniform)
  sort_pass<BitSet<1, 2, 3, 4, 5>, NB, 8, XS_ST(8, 5), XS_ST(8, 4), XS_ST(8, 3), XS_ST(8, 2), XS_ST(8, 1)>(hc, blk);
  group_barrier(5 + hb, 512);
  // pass 6: phase 8 bit 0 (direction bit 8 in S), phase 9 bits 8..5 (ascending)
  sort_pass<BitSet<0, 5, 6, 7, 8>, NB, -1, XS_ST(8, 0), XS_ST(9, 8), XS_ST(9, 7), XS_ST(9, 6), XS_ST(9, 5)>(hc, blk);
  group_barrier(5 + hb, 512);
  // pass 7: phase 9 bits 4..0 (ascending)
  sort_pass<BitSet<0, 1, 2, 3, 4>, NB, -1, XS_ST(9, 4), XS_ST(9, 3), XS_ST(9, 2), XS_ST(9, 1), XS_ST(9, 0)>(hc, blk);
  __syncthreads();
}

// Order statistics i and i+1 (0-based) of the union of two ascending runs A[0..nA), B[0..nB) of one
// column (element k of a run at run[k*32]).  Requires 0 <= i < nA + nB.  v1 = +inf if i+1 == nA+nB.
__device__ __forceinline__ void two_run_pair(const float* __restrict__ A, int nA, const float* __restrict__ B, int nB,
                                             int i, float& v0, float& v1) {
  const int k = i + 1;  // number of elements <= the wanted one
  int lo = k - nB > 0 ? k - nB : 0;
  int hi = k < nA ? k : nA;
  while (lo < hi) {  // smallest a with A[a] >= B[k-a-1]  (a = how many of the k come from A)
    const int mid = (lo + hi) >> 1;
    if (A[mid * 32] < B[(k - mid - 1) * 32]) lo = mid + 1; else hi = mid;
  }
  const int a = lo, b = k - lo;
  const float inf = __int_as_float(0x7f800000);
  const float a_prev = a > 0 ? A[(a - 1) * 32] : -inf;
  const float b_prev = b > 0 ? B[(b - 1) * 32] : -inf;
  const float a_next = a < nA ? A[a * 32] : inf;
  const float b_next = b < nB ? B[b * 32] : inf;
  v0 = fmaxf(a_prev, b_prev);
  v1 = fminf(a_next, b_next);
}

// single order statistic
__device__ __forceinline__ float two_run_at(const float* A, int nA, const float* B, int nB, int i) {
  float v0, v1;
  two_run_pair(A, nA, B, nB, i, v0, v1);
  return v0;
}

}  // namespace xsdba
