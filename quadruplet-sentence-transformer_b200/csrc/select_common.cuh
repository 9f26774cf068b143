// Warp-level selection shared by K2's epilogue (score_select.cu) and the per-row candidate selection of
// the sharded path (finalize.cu).
#pragma once
#include "qst_common.cuh"

namespace qst {

// ------------------------------------------------------------------------------------------
// Warp-cooperative exact selection of the k-th largest key of buf[0..n) (keys in .x) followed by
// an in-place compaction that keeps exactly k entries (all keys > T and enough == T).
// Returns T.  hist: 256 ints of warp-private shared memory.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_select_compact(uint2* __restrict__ buf, int n, int k, int* hist, int lane) {
  uint32_t prefix = 0, mask = 0;
  int krem = k;
#pragma unroll 1
  for (int shift = 24; shift >= 0; shift -= 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) hist[lane * 8 + i] = 0;
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
      const uint32_t key = buf[i].x;
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1);
    }
    __syncwarp();
    int c[8];
    int ls = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i] = hist[lane * 8 + i]; ls += c[i]; }
    // above = number of matching keys in bins owned by higher lanes (inclusive suffix - own)
    int incl = ls;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_down_sync(0xffffffffu, incl, o);
      if (lane + o < 32) incl += t;
    }
    const int above = incl - ls;
    const bool mine = (above < krem) && (krem <= above + ls);
    const unsigned who = __ballot_sync(0xffffffffu, mine);
    const int src = 31 - __clz(who);  // exactly one lane satisfies it when n >= krem
    int digit = 0, cnt_above = 0;
    if (lane == src) {
      int run = above;
#pragma unroll
      for (int i = 7; i >= 0; --i) {
        if (run < krem && krem <= run + c[i]) { digit = lane * 8 + i; cnt_above = run; }
        run += c[i];
      }
    }
    digit = __shfl_sync(0xffffffffu, digit, src);
    cnt_above = __shfl_sync(0xffffffffu, cnt_above, src);
    krem -= cnt_above;
    prefix |= (uint32_t)digit << shift;
    mask |= 255u << shift;
    __syncwarp();
  }
  const uint32_t T = prefix;
  // in-place stable compaction, 32 entries per step (reads of a step precede its writes)
  int w = 0, eq_taken = 0;
  const unsigned lt = (1u << lane) - 1u;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    uint2 e = make_uint2(0u, 0u);
    if (i < n) e = buf[i];
    const bool gt = (i < n) && (e.x > T);
    const bool eq = (i < n) && (e.x == T);
    const unsigned eqm = __ballot_sync(0xffffffffu, eq);
    const bool keep = gt || (eq && (eq_taken + __popc(eqm & lt) < krem));
    const unsigned km = __ballot_sync(0xffffffffu, keep);
    __syncwarp();
    if (keep) buf[w + __popc(km & lt)] = e;
    w += __popc(km);
    eq_taken += __popc(eqm);
  }
  __syncwarp();
  return T;
}


}  // namespace qst
