// K5 kernels and launch templates (included by quad_loss_{f32,f16,bf16}.cu, one translation unit per dtype:
// the template instantiations of one file took six minutes to compile).
// K5: fused gamma-quadruplet loss forward / backward / forward+backward (HBM-bound).
//
// Replaces /root/reference/models/losses/losses.py:9-69 (three F.triplet_margin_loss calls and
// the reductions) and the autograd graph behind them.  One warp owns one row: 128-bit coalesced
// loads of the four embeddings, up to six p-norm distances accumulated in registers, warp-shuffle
// reduction, then (fused / backward) the gradient of every input written once.
//
// Algorithmic traffic: forward 4*B*D*sizeof(T) read; fused forward+backward 8*B*D*sizeof(T)
// (4 reads + 4 writes; the second look at the row comes from L1/L2).
#pragma once
#include "qst_common.cuh"

namespace qst {

enum PMode { PM_2 = 0, PM_1 = 1, PM_INF = 2, PM_GEN = 3 };
enum Kind { K_FWD = 0, K_BWD = 1, K_FUSED = 2 };

constexpr int kQuadThreads = 128;          // 4 warps = 4 rows in flight per CTA
constexpr int kQuadMaxBlocks = 148 * 16;   // grid cap (persistent over rows) -> fixed-size workspace

constexpr int kQuadLimbs = 7;                    // 7 x 32 bits: 64 fractional + 160 integer bits, |x| < 2^160
struct QuadWorkspace {
  unsigned int counter;                          // ticket protocol (generic kernels)
  unsigned int pad;
  double partial[kQuadMaxBlocks];
  // limb protocol (register-resident fused kernel), see limbs_issue / limbs_finish:
  // words 0..6 sum the limbs of the non-negative partials, 7..13 those of the negative ones (magnitudes);
  // every word: [arrivals:10 | special partials:11 | sum of limbs:43]
  unsigned long long limb[16];
};
constexpr int kQuadLimbMaxBlocks = 1023;         // 10-bit arrival count

template <int PM>
__device__ __forceinline__ float acc_term(float acc, float e, float p) {
  if (PM == PM_2) return fmaf(e, e, acc);
  if (PM == PM_1) return acc + fabsf(e);
  if (PM == PM_INF) return fmaxf(acc, fabsf(e));
  return acc + powf(fabsf(e), p);
}
template <int PM>
__device__ __forceinline__ float acc_reduce(float acc) {
  return PM == PM_INF ? warp_max(acc) : warp_sum(acc);
}
template <int PM>
__device__ __forceinline__ float acc_finish(float acc, float p) {
  if (PM == PM_2) return sqrtf(acc);
  if (PM == PM_GEN) return powf(acc, 1.0f / p);
  return acc;
}

__device__ __forceinline__ float sgn(float e) { return e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f); }

// d(norm)/d(e) up to the row scalar s (see row_scale): torch norm_backward semantics.
template <int PM>
__device__ __forceinline__ float phi(float e, float d, float s, float p) {
  if (PM == PM_2) return e * s;
  if (PM == PM_1) return sgn(e) * s;
  if (PM == PM_INF) return fabsf(e) == d ? sgn(e) * s : 0.f;
  return e == 0.f ? 0.f : sgn(e) * powf(fabsf(e), p - 1.0f) * s;
}
// s: P2 -> 1/d (0 if d==0); P1 -> unused; PINF -> 1/count(|e|==d); PGEN -> d^(1-p) (0 if d==0)
template <int PM>
__device__ __forceinline__ float row_scale(float d, float p, float cnt) {
  if (PM == PM_2) return d > 0.f ? 1.0f / d : 0.f;
  if (PM == PM_1) return 1.f;
  if (PM == PM_INF) return cnt > 0.f ? 1.0f / cnt : 0.f;
  return d > 0.f ? powf(d, 1.0f - p) : 0.f;
}

// torch.minimum backward: the smaller operand takes the gradient, an exact tie splits it.
__device__ __forceinline__ void min_sel(float x, float y, float& m, float& sx, float& sy) {
  m = fminf(x, y);
  sx = x < y ? 1.f : (x == y ? 0.5f : 0.f);
  sy = y < x ? 1.f : (x == y ? 0.5f : 0.f);
}

struct RowTerms {
  float loss;
  float w[6];  // d(loss)/d(d_k)
};

// d0=d(a,pos) d1=d(a,part) d2=d(a,neg) d3=d(pos,neg) d4=d(part,neg) d5=d(pos,part)
__device__ __forceinline__ RowTerms row_terms(const float d[6], const qst_quad_params& q) {
  float dnA = d[2], dnB = d[2], dnC = d[1];
  float sA2 = 1.f, sA3 = 0.f, sB2 = 1.f, sB4 = 0.f, sC1 = 1.f, sC5 = 0.f;
  if (q.swap) {
    min_sel(d[2], d[3], dnA, sA2, sA3);
    min_sel(d[2], d[4], dnB, sB2, sB4);
    min_sel(d[1], d[5], dnC, sC1, sC5);
  }
  const float tA = (q.margin_pos_neg + d[0]) - dnA;
  const float tB = (q.margin_part_neg + d[1]) - dnB;
  const float tC = (q.margin_pos_part + d[0]) - dnC;
  // clamp_min keeps NaN (fmaxf would drop it): a NaN input makes the loss NaN, as in the reference
  const float A = tA < 0.f ? 0.f : tA, Bv = tB < 0.f ? 0.f : tB, C = tC < 0.f ? 0.f : tC;
  // clamp_min backward passes the gradient where input >= min
  const float aA = tA >= 0.f ? 1.f : 0.f;
  const float aB = tB >= 0.f ? q.gamma : 0.f;
  const float aC = tC >= 0.f ? q.one_minus_gamma : 0.f;
  RowTerms r;
  r.loss = A + q.gamma * Bv + q.one_minus_gamma * C;
  r.w[0] = aA + aC;
  r.w[1] = aB - aC * sC1;
  r.w[2] = -aA * sA2 - aB * sB2;
  r.w[3] = -aA * sA3;
  r.w[4] = -aB * sB4;
  r.w[5] = -aC * sC5;
  return r;
}

template <typename T, int VEC>
struct RowLoader {
  // loads VEC consecutive elements starting at i (i multiple of VEC when VEC > 1)
  __device__ __forceinline__ static void load(const T* __restrict__ row, int64_t i, float out[VEC]) {
    if (VEC == 1) {
      out[0] = to_f32<T>(row[i]);
    } else {
      Vec16<T> v = ld_vec16<T>(row + i);
#pragma unroll
      for (int j = 0; j < VEC; ++j) out[j] = to_f32<T>(v.v[j]);
    }
  }
  __device__ __forceinline__ static void store(T* __restrict__ row, int64_t i, const float in[VEC]) {
    if (VEC == 1) {
      row[i] = from_f32<T>(in[0]);
    } else {
      Vec16<T> v;
#pragma unroll
      for (int j = 0; j < VEC; ++j) v.v[j] = from_f32<T>(in[j]);
      st_vec16<T>(row + i, v);
    }
  }
};

template <typename T, int VEC, int PM>
__device__ __forceinline__ void row_distances(const T* __restrict__ a, const T* __restrict__ po,
                                              const T* __restrict__ pa, const T* __restrict__ ne,
                                              int64_t D, const qst_quad_params& q, int lane, float d[6]) {
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float eps = q.eps, p = q.p;
  const bool swap = q.swap != 0;
  for (int64_t i = (int64_t)lane * VEC; i < D; i += 32 * VEC) {
    float va[VEC], vp[VEC], vq[VEC], vn[VEC];
    RowLoader<T, VEC>::load(a, i, va);
    RowLoader<T, VEC>::load(po, i, vp);
    RowLoader<T, VEC>::load(pa, i, vq);
    RowLoader<T, VEC>::load(ne, i, vn);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      acc[0] = acc_term<PM>(acc[0], va[j] - vp[j] + eps, p);
      acc[1] = acc_term<PM>(acc[1], va[j] - vq[j] + eps, p);
      acc[2] = acc_term<PM>(acc[2], va[j] - vn[j] + eps, p);
      if (swap) {
        acc[3] = acc_term<PM>(acc[3], vp[j] - vn[j] + eps, p);
        acc[4] = acc_term<PM>(acc[4], vq[j] - vn[j] + eps, p);
        acc[5] = acc_term<PM>(acc[5], vp[j] - vq[j] + eps, p);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) d[k] = (k < 3 || swap) ? acc_finish<PM>(acc_reduce<PM>(acc[k]), p) : 0.f;
}

// w[k] already carries the upstream gradient of the row.
template <typename T, int VEC, int PM>
__device__ __forceinline__ void row_gradients(const T* __restrict__ a, const T* __restrict__ po,
                                              const T* __restrict__ pa, const T* __restrict__ ne,
                                              int64_t D, const qst_quad_params& q, int lane,
                                              const float d[6], const float w[6],
                                              T* __restrict__ ga, T* __restrict__ gp,
                                              T* __restrict__ gq, T* __restrict__ gn) {
  const float eps = q.eps, p = q.p;
  const bool swap = q.swap != 0;
  float cnt[6] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
  if (PM == PM_INF) {  // number of maximal elements per distance (ties share the gradient)
    float c[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int64_t i = (int64_t)lane * VEC; i < D; i += 32 * VEC) {
      float va[VEC], vp[VEC], vq[VEC], vn[VEC];
      RowLoader<T, VEC>::load(a, i, va);
      RowLoader<T, VEC>::load(po, i, vp);
      RowLoader<T, VEC>::load(pa, i, vq);
      RowLoader<T, VEC>::load(ne, i, vn);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        c[0] += fabsf(va[j] - vp[j] + eps) == d[0] ? 1.f : 0.f;
        c[1] += fabsf(va[j] - vq[j] + eps) == d[1] ? 1.f : 0.f;
        c[2] += fabsf(va[j] - vn[j] + eps) == d[2] ? 1.f : 0.f;
        if (swap) {
          c[3] += fabsf(vp[j] - vn[j] + eps) == d[3] ? 1.f : 0.f;
          c[4] += fabsf(vq[j] - vn[j] + eps) == d[4] ? 1.f : 0.f;
          c[5] += fabsf(vp[j] - vq[j] + eps) == d[5] ? 1.f : 0.f;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) cnt[k] = warp_sum(c[k]);
  }
  float s[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) s[k] = row_scale<PM>(d[k], p, cnt[k]) * w[k];

  for (int64_t i = (int64_t)lane * VEC; i < D; i += 32 * VEC) {
    float va[VEC], vp[VEC], vq[VEC], vn[VEC];
    RowLoader<T, VEC>::load(a, i, va);
    RowLoader<T, VEC>::load(po, i, vp);
    RowLoader<T, VEC>::load(pa, i, vq);
    RowLoader<T, VEC>::load(ne, i, vn);
    float oa[VEC], op[VEC], oq[VEC], on[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float f0 = phi<PM>(va[j] - vp[j] + eps, d[0], s[0], p);
      const float f1 = phi<PM>(va[j] - vq[j] + eps, d[1], s[1], p);
      const float f2 = phi<PM>(va[j] - vn[j] + eps, d[2], s[2], p);
      float f3 = 0.f, f4 = 0.f, f5 = 0.f;
      if (swap) {
        f3 = phi<PM>(vp[j] - vn[j] + eps, d[3], s[3], p);
        f4 = phi<PM>(vq[j] - vn[j] + eps, d[4], s[4], p);
        f5 = phi<PM>(vp[j] - vq[j] + eps, d[5], s[5], p);
      }
      oa[j] = f0 + f1 + f2;
      op[j] = -f0 + f3 + f5;
      oq[j] = -f1 + f4 - f5;
      on[j] = -f2 - f3 - f4;
    }
    if (ga) RowLoader<T, VEC>::store(ga, i, oa);
    if (gp) RowLoader<T, VEC>::store(gp, i, op);
    if (gq) RowLoader<T, VEC>::store(gq, i, oq);
    if (gn) RowLoader<T, VEC>::store(gn, i, on);
  }
}

struct QuadArgs {
  const void *a, *po, *pa, *ne;
  void *ga, *gp, *gq, *gn;
  int64_t B, D;
  qst_quad_params prm;
  int reduction;
  float upstream;         // fused: scalar upstream gradient
  float* loss_out;        // fwd / fused
  float* saved;           // fwd: out (may be null); bwd: in
  const float* grad_out;  // bwd
  QuadWorkspace* ws;
  int reduce_mode;        // fused register kernel: 0 limb reduction, 1 fence + ticket (rounds 1-2; QST_LOSS_REDUCE=ticket)
};

template <typename T, int VEC, int PM, int KIND>
__global__ void __launch_bounds__(kQuadThreads) quad_kernel(const QuadArgs g) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  constexpr int kWarps = kQuadThreads / 32;
  const int64_t D = g.D;
  const float inv_b = g.reduction == QST_RED_MEAN ? 1.0f / (float)g.B : 1.0f;
  double block_sum = 0.0;  // meaningful in lane 0 of each warp

  for (int64_t row = (int64_t)blockIdx.x * kWarps + warp; row < g.B; row += (int64_t)gridDim.x * kWarps) {
    const T* a = reinterpret_cast<const T*>(g.a) + row * D;
    const T* po = reinterpret_cast<const T*>(g.po) + row * D;
    const T* pa = reinterpret_cast<const T*>(g.pa) + row * D;
    const T* ne = reinterpret_cast<const T*>(g.ne) + row * D;
    float d[6];
    if (KIND == K_BWD) {
#pragma unroll
      for (int k = 0; k < 6; ++k) d[k] = g.saved[row * QST_QUAD_SAVED_PER_ROW + k];
    } else {
      row_distances<T, VEC, PM>(a, po, pa, ne, D, g.prm, lane, d);
    }
    RowTerms t = row_terms(d, g.prm);
    if (KIND != K_BWD) {
      if (lane == 0) {
        if (g.saved) {
#pragma unroll
          for (int k = 0; k < 6; ++k) g.saved[row * QST_QUAD_SAVED_PER_ROW + k] = d[k];
        }
        if (g.reduction == QST_RED_NONE) g.loss_out[row] = t.loss;
      }
      block_sum += (double)t.loss;
    }
    if (KIND != K_FWD) {
      float up;
      if (KIND == K_BWD) up = (g.reduction == QST_RED_NONE ? g.grad_out[row] : g.grad_out[0] * inv_b);
      else up = g.upstream * inv_b;
      float w[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) w[k] = t.w[k] * up;
      T* ga = g.ga ? reinterpret_cast<T*>(g.ga) + row * D : nullptr;
      T* gp = g.gp ? reinterpret_cast<T*>(g.gp) + row * D : nullptr;
      T* gq = g.gq ? reinterpret_cast<T*>(g.gq) + row * D : nullptr;
      T* gn = g.gn ? reinterpret_cast<T*>(g.gn) + row * D : nullptr;
      row_gradients<T, VEC, PM>(a, po, pa, ne, D, g.prm, lane, d, w, ga, gp, gq, gn);
    }
  }

  if (KIND != K_BWD && g.reduction != QST_RED_NONE) {
    // deterministic two-level reduction: warps -> CTA partial -> last CTA sums partials in order
    __shared__ double s_part[kWarps];
    __shared__ bool s_last;
    if (lane == 0) s_part[warp] = block_sum;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) tot += s_part[w];
      g.ws->partial[blockIdx.x] = tot;
      __threadfence();
      const unsigned int ticket = atomicAdd(&g.ws->counter, 1u);
      s_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && warp == 0) {
      __threadfence();
      double acc = 0.0;
      for (int i = lane; i < (int)gridDim.x; i += 32) acc += __ldcg(&g.ws->partial[i]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) {
        if (g.reduction == QST_RED_MEAN) acc /= (double)g.B;
        g.loss_out[0] = (float)acc;
        g.ws->counter = 0u;  // leave the workspace zeroed for the next launch
      }
    }
  }
}

// Cross-CTA loss reduction for the persistent fused kernel, placed BEFORE a warp's last gradient
// stores so that its fence / ticket latency overlaps them.  Warps 1..3 only arrive on a named
// barrier and move on; warp 0 collects the CTA sum, publishes it and takes a ticket; the CTA that
// draws the last ticket sums all partials in a fixed order (deterministic result).
__device__ __forceinline__ void post_cta_loss(const QuadArgs& g, double warp_sum_d, double* s_part, int warp, int lane) {
  constexpr int kWarps = kQuadThreads / 32;
  if (lane == 0) s_part[warp] = warp_sum_d;
  if (warp != 0) {
    __threadfence_block();
    asm volatile("bar.arrive 1, %0;" ::"n"(kQuadThreads) : "memory");
    return;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kQuadThreads) : "memory");
  unsigned int ticket = 0;
  if (lane == 0) {
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) tot += s_part[w];
    g.ws->partial[blockIdx.x] = tot;
    __threadfence();
    ticket = atomicAdd(&g.ws->counter, 1u);
  }
  ticket = __shfl_sync(0xffffffffu, ticket, 0);
  if (ticket == gridDim.x - 1) {
    __threadfence();
    double acc2 = 0.0;
    for (int i = lane; i < (int)gridDim.x; i += 32) acc2 += __ldcg(&g.ws->partial[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc2 += __shfl_xor_sync(0xffffffffu, acc2, o);
    if (lane == 0) {
      if (g.reduction == QST_RED_MEAN) acc2 /= (double)g.B;
      g.loss_out[0] = (float)acc2;
      g.ws->counter = 0u;
    }
  }
}

// Cross-CTA loss reduction without a fence, without a partial array to re-read and without anybody waiting
// for a reply while the memory system is busy: the CTA's partial sum travels inside the atomics themselves.
// The partial (a double) is written as a sign + 224-bit fixed-point magnitude with 64 fractional bits
// (|x| < 2^160: more than any sum of fp32 row losses can reach), cut into seven 32-bit limbs; lane i of
// warp 0 adds limb i to word i (non-negative partials) or word 7 + i (negative ones), every lane adds an
// arrival to the top ten bits of its word -- ONE atom instruction per CTA.  Integer addition is associative,
// so the fourteen sums, and the loss computed from them, do not depend on the order in which CTAs arrive:
// bitwise reproducible; the sums are exact, their evaluation rounds to double a few times.  NaN, +inf and -inf partials (and magnitudes of
// 2^160 and more, which fp32 cannot hold either) are counted in the spare bits of words 0, 1 and 2 and come
// out as float addition gives them.
//   limbs_issue   every warp when its share of the loss is complete, i.e. BEFORE its last gradient stores.
//                 Warps 1..3 post their sums and go on (named barrier, arrive only); warp 0 issues the atom.
//                 Nothing is ordered against the gradient stores, so nobody waits for stores to drain (a
//                 fence would), and the reply travels while warp 0 issues its own stores.
//   limbs_finish  warp 0, after its stores: the CTA whose reply from word 0 says "all others have arrived"
//                 makes sure the other words show the full count too (normally the replies already do),
//                 evaluates the limbs in a fixed order, writes the loss and zeroes the words for the next
//                 launch.
// What this replaced, and what the measurements said (profiles/r02_loss_limbs_vs_ticket.txt): partial store +
// fence + ticket + last CTA re-reads the partials cost 2.4 us on top of the 18.7 us of reduction='none';
// issuing these atomics AFTER the last stores and waiting for the reply there cost the same (a reply queued
// behind the kernel's last stores takes ~2 us); and a first version that tagged per-CTA fallback slots
// with a launch epoch lost 1.4 us to the single strong load that fetched the epoch at kernel start.
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long atom_add_relaxed_u64(unsigned long long* p, unsigned long long v) {
  unsigned long long old;
  asm volatile("atom.relaxed.gpu.global.add.u64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(v) : "memory");
  return old;
}
constexpr unsigned long long kLimbOne = 1ull << 54, kLimbSpecial = 1ull << 43, kLimbSum = (1ull << 43) - 1;

// limb `idx` (bits [32 idx, 32 idx + 32) of |x| * 2^64) of a finite double below 2^160
__device__ __forceinline__ unsigned long long limb_of(unsigned long long mant, int expo, int idx) {
  const int sh = expo + 12 - 32 * idx;   // position of the mantissa's lowest bit relative to the limb's
  if (sh >= 32 || sh <= -53) return 0ull;
  return (sh >= 0 ? (mant << sh) : (mant >> -sh)) & 0xffffffffull;
}

// `reply` (the word before this CTA's add) is still in flight when the function returns -- do not look at it
// before the gradient stores have been issued; `added` is what the lane added.  Per-lane addresses also keep
// ptxas from wrapping the atomic in its warp-aggregation code (it does that for a provably warp-uniform
// address), whose closing shuffle of the reply would make the warp wait on the spot.
__device__ __forceinline__ void limbs_issue(const QuadArgs& g, double warp_sum_d, volatile double* s_part, int warp, int lane,
                                            unsigned long long& reply, unsigned long long& added) {
  constexpr int kWarps = kQuadThreads / 32;
  if (lane == 0) s_part[warp] = warp_sum_d;
  // barrier.arrive / barrier.sync order the shared-memory writes before the reads (PTX: when the barrier
  // completes, prior accesses of the arriving threads are performed relative to the participants)
  if (warp != 0) {
    asm volatile("bar.arrive 1, %0;" ::"n"(kQuadThreads) : "memory");
    return;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kQuadThreads) : "memory");
  double tot = 0.0;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) tot += s_part[w];
  if (lane < 2 * kQuadLimbs) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(tot);
    const bool neg = (bits >> 63) != 0;
    const int biased = (int)((bits >> 52) & 0x7ffull);
    const unsigned long long frac = bits & ((1ull << 52) - 1);
    const int expo = biased - 1023;
    const bool is_nan = biased == 0x7ff && frac != 0;
    const bool is_big = !is_nan && expo >= 160;            // +-inf, or beyond the limbs (and beyond fp32)
    const int idx = lane < kQuadLimbs ? lane : lane - kQuadLimbs;
    unsigned long long limb = 0ull;
    if (biased != 0 && !is_nan && !is_big && neg == (lane >= kQuadLimbs)) limb = limb_of(frac | (1ull << 52), expo, idx);
    const bool special = (lane == 0 && is_nan) || (lane == 1 && is_big && !neg) || (lane == 2 && is_big && neg);
    added = kLimbOne | limb | (special ? kLimbSpecial : 0ull);
    reply = atom_add_relaxed_u64(&g.ws->limb[lane], added);
  }
}

// warp 0 only, all lanes, after the warp's last gradient stores
__device__ __forceinline__ void limbs_finish(const QuadArgs& g, unsigned long long reply, unsigned long long added, int lane) {
  const unsigned int n = gridDim.x;
  const unsigned int last = __shfl_sync(0xffffffffu, (unsigned int)(reply >> 54) == n - 1 ? 1u : 0u, 0);
  if (!last) return;
  QuadWorkspace* ws = g.ws;
  unsigned long long v = reply + added;     // the word right after this CTA's add
  if (lane < 2 * kQuadLimbs) {
    while ((unsigned int)(v >> 54) != n) v = ld_volatile_u64(&ws->limb[lane]);   // a few adds may still be under way
  }
  // fixed evaluation order: Horner from the top limb, positive and negative side, then the difference
  double pos = 0.0, neg = 0.0;
#pragma unroll
  for (int i = kQuadLimbs - 1; i >= 0; --i) {
    pos = pos * 4294967296.0 + (double)(__shfl_sync(0xffffffffu, v, i) & kLimbSum);
    neg = neg * 4294967296.0 + (double)(__shfl_sync(0xffffffffu, v, kQuadLimbs + i) & kLimbSum);
  }
  const bool any_nan = ((__shfl_sync(0xffffffffu, v, 0) >> 43) & 0x7ffull) != 0;
  const bool any_pinf = ((__shfl_sync(0xffffffffu, v, 1) >> 43) & 0x7ffull) != 0;
  const bool any_ninf = ((__shfl_sync(0xffffffffu, v, 2) >> 43) & 0x7ffull) != 0;
  if (lane == 0) {
    double result = (pos - neg) * (1.0 / 18446744073709551616.0);
    if (any_nan || (any_pinf && any_ninf)) result = __longlong_as_double(0x7ff8000000000000ll);
    else if (any_pinf) result = __longlong_as_double(0x7ff0000000000000ll);
    else if (any_ninf) result = __longlong_as_double((long long)0xfff0000000000000ull);
    if (g.reduction == QST_RED_MEAN) result /= (double)g.B;
    g.loss_out[0] = (float)result;
  }
  if (lane < 2 * kQuadLimbs) ws->limb[lane] = 0ull;   // every CTA's adds have been counted: nobody touches the words again
}

// ------------------------------------------------------------------------------------------
// Register-resident fused forward+backward for rows of at most 32*VEC*kRegChunks elements
// (1024 fp32 / 2048 half): the four rows are loaded ONCE with every 128-bit load issued up front
// (kRegChunks*4 independent loads per thread), distances, loss terms and all four gradients are
// computed from registers.  HBM traffic = the algorithmic 8*B*D*sizeof(T), nothing re-read.
// ------------------------------------------------------------------------------------------
constexpr int kRegChunksMax = 8;

// NCH = 16-byte chunks per lane and input row (row length <= 32*VEC*NCH): sized to the row so that
// short rows do not pay registers (occupancy) for the longest supported one
template <typename T, int PM, int NCH>
__global__ void __launch_bounds__(kQuadThreads, NCH <= 6 ? 3 : 2) quad_fused_reg_kernel(const QuadArgs g) {
  constexpr int kRegChunks = NCH;
  constexpr int VEC = 16 / sizeof(T);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  constexpr int kWarps = kQuadThreads / 32;
  const int64_t D = g.D;
  const float inv_b = g.reduction == QST_RED_MEAN ? 1.0f / (float)g.B : 1.0f;
  const float eps = g.prm.eps, p = g.prm.p;
  const bool swap = g.prm.swap != 0;
  double block_sum = 0.0;
  __shared__ double s_part[kWarps];
  bool posted = false;
  unsigned long long reply = 0ull, added = 0ull;   // limb protocol, lanes 0..13 of warp 0: see limbs_issue
  const int64_t row_stride = (int64_t)gridDim.x * kWarps;
  const bool limbs = g.reduction != QST_RED_NONE && !(g.reduce_mode & 1);

  for (int64_t row = (int64_t)blockIdx.x * kWarps + warp; row < g.B; row += (int64_t)gridDim.x * kWarps) {
    const T* a = reinterpret_cast<const T*>(g.a) + row * D;
    const T* po = reinterpret_cast<const T*>(g.po) + row * D;
    const T* pa = reinterpret_cast<const T*>(g.pa) + row * D;
    const T* ne = reinterpret_cast<const T*>(g.ne) + row * D;
    Vec16<T> ra[kRegChunks], rp[kRegChunks], rq[kRegChunks], rn[kRegChunks];
#pragma unroll
    for (int c = 0; c < kRegChunks; ++c) {
      const int64_t i = ((int64_t)c * 32 + lane) * VEC;
      if (i < D) {
        ra[c] = ld_vec16<T>(a + i);
        rp[c] = ld_vec16<T>(po + i);
        rq[c] = ld_vec16<T>(pa + i);
        rn[c] = ld_vec16<T>(ne + i);
      }
    }
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < kRegChunks; ++c) {
      if (((int64_t)c * 32 + lane) * VEC < D) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float va = to_f32<T>(ra[c].v[j]), vp = to_f32<T>(rp[c].v[j]);
          const float vq = to_f32<T>(rq[c].v[j]), vn = to_f32<T>(rn[c].v[j]);
          acc[0] = acc_term<PM>(acc[0], va - vp + eps, p);
          acc[1] = acc_term<PM>(acc[1], va - vq + eps, p);
          acc[2] = acc_term<PM>(acc[2], va - vn + eps, p);
          if (swap) {
            acc[3] = acc_term<PM>(acc[3], vp - vn + eps, p);
            acc[4] = acc_term<PM>(acc[4], vq - vn + eps, p);
            acc[5] = acc_term<PM>(acc[5], vp - vq + eps, p);
          }
        }
      }
    }
    float d[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) d[k] = (k < 3 || swap) ? acc_finish<PM>(acc_reduce<PM>(acc[k]), p) : 0.f;
    const RowTerms t = row_terms(d, g.prm);
    if (lane == 0 && g.reduction == QST_RED_NONE) g.loss_out[row] = t.loss;
    block_sum += (double)t.loss;
    if (g.reduction != QST_RED_NONE && row + row_stride >= g.B) {   // this warp's last row
      if (limbs) limbs_issue(g, block_sum, s_part, warp, lane, reply, added);   // warp 0 reads the reply below the loop
      else post_cta_loss(g, block_sum, s_part, warp, lane);
      posted = true;
    }

    const float up = g.upstream * inv_b;
    float cnt[6] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
    if (PM == PM_INF) {
      float cc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < kRegChunks; ++c) {
        if (((int64_t)c * 32 + lane) * VEC < D) {
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            const float va = to_f32<T>(ra[c].v[j]), vp = to_f32<T>(rp[c].v[j]);
            const float vq = to_f32<T>(rq[c].v[j]), vn = to_f32<T>(rn[c].v[j]);
            cc[0] += fabsf(va - vp + eps) == d[0] ? 1.f : 0.f;
            cc[1] += fabsf(va - vq + eps) == d[1] ? 1.f : 0.f;
            cc[2] += fabsf(va - vn + eps) == d[2] ? 1.f : 0.f;
            if (swap) {
              cc[3] += fabsf(vp - vn + eps) == d[3] ? 1.f : 0.f;
              cc[4] += fabsf(vq - vn + eps) == d[4] ? 1.f : 0.f;
              cc[5] += fabsf(vp - vq + eps) == d[5] ? 1.f : 0.f;
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) cnt[k] = warp_sum(cc[k]);
    }
    float sc[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) sc[k] = row_scale<PM>(d[k], p, cnt[k]) * (t.w[k] * up);
    T* ga = g.ga ? reinterpret_cast<T*>(g.ga) + row * D : nullptr;
    T* gp = g.gp ? reinterpret_cast<T*>(g.gp) + row * D : nullptr;
    T* gq = g.gq ? reinterpret_cast<T*>(g.gq) + row * D : nullptr;
    T* gn = g.gn ? reinterpret_cast<T*>(g.gn) + row * D : nullptr;
#pragma unroll
    for (int c = 0; c < kRegChunks; ++c) {
      const int64_t i = ((int64_t)c * 32 + lane) * VEC;
      if (i < D) {
        Vec16<T> oa, op, oq, on;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float va = to_f32<T>(ra[c].v[j]), vp = to_f32<T>(rp[c].v[j]);
          const float vq = to_f32<T>(rq[c].v[j]), vn = to_f32<T>(rn[c].v[j]);
          const float f0 = phi<PM>(va - vp + eps, d[0], sc[0], p);
          const float f1 = phi<PM>(va - vq + eps, d[1], sc[1], p);
          const float f2 = phi<PM>(va - vn + eps, d[2], sc[2], p);
          float f3 = 0.f, f4 = 0.f, f5 = 0.f;
          if (swap) {
            f3 = phi<PM>(vp - vn + eps, d[3], sc[3], p);
            f4 = phi<PM>(vq - vn + eps, d[4], sc[4], p);
            f5 = phi<PM>(vp - vq + eps, d[5], sc[5], p);
          }
          oa.v[j] = from_f32<T>(f0 + f1 + f2);
          op.v[j] = from_f32<T>(-f0 + f3 + f5);
          oq.v[j] = from_f32<T>(-f1 + f4 - f5);
          on.v[j] = from_f32<T>(-f2 - f3 - f4);
        }
        if (ga) st_vec16<T>(ga + i, oa);
        if (gp) st_vec16<T>(gp + i, op);
        if (gq) st_vec16<T>(gq + i, oq);
        if (gn) st_vec16<T>(gn + i, on);
      }
    }
  }

  if (g.reduction != QST_RED_NONE && !posted) {   // warp had no row
    if (limbs) limbs_issue(g, block_sum, s_part, warp, lane, reply, added);
    else post_cta_loss(g, block_sum, s_part, warp, lane);
  }
  if (limbs && warp == 0) limbs_finish(g, reply, added, lane);
}

template <typename T, int NCH>
static void launch_fused_reg_n(const QuadArgs& a, int pm, int grid, cudaStream_t st) {
  switch (pm) {
    case PM_2: quad_fused_reg_kernel<T, PM_2, NCH><<<grid, kQuadThreads, 0, st>>>(a); break;
    case PM_1: quad_fused_reg_kernel<T, PM_1, NCH><<<grid, kQuadThreads, 0, st>>>(a); break;
    case PM_INF: quad_fused_reg_kernel<T, PM_INF, NCH><<<grid, kQuadThreads, 0, st>>>(a); break;
    default: quad_fused_reg_kernel<T, PM_GEN, NCH><<<grid, kQuadThreads, 0, st>>>(a); break;
  }
}

template <typename T>
static void launch_fused_reg(const QuadArgs& a, int pm, int grid, cudaStream_t st) {
  const int64_t per_chunk = 32 * (16 / sizeof(T));
  const int nch = (int)ceil_div(a.D, per_chunk);
  if (nch <= 2) launch_fused_reg_n<T, 2>(a, pm, grid, st);
  else if (nch <= 4) launch_fused_reg_n<T, 4>(a, pm, grid, st);
  else if (nch <= 6) launch_fused_reg_n<T, 6>(a, pm, grid, st);
  else launch_fused_reg_n<T, 8>(a, pm, grid, st);
}

template <typename T, int VEC, int KIND>
static void launch_pm(const QuadArgs& a, int pm, int grid, cudaStream_t st) {
  switch (pm) {
    case PM_2: quad_kernel<T, VEC, PM_2, KIND><<<grid, kQuadThreads, 0, st>>>(a); break;
    case PM_1: quad_kernel<T, VEC, PM_1, KIND><<<grid, kQuadThreads, 0, st>>>(a); break;
    case PM_INF: quad_kernel<T, VEC, PM_INF, KIND><<<grid, kQuadThreads, 0, st>>>(a); break;
    default: quad_kernel<T, VEC, PM_GEN, KIND><<<grid, kQuadThreads, 0, st>>>(a); break;
  }
}

template <typename T, int KIND>
static void launch_vec(const QuadArgs& a, int pm, bool vec_ok, int grid, cudaStream_t st) {
  if (vec_ok) launch_pm<T, 16 / sizeof(T), KIND>(a, pm, grid, st);
  else launch_pm<T, 1, KIND>(a, pm, grid, st);
}

// one entry per dtype, defined in quad_loss_<dtype>.cu
void quad_launch_f32(int kind, const QuadArgs& a, int pm, bool vec_ok, bool reg_path, int grid, cudaStream_t st);
void quad_launch_f16(int kind, const QuadArgs& a, int pm, bool vec_ok, bool reg_path, int grid, cudaStream_t st);
void quad_launch_bf16(int kind, const QuadArgs& a, int pm, bool vec_ok, bool reg_path, int grid, cudaStream_t st);

}  // namespace qst
