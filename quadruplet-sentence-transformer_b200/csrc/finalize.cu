// K3: candidate selection, exact fp32 rescoring, final ordering and exactness certificate.
// K6: merge of per-shard top-k lists.   Exact re-scan for uncertified queries.
//
// Replaces the tail of sentence-transformers 2.2.2 InformationRetrievalEvaluator:
//   torch.topk(...)  +  per-query  sorted(hits, key=score, reverse=True)
// (reference construction sites: /root/reference/ir_evauation_script.py:107-123,
// models/evaluators.py:572-588).  The fp32 rescoring makes the final order the one the
// reference's fp32 cos_sim / dot_score produces (up to ties within 1e-6).
#include "qst_common.cuh"
#include "select_common.cuh"
#include <stdlib.h>

namespace qst {

constexpr int kFinThreads = 256;
constexpr int kFinWarps = kFinThreads / 32;
constexpr int kMaxStripes = 160;  // must cover qst_topk_plan_make's s_max

// ------------------------------------------------------------------------------------------
// CTA-wide: keep the k largest keys of (keys, idx)[0..n) -> compacted to the front (order
// arbitrary).  Returns the k-th largest key.  Requires n >= k.  tmp_* hold k entries.
// ------------------------------------------------------------------------------------------
__device__ uint32_t block_select_topk(uint32_t* keys, int32_t* idx, int n, int k, uint32_t* tmp_keys,
                                      int32_t* tmp_idx, int* hist, int* s_misc) {
  const int tid = threadIdx.x;
  uint32_t prefix = 0, mask = 0;
  int krem = k;
  for (int shift = 24; shift >= 0; shift -= 8) {
    hist[tid] = 0;  // kFinThreads == 256 bins
    __syncthreads();
    for (int i = tid; i < n; i += kFinThreads) {
      const uint32_t key = keys[i];
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1);
    }
    __syncthreads();
    if (tid < 32) {
      int c[8], ls = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) { c[i] = hist[tid * 8 + i]; ls += c[i]; }
      int incl = ls;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_down_sync(0xffffffffu, incl, o);
        if (tid + o < 32) incl += t;
      }
      const int above = incl - ls;
      if (above < krem && krem <= above + ls) {
        int run = above;
#pragma unroll
        for (int i = 7; i >= 0; --i) {
          if (run < krem && krem <= run + c[i]) { s_misc[0] = tid * 8 + i; s_misc[1] = run; }
          run += c[i];
        }
      }
    }
    __syncthreads();
    krem -= s_misc[1];
    prefix |= (uint32_t)s_misc[0] << shift;
    mask |= 255u << shift;
    __syncthreads();
  }
  const uint32_t T = prefix;
  if (tid == 0) { s_misc[2] = 0; s_misc[3] = 0; }
  __syncthreads();
  for (int i = tid; i < n; i += kFinThreads) {
    const uint32_t key = keys[i];
    bool keep = key > T;
    if (key == T) keep = atomicAdd(&s_misc[3], 1) < krem;
    if (keep) {
      const int p = atomicAdd(&s_misc[2], 1);
      tmp_keys[p] = key;
      tmp_idx[p] = idx[i];
    }
  }
  __syncthreads();
  for (int i = tid; i < k; i += kFinThreads) { keys[i] = tmp_keys[i]; idx[i] = tmp_idx[i]; }
  __syncthreads();
  return T;
}

// order: score descending, ties -> lower index first; empty slots (idx < 0) last
__device__ __forceinline__ bool beats(float sa, int64_t ia, float sb, int64_t ib) {
  if (ia < 0) return false;
  if (ib < 0) return true;
  return sa > sb || (sa == sb && ia < ib);
}

// CTA-wide bitonic sort of n2 (power of two) (score, idx) pairs into `beats` order.
__device__ void block_bitonic_sort(float* sc, int32_t* ix, int n2) {
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < n2 / 2; t += kFinThreads) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;  // first half of each `size` block: best first
        const bool hi_first = beats(sc[hi], ix[hi], sc[lo], ix[lo]);
        if (hi_first == up) {
          const float ts = sc[lo]; sc[lo] = sc[hi]; sc[hi] = ts;
          const int32_t ti = ix[lo]; ix[lo] = ix[hi]; ix[hi] = ti;
        }
      }
    }
  }
  __syncthreads();
}

// Exact fp32 dot product of a shared-memory query row with a global corpus row, one warp.
// Both K3 (rescoring) and the exact re-scan call this, so a row gets bit-identical scores in both.
// SQ = true accumulates (q_i - c_i)^2 instead of q_i * c_i (euclidean_score).
template <bool SQ>
__device__ __forceinline__ float acc1(float a, float c, float acc) {
  if (SQ) { const float d = a - c; return fmaf(d, d, acc); }
  return fmaf(a, c, acc);
}

template <bool SQ>
__device__ __forceinline__ float warp_dot_f32(const float* __restrict__ qrow, const float* __restrict__ crow, int D,
                                              bool vec4, int lane) {
  float acc = 0.f;
  if (vec4) {
    const float4* c4 = reinterpret_cast<const float4*>(crow);
    const float4* q4 = reinterpret_cast<const float4*>(qrow);
    for (int i = lane; i < D / 4; i += 32) {
      const float4 c = __ldg(c4 + i);
      const float4 a = q4[i];
      acc = acc1<SQ>(a.x, c.x, acc); acc = acc1<SQ>(a.y, c.y, acc);
      acc = acc1<SQ>(a.z, c.z, acc); acc = acc1<SQ>(a.w, c.w, acc);
    }
  } else {
    for (int i = lane; i < D; i += 32) acc = acc1<SQ>(qrow[i], __ldg(crow + i), acc);
  }
  return warp_sum(acc);
}

// Two rows at once: same per-row operation order as warp_dot_f32 (bit-identical results), twice the
// loads in flight.
template <bool SQ>
__device__ __forceinline__ void warp_dot_f32_x2(const float* __restrict__ qrow, const float* __restrict__ c0,
                                                const float* __restrict__ c1, int D, bool vec4, int lane, float& o0,
                                                float& o1) {
  float a0 = 0.f, a1 = 0.f;
  if (vec4) {
    const float4* p0 = reinterpret_cast<const float4*>(c0);
    const float4* p1 = reinterpret_cast<const float4*>(c1);
    const float4* q4 = reinterpret_cast<const float4*>(qrow);
    for (int i = lane; i < D / 4; i += 32) {
      const float4 x = __ldg(p0 + i);
      const float4 y = __ldg(p1 + i);
      const float4 a = q4[i];
      a0 = acc1<SQ>(a.x, x.x, a0); a0 = acc1<SQ>(a.y, x.y, a0); a0 = acc1<SQ>(a.z, x.z, a0); a0 = acc1<SQ>(a.w, x.w, a0);
      a1 = acc1<SQ>(a.x, y.x, a1); a1 = acc1<SQ>(a.y, y.y, a1); a1 = acc1<SQ>(a.z, y.z, a1); a1 = acc1<SQ>(a.w, y.w, a1);
    }
  } else {
    for (int i = lane; i < D; i += 32) {
      const float a = qrow[i];
      a0 = acc1<SQ>(a, __ldg(c0 + i), a0);
      a1 = acc1<SQ>(a, __ldg(c1 + i), a1);
    }
  }
  o0 = warp_sum(a0);
  o1 = warp_sum(a1);
}

__device__ __forceinline__ float apply_score(float dot, int score, float q_inv, const float* c_inv, int row) {
  if (score == QST_SCORE_EUCLID) return 1.0f / (1.0f + sqrtf(dot));   // `dot` is ||q-c||^2 here
  return score == QST_SCORE_COS ? (dot * q_inv) * (c_inv ? c_inv[row] : 1.0f) : dot;
}

// Destination of a producer kernel of the sharded path whose output rows are grouped in `world` blocks of
// `rows_per_block` rows, block b being meant for rank b (candidate lists -> owners, requests -> shards,
// exact scores -> owners).  world == 0: plain local output.  Otherwise row (b, i) is stored at row
// (rank * rows_per_block + i) of base[b] -- rank b's peer-mapped receive buffer, written straight over
// NVLink by the kernel that produces the row (the all-to-all that would follow is a flag barrier).
struct Scatter {
  void* base[QST_MAX_WORLD];
  int world, rank;
  long long rows_per_block;
};

template <typename T>
__device__ __forceinline__ T* scatter_row(const Scatter& sc, T* local, long long row, long long pitch) {
  if (sc.world == 0) return local + row * pitch;
  const long long b = row / sc.rows_per_block;
  const long long i = row - b * sc.rows_per_block;
  return reinterpret_cast<T*>(sc.base[b]) + ((long long)sc.rank * sc.rows_per_block + i) * pitch;
}

struct FinParams {
  int Q, N, D;
  int k, kprime, cap, m_tiles, stripes, score, rows_per_unit;
  int sm_cap;  // smem candidate capacity (entries)
  const uint32_t* thr_hint;
  const int* unit_cnt;
  const uint32_t* unit_thr;
  const uint2* unit_cand;
  const float *q_f32, *q_inv, *q_err, *c_f32, *c_inv, *c_stats;
  int64_t idx_offset;
  float* out_val;
  int64_t* out_idx;
  float* out_margin;
  // select-only mode (sharded retrieval, step 1): write the m best candidates by bf16 key instead
  // of rescoring; entry m of each row carries (bound key, count)
  uint2* sel_out;
  Scatter scat;   // where sel_out / req_out rows go when the exchange is fused into this kernel
  // request mode (fully sharded master, owner side): instead of rescoring, the k' best candidates are
  // grouped by the shard that holds them: req_out [G, Q, req_m] local row ids (-1 = none),
  // bound_out [Q] = bf16 key bounding everything that is NOT requested
  int32_t* req_out;
  uint32_t* bound_out;
  int req_G, req_m;
  int64_t n_total;
  // refine mode: queries whose certificate already holds (out_margin[q] > 0 on entry) are left alone
  int refine;
};

// balanced contiguous shards (sharded.shard_bounds): first n % G shards hold one row more
__device__ __forceinline__ int shard_of_row(int64_t id, int64_t n_total, int G, int64_t* start) {
  const int64_t base = n_total / G, rem = n_total % G;
  const int64_t big = rem * (base + 1);
  int g;
  if (id < big) { g = (int)(id / (base + 1)); *start = (int64_t)g * (base + 1); }
  else { g = (int)(rem + (id - big) / (base > 0 ? base : 1)); *start = big + (int64_t)(g - rem) * base; }
  return g;
}
__device__ __forceinline__ int64_t shard_start_of(int g, int64_t n_total, int G) {
  const int64_t base = n_total / G, rem = n_total % G;
  return (int64_t)g * base + (g < rem ? g : rem);
}

__global__ void __launch_bounds__(kFinThreads) finalize_kernel(const FinParams P) {
  extern __shared__ uint8_t fsm[];
  // layout: keys[sm_cap] | idx[sm_cap] | tmp_keys[kprime] | tmp_idx[kprime] | qrow[D] | exact[kprime2]
  uint32_t* keys = reinterpret_cast<uint32_t*>(fsm);
  int32_t* idx = reinterpret_cast<int32_t*>(keys + P.sm_cap);
  uint32_t* tmp_keys = reinterpret_cast<uint32_t*>(idx + P.sm_cap);
  int32_t* tmp_idx = reinterpret_cast<int32_t*>(tmp_keys + P.kprime);
  float* qrow = reinterpret_cast<float*>(tmp_idx + P.kprime);
  __shared__ int hist[256];
  __shared__ int s_misc[4];
  __shared__ float s_red[kFinWarps];

  const int q = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m = q / P.rows_per_unit, r = q % P.rows_per_unit;
  if (P.refine && P.out_margin[q] > 0.f) return;   // certified by the first, narrower pass

  // stage the fp32 query row and its squared norm
  float qq = 0.f;
  for (int i = tid; i < P.D; i += kFinThreads) {
    const float v = P.q_f32[(size_t)q * P.D + i];
    qrow[i] = v;
    qq = fmaf(v, v, qq);
  }
  qq = warp_sum(qq);
  if (lane == 0) s_red[warp] = qq;
  __syncthreads();
  float qnorm2 = 0.f;
  for (int w = 0; w < kFinWarps; ++w) qnorm2 += s_red[w];

  // 1. gather the candidates of every stripe; compact to the best k' whenever smem fills up
  bool reduced = false;
  uint32_t T = 0;      // k'-th largest key among the gathered candidates (once anything is dropped here)
  uint32_t Tstar = 0;  // largest final unit threshold: upper bound of everything K2 discarded
  int fill = 0;
  __shared__ int s_cnt[kMaxStripes + 1];
  __shared__ uint32_t s_tstar;
  // all stripe counts at once, then every warp copies whole unit buffers in parallel
  if (tid == 0) s_tstar = 0u;
  __syncthreads();
  for (int st = tid; st < P.stripes; st += kFinThreads) {
    const size_t urow = (size_t)(st * P.m_tiles + m) * P.rows_per_unit + r;
    s_cnt[st] = P.unit_cnt[urow];
    atomicMax(&s_tstar, P.unit_thr[urow]);
  }
  __syncthreads();
  if (tid == 0) {  // exclusive prefix (at most kMaxStripes terms)
    int run = 0;
    for (int st = 0; st < P.stripes; ++st) { const int c = s_cnt[st]; s_cnt[st] = run; run += c; }
    s_cnt[P.stripes] = run;
  }
  __syncthreads();
  Tstar = s_tstar;
  const int total = s_cnt[P.stripes];
  if (total <= P.sm_cap) {
    for (int st = warp; st < P.stripes; st += kFinWarps) {
      const size_t urow = (size_t)(st * P.m_tiles + m) * P.rows_per_unit + r;
      const int off = s_cnt[st], cnt = s_cnt[st + 1] - off;
      const uint2* src = P.unit_cand + urow * (size_t)P.cap;
      for (int i = lane; i < cnt; i += 32) {
        const uint2 e = src[i];
        keys[off + i] = e.x;
        idx[off + i] = (int32_t)e.y;
      }
    }
    fill = total;
    __syncthreads();
  } else {
    // rare (very large k'): stream the unit buffers through shared memory, compacting to the best
    // k' whenever it fills up
    for (int st = 0; st < P.stripes; ++st) {
      const size_t urow = (size_t)(st * P.m_tiles + m) * P.rows_per_unit + r;
      const int cnt = s_cnt[st + 1] - s_cnt[st];
      if (fill + cnt > P.sm_cap) {
        T = block_select_topk(keys, idx, fill, P.kprime, tmp_keys, tmp_idx, hist, s_misc);
        fill = P.kprime;
        reduced = true;
      }
      const uint2* src = P.unit_cand + urow * (size_t)P.cap;
      for (int i = tid; i < cnt; i += kFinThreads) {
        const uint2 e = src[i];
        keys[fill + i] = e.x;
        idx[fill + i] = (int32_t)e.y;
      }
      fill += cnt;
      __syncthreads();
    }
  }
  if (fill > P.kprime) {
    T = block_select_topk(keys, idx, fill, P.kprime, tmp_keys, tmp_idx, hist, s_misc);
    fill = P.kprime;
    reduced = true;
  }
  const int ncand = fill;  // <= kprime
  if (P.sel_out) {
    // everything this shard did not put in the list has a bf16 key <= bound
    const uint32_t bound = reduced ? max(T, Tstar) : Tstar;
    uint2* dst = scatter_row(P.scat, P.sel_out, q, P.kprime + 1);
    for (int i = tid; i < P.kprime; i += kFinThreads)
      dst[i] = i < ncand ? make_uint2(keys[i], (uint32_t)((int64_t)idx[i] + P.idx_offset)) : make_uint2(0u, 0xffffffffu);
    if (tid == 0) dst[P.kprime] = make_uint2(bound, (uint32_t)ncand);
    return;
  }
  if (P.req_out) {
    // hand every selected candidate to the shard that owns its row; slot order inside a shard's list
    // is arbitrary (the final ordering happens after the exact scores are back)
    int* s_slot = hist;   // G <= 256 counters
    for (int g = tid; g < P.req_G; g += kFinThreads) s_slot[g] = 0;
    for (int g = 0; g < P.req_G; ++g)
      for (int j = tid; j < P.req_m; j += kFinThreads) P.req_out[((size_t)g * P.Q + q) * P.req_m + j] = -1;
    __syncthreads();
    for (int i = tid; i < ncand; i += kFinThreads) {
      int64_t start;
      const int64_t id = (int64_t)(uint32_t)idx[i];
      const int g = shard_of_row(id, P.n_total, P.req_G, &start);
      const int slot = atomicAdd(&s_slot[g], 1);
      // a shard listed at most req_m candidates, so its slots cannot overflow
      if (slot < P.req_m) P.req_out[((size_t)g * P.Q + q) * P.req_m + slot] = (int32_t)(id - start);
    }
    if (tid == 0) P.bound_out[q] = reduced ? max(T, Tstar) : Tstar;
    if (P.scat.world) {
      // the rows just written (scattered 4-byte stores, kept: finalize_exact reads them) go to their
      // shards as whole rows, coalesced, straight into the shard's receive buffer
      __syncthreads();
      for (int g = 0; g < P.req_G; ++g) {
        const int32_t* src = P.req_out + ((size_t)g * P.Q + q) * P.req_m;
        int32_t* dst = scatter_row(P.scat, static_cast<int32_t*>(nullptr), (long long)g * P.Q + q, P.req_m);
        for (int j = tid; j < P.req_m; j += kFinThreads) dst[j] = __ldcg(src + j);
      }
    }
    return;
  }
  // every document that is NOT rescored below has a bf16 score <= t_bf
  const uint32_t Tmax = reduced ? max(T, Tstar) : Tstar;
  const float t_bf = key_to_float(Tmax);

  // 2. exact fp32 rescoring: one warp per candidate, 128-bit loads of the corpus row
  float* exact = reinterpret_cast<float*>(keys);  // keys are no longer needed after selection
  __syncthreads();
  const float qi = (P.score == QST_SCORE_COS && P.q_inv) ? P.q_inv[q] : 1.0f;
  const bool vec4 = (P.D % 4) == 0 && ((reinterpret_cast<uintptr_t>(P.c_f32) & 15u) == 0);
  const bool sq = P.score == QST_SCORE_EUCLID;
  for (int j = warp; j < ncand; j += 2 * kFinWarps) {
    const int j1 = j + kFinWarps;
    const int c0 = idx[j];
    const float* r0 = P.c_f32 + (size_t)c0 * P.D;
    if (j1 < ncand) {
      const int c1 = idx[j1];
      const float* r1 = P.c_f32 + (size_t)c1 * P.D;
      float d0, d1;
      if (sq) warp_dot_f32_x2<true>(qrow, r0, r1, P.D, vec4, lane, d0, d1);
      else warp_dot_f32_x2<false>(qrow, r0, r1, P.D, vec4, lane, d0, d1);
      if (lane == 0) {
        exact[j] = apply_score(d0, P.score, qi, P.c_inv, c0);
        exact[j1] = apply_score(d1, P.score, qi, P.c_inv, c1);
      }
    } else {
      const float d0 = sq ? warp_dot_f32<true>(qrow, r0, P.D, vec4, lane) : warp_dot_f32<false>(qrow, r0, P.D, vec4, lane);
      if (lane == 0) exact[j] = apply_score(d0, P.score, qi, P.c_inv, c0);
    }
  }
  // pad to a power of two for the sort
  int n2 = 1;
  while (n2 < ncand) n2 <<= 1;
  __syncthreads();
  for (int i = ncand + tid; i < n2; i += kFinThreads) { exact[i] = -INFINITY; idx[i] = -1; }
  block_bitonic_sort(exact, idx, n2);

  // 3. emit
  const int kk = P.k;
  for (int j = tid; j < kk; j += kFinThreads) {
    const bool ok = j < ncand;
    P.out_val[(size_t)q * kk + j] = ok ? exact[j] : -INFINITY;
    P.out_idx[(size_t)q * kk + j] = ok ? (int64_t)idx[j] + P.idx_offset : (int64_t)-1;
  }
  if (P.out_margin) {
    // euclid: the tensor-core keys live in "2 q.c - ||c||^2" space; bring the k-th best back there
    float kth = ncand >= kk ? exact[kk - 1] : 0.f;
    if (sq && ncand >= kk && warp == 0) {
      const float d2 = warp_dot_f32<true>(qrow, P.c_f32 + (size_t)idx[kk - 1] * P.D, P.D, vec4, lane);
      kth = qnorm2 - d2;
    }
    if (tid == 0) {
      float margin = INFINITY;
      if (ncand >= kk && t_bf > -INFINITY) {
        // rigorous bound on |bf16 tensor-core score - exact score| for this query against any row:
        //   |dq.c| + |q.dc| + |dq.dc| <= eq*cn + qn*ec + eq*ec      (Cauchy-Schwarz)
        //   + fp32 accumulation inside the tensor core: <= D * 2^-23 * (sum of |terms|)
        float eps = 0.f;
        if (P.q_err && P.c_stats) {
          const float eq = P.q_err[q], ec = P.c_stats[0], cn = P.c_stats[1];
          float qn = sqrtf(qnorm2);
          if (P.score == QST_SCORE_COS) qn = qnorm2 > 0.f ? 1.0f : 0.f;
          if (sq) qn *= 2.0f;   // the query operand is 2q
          eps = eq * cn + qn * ec + eq * ec + (float)P.D * 1.2e-7f * (qn * cn + (sq ? cn * cn : 0.f));
          if (sq) eps += 1e-5f * cn * cn;   // fp32 evaluation of ||c||^2 and of the exact distance
        }
        margin = kth - (t_bf + eps);
        // exactly 0 (e.g. an all-zero query: every score is 0) is "not certified" like any other
        // non-positive margin; the value 0 itself is reserved for rescan_emit_kernel's mark
        if (margin == 0.f) margin = -1e-30f;
      }
      P.out_margin[q] = margin;
    }
  }
}

// ------------------------------------------------------------------------------------------
// K6: merge G descending lists of k entries per query by rank computation (binary searches).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) merge_topk_kernel(const float* __restrict__ vals, const int64_t* __restrict__ idx,
                                                         int G, int64_t Q, int k, float* __restrict__ out_val,
                                                         int64_t* __restrict__ out_idx) {
  const int64_t q = blockIdx.x;
  const int total = G * k;
  // default-fill (lists may hold fewer than k valid entries in total)
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    out_val[q * k + j] = -INFINITY;
    out_idx[q * k + j] = -1;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int g = e / k, j = e - g * k;
    const size_t base = ((size_t)g * Q + q) * k;
    const float s = vals[base + j];
    const int64_t id = idx[base + j];
    if (id < 0) continue;
    int rank = j;
    for (int h = 0; h < G; ++h) {
      if (h == g) continue;
      const size_t hb = ((size_t)h * Q + q) * k;
      // number of entries of list h that beat (s, id): lists are sorted in `beats` order
      int lo = 0, hi = k;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (beats(vals[hb + mid], idx[hb + mid], s, id)) lo = mid + 1; else hi = mid;
      }
      rank += lo;
    }
    if (rank < k) { out_val[q * k + rank] = s; out_idx[q * k + rank] = id; }
  }
}

// ------------------------------------------------------------------------------------------
// Exact re-scan of uncertified queries (margin <= 0).  Brute force fp32 on CUDA cores with the
// already known k-th best exact score as the acceptance threshold, so only a handful of rows
// per query are appended; the collected rows are then sorted and the top k re-emitted.
//   step 1  rescan_collect:  grid (corpus blocks, flagged-query batches)
//   step 2  rescan_emit:     one CTA per query
// ------------------------------------------------------------------------------------------
constexpr int kRescanCap = 2048;        // collected rows per flagged query (>= k + ties at the threshold)
constexpr int kRescanMaxFlagged = 8192; // flagged queries served per call (bounds the scratch to 128 MB);
                                        // any beyond that keep margin <= 0 and can be served by another call

struct RescanScratch {
  int n_flagged;
  int pad[3];
};

__global__ void rescan_list_kernel(const float* margin, int Q, int* flagged, RescanScratch* hdr, int* counts) {
  // single CTA: list of the flagged queries (at most kRescanMaxFlagged; a query this pass repairs
  // gets margin = +inf, so the next pass of qst_exact_rescan lists the ones left over)
  __shared__ int s_n;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  for (int base = 0; base < Q; base += blockDim.x) {
    const int q = base + threadIdx.x;
    // margin == 0 exactly is the "cannot be repaired" mark of rescan_emit_kernel (more than
    // kRescanCap rows tie at the k-th score): not listed again
    const bool f = q < Q && !(margin[q] > 0.f) && margin[q] != 0.f;
    if (f) { const int p = atomicAdd(&s_n, 1); if (p < kRescanMaxFlagged) flagged[p] = q; }
  }
  __syncthreads();
  if (threadIdx.x == 0) hdr->n_flagged = s_n < kRescanMaxFlagged ? s_n : kRescanMaxFlagged;
  for (int i = threadIdx.x; i < Q; i += blockDim.x) counts[i] = 0;
}

constexpr int kRescanBatch = 16;   // most flagged queries scored per corpus pass (register accumulators)

// One corpus pass per batch of kRescanBatch flagged queries: the batch's query rows sit in shared
// memory, every warp streams corpus rows (each row read once per batch) and keeps one accumulator
// per query.  The per-query operation order is exactly warp_dot_f32's, so scores are bit-identical
// to the ones K3 produced for the same (query, row).
template <int kRescanBatch>
__global__ void __launch_bounds__(256) rescan_collect_kernel(int N, int D, int k, int score, const float* __restrict__ q_f32,
                                                             const float* __restrict__ q_inv, const float* __restrict__ c_f32,
                                                             const float* __restrict__ c_inv, const float* __restrict__ thr_base,
                                                             int thr_stride, int thr_off,
                                                             const int* __restrict__ flagged, const RescanScratch* hdr,
                                                             int* counts, float* coll_val, int* coll_idx) {
  extern __shared__ float s_q[];  // [kRescanBatch][D]
  __shared__ float s_thr[kRescanBatch], s_qi[kRescanBatch];
  __shared__ int s_qid[kRescanBatch];
  const int nf = hdr->n_flagged;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool vec4 = (D % 4) == 0 && ((reinterpret_cast<uintptr_t>(c_f32) & 15u) == 0);
  const bool sq = score == QST_SCORE_EUCLID;
  for (int f0 = 0; f0 < nf; f0 += kRescanBatch) {
    const int nb = min(kRescanBatch, nf - f0);
    __syncthreads();
    for (int i = threadIdx.x; i < kRescanBatch * D; i += blockDim.x) {
      const int b = i / D, j = i - b * D;
      s_q[i] = b < nb ? q_f32[(size_t)flagged[f0 + b] * D + j] : 0.f;
    }
    if (threadIdx.x < kRescanBatch) {
      const int b = threadIdx.x;
      const int q = b < nb ? flagged[f0 + b] : 0;
      s_qid[b] = q;
      s_thr[b] = b < nb ? thr_base[(size_t)q * thr_stride + thr_off] : INFINITY;  // exact k-th best so far
      s_qi[b] = (score == QST_SCORE_COS && q_inv) ? q_inv[q] : 1.0f;
    }
    __syncthreads();
    for (int row = blockIdx.x * 8 + warp; row < N; row += gridDim.x * 8) {
      const float* crow = c_f32 + (size_t)row * D;
      float acc[kRescanBatch];
#pragma unroll
      for (int b = 0; b < kRescanBatch; ++b) acc[b] = 0.f;
      if (vec4) {
        const float4* c4 = reinterpret_cast<const float4*>(crow);
        for (int i = lane; i < D / 4; i += 32) {
          const float4 c = __ldg(c4 + i);
#pragma unroll
          for (int b = 0; b < kRescanBatch; ++b) {
            const float4 a = reinterpret_cast<const float4*>(s_q + (size_t)b * D)[i];
            if (sq) {
              acc[b] = acc1<true>(a.x, c.x, acc[b]); acc[b] = acc1<true>(a.y, c.y, acc[b]);
              acc[b] = acc1<true>(a.z, c.z, acc[b]); acc[b] = acc1<true>(a.w, c.w, acc[b]);
            } else {
              acc[b] = acc1<false>(a.x, c.x, acc[b]); acc[b] = acc1<false>(a.y, c.y, acc[b]);
              acc[b] = acc1<false>(a.z, c.z, acc[b]); acc[b] = acc1<false>(a.w, c.w, acc[b]);
            }
          }
        }
      } else {
        for (int i = lane; i < D; i += 32) {
          const float c = __ldg(crow + i);
#pragma unroll
          for (int b = 0; b < kRescanBatch; ++b)
            acc[b] = sq ? acc1<true>(s_q[(size_t)b * D + i], c, acc[b]) : acc1<false>(s_q[(size_t)b * D + i], c, acc[b]);
        }
      }
#pragma unroll
      for (int b = 0; b < kRescanBatch; ++b) {
        const float dot = warp_sum(acc[b]);
        if (lane == 0 && b < nb) {
          const float sc = apply_score(dot, score, s_qi[b], c_inv, row);
          if (sc >= s_thr[b]) {
            const int q = s_qid[b];
            const int p = atomicAdd(&counts[q], 1);
            if (p < kRescanCap) {
              coll_val[(size_t)(f0 + b) * kRescanCap + p] = sc;
              coll_idx[(size_t)(f0 + b) * kRescanCap + p] = row;
            }
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(kFinThreads) rescan_emit_kernel(int k, int64_t idx_offset, const int* __restrict__ flagged,
                                                                  const RescanScratch* hdr, const int* __restrict__ counts,
                                                                  float* coll_val, int* coll_idx, float* out_val,
                                                                  int64_t* out_idx, float* margin, int partial,
                                                                  int* overflow) {
  __shared__ float sc[kRescanCap];
  __shared__ int32_t ix[kRescanCap];
  const int nf = hdr->n_flagged;
  for (int f = blockIdx.x; f < nf; f += gridDim.x) {
    const int q = flagged[f];
    int n = counts[q];
    if (n > kRescanCap) n = kRescanCap;
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    __syncthreads();
    for (int i = threadIdx.x; i < n2; i += kFinThreads) {
      sc[i] = i < n ? coll_val[(size_t)f * kRescanCap + i] : -INFINITY;
      ix[i] = i < n ? coll_idx[(size_t)f * kRescanCap + i] : -1;
    }
    block_bitonic_sort(sc, ix, n2);
    if (partial) {
      // one SHARD of a sharded corpus: emit whatever this shard holds at or above the threshold
      // (possibly fewer than k rows), padded; the owner of the query merges the shards' lists
      for (int j = threadIdx.x; j < k; j += kFinThreads) {
        out_val[(size_t)q * k + j] = j < n ? sc[j] : -INFINITY;
        out_idx[(size_t)q * k + j] = j < n ? (int64_t)ix[j] + idx_offset : (int64_t)-1;
      }
      if (threadIdx.x == 0 && overflow) overflow[q] = counts[q] > kRescanCap ? 1 : 0;
    } else if (n >= k) {
      for (int j = threadIdx.x; j < k; j += kFinThreads) {
        out_val[(size_t)q * k + j] = sc[j];
        out_idx[(size_t)q * k + j] = (int64_t)ix[j] + idx_offset;
      }
      // exact now (unless more than kRescanCap rows tie at the threshold)
      if (threadIdx.x == 0) margin[q] = counts[q] <= kRescanCap ? INFINITY : 0.f;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// Fully sharded fp32 master (every GPU keeps ITS shard only).
//   shard side : rescore_requests_kernel  -- exact fp32 scores of the rows an owner asked for, from the
//                                            shard's own fp32 rows and the all-gathered fp32 queries
//   owner side : finalize_exact_kernel    -- order the exact scores that came back, emit top k + margin
// Scores are produced by warp_dot_f32 / apply_score exactly as K3 does, so a ranking is bit-identical
// to the single-GPU one.  euclid_score: the shard returns ||q-c||^2 and the owner applies
// 1/(1+sqrt(.)) (it needs the squared distance of the k-th document for the certificate).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFinThreads) rescore_requests_kernel(int m, int D, int score, const int32_t* __restrict__ req,
                                                                       const float* __restrict__ q_f32,
                                                                       const float* __restrict__ q_inv,
                                                                       const float* __restrict__ c_f32,
                                                                       const float* __restrict__ c_inv, float* __restrict__ out,
                                                                       const Scatter scat) {
  extern __shared__ float rq_row[];
  const int64_t q = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < D; i += kFinThreads) rq_row[i] = q_f32[(size_t)q * D + i];
  __syncthreads();
  const float qi = (score == QST_SCORE_COS && q_inv) ? q_inv[q] : 1.0f;
  const bool vec4 = (D % 4) == 0 && ((reinterpret_cast<uintptr_t>(c_f32) & 15u) == 0);
  const bool sq = score == QST_SCORE_EUCLID;
  const int32_t* rq = req + (size_t)q * m;
  float* o = out + (size_t)q * m;
  for (int j = warp; j < m; j += 2 * kFinWarps) {
    const int j1 = j + kFinWarps;
    const int r0 = rq[j];
    const int r1 = j1 < m ? rq[j1] : -1;
    if (r0 >= 0 && r1 >= 0) {
      float d0, d1;
      if (sq) warp_dot_f32_x2<true>(rq_row, c_f32 + (size_t)r0 * D, c_f32 + (size_t)r1 * D, D, vec4, lane, d0, d1);
      else warp_dot_f32_x2<false>(rq_row, c_f32 + (size_t)r0 * D, c_f32 + (size_t)r1 * D, D, vec4, lane, d0, d1);
      if (lane == 0) {
        o[j] = sq ? d0 : apply_score(d0, score, qi, c_inv, r0);
        o[j1] = sq ? d1 : apply_score(d1, score, qi, c_inv, r1);
      }
    } else {
      if (r0 >= 0) {
        const float d0 = sq ? warp_dot_f32<true>(rq_row, c_f32 + (size_t)r0 * D, D, vec4, lane)
                            : warp_dot_f32<false>(rq_row, c_f32 + (size_t)r0 * D, D, vec4, lane);
        if (lane == 0) o[j] = sq ? d0 : apply_score(d0, score, qi, c_inv, r0);
      } else if (lane == 0) {
        o[j] = -INFINITY;
      }
      if (j1 < m) {
        if (r1 >= 0) {
          const float d1 = sq ? warp_dot_f32<true>(rq_row, c_f32 + (size_t)r1 * D, D, vec4, lane)
                              : warp_dot_f32<false>(rq_row, c_f32 + (size_t)r1 * D, D, vec4, lane);
          if (lane == 0) o[j1] = sq ? d1 : apply_score(d1, score, qi, c_inv, r1);
        } else if (lane == 0) {
          o[j1] = -INFINITY;
        }
      }
    }
  }
  if (scat.world) {   // the finished row goes to the owner of the query, coalesced (see Scatter)
    __syncthreads();
    float* dst = scatter_row(scat, static_cast<float*>(nullptr), q, m);
    for (int j = tid; j < m; j += kFinThreads) dst[j] = __ldcg(o + j);
  }
}

// Same work, ONE WARP per (owner, query) row with the query row held in registers (D <= 1024, D % 4 == 0):
// a shard sees G * q_own rows with only ~k'/G requested entries each, so a CTA per row mostly pays for
// staging the query in shared memory and for its barriers.  The per-row operation order is exactly
// warp_dot_f32's (lane-strided float4 chunks, four fmaf per chunk in x, y, z, w order, xor-tree sum), so
// the scores are bit-identical to the CTA kernel's and to K3's.
constexpr int kRescoreRegChunks = 8;   // float4 chunks per lane: D <= 32 * 4 * 8 = 1024

template <bool SQ>
__global__ void __launch_bounds__(kFinThreads) rescore_requests_warp_kernel(int64_t rows, int m, int D, int score,
                                                                            const int32_t* __restrict__ req,
                                                                            const float* __restrict__ q_f32,
                                                                            const float* __restrict__ q_inv,
                                                                            const float* __restrict__ c_f32,
                                                                            const float* __restrict__ c_inv,
                                                                            float* __restrict__ out, const Scatter scat) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t q = (int64_t)blockIdx.x * kFinWarps + warp;
  if (q >= rows) return;
  const int n4 = D / 4;
  float4 qa[kRescoreRegChunks];
  const float4* q4 = reinterpret_cast<const float4*>(q_f32 + (size_t)q * D);
#pragma unroll
  for (int c = 0; c < kRescoreRegChunks; ++c) {
    const int i = lane + 32 * c;
    qa[c] = i < n4 ? __ldg(q4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float qi = (score == QST_SCORE_COS && q_inv) ? q_inv[q] : 1.0f;
  const int32_t* rq = req + (size_t)q * m;
  float* o = scatter_row(scat, out, q, m);   // local row, or the row in the query owner's receive buffer
  for (int j0 = 0; j0 < m; j0 += 32) {
    const int jj = j0 + lane;
    const int my = jj < m ? rq[jj] : -1;
    float res = -INFINITY;                   // lane j keeps the score of entry j0 + j: one coalesced store per 32
    unsigned valid = __ballot_sync(0xffffffffu, my >= 0);
    while (valid) {
      // two requested rows per trip: twice the loads in flight
      const int l0 = __ffs(valid) - 1;
      valid &= valid - 1;
      const int l1 = valid ? __ffs(valid) - 1 : -1;
      if (l1 >= 0) valid &= valid - 1;
      const int r0 = __shfl_sync(0xffffffffu, my, l0);
      const int r1 = l1 >= 0 ? __shfl_sync(0xffffffffu, my, l1) : r0;
      const float4* p0 = reinterpret_cast<const float4*>(c_f32 + (size_t)r0 * D);
      const float4* p1 = reinterpret_cast<const float4*>(c_f32 + (size_t)r1 * D);
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int c = 0; c < kRescoreRegChunks; ++c) {
        const int i = lane + 32 * c;
        if (i < n4) {
          const float4 x = __ldg(p0 + i);
          const float4 y = __ldg(p1 + i);
          const float4 a = qa[c];
          a0 = acc1<SQ>(a.x, x.x, a0); a0 = acc1<SQ>(a.y, x.y, a0); a0 = acc1<SQ>(a.z, x.z, a0); a0 = acc1<SQ>(a.w, x.w, a0);
          a1 = acc1<SQ>(a.x, y.x, a1); a1 = acc1<SQ>(a.y, y.y, a1); a1 = acc1<SQ>(a.z, y.z, a1); a1 = acc1<SQ>(a.w, y.w, a1);
        }
      }
      a0 = warp_sum(a0);
      a1 = warp_sum(a1);
      if (lane == l0) res = SQ ? a0 : apply_score(a0, score, qi, c_inv, r0);
      if (lane == l1) res = SQ ? a1 : apply_score(a1, score, qi, c_inv, r1);
    }
    if (jj < m) o[jj] = res;
  }
}

// ------------------------------------------------------------------------------------------
// Dense exact scores [Q, N] in fp32 on CUDA cores, with the arithmetic of K3 (warp_dot_f32 +
// apply_score): the direct-call form of cos_sim / dot_score / euclidean_score
// (/root/reference/dataset/positive_examples_selection.py:55, dataset/quadruplet_dataset.py:229-234,
// training/main.py:57 call them on small inputs and want the matrix).  Not a hot path: retrieval-sized
// inputs go through K2/K3 and never materialise the matrix.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFinThreads) dense_scores_kernel(int64_t N, int D, int score, const float* __restrict__ q_f32,
                                                                   const float* __restrict__ q_inv,
                                                                   const float* __restrict__ c_f32,
                                                                   const float* __restrict__ c_inv, float* __restrict__ out) {
  extern __shared__ float dq_row[];
  const int64_t q = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < D; i += kFinThreads) dq_row[i] = q_f32[(size_t)q * D + i];
  __syncthreads();
  const float qi = (score == QST_SCORE_COS && q_inv) ? q_inv[q] : 1.0f;
  const bool vec4 = (D % 4) == 0 && ((reinterpret_cast<uintptr_t>(c_f32) & 15u) == 0);
  const bool sq = score == QST_SCORE_EUCLID;
  for (int64_t row = (int64_t)blockIdx.x * kFinWarps + warp; row < N; row += (int64_t)gridDim.x * kFinWarps) {
    const float* crow = c_f32 + (size_t)row * D;
    const float d = sq ? warp_dot_f32<true>(dq_row, crow, D, vec4, lane) : warp_dot_f32<false>(dq_row, crow, D, vec4, lane);
    if (lane == 0) out[(size_t)q * N + row] = apply_score(d, score, qi, c_inv, (int)row);
  }
}

constexpr int kExactMax = 2048;   // entries one query can get back (k' <= 2048)

__global__ void __launch_bounds__(kFinThreads) finalize_exact_kernel(int Q, int G, int m, int k, int score, int D, int64_t n_total,
                                                                     const int32_t* __restrict__ req,
                                                                     const float* __restrict__ exact_in,
                                                                     const uint32_t* __restrict__ bound,
                                                                     const float* __restrict__ q_f32,
                                                                     const float* __restrict__ q_err,
                                                                     const float* __restrict__ c_stats, float* out_val,
                                                                     int64_t* out_idx, float* out_margin) {
  __shared__ float sc[kExactMax];
  __shared__ int32_t ix[kExactMax];
  __shared__ float raw[kExactMax];     // euclid: squared distances in arrival order
  __shared__ int32_t raw_id[kExactMax];
  __shared__ int s_n;
  __shared__ float s_red[kFinWarps];
  __shared__ float s_kth_d2;
  const int q = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool sq = score == QST_SCORE_EUCLID;
  if (tid == 0) s_n = 0;
  float qq = 0.f;
  for (int i = tid; i < D; i += kFinThreads) { const float v = q_f32[(size_t)q * D + i]; qq = fmaf(v, v, qq); }
  qq = warp_sum(qq);
  if (lane == 0) s_red[warp] = qq;
  __syncthreads();
  float qnorm2 = 0.f;
  for (int w = 0; w < kFinWarps; ++w) qnorm2 += s_red[w];
  for (int e = tid; e < G * m; e += kFinThreads) {
    const int g = e / m, j = e - g * m;
    const size_t at = ((size_t)g * Q + q) * m + j;
    const int32_t r = req[at];
    if (r < 0) continue;
    const int p = atomicAdd(&s_n, 1);
    if (p < kExactMax) {
      const float x = exact_in[at];
      const int32_t id = (int32_t)(shard_start_of(g, n_total, G) + r);
      sc[p] = sq ? 1.0f / (1.0f + sqrtf(x)) : x;
      ix[p] = id;
      if (sq) { raw[p] = x; raw_id[p] = id; }
    }
  }
  __syncthreads();
  const int n = s_n < kExactMax ? s_n : kExactMax;
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  for (int i = n + tid; i < n2; i += kFinThreads) { sc[i] = -INFINITY; ix[i] = -1; }
  block_bitonic_sort(sc, ix, n2);
  for (int j = tid; j < k; j += kFinThreads) {
    const bool ok = j < n;
    out_val[(size_t)q * k + j] = ok ? sc[j] : -INFINITY;
    out_idx[(size_t)q * k + j] = ok ? (int64_t)ix[j] : (int64_t)-1;
  }
  if (out_margin) {
    if (sq && n >= k) {
      const int32_t want = ix[k - 1];
      for (int i = tid; i < n; i += kFinThreads)
        if (raw_id[i] == want) s_kth_d2 = raw[i];
    }
    __syncthreads();
    if (tid == 0) {
      const float t_bf = key_to_float(bound[q]);
      float margin = INFINITY;
      if (n >= k && t_bf > -INFINITY) {
        const float kth = sq ? qnorm2 - s_kth_d2 : sc[k - 1];
        float eps = 0.f;
        if (q_err && c_stats) {   // same bound as finalize_kernel
          const float eq = q_err[q], ec = c_stats[0], cn = c_stats[1];
          float qn = sqrtf(qnorm2);
          if (score == QST_SCORE_COS) qn = qnorm2 > 0.f ? 1.0f : 0.f;
          if (sq) qn *= 2.0f;
          eps = eq * cn + qn * ec + eq * ec + (float)D * 1.2e-7f * (qn * cn + (sq ? cn * cn : 0.f));
          if (sq) eps += 1e-5f * cn * cn;
        }
        margin = kth - (t_bf + eps);
        if (margin == 0.f) margin = -1e-30f;   // 0 is the re-scan's "cannot be repaired" mark
      }
      out_margin[q] = margin;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Candidate lists of the sharded path, ONE WARP per query row (qst_select_candidates for m <= 256).
// A shard scores all G*q_own queries, so this kernel runs over G times more rows than anything on the
// owner's side, each holding only a few dozen entries (S unit buffers of ~kunit entries): a CTA per row
// spends its time on launch and barriers.  A warp gathers the row's unit buffers into a private
// shared-memory window, keeps the m largest keys (exact radix select whenever the window fills up) and
// writes the list + trailer (bound of everything not listed, count).  Entry order is arbitrary.
// ------------------------------------------------------------------------------------------
constexpr int kSelWarpWindow = 512;   // entries per warp window (m <= 256 leaves room to append between selects)
constexpr int kSelWarps = 8;

__global__ void __launch_bounds__(kSelWarps * 32) select_rows_warp_kernel(const FinParams P) {
  __shared__ uint2 s_win[kSelWarps][kSelWarpWindow];
  __shared__ int s_hist[kSelWarps][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * kSelWarps + warp;
  if (q >= P.Q) return;
  const int mt = q / P.rows_per_unit, r = q % P.rows_per_unit;
  const int m = P.kprime;
  uint2* win = s_win[warp];
  int* hist = s_hist[warp];
  int fill = 0;
  bool reduced = false;
  uint32_t T = 0u, Tstar = 0u;
  for (int st = 0; st < P.stripes; ++st) {
    const size_t urow = (size_t)(st * P.m_tiles + mt) * P.rows_per_unit + r;
    const int cnt = P.unit_cnt[urow];
    Tstar = max(Tstar, P.unit_thr[urow]);
    const uint2* src = P.unit_cand + urow * (size_t)P.cap;
    int pos = 0;
    while (pos < cnt) {
      if (fill == kSelWarpWindow) {
        __syncwarp();
        T = warp_select_compact(win, fill, m, hist, lane);
        fill = m;
        reduced = true;
      }
      const int take = min(kSelWarpWindow - fill, cnt - pos);
      for (int i = lane; i < take; i += 32) win[fill + i] = src[pos + i];
      fill += take;
      pos += take;
    }
  }
  __syncwarp();
  if (fill > m) {
    T = warp_select_compact(win, fill, m, hist, lane);
    fill = m;
    reduced = true;
  }
  __syncwarp();
  uint2* dst = scatter_row(P.scat, P.sel_out, q, m + 1);
  for (int i = lane; i < m; i += 32) {
    uint2 e = make_uint2(0u, 0xffffffffu);
    if (i < fill) { e = win[i]; e.y = (uint32_t)((int64_t)(int32_t)e.y + P.idx_offset); }
    dst[i] = e;
  }
  if (lane == 0) dst[m] = make_uint2(reduced ? max(T, Tstar) : Tstar, (uint32_t)fill);
}

__global__ void unpack_list_trailers_kernel(const uint2* __restrict__ lists, int64_t n, int m, uint32_t* thr, int* cnt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint2 t = lists[i * (m + 1) + m];
  thr[i] = t.x;
  cnt[i] = (int)t.y;
}

// Barrier over peer memory between the ranks of a node, stream-ordered: thread t signals rank t
// (release at system scope: everything this stream wrote before, peer writes included, is visible to
// whoever sees the flag) and waits for rank t's signal in the local flag array.  Epochs only grow, so a
// rank that is already one barrier ahead cannot be missed.  A rank that never arrives (crashed peer) would
// hang the GPU: after `timeout_ns` (120 s unless QST_BARRIER_TIMEOUT_S says otherwise) the kernel traps instead.
__global__ void peer_barrier_kernel(const Scatter flags, uint32_t epoch, unsigned long long timeout_ns) {
  const int t = threadIdx.x;
  if (t >= flags.world) return;
  __threadfence_system();
  uint32_t* remote = reinterpret_cast<uint32_t*>(flags.base[t]) + flags.rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
  const uint32_t* mine = reinterpret_cast<const uint32_t*>(flags.base[flags.rank]) + t;
  unsigned long long t0 = 0, now = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if ((int32_t)(v - epoch) >= 0) break;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (now - t0 > timeout_ns) __trap();
    __nanosleep(200);
  }
}

}  // namespace qst

using namespace qst;

static int scatter_from(const qst_scatter* dst, Scatter* out) {
  *out = Scatter{};
  if (dst == nullptr) return QST_OK;
  QST_CHECK_ARG(dst->world >= 1 && dst->world <= QST_MAX_WORLD && dst->rank >= 0 && dst->rank < dst->world &&
                    dst->rows_per_block >= 1,
                "scatter: bad descriptor world=%d rank=%d rows_per_block=%lld", dst->world, dst->rank,
                (long long)dst->rows_per_block);
  for (int r = 0; r < dst->world; ++r) {
    QST_CHECK_ARG(dst->base[r] != nullptr, "scatter: null base pointer for rank %d", r);
    out->base[r] = dst->base[r];
  }
  out->world = dst->world; out->rank = dst->rank; out->rows_per_block = dst->rows_per_block;
  return QST_OK;
}

extern "C" int qst_peer_barrier(const qst_scatter* flags, uint32_t epoch, qst_stream_t stream) {
  QST_CHECK_ARG(flags != nullptr, "peer_barrier: null descriptor");
  Scatter f;
  if (int rc = scatter_from(flags, &f)) return rc;
  static unsigned long long timeout_ns = 0;
  if (timeout_ns == 0) {
    const char* e = getenv("QST_BARRIER_TIMEOUT_S");
    const double sec = e ? atof(e) : 120.0;
    timeout_ns = (unsigned long long)((sec > 0.001 ? sec : 120.0) * 1e9);
  }
  peer_barrier_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(f, epoch, timeout_ns);
  QST_LAUNCH_CHECK();
  return QST_OK;
}

static int finalize_topk_impl(const qst_topk_plan* plan, int kprime, int refine, const void* workspace, const float* q_f32,
                              const float* q_inv, const float* q_err, const float* c_f32, const float* c_inv,
                              const float* c_stats, int64_t idx_offset, float* out_val, int64_t* out_idx,
                              float* out_margin, qst_stream_t stream);

extern "C" int qst_finalize_topk(const qst_topk_plan* plan, const void* workspace, const float* q_f32,
                                 const float* q_inv, const float* q_err, const float* c_f32, const float* c_inv,
                                 const float* c_stats, int64_t idx_offset, float* out_val, int64_t* out_idx,
                                 float* out_margin, qst_stream_t stream) {
  QST_CHECK_ARG(plan != nullptr, "finalize_topk: null plan");
  return finalize_topk_impl(plan, plan->kprime, 0, workspace, q_f32, q_inv, q_err, c_f32, c_inv, c_stats, idx_offset,
                            out_val, out_idx, out_margin, stream);
}

// Two-pass K3: the first pass rescoring only `kprime_first` (< plan->kprime) candidates per query certifies
// all but a fraction of a percent of the queries at config 3 (k' = 176: 33 of 10 000 left, against 14 % less
// gathered data than k' = 224); the second pass repeats the work with the plan's full k' for the queries
// whose margin is still <= 0 and leaves the others alone.  Same result as one pass with the full k':
// a certified ranking is THE exact ranking whatever k' produced it.
extern "C" int qst_finalize_topk_adaptive(const qst_topk_plan* plan, int kprime_first, const void* workspace,
                                          const float* q_f32, const float* q_inv, const float* q_err,
                                          const float* c_f32, const float* c_inv, const float* c_stats,
                                          int64_t idx_offset, float* out_val, int64_t* out_idx, float* out_margin,
                                          qst_stream_t stream) {
  QST_CHECK_ARG(plan != nullptr && out_margin != nullptr, "finalize_topk_adaptive: null plan / margin");
  QST_CHECK_ARG(q_err != nullptr && c_stats != nullptr, "finalize_topk_adaptive: needs the certificate inputs");
  if (kprime_first < plan->k || kprime_first >= plan->kprime)
    return finalize_topk_impl(plan, plan->kprime, 0, workspace, q_f32, q_inv, q_err, c_f32, c_inv, c_stats, idx_offset,
                              out_val, out_idx, out_margin, stream);
  int rc = finalize_topk_impl(plan, kprime_first, 0, workspace, q_f32, q_inv, q_err, c_f32, c_inv, c_stats, idx_offset,
                              out_val, out_idx, out_margin, stream);
  if (rc) return rc;
  return finalize_topk_impl(plan, plan->kprime, 1, workspace, q_f32, q_inv, q_err, c_f32, c_inv, c_stats, idx_offset,
                            out_val, out_idx, out_margin, stream);
}

static int finalize_topk_impl(const qst_topk_plan* plan, int kprime, int refine, const void* workspace, const float* q_f32,
                              const float* q_inv, const float* q_err, const float* c_f32, const float* c_inv,
                              const float* c_stats, int64_t idx_offset, float* out_val, int64_t* out_idx,
                              float* out_margin, qst_stream_t stream) {
  QST_CHECK_ARG(plan && workspace && q_f32 && c_f32 && out_val && out_idx, "finalize_topk: null argument");
  QST_CHECK_ARG(plan->score != QST_SCORE_COS || (q_inv && c_inv), "finalize_topk: cos score needs inverse norms");
  QST_CHECK_ARG(plan->stripes <= kMaxStripes, "finalize_topk: too many stripes (%d)", plan->stripes);
  const uint8_t* ws = reinterpret_cast<const uint8_t*>(workspace);
  FinParams P{};
  P.Q = (int)plan->Q; P.N = (int)plan->N; P.D = (int)plan->D;
  P.k = plan->k; P.kprime = kprime; P.cap = plan->cap;
  P.refine = refine;
  P.m_tiles = plan->m_tiles; P.stripes = plan->stripes; P.score = plan->score;
  P.rows_per_unit = plan->rows_per_unit;
  int sm_cap = kprime + plan->cap;
  if (sm_cap < 4096) sm_cap = 4096;
  // many short stripes (small query batches): room for what every unit may leave per row, so the
  // gather stays a single parallel pass
  const int expect = plan->stripes * (plan->kunit + 16);
  if (sm_cap < expect) sm_cap = expect < 12288 ? expect : 12288;
  P.sm_cap = sm_cap;
  P.thr_hint = reinterpret_cast<const uint32_t*>(ws + plan->off_thr);
  P.unit_cnt = reinterpret_cast<const int*>(ws + plan->off_cnt);
  P.unit_thr = reinterpret_cast<const uint32_t*>(ws + plan->off_uthr);
  P.unit_cand = reinterpret_cast<const uint2*>(ws + plan->off_cand);
  P.q_f32 = q_f32; P.q_inv = q_inv; P.q_err = q_err; P.c_f32 = c_f32; P.c_inv = c_inv; P.c_stats = c_stats;
  P.idx_offset = idx_offset; P.out_val = out_val; P.out_idx = out_idx; P.out_margin = out_margin;
  const size_t smem = (size_t)sm_cap * 8 + (size_t)kprime * 8 + round_up((size_t)plan->D * 4, 16);
  QST_CHECK_ARG(smem <= 200 * 1024, "finalize_topk: D=%lld too large for the rescoring stage", (long long)plan->D);
  QST_CUDA(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  finalize_kernel<<<(unsigned)plan->Q, kFinThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(P);
  QST_LAUNCH_CHECK();
  return QST_OK;
}

static int select_candidates_impl(const qst_topk_plan* plan, const void* workspace, int m, int64_t idx_offset,
                                  void* out_lists, const qst_scatter* dst, qst_stream_t stream) {
  QST_CHECK_ARG(plan && workspace && (out_lists || dst), "select_candidates: null argument");
  QST_CHECK_ARG(m >= 1 && m <= 2048, "select_candidates: m=%d out of range", m);
  QST_CHECK_ARG(!dst || (int64_t)dst->world * dst->rows_per_block == plan->Q,
                "select_candidates: scatter blocks do not cover the %lld query rows", (long long)plan->Q);
  QST_CHECK_ARG(plan->stripes <= kMaxStripes, "select_candidates: too many stripes (%d)", plan->stripes);
  const uint8_t* ws = reinterpret_cast<const uint8_t*>(workspace);
  FinParams P{};
  P.Q = (int)plan->Q; P.N = (int)plan->N; P.D = 0;
  P.k = m; P.kprime = m; P.cap = plan->cap;
  P.m_tiles = plan->m_tiles; P.stripes = plan->stripes; P.score = plan->score;
  P.rows_per_unit = plan->rows_per_unit;
  int sm_cap = m + plan->cap;
  if (sm_cap < 4096) sm_cap = 4096;
  const int expect = plan->stripes * (plan->kunit + 16);
  if (sm_cap < expect) sm_cap = expect < 12288 ? expect : 12288;
  P.sm_cap = sm_cap;
  P.thr_hint = reinterpret_cast<const uint32_t*>(ws + plan->off_thr);
  P.unit_cnt = reinterpret_cast<const int*>(ws + plan->off_cnt);
  P.unit_thr = reinterpret_cast<const uint32_t*>(ws + plan->off_uthr);
  P.unit_cand = reinterpret_cast<const uint2*>(ws + plan->off_cand);
  P.idx_offset = idx_offset;
  P.sel_out = reinterpret_cast<uint2*>(out_lists);
  if (int rc = scatter_from(dst, &P.scat)) return rc;
  if (dst && !P.sel_out) P.sel_out = reinterpret_cast<uint2*>(dst->base[dst->rank]);   // marks the mode only
  {
    const char* e = getenv("QST_SELECT_CTA");   // 1 forces the CTA-per-row kernel (comparison / debugging)
    if (m <= kSelWarpWindow / 2 && !(e && e[0] == '1')) {
      select_rows_warp_kernel<<<(unsigned)ceil_div(plan->Q, kSelWarps), kSelWarps * 32, 0,
                                reinterpret_cast<cudaStream_t>(stream)>>>(P);
      QST_LAUNCH_CHECK();
      return QST_OK;
    }
  }
  const size_t smem = (size_t)sm_cap * 8 + (size_t)m * 8 + 16;
  QST_CUDA(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  finalize_kernel<<<(unsigned)plan->Q, kFinThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(P);
  QST_LAUNCH_CHECK();
  return QST_OK;
}

extern "C" int qst_select_candidates(const qst_topk_plan* plan, const void* workspace, int m, int64_t idx_offset,
                                     void* out_lists, qst_stream_t stream) {
  QST_CHECK_ARG(out_lists != nullptr, "select_candidates: null output");
  return select_candidates_impl(plan, workspace, m, idx_offset, out_lists, nullptr, stream);
}

extern "C" int qst_select_candidates_scatter(const qst_topk_plan* plan, const void* workspace, int m, int64_t idx_offset,
                                             const qst_scatter* dst, qst_stream_t stream) {
  QST_CHECK_ARG(dst != nullptr, "select_candidates_scatter: null descriptor");
  return select_candidates_impl(plan, workspace, m, idx_offset, nullptr, dst, stream);
}

extern "C" int qst_finalize_lists(int64_t Q, int G, int m, int k, int kprime, int score, int64_t D,
                                  const void* lists, const float* q_f32, const float* q_inv, const float* q_err,
                                  const float* c_f32, const float* c_inv, const float* c_stats, float* out_val,
                                  int64_t* out_idx, float* out_margin, void* scratch, qst_stream_t stream) {
  QST_CHECK_ARG(lists && q_f32 && c_f32 && out_val && out_idx && scratch, "finalize_lists: null argument");
  QST_CHECK_ARG(Q >= 1 && G >= 1 && G <= kMaxStripes && m >= 1 && k >= 1 && kprime >= k && kprime <= 2048,
                "finalize_lists: bad shape Q=%lld G=%d m=%d k=%d kprime=%d", (long long)Q, G, m, k, kprime);
  QST_CHECK_ARG(score != QST_SCORE_COS || (q_inv && c_inv), "finalize_lists: cos score needs inverse norms");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // unpack the (bound, count) trailer of every received list into the arrays K3 reads
  uint32_t* thr = reinterpret_cast<uint32_t*>(scratch);
  int* cnt = reinterpret_cast<int*>(thr + (size_t)G * Q);
  unpack_list_trailers_kernel<<<(unsigned)ceil_div((int64_t)G * Q, 256), 256, 0, st>>>(
      reinterpret_cast<const uint2*>(lists), (int64_t)G * Q, m, thr, cnt);
  QST_LAUNCH_CHECK();
  FinParams P{};
  P.Q = (int)Q; P.N = 0; P.D = (int)D;
  P.k = k; P.kprime = kprime; P.cap = m + 1;          // row pitch of a received list
  P.m_tiles = 1; P.stripes = G; P.score = score; P.rows_per_unit = (int)Q;
  int sm_cap = kprime + m + 1;
  if (sm_cap < G * m) sm_cap = G * m;                  // gather everything in one pass when it fits
  if (sm_cap < 4096) sm_cap = 4096;
  if (sm_cap > 16384) sm_cap = 16384;
  P.sm_cap = sm_cap;
  P.unit_cnt = cnt; P.unit_thr = thr; P.unit_cand = reinterpret_cast<const uint2*>(lists);
  P.q_f32 = q_f32; P.q_inv = q_inv; P.q_err = q_err; P.c_f32 = c_f32; P.c_inv = c_inv; P.c_stats = c_stats;
  P.idx_offset = 0; P.out_val = out_val; P.out_idx = out_idx; P.out_margin = out_margin;
  const size_t smem = (size_t)sm_cap * 8 + (size_t)kprime * 8 + round_up((size_t)D * 4, 16);
  QST_CHECK_ARG(smem <= 200 * 1024, "finalize_lists: D=%lld too large", (long long)D);
  QST_CUDA(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  finalize_kernel<<<(unsigned)Q, kFinThreads, smem, st>>>(P);
  QST_LAUNCH_CHECK();
  return QST_OK;
}

extern "C" size_t qst_finalize_lists_scratch_bytes(int64_t Q, int G) { return (size_t)G * Q * 8; }

static int select_requests_impl(int64_t Q, int G, int m, int kprime, int64_t n_total, const void* lists,
                                int32_t* out_req, uint32_t* out_bound, void* scratch, const qst_scatter* dst,
                                qst_stream_t stream) {
  QST_CHECK_ARG(lists && out_req && out_bound && scratch, "select_requests: null argument");
  QST_CHECK_ARG(!dst || (dst->world == G && dst->rows_per_block == Q),
                "select_requests: scatter descriptor must have one block of Q rows per shard");
  QST_CHECK_ARG(Q >= 1 && G >= 1 && G <= kMaxStripes && m >= 1 && kprime >= 1 && kprime <= 2048 && n_total >= G,
                "select_requests: bad shape Q=%lld G=%d m=%d kprime=%d n_total=%lld", (long long)Q, G, m, kprime,
                (long long)n_total);
  QST_CHECK_ARG(n_total < (1ll << 31), "select_requests: corpus ids must fit in 31 bits");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint32_t* thr = reinterpret_cast<uint32_t*>(scratch);
  int* cnt = reinterpret_cast<int*>(thr + (size_t)G * Q);
  unpack_list_trailers_kernel<<<(unsigned)ceil_div((int64_t)G * Q, 256), 256, 0, st>>>(
      reinterpret_cast<const uint2*>(lists), (int64_t)G * Q, m, thr, cnt);
  QST_LAUNCH_CHECK();
  FinParams P{};
  P.Q = (int)Q; P.N = 0; P.D = 0;
  P.k = kprime; P.kprime = kprime; P.cap = m + 1;
  P.m_tiles = 1; P.stripes = G; P.score = QST_SCORE_DOT; P.rows_per_unit = (int)Q;
  int sm_cap = kprime + m + 1;
  if (sm_cap < G * m) sm_cap = G * m;
  if (sm_cap < 4096) sm_cap = 4096;
  if (sm_cap > 16384) sm_cap = 16384;
  P.sm_cap = sm_cap;
  P.unit_cnt = cnt; P.unit_thr = thr; P.unit_cand = reinterpret_cast<const uint2*>(lists);
  P.req_out = out_req; P.bound_out = out_bound; P.req_G = G; P.req_m = m; P.n_total = n_total;
  if (int rc = scatter_from(dst, &P.scat)) return rc;
  const size_t smem = (size_t)sm_cap * 8 + (size_t)kprime * 8 + 16;
  QST_CUDA(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  finalize_kernel<<<(unsigned)Q, kFinThreads, smem, st>>>(P);
  QST_LAUNCH_CHECK();
  return QST_OK;
}

extern "C" int qst_select_requests(int64_t Q, int G, int m, int kprime, int64_t n_total, const void* lists,
                                   int32_t* out_req, uint32_t* out_bound, void* scratch, qst_stream_t stream) {
  return select_requests_impl(Q, G, m, kprime, n_total, lists, out_req, out_bound, scratch, nullptr, stream);
}

extern "C" int qst_select_requests_scatter(int64_t Q, int G, int m, int kprime, int64_t n_total, const void* lists,
                                           int32_t* out_req, uint32_t* out_bound, void* scratch,
                                           const qst_scatter* dst, qst_stream_t stream) {
  QST_CHECK_ARG(dst != nullptr, "select_requests_scatter: null descriptor");
  return select_requests_impl(Q, G, m, kprime, n_total, lists, out_req, out_bound, scratch, dst, stream);
}

static int rescore_requests_impl(int64_t rows, int m, int64_t D, int score, const int32_t* req, const float* q_f32,
                                 const float* q_inv, const float* c_f32, const float* c_inv, float* out,
                                 const qst_scatter* dst, qst_stream_t stream) {
  QST_CHECK_ARG(req && q_f32 && c_f32 && out, "rescore_requests: null argument");
  QST_CHECK_ARG(!dst || (int64_t)dst->world * dst->rows_per_block == rows,
                "rescore_requests: scatter blocks do not cover the %lld rows", (long long)rows);
  Scatter scat;
  if (int rc = scatter_from(dst, &scat)) return rc;
  QST_CHECK_ARG(rows >= 1 && m >= 1 && D >= 1 && D * 4 <= 200 * 1024, "rescore_requests: bad shape rows=%lld m=%d D=%lld",
                (long long)rows, m, (long long)D);
  QST_CHECK_ARG(score >= QST_SCORE_COS && score <= QST_SCORE_EUCLID, "rescore_requests: unknown score %d", score);
  QST_CHECK_ARG(score != QST_SCORE_COS || (q_inv && c_inv), "rescore_requests: cos score needs inverse norms");
  {
    const char* e = getenv("QST_RESCORE_CTA");   // 1 forces the CTA-per-row kernel (comparison / debugging)
    const bool aligned = ((reinterpret_cast<uintptr_t>(c_f32) | reinterpret_cast<uintptr_t>(q_f32)) & 15u) == 0;
    if (D % 4 == 0 && D <= 128 * kRescoreRegChunks && aligned && !(e && e[0] == '1')) {
      const unsigned grid = (unsigned)ceil_div(rows, kFinWarps);
      cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
      if (score == QST_SCORE_EUCLID)
        rescore_requests_warp_kernel<true><<<grid, kFinThreads, 0, st>>>(rows, m, (int)D, score, req, q_f32, q_inv, c_f32, c_inv, out, scat);
      else
        rescore_requests_warp_kernel<false><<<grid, kFinThreads, 0, st>>>(rows, m, (int)D, score, req, q_f32, q_inv, c_f32, c_inv, out, scat);
      QST_LAUNCH_CHECK();
      return QST_OK;
    }
  }
  const size_t smem = round_up((size_t)D * 4, 16);
  QST_CUDA(cudaFuncSetAttribute(rescore_requests_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rescore_requests_kernel<<<(unsigned)rows, kFinThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      m, (int)D, score, req, q_f32, q_inv, c_f32, c_inv, out, scat);
  QST_LAUNCH_CHECK();
  return QST_OK;
}

extern "C" int qst_rescore_requests(int64_t rows, int m, int64_t D, int score, const int32_t* req, const float* q_f32,
                                    const float* q_inv, const float* c_f32, const float* c_inv, float* out,
                                    qst_stream_t stream) {
  return rescore_requests_impl(rows, m, D, score, req, q_f32, q_inv, c_f32, c_inv, out, nullptr, stream);
}

extern "C" int qst_rescore_requests_scatter(int64_t rows, int m, int64_t D, int score, const int32_t* req,
                                            const float* q_f32, const float* q_inv, const float* c_f32,
                                            const float* c_inv, float* staging, const qst_scatter* dst,
                                            qst_stream_t stream) {
  QST_CHECK_ARG(dst != nullptr, "rescore_requests_scatter: null descriptor");
  return rescore_requests_impl(rows, m, D, score, req, q_f32, q_inv, c_f32, c_inv, staging, dst, stream);
}

extern "C" int qst_dense_scores(int64_t Q, int64_t N, int64_t D, int score, const float* q_f32, const float* q_inv,
                                const float* c_f32, const float* c_inv, float* out, qst_stream_t stream) {
  QST_CHECK_ARG(q_f32 && c_f32 && out, "dense_scores: null argument");
  QST_CHECK_ARG(Q >= 1 && N >= 1 && D >= 1 && Q <= 65535 && N < (1ll << 31) && D * 4 <= 200 * 1024,
                "dense_scores: bad shape Q=%lld N=%lld D=%lld (Q <= 65535 per call)", (long long)Q, (long long)N, (long long)D);
  QST_CHECK_ARG(score >= QST_SCORE_COS && score <= QST_SCORE_EUCLID, "dense_scores: unknown score %d", score);
  QST_CHECK_ARG(score != QST_SCORE_COS || (q_inv && c_inv), "dense_scores: cos score needs inverse norms");
  const size_t smem = round_up((size_t)D * 4, 16);
  QST_CUDA(cudaFuncSetAttribute(dense_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t gx = ceil_div(N, kFinWarps);
  if (gx > 1024) gx = 1024;
  dense_scores_kernel<<<dim3((unsigned)gx, (unsigned)Q), kFinThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      N, (int)D, score, q_f32, q_inv, c_f32, c_inv, out);
  QST_LAUNCH_CHECK();
  return QST_OK;
}

extern "C" int qst_finalize_exact(int64_t Q, int G, int m, int k, int score, int64_t D, int64_t n_total,
                                  const int32_t* req, const float* exact, const uint32_t* bound, const float* q_f32,
                                  const float* q_err, const float* c_stats, float* out_val, int64_t* out_idx,
                                  float* out_margin, qst_stream_t stream) {
  QST_CHECK_ARG(req && exact && bound && q_f32 && out_val && out_idx, "finalize_exact: null argument");
  QST_CHECK_ARG(Q >= 1 && G >= 1 && m >= 1 && k >= 1 && D >= 1, "finalize_exact: bad shape");
  QST_CHECK_ARG(score >= QST_SCORE_COS && score <= QST_SCORE_EUCLID, "finalize_exact: unknown score %d", score);
  finalize_exact_kernel<<<(unsigned)Q, kFinThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (int)Q, G, m, k, score, (int)D, n_total, req, exact, bound, q_f32, q_err, c_stats, out_val, out_idx, out_margin);
  QST_LAUNCH_CHECK();
  return QST_OK;
}

extern "C" int qst_merge_topk(const float* vals, const int64_t* idx, int G, int64_t Q, int k, float* out_val,
                              int64_t* out_idx, qst_stream_t stream) {
  QST_CHECK_ARG(vals && idx && out_val && out_idx, "merge_topk: null argument");
  QST_CHECK_ARG(G >= 1 && Q >= 0 && k >= 1, "merge_topk: bad shape G=%d Q=%lld k=%d", G, (long long)Q, k);
  if (Q == 0) return QST_OK;
  merge_topk_kernel<<<(unsigned)Q, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(vals, idx, G, Q, k, out_val, out_idx);
  QST_LAUNCH_CHECK();
  return QST_OK;
}

extern "C" size_t qst_exact_rescan_workspace_bytes(int64_t Q, int k) {
  (void)k;
  // header | flagged[Q] | counts[Q] | coll_val[Q*cap] | coll_idx[Q*cap]  (worst case: all flagged)
  const size_t nf = (size_t)(Q < kRescanMaxFlagged ? Q : kRescanMaxFlagged);
  return 256 + round_up((size_t)Q * 4, 256) * 2 + nf * kRescanCap * 8;
}

static int exact_rescan_impl(int64_t Q, int64_t N, int64_t D, int k, int score, const float* q_f32,
                             const float* q_inv, const float* c_f32, const float* c_inv, int64_t idx_offset,
                             float* out_val, int64_t* out_idx, float* margin_inout, void* scratch,
                             const float* thr_base, int thr_stride, int thr_off, int partial, int* overflow,
                             qst_stream_t stream) {
  QST_CHECK_ARG(q_f32 && c_f32 && out_val && out_idx && margin_inout && scratch, "exact_rescan: null argument");
  QST_CHECK_ARG(score >= QST_SCORE_COS && score <= QST_SCORE_EUCLID, "exact_rescan: unknown score %d", score);
  // query rows of one batch live in shared memory: 16 per corpus pass up to D = 3200, fewer beyond
  int batch = kRescanBatch;
  while (batch > 1 && (size_t)batch * D * 4 > 200 * 1024) batch >>= 1;
  QST_CHECK_ARG((size_t)batch * D * 4 <= 200 * 1024, "exact_rescan: D=%lld too large", (long long)D);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* p = reinterpret_cast<uint8_t*>(scratch);
  RescanScratch* hdr = reinterpret_cast<RescanScratch*>(p); p += 256;
  int* flagged = reinterpret_cast<int*>(p); p += round_up((size_t)Q * 4, 256);
  int* counts = reinterpret_cast<int*>(p); p += round_up((size_t)Q * 4, 256);
  const size_t nf_max = (size_t)(Q < kRescanMaxFlagged ? Q : kRescanMaxFlagged);
  float* coll_val = reinterpret_cast<float*>(p); p += nf_max * kRescanCap * 4;
  int* coll_idx = reinterpret_cast<int*>(p);
  const size_t smem = (size_t)batch * D * 4;
  int sms = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  const int grid = sms * 2;
  // One pass serves at most kRescanMaxFlagged queries (that bounds the scratch); the flagged list is
  // consumed on the device (no host read of its length), so enough passes are queued to serve EVERY
  // query even if all of them are flagged -- a pass that finds nothing listed returns at once
  // (three empty launches).  After the last pass the only queries left with margin <= 0 are those
  // with more than kRescanCap rows tied at their k-th score (margin == 0).
  const int passes = (int)ceil_div(Q, kRescanMaxFlagged);
  for (int pass = 0; pass < passes; ++pass) {
    rescan_list_kernel<<<1, 1024, 0, st>>>(margin_inout, (int)Q, flagged, hdr, counts);
    QST_LAUNCH_CHECK();
#define QST_RESCAN_COLLECT(B)                                                                                          \
  do {                                                                                                                 \
    QST_CUDA(cudaFuncSetAttribute(rescan_collect_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
    rescan_collect_kernel<B><<<grid, 256, smem, st>>>((int)N, (int)D, k, score, q_f32, q_inv, c_f32, c_inv, thr_base,  \
                                                      thr_stride, thr_off, flagged, hdr, counts, coll_val, coll_idx);  \
  } while (0)
    switch (batch) {
      case 16: QST_RESCAN_COLLECT(16); break;
      case 8: QST_RESCAN_COLLECT(8); break;
      case 4: QST_RESCAN_COLLECT(4); break;
      case 2: QST_RESCAN_COLLECT(2); break;
      default: QST_RESCAN_COLLECT(1); break;
    }
#undef QST_RESCAN_COLLECT
    QST_LAUNCH_CHECK();
    rescan_emit_kernel<<<(unsigned)(Q < 1024 ? Q : 1024), kFinThreads, 0, st>>>(k, idx_offset, flagged, hdr, counts,
                                                                                coll_val, coll_idx, out_val, out_idx,
                                                                                margin_inout, partial, overflow);
    QST_LAUNCH_CHECK();
    if (partial) break;   // the margins are not updated in this mode: one pass (at most 8192 flagged queries)
  }
  return QST_OK;
}

extern "C" int qst_exact_rescan(int64_t Q, int64_t N, int64_t D, int k, int score, const float* q_f32,
                                const float* q_inv, const float* c_f32, const float* c_inv, int64_t idx_offset,
                                float* out_val, int64_t* out_idx, float* margin_inout, void* scratch,
                                qst_stream_t stream) {
  return exact_rescan_impl(Q, N, D, k, score, q_f32, q_inv, c_f32, c_inv, idx_offset, out_val, out_idx, margin_inout,
                           scratch, out_val, k, k - 1, 0, nullptr, stream);
}

extern "C" int qst_exact_rescan_lists(int64_t Q, int64_t N, int64_t D, int k, int score, const float* q_f32,
                                      const float* q_inv, const float* c_f32, const float* c_inv, int64_t idx_offset,
                                      const float* kth_val, const float* margin, float* out_val, int64_t* out_idx,
                                      int* overflow, void* scratch, qst_stream_t stream) {
  QST_CHECK_ARG(kth_val && margin, "exact_rescan_lists: null argument");
  return exact_rescan_impl(Q, N, D, k, score, q_f32, q_inv, c_f32, c_inv, idx_offset, out_val, out_idx,
                           const_cast<float*>(margin), scratch, kth_val, 1, 0, 1, overflow, stream);
}
