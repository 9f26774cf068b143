// K4: per-query IR metrics in fp64 with the reference's operation order.
//
// Replaces the Python loops of sentence-transformers 2.2.2
// InformationRetrievalEvaluator.compute_metrics / compute_dcg_at_k (float64; 5.9 s of 6.7 s at
// /root/reference/ir_evauation_script.py:163-173's default k-lists).  One warp per query:
// lanes test ranks for membership in the query's relevant set (binary search in the CSR row),
// ballots turn that into a hit bitmap, then one lane per cut-off k walks the bitmap sequentially
// so every sum is accumulated in the same order as the Python code.  Only IEEE add/div are used
// (no FMA contraction possible), so per-query values are bit-identical to the reference's.
#include "qst_common.cuh"

namespace qst {

constexpr int kMetThreads = 128;  // 4 queries per CTA
constexpr int kMaxK = 1024;

__global__ void __launch_bounds__(kMetThreads)
ir_metrics_kernel(const int64_t* __restrict__ ranked, int64_t Q, int K, const int64_t* __restrict__ rowptr,
                  const int64_t* __restrict__ cols, const int32_t* __restrict__ ks, int n_ks,
                  const double* __restrict__ log2_tab, const double* __restrict__ idcg_tab, double* __restrict__ out) {
  __shared__ uint32_t s_hits[kMetThreads / 32][kMaxK / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t q = (int64_t)blockIdx.x * (kMetThreads / 32) + warp;
  if (q >= Q) return;
  const int64_t r0 = rowptr[q], r1 = rowptr[q + 1];
  const int64_t n_rel = r1 - r0;
  uint32_t* hits = s_hits[warp];
  const int words = (K + 31) / 32;
  for (int w = 0; w < words; ++w) {
    const int rank = w * 32 + lane;
    bool hit = false;
    if (rank < K) {
      const int64_t id = ranked[q * K + rank];
      if (id >= 0) {
        int64_t lo = r0, hi = r1;
        while (lo < hi) {
          const int64_t mid = (lo + hi) >> 1;
          const int64_t c = cols[mid];
          if (c < id) lo = mid + 1; else hi = mid;
        }
        hit = lo < r1 && cols[lo] == id;
      }
    }
    const uint32_t b = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) hits[w] = b;
  }
  __syncwarp();
  for (int ki = lane; ki < n_ks; ki += 32) {
    const int k = ks[ki];
    int num_correct = 0, first = -1;
    double dcg = 0.0, sum_prec = 0.0;
    for (int rank = 0; rank < k && rank < K; ++rank) {
      if ((hits[rank >> 5] >> (rank & 31)) & 1u) {
        ++num_correct;
        if (first < 0) first = rank;
        dcg = __dadd_rn(dcg, __ddiv_rn(1.0, log2_tab[rank]));
        sum_prec = __dadd_rn(sum_prec, __ddiv_rn((double)num_correct, (double)(rank + 1)));
      }
    }
    const int64_t ideal_n = n_rel < (int64_t)k ? n_rel : (int64_t)k;
    const size_t o = (size_t)ki * Q + q;
    const size_t plane = (size_t)n_ks * Q;
    out[0 * plane + o] = num_correct > 0 ? 1.0 : 0.0;
    out[1 * plane + o] = __ddiv_rn((double)num_correct, (double)k);
    out[2 * plane + o] = __ddiv_rn((double)num_correct, (double)n_rel);
    out[3 * plane + o] = first >= 0 ? __ddiv_rn(1.0, (double)(first + 1)) : 0.0;
    out[4 * plane + o] = __ddiv_rn(dcg, idcg_tab[ideal_n]);
    out[5 * plane + o] = __ddiv_rn(sum_prec, (double)ideal_n);
  }
}

}  // namespace qst

using namespace qst;

extern "C" int qst_ir_metrics(const int64_t* ranked_idx, int64_t Q, int K, const int64_t* rel_rowptr,
                              const int64_t* rel_cols, const int32_t* ks, int n_ks, const double* log2_tab,
                              const double* idcg_tab, double* out, qst_stream_t stream) {
  QST_CHECK_ARG(ranked_idx && rel_rowptr && rel_cols && ks && log2_tab && idcg_tab && out, "ir_metrics: null argument");
  QST_CHECK_ARG(K >= 1 && K <= kMaxK, "ir_metrics: K must be in [1, %d], %d given", kMaxK, K);
  QST_CHECK_ARG(n_ks >= 1, "ir_metrics: no cut-offs");
  if (Q == 0) return QST_OK;
  const unsigned grid = (unsigned)ceil_div(Q, kMetThreads / 32);
  ir_metrics_kernel<<<grid, kMetThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      ranked_idx, Q, K, rel_rowptr, rel_cols, ks, n_ks, log2_tab, idcg_tab, out);
  QST_LAUNCH_CHECK();
  return QST_OK;
}
