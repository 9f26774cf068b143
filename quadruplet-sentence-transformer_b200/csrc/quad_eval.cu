// Paired distances of the QuadrupletEvaluator.
//
// Replaces the arithmetic of /root/reference/models/evaluators.py:130-389: three
// sentence-transformers 2.2.2 TripletEvaluators (pos/part, pos/neg, part/neg), each of which computes
// sklearn paired cosine / manhattan / euclidean distances anchor<->x on the CPU and counts
// d(anchor, first) < d(anchor, second).  Here one warp owns one quadruplet row: the four embeddings
// are read once (128-bit loads), the nine distances are accumulated in registers, and the nine
// comparison counts are reduced per CTA and added to the output with integer atomics
// (deterministic).  HBM-bound: 4*B*D*sizeof(T) bytes.
#include "qst_common.cuh"

namespace qst {

constexpr int kEvalThreads = 128;

struct EvalArgs {
  const void *a, *po, *pa, *ne;
  int64_t B, D;
  float* dist;              // [B, 9] or null: (cos, manhattan, euclid) x (pos, part, neg)
  unsigned long long* cnt;  // [9]: (cos, manhattan, euclid) x (pos<part, pos<neg, part<neg)
};

template <typename T, int VEC>
__global__ void __launch_bounds__(kEvalThreads) quad_eval_kernel(const EvalArgs g) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = kEvalThreads / 32;
  __shared__ unsigned int s_cnt[9];
  if (threadIdx.x < 9) s_cnt[threadIdx.x] = 0u;
  __syncthreads();
  const int64_t D = g.D;
  for (int64_t row = (int64_t)blockIdx.x * kWarps + warp; row < g.B; row += (int64_t)gridDim.x * kWarps) {
    const T* a = reinterpret_cast<const T*>(g.a) + row * D;
    const T* x[3] = {reinterpret_cast<const T*>(g.po) + row * D, reinterpret_cast<const T*>(g.pa) + row * D,
                     reinterpret_cast<const T*>(g.ne) + row * D};
    float aa = 0.f, dot[3] = {0.f, 0.f, 0.f}, xx[3] = {0.f, 0.f, 0.f}, l1[3] = {0.f, 0.f, 0.f},
          l2[3] = {0.f, 0.f, 0.f};
    for (int64_t i = (int64_t)lane * VEC; i < D; i += 32 * VEC) {
      float va[VEC], vx[3][VEC];
      if (VEC == 1) {
        va[0] = to_f32<T>(a[i]);
#pragma unroll
        for (int k = 0; k < 3; ++k) vx[k][0] = to_f32<T>(x[k][i]);
      } else {
        Vec16<T> ra = ld_vec16<T>(a + i);
#pragma unroll
        for (int j = 0; j < VEC; ++j) va[j] = to_f32<T>(ra.v[j]);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          Vec16<T> rx = ld_vec16<T>(x[k] + i);
#pragma unroll
          for (int j = 0; j < VEC; ++j) vx[k][j] = to_f32<T>(rx.v[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        aa = fmaf(va[j], va[j], aa);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float d = va[j] - vx[k][j];
          dot[k] = fmaf(va[j], vx[k][j], dot[k]);
          xx[k] = fmaf(vx[k][j], vx[k][j], xx[k]);
          l1[k] += fabsf(d);
          l2[k] = fmaf(d, d, l2[k]);
        }
      }
    }
    aa = warp_sum(aa);
    float dc[3], dm[3], de[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float dt = warp_sum(dot[k]), nx = warp_sum(xx[k]);
      dm[k] = warp_sum(l1[k]);
      de[k] = sqrtf(warp_sum(l2[k]));
      // sklearn: 0.5 * || a/||a|| - x/||x|| ||^2 = 1 - cos (rows of zero norm are left as they are)
      const float na = sqrtf(aa), nn = sqrtf(nx);
      const float cs = dt / ((na > 0.f ? na : 1.f) * (nn > 0.f ? nn : 1.f));
      dc[k] = 0.5f * ((na > 0.f ? 1.f : 0.f) + (nn > 0.f ? 1.f : 0.f)) - cs;
    }
    if (lane == 0) {
      if (g.dist) {
        float* o = g.dist + row * 9;
#pragma unroll
        for (int k = 0; k < 3; ++k) { o[k] = dc[k]; o[3 + k] = dm[k]; o[6 + k] = de[k]; }
      }
      const float* m[3] = {dc, dm, de};
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        if (m[f][0] < m[f][1]) atomicAdd(&s_cnt[f * 3 + 0], 1u);   // pos  < part
        if (m[f][0] < m[f][2]) atomicAdd(&s_cnt[f * 3 + 1], 1u);   // pos  < neg
        if (m[f][1] < m[f][2]) atomicAdd(&s_cnt[f * 3 + 2], 1u);   // part < neg
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 9 && s_cnt[threadIdx.x]) atomicAdd(&g.cnt[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

}  // namespace qst

using namespace qst;

extern "C" int qst_quadruplet_eval(const void* x_anchor, const void* x_pos, const void* x_part, const void* x_neg,
                                   int dtype, int64_t B, int64_t D, float* out_dist, unsigned long long* out_counts,
                                   qst_stream_t stream) {
  QST_CHECK_ARG(out_counts != nullptr, "quadruplet_eval: null out_counts");
  QST_CHECK_ARG(B >= 0 && D >= 1, "quadruplet_eval: bad shape B=%lld D=%lld", (long long)B, (long long)D);
  QST_CHECK_ARG(dtype == QST_F32 || dtype == QST_F16 || dtype == QST_BF16, "quadruplet_eval: bad dtype %d", dtype);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  QST_CUDA(cudaMemsetAsync(out_counts, 0, 9 * sizeof(unsigned long long), st));
  if (B == 0) return QST_OK;
  QST_CHECK_ARG(x_anchor && x_pos && x_part && x_neg, "quadruplet_eval: null input pointer");
  EvalArgs a{x_anchor, x_pos, x_part, x_neg, B, D, out_dist, out_counts};
  const size_t esz = dtype == QST_F32 ? 4 : 2;
  auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  const bool vec_ok = (D % (16 / esz)) == 0 && aligned(x_anchor) && aligned(x_pos) && aligned(x_part) && aligned(x_neg);
  const int64_t want = ceil_div(B, kEvalThreads / 32);
  const int grid = (int)(want < 148 * 16 ? want : 148 * 16);
#define QST_EVAL(T)                                                                          \
  do {                                                                                       \
    if (vec_ok) quad_eval_kernel<T, 16 / sizeof(T)><<<grid, kEvalThreads, 0, st>>>(a);       \
    else quad_eval_kernel<T, 1><<<grid, kEvalThreads, 0, st>>>(a);                           \
  } while (0)
  if (dtype == QST_F32) QST_EVAL(float);
  else if (dtype == QST_F16) QST_EVAL(__half);
  else QST_EVAL(__nv_bfloat16);
#undef QST_EVAL
  QST_LAUNCH_CHECK();
  return QST_OK;
}
