// K1: row preparation -- L2 norm (eps 1e-12), optional normalisation, bf16 cast, zero padding to
// a multiple of 64 columns, and the rounding residual the exactness certificate needs.
//
// Replaces the two F.normalize(p=2, dim=1) calls of sentence_transformers.util.cos_sim
// (used at /root/reference/ir_evauation_script.py:70 and models/evaluators.py:545) and feeds
// the tensor-core pass.  HBM-bound: reads n*d*sizeof(T), writes n*d_pad*2 bytes.
#include "qst_common.cuh"

namespace qst {

constexpr int kPrepThreads = 256;  // 8 rows per CTA

template <typename T, int VEC>
__global__ void __launch_bounds__(kPrepThreads)
prep_rows_kernel(const T* __restrict__ x, int64_t n, int64_t d, int64_t d_pad, int mode,
                 __nv_bfloat16* __restrict__ out, float* __restrict__ out_inv, float* __restrict__ out_sq,
                 float* __restrict__ out_err, float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  constexpr int kWarps = kPrepThreads / 32;
  float max_err = 0.f, max_norm = 0.f;
  for (int64_t row = (int64_t)blockIdx.x * kWarps + warp; row < n; row += (int64_t)gridDim.x * kWarps) {
    const T* src = x + row * d;
    float ss = 0.f;
    for (int64_t i = (int64_t)lane * VEC; i < d; i += 32 * VEC) {
      if (VEC == 1) {
        const float v = to_f32<T>(src[i]);
        ss = fmaf(v, v, ss);
      } else {
        Vec16<T> v = ld_vec16<T>(src + i);
#pragma unroll
        for (int j = 0; j < VEC; ++j) { const float f = to_f32<T>(v.v[j]); ss = fmaf(f, f, ss); }
      }
    }
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss);
    const float inv = 1.0f / fmaxf(nrm, 1e-12f);
    // mode 0 raw (dot) | 1 L2-normalised (cos) | 2 euclid corpus: raw + ||x||^2 split over three
    // bf16 columns | 3 euclid query: 2x and -1 in those columns, so that the tensor-core dot product
    // is 2 q.c - ||c||^2 = ||q||^2 - ||q-c||^2 (same ranking as 1/(1+||q-c||))
    const float scale = mode == QST_PREP_COS ? inv : (mode == QST_PREP_EUCLID_QUERY ? 2.0f : 1.0f);
    const __nv_bfloat16 sq_hi = __float2bfloat16_rn(ss);
    const __nv_bfloat16 sq_mid = __float2bfloat16_rn(ss - __bfloat162float(sq_hi));
    const __nv_bfloat16 sq_lo = __float2bfloat16_rn(ss - __bfloat162float(sq_hi) - __bfloat162float(sq_mid));
    __nv_bfloat16* dst = out ? out + row * d_pad : nullptr;
    float es = 0.f;
    // d_pad is a multiple of 64 -> every lane writes whole 8-element (16 B) groups
    for (int64_t i = (int64_t)lane * 8; i < d_pad; i += 32 * 8) {
      Vec16<__nv_bfloat16> o;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int64_t c = i + j;
        const float v = c < d ? to_f32<T>(src[c]) * scale : 0.f;
        __nv_bfloat16 b = __float2bfloat16_rn(v);
        const float r = __bfloat162float(b) - v;
        es = fmaf(r, r, es);
        if (c >= d && c < d + 3) {
          if (mode == QST_PREP_EUCLID_CORPUS) b = c == d ? sq_hi : (c == d + 1 ? sq_mid : sq_lo);
          else if (mode == QST_PREP_EUCLID_QUERY) b = __float2bfloat16_rn(-1.0f);
        }
        o.v[j] = b;
      }
      if (dst) st_vec16<__nv_bfloat16>(dst + i, o);
    }
    es = warp_sum(es);
    const float err = sqrtf(es);
    const float used_norm = mode == QST_PREP_COS ? nrm * inv : (mode == QST_PREP_EUCLID_QUERY ? 2.0f * nrm : nrm);
    if (lane == 0) {
      if (out_inv) out_inv[row] = inv;
      if (out_sq) out_sq[row] = ss;
      if (out_err) out_err[row] = err;
    }
    max_err = fmaxf(max_err, err);
    max_norm = fmaxf(max_norm, used_norm);
  }
  if (stats && lane == 0) {
    // non-negative floats order like their bit patterns
    atomicMax(reinterpret_cast<int*>(&stats[0]), __float_as_int(max_err));
    atomicMax(reinterpret_cast<int*>(&stats[1]), __float_as_int(max_norm));
  }
}

}  // namespace qst

using namespace qst;

extern "C" int64_t qst_padded_dim(int64_t d) { return round_up(d, 64); }
extern "C" int64_t qst_padded_dim_for(int64_t d, int mode) {
  return round_up(d + (mode == QST_PREP_EUCLID_CORPUS || mode == QST_PREP_EUCLID_QUERY ? 3 : 0), 64);
}

extern "C" int qst_prep_rows(const void* x, int dtype, int64_t n, int64_t d, int mode, void* out_bf16,
                             float* out_inv_norm, float* out_sq_norm, float* out_err, float* stats,
                             qst_stream_t stream) {
  QST_CHECK_ARG(n >= 0 && d >= 1, "prep_rows: bad shape n=%lld d=%lld", (long long)n, (long long)d);
  QST_CHECK_ARG(dtype == QST_F32 || dtype == QST_F16 || dtype == QST_BF16, "prep_rows: bad dtype %d", dtype);
  QST_CHECK_ARG(x != nullptr || n == 0, "prep_rows: null input");
  QST_CHECK_ARG((reinterpret_cast<uintptr_t>(out_bf16) & 15u) == 0, "prep_rows: out_bf16 must be 16-byte aligned");
  if (n == 0) return QST_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  QST_CHECK_ARG(mode >= QST_PREP_RAW && mode <= QST_PREP_EUCLID_QUERY, "prep_rows: bad mode %d", mode);
  const int64_t d_pad = qst_padded_dim_for(d, mode);
  const int warps = kPrepThreads / 32;
  const int64_t want = ceil_div(n, warps);
  const int grid = (int)(want < 148 * 8 ? want : 148 * 8);
  const size_t esz = dtype == QST_F32 ? 4 : 2;
  const bool vec_ok = (d % (16 / esz)) == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0;
  __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(out_bf16);
#define QST_PREP(T)                                                                                        \
  do {                                                                                                     \
    if (vec_ok)                                                                                            \
      prep_rows_kernel<T, 16 / sizeof(T)><<<grid, kPrepThreads, 0, st>>>(                                  \
          reinterpret_cast<const T*>(x), n, d, d_pad, mode, ob, out_inv_norm, out_sq_norm, out_err, stats); \
    else                                                                                                   \
      prep_rows_kernel<T, 1><<<grid, kPrepThreads, 0, st>>>(                                               \
          reinterpret_cast<const T*>(x), n, d, d_pad, mode, ob, out_inv_norm, out_sq_norm, out_err, stats); \
  } while (0)
  if (dtype == QST_F32) QST_PREP(float);
  else if (dtype == QST_F16) QST_PREP(__half);
  else QST_PREP(__nv_bfloat16);
#undef QST_PREP
  QST_LAUNCH_CHECK();
  return QST_OK;
}
