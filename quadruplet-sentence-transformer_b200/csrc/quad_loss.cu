// K5 host side: argument checks, dispatch to the per-dtype translation units, C ABI.
//
// Replaces /root/reference/models/losses/losses.py:9-69 (three F.triplet_margin_loss calls and the
// reductions) and the autograd graph behind them; kernels in quad_loss_kernels.cuh.
#include <cstdlib>
#include <cstring>

#include "quad_loss_kernels.cuh"

namespace qst {

__global__ void fill_scalar_kernel(float* out, float v) { *out = v; }

static int quad_dispatch(int kind, QuadArgs& a, int dtype, cudaStream_t st) {
  QST_CHECK_ARG(a.B >= 0 && a.D >= 1, "quadruplet: bad shape B=%lld D=%lld", (long long)a.B, (long long)a.D);
  QST_CHECK_ARG(dtype == QST_F32 || dtype == QST_F16 || dtype == QST_BF16, "quadruplet: bad dtype %d", dtype);
  QST_CHECK_ARG(a.reduction >= QST_RED_NONE && a.reduction <= QST_RED_MEAN, "quadruplet: bad reduction %d", a.reduction);
  QST_CHECK_ARG(a.prm.p > 0.f, "p must be positive, %g given", (double)a.prm.p);
  if (kind != K_BWD) QST_CHECK_ARG(a.loss_out != nullptr || (a.B == 0 && a.reduction == QST_RED_NONE), "quadruplet: null loss_out");
  if (a.B == 0) {
    if (kind != K_BWD && a.reduction != QST_RED_NONE) {
      // sum over nothing = 0, mean over nothing = nan (torch semantics)
      // written by a kernel (not a host-to-device copy of a stack variable): capturable in a CUDA graph
      fill_scalar_kernel<<<1, 1, 0, st>>>(a.loss_out, a.reduction == QST_RED_MEAN ? NAN : 0.f);
      QST_LAUNCH_CHECK();
    }
    return QST_OK;
  }
  QST_CHECK_ARG(a.a && a.po && a.pa && a.ne, "quadruplet: null input pointer");
  if (kind != K_BWD && a.reduction != QST_RED_NONE) QST_CHECK_ARG(a.ws != nullptr, "quadruplet: null workspace");
  if (kind == K_BWD) QST_CHECK_ARG(a.saved && a.grad_out, "quadruplet bwd: null saved/grad_out");
  const int pm = a.prm.p == 2.0f ? PM_2 : (a.prm.p == 1.0f ? PM_1 : (isinf(a.prm.p) ? PM_INF : PM_GEN));
  const size_t esz = dtype == QST_F32 ? 4 : 2;
  const int vec = (int)(16 / esz);
  auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  bool vec_ok = (a.D % vec) == 0 && aligned(a.a) && aligned(a.po) && aligned(a.pa) && aligned(a.ne) &&
                aligned(a.ga) && aligned(a.gp) && aligned(a.gq) && aligned(a.gn);
  const int warps = kQuadThreads / 32;
  int grid = (int)(ceil_div(a.B, warps) < kQuadMaxBlocks ? ceil_div(a.B, warps) : kQuadMaxBlocks);
  const bool reg_path = kind == K_FUSED && vec_ok && a.D <= (int64_t)32 * vec * kRegChunksMax;
  if (reg_path) {
    // exactly one resident wave of CTAs (3 per SM at <= 168 registers), each looping over rows:
    // fewer partials and tickets in the cross-CTA reduction, no ragged last wave
    // SM count of the CURRENT device, cached per device ordinal (a process may drive several GPUs)
    static int sms_of[64] = {0};
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64) sms = sms_of[dev];
    if (sms == 0) {
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (sms <= 0) sms = 148;
      if (dev >= 0 && dev < 64) sms_of[dev] = sms;
    }
    if (grid > sms * 3) grid = sms * 3;
    if (grid > kQuadLimbMaxBlocks) grid = kQuadLimbMaxBlocks;
    // A/B switch for profiles/loss_probe.py, read per call (tens of nanoseconds) so that the probe can switch
    // inside one process
    const char* e = getenv("QST_LOSS_REDUCE");
    const int mode = (e && !strcmp(e, "ticket")) ? 1 : 0;
    a.reduce_mode = mode;
  }
  if (dtype == QST_F32) quad_launch_f32(kind, a, pm, vec_ok, reg_path, grid, st);
  else if (dtype == QST_F16) quad_launch_f16(kind, a, pm, vec_ok, reg_path, grid, st);
  else quad_launch_bf16(kind, a, pm, vec_ok, reg_path, grid, st);
  QST_LAUNCH_CHECK();
  return QST_OK;
}

}  // namespace qst

using namespace qst;

extern "C" size_t qst_quadruplet_workspace_bytes(void) { return sizeof(QuadWorkspace); }

extern "C" int qst_quadruplet_fwd(const void* x_anchor, const void* x_pos, const void* x_part, const void* x_neg,
                                  int dtype, int64_t B, int64_t D, const qst_quad_params* prm, int reduction,
                                  float* loss_out, float* saved, void* workspace, qst_stream_t stream) {
  QST_CHECK_ARG(prm != nullptr, "quadruplet: null params");
  QuadArgs a{};
  a.a = x_anchor; a.po = x_pos; a.pa = x_part; a.ne = x_neg;
  a.B = B; a.D = D; a.prm = *prm; a.reduction = reduction;
  a.loss_out = loss_out; a.saved = saved; a.ws = reinterpret_cast<QuadWorkspace*>(workspace);
  return quad_dispatch(K_FWD, a, dtype, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int qst_quadruplet_bwd(const void* x_anchor, const void* x_pos, const void* x_part, const void* x_neg,
                                  int dtype, int64_t B, int64_t D, const qst_quad_params* prm, int reduction,
                                  const float* saved, const float* grad_out,
                                  void* g_anchor, void* g_pos, void* g_part, void* g_neg, qst_stream_t stream) {
  QST_CHECK_ARG(prm != nullptr, "quadruplet: null params");
  QuadArgs a{};
  a.a = x_anchor; a.po = x_pos; a.pa = x_part; a.ne = x_neg;
  a.ga = g_anchor; a.gp = g_pos; a.gq = g_part; a.gn = g_neg;
  a.B = B; a.D = D; a.prm = *prm; a.reduction = reduction;
  a.saved = const_cast<float*>(saved); a.grad_out = grad_out;
  return quad_dispatch(K_BWD, a, dtype, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int qst_quadruplet_fwd_bwd(const void* x_anchor, const void* x_pos, const void* x_part, const void* x_neg,
                                      int dtype, int64_t B, int64_t D, const qst_quad_params* prm, int reduction,
                                      float upstream, float* loss_out,
                                      void* g_anchor, void* g_pos, void* g_part, void* g_neg,
                                      void* workspace, qst_stream_t stream) {
  QST_CHECK_ARG(prm != nullptr, "quadruplet: null params");
  QuadArgs a{};
  a.a = x_anchor; a.po = x_pos; a.pa = x_part; a.ne = x_neg;
  a.ga = g_anchor; a.gp = g_pos; a.gq = g_part; a.gn = g_neg;
  a.B = B; a.D = D; a.prm = *prm; a.reduction = reduction; a.upstream = upstream;
  a.loss_out = loss_out; a.ws = reinterpret_cast<QuadWorkspace*>(workspace);
  return quad_dispatch(K_FUSED, a, dtype, reinterpret_cast<cudaStream_t>(stream));
}
