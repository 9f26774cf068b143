// K5 instantiations for float rows (see quad_loss_kernels.cuh).
#include "quad_loss_kernels.cuh"

namespace qst {

void quad_launch_f32(int kind, const QuadArgs& a, int pm, bool vec_ok, bool reg_path, int grid, cudaStream_t st) {
  if (reg_path) launch_fused_reg<float>(a, pm, grid, st);
  else if (kind == K_FWD) launch_vec<float, K_FWD>(a, pm, vec_ok, grid, st);
  else if (kind == K_BWD) launch_vec<float, K_BWD>(a, pm, vec_ok, grid, st);
  else launch_vec<float, K_FUSED>(a, pm, vec_ok, grid, st);
}

}  // namespace qst
