/* A plain C99 caller of the ABI (tests/test_abi.py compiles, links and runs it on the CPU box):
 * include/qst.h must be a C header, libqst.so must link from C without the C++ or CUDA runtimes on the
 * command line, and the host-only entries must work without a device. */
#include <stdio.h>
#include <string.h>

#include "qst.h"

int main(void) {
  qst_topk_plan plan;
  qst_quad_params prm;
  memset(&prm, 0, sizeof prm);
  if (qst_version() < 100) return 1;
  /* config 3 of BASELINE.json on a 148-SM device: CTA pairs, query-stationary tiles, k' = 224 */
  if (qst_topk_plan_make(10000, 1000000, 768, 100, 0, QST_SCORE_COS, 148, &plan) != QST_OK) return 2;
  if (plan.ctas != 2 || plan.qs != 1 || plan.kprime != 224 || plan.D_pad != 768) return 3;
  if (plan.units != plan.m_tiles * plan.stripes || plan.ws_bytes <= plan.off_cand) return 4;
  /* errors come back as a code and a message, never as a crash */
  if (qst_topk_plan_make(0, 10, 8, 5, 0, QST_SCORE_COS, 148, &plan) == QST_OK) return 5;
  if (strstr(qst_last_error(), "bad shape") == NULL) return 6;
  if (qst_quadruplet_fwd(NULL, NULL, NULL, NULL, QST_F32, 4, 8, &prm, QST_RED_MEAN, NULL, NULL, NULL, NULL) == QST_OK) return 7;
  if (qst_padded_dim(385) != 448 || qst_quadruplet_workspace_bytes() == 0) return 8;
  printf("c_abi_probe ok: version %d, plan units %d, grid %d, workspace %zu bytes\n", qst_version(), plan.units,
         plan.grid, plan.ws_bytes);
  return 0;
}
