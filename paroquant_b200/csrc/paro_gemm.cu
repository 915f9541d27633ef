// Large-M path (M > 16).  INTERIM: rows are pushed through the fused small-M kernel 16 at a
// time (correct, weight-stream bound); the tcgen05 GEMM replaces this body.
#include "paro_common.cuh"
#include "paro_layout.h"

namespace paro {

size_t decode_workspace_bytes(const Layout &L, int64_t max_m);
int decode_forward(const paro_linear_shape &s, const Layout &L, const void *packed, const void *x, int64_t M,
                   const void *bias, void *y, void *workspace, size_t workspace_bytes, cudaStream_t stream);

size_t gemm_workspace_bytes(const Layout &L, int64_t max_m) { return decode_workspace_bytes(L, 16); }

int gemm_forward(const paro_linear_shape &s, const Layout &L, const void *packed, const void *x, int64_t M,
                 const void *bias, void *y, void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  const size_t esz = 2;
  for (int64_t m0 = 0; m0 < M; m0 += 16) {
    const int64_t mc = M - m0 < 16 ? M - m0 : 16;
    const int rc = decode_forward(s, L, packed, static_cast<const uint8_t *>(x) + m0 * L.K * esz, mc, bias,
                                  static_cast<uint8_t *>(y) + m0 * L.N * esz, workspace, workspace_bytes, stream);
    if (rc) return rc;
  }
  return PARO_OK;
}

}  // namespace paro
