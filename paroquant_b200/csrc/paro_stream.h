// Host-side description of one step of a small-M chain (paro_capi.cu fills it from paro_chain_step, paro_stream.cu plans and launches).
#pragma once
#include "paro_layout.h"

namespace paro {

struct HostStep {
  const paro_linear_shape *shape;
  Layout L;
  const void *packed, *bias, *x;
  void *y;
  int x_op, epi_op;
  const void *res_in;
  void *res_out;
  const void *norm_w;
  float eps;
  const paro_tp_info *tp;
};

}  // namespace paro
