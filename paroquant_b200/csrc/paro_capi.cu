// extern "C" surface of libparo_b200.so (declared in include/paro_b200.h): argument checks with
// the reference's error behaviour (TORCH_CHECK -> message, here: code + paro_last_error()),
// dispatch to the sm_100a kernels.  No allocation, no host synchronisation.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "paro_common.cuh"
#include "paro_layout.h"
#include "paro_stream.h"

namespace paro {

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void note_launches(int n) { g_launches += n; }

// implemented in the kernel translation units
int rotate_launch(const void *x, void *out, const int16_t *idx, const void *theta, int theta_dtype, const void *scales,
                  int scales_dtype, int64_t M, int K, int krot, int G, int dtype, cudaStream_t stream);
int rotate_backward_launch(const void *y, const void *gout, const void *x, const int16_t *idx, const void *theta, int theta_dtype, const void *scales,
                           int scales_dtype, void *grad_x, float *grad_theta, float *grad_scale, int64_t M, int K, int krot, int G, int dtype,
                           int ref_formula, cudaStream_t stream);
int prepack_launch(const paro_linear_shape &s, const Layout &L, const int32_t *qweight, const int32_t *qzeros,
                   const void *scales, int scales_dtype, const int16_t *pairs, const void *theta, int theta_dtype,
                   const void *cscales, int cs_dtype, void *packed, cudaStream_t stream);
int unpack_dense_launch(const paro_linear_shape &s, const Layout &L, const void *packed, void *W, cudaStream_t stream);
bool decode_supported(const Layout &L, int64_t M);
int decode_forward(const paro_linear_shape &s, const Layout &L, const void *packed, const void *x, int64_t M,
                   const void *bias, void *y, void *scratch, size_t scratch_bytes, cudaStream_t stream);
size_t decode_scratch_bytes(const Layout &L, int64_t max_m);
int decode_trace_read(unsigned long long *host, int max_ctas);
int decode_debug_plan(const Layout &L, int64_t M, int sets, int sms, const int32_t *resident4, int32_t *out20);
bool stream_supported(const Layout &L, int64_t M);
size_t stream_sync_bytes(const Layout &L);
size_t stream_workspace_bytes(const Layout &L, int64_t max_m);
size_t stream_chain_workspace_bytes(const HostStep *steps, int n, int64_t M);
int stream_forward(const HostStep *steps, int n, int64_t M, void *workspace, size_t workspace_bytes, cudaStream_t stream);
int stream_linear_forward(const paro_linear_shape &s, const Layout &L, const void *packed, const void *x, int64_t M, const void *bias, void *y,
                          void *workspace, size_t workspace_bytes, cudaStream_t stream);
int stream_trace_read(unsigned long long *host, int max_ctas);
int stream_debug_plan(const Layout &L, int64_t M, int sets, int ctas, int32_t *out);
size_t gemm_workspace_bytes(const Layout &L, int64_t max_m);
int gemm_forward(const paro_linear_shape &s, const Layout &L, const void *packed, const void *x, int64_t M,
                 const void *bias, void *y, void *workspace, size_t workspace_bytes, cudaStream_t stream);

// rows up to which the persistent small-M kernel serves a linear; above, the rotation pre-pass + GEMM path (knob for A/B runs)
static int small_m_max() {
  static const int v = [] {
    const char *e = getenv("PARO_SMALL_M_MAX");
    const int n = e && *e ? atoi(e) : 16;
    return n < 0 ? 0 : n > 16 ? 16 : n;
  }();
  return v;
}

static bool valid_dtype(int d) { return d == PARO_F32 || d == PARO_F16 || d == PARO_BF16; }
static bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

}  // namespace paro

using namespace paro;

extern "C" {

int paro_abi_version(void) { return PARO_ABI_VERSION; }
const char *paro_last_error(void) { return g_err; }
int paro_last_launch_count(void) { return g_launches; }

int paro_rotate(const void *x, void *out, const int16_t *idx_ij, const void *theta, int32_t theta_dtype,
                const void *scales, int32_t scales_dtype, int64_t M, int32_t K, int32_t krot, int32_t group_size,
                int32_t dtype, paro_stream_t stream) {
  g_launches = 0;
  if (!x || !out || !idx_ij || !theta) { set_error("rotate: null pointer argument"); return PARO_EINVAL; }
  if (!valid_dtype(dtype)) { set_error("rotate supports Float, Half, and BFloat16, got dtype code %d", dtype); return PARO_EINVAL; }
  if (!valid_dtype(theta_dtype) || (scales && !valid_dtype(scales_dtype))) { set_error("rotate: bad theta/scales dtype"); return PARO_EINVAL; }
  if (group_size != 64 && group_size != 128) {  // rotation.cu:123
    set_error("Unsupported group_size: %d; expected 64 or 128", group_size);
    return PARO_EUNSUPPORTED;
  }
  if (K <= 0 || K % group_size) { set_error("h must be divisible by GROUP_SIZE"); return PARO_EINVAL; }  // rotation.cu:66
  if (krot < 1 || krot > 16) { set_error("Unsupported KROT = %d; supported: 1..16", krot); return PARO_EUNSUPPORTED; }
  if (M < 0) { set_error("rotate: negative row count"); return PARO_EINVAL; }
  if (!aligned(x, 16) || !aligned(out, 16) || !aligned(idx_ij, 4)) { set_error("rotate: x/out must be 16-byte aligned, idx_ij 4-byte"); return PARO_EINVAL; }
  return rotate_launch(x, out, idx_ij, theta, theta_dtype, scales, scales_dtype, M, K, krot, group_size, dtype,
                       static_cast<cudaStream_t>(stream));
}

int paro_rotate_backward(const void *y, const void *grad_out, const void *x, const int16_t *idx_ij, const void *theta, int32_t theta_dtype,
                         const void *scales, int32_t scales_dtype, void *grad_x, float *grad_theta, float *grad_scale, int64_t M, int32_t K,
                         int32_t krot, int32_t group_size, int32_t dtype, int32_t theta_formula, paro_stream_t stream) {
  g_launches = 0;
  if (theta_formula != PARO_THETA_GRADIENT && theta_formula != PARO_THETA_REFERENCE_EXPRESSION) { set_error("rotate_backward: unknown theta_formula %d", theta_formula); return PARO_EINVAL; }
  if (!y || !grad_out || !idx_ij || !theta || !grad_x || !grad_theta) { set_error("rotate_backward: null pointer argument"); return PARO_EINVAL; }
  if (grad_scale && (!x || !scales)) { set_error("rotate_backward: grad_scale needs x and scales"); return PARO_EINVAL; }
  if (!valid_dtype(dtype)) { set_error("rotate supports Float, Half, and BFloat16, got dtype code %d", dtype); return PARO_EINVAL; }
  if (!valid_dtype(theta_dtype) || (scales && !valid_dtype(scales_dtype))) { set_error("rotate_backward: bad theta/scales dtype"); return PARO_EINVAL; }
  if (group_size != 64 && group_size != 128) { set_error("Unsupported group_size: %d; expected 64 or 128", group_size); return PARO_EUNSUPPORTED; }
  if (K <= 0 || K % group_size) { set_error("h must be divisible by GROUP_SIZE"); return PARO_EINVAL; }
  if (krot < 1 || krot > 16) { set_error("Unsupported KROT = %d; supported: 1..16", krot); return PARO_EUNSUPPORTED; }
  if (M < 0) { set_error("rotate_backward: negative row count"); return PARO_EINVAL; }
  if (grad_x == y || grad_x == grad_out) { set_error("rotate_backward: grad_x must not alias y / grad_out"); return PARO_EINVAL; }
  if (!aligned(idx_ij, 4)) { set_error("rotate_backward: idx_ij must be 4-byte aligned"); return PARO_EINVAL; }
  return rotate_backward_launch(y, grad_out, x, idx_ij, theta, theta_dtype, scales, scales_dtype, grad_x, grad_theta, grad_scale, M, K, krot, group_size,
                                dtype, theta_formula, static_cast<cudaStream_t>(stream));
}

size_t paro_packed_bytes(const paro_linear_shape *shape) {
  Layout L;
  const char *why = "";
  if (!shape || !make_layout(*shape, L, &why)) { set_error("packed_bytes: %s", shape ? why : "null shape"); return 0; }
  return L.total_bytes;
}

int paro_prepack(const paro_linear_shape *shape, const int32_t *qweight, const int32_t *qzeros, const void *scales,
                 int32_t scales_dtype, const int16_t *pairs, const void *theta, int32_t theta_dtype,
                 const void *channel_scales, int32_t cs_dtype, void *packed, paro_stream_t stream) {
  g_launches = 0;
  Layout L;
  const char *why = "";
  if (!shape || !make_layout(*shape, L, &why)) { set_error("prepack: %s", shape ? why : "null shape"); return PARO_EINVAL; }
  if (!qweight || !qzeros || !scales || !pairs || !theta || !channel_scales || !packed) { set_error("prepack: null pointer argument"); return PARO_EINVAL; }
  if (!valid_dtype(scales_dtype) || !valid_dtype(theta_dtype) || !valid_dtype(cs_dtype)) { set_error("prepack: bad parameter dtype"); return PARO_EINVAL; }
  if (!aligned(packed, 128)) { set_error("prepack: packed buffer must be 128-byte aligned"); return PARO_EINVAL; }
  return prepack_launch(*shape, L, qweight, qzeros, scales, scales_dtype, pairs, theta, theta_dtype, channel_scales,
                        cs_dtype, packed, static_cast<cudaStream_t>(stream));
}

size_t paro_workspace_bytes(const paro_linear_shape *shape, int64_t max_m) {
  Layout L;
  const char *why = "";
  if (!shape || !make_layout(*shape, L, &why)) { set_error("workspace_bytes: %s", shape ? why : "null shape"); return 0; }
  // head: sync words + block counters of the small-M kernel (zero between calls); behind them scratch: the partial slots
  // of the small-M kernel or the rotated activations of the M > 16 path
  size_t b = stream_workspace_bytes(L, max_m);
  {
    const size_t d = stream_sync_bytes(L) + decode_scratch_bytes(L, max_m);
    if (d > b) b = d;
  }
  if (max_m > small_m_max()) {
    const size_t g = stream_sync_bytes(L) + gemm_workspace_bytes(L, max_m);
    if (g > b) b = g;
  }
  return b > 256 ? b : 256;
}

int paro_linear_forward(const paro_linear_shape *shape, const void *packed, const void *x, int64_t M, const void *bias,
                        void *y, void *workspace, size_t workspace_bytes, paro_stream_t stream) {
  g_launches = 0;
  Layout L;
  const char *why = "";
  if (!shape || !make_layout(*shape, L, &why)) { set_error("linear_forward: %s", shape ? why : "null shape"); return PARO_EINVAL; }
  if (M < 0) { set_error("linear_forward: negative row count"); return PARO_EINVAL; }
  if (M == 0) return PARO_OK;
  if (!packed || !x || !y || !workspace) { set_error("linear_forward: null pointer argument"); return PARO_EINVAL; }
  if (!aligned(packed, 128) || !aligned(x, 16) || !aligned(y, 16) || !aligned(workspace, 256)) {
    set_error("linear_forward: packed must be 128-byte, workspace 256-byte, x / y 16-byte aligned");
    return PARO_EINVAL;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (M <= small_m_max()) {
    // A single linear: the cluster kernel (paro_decode.cu: K slices reduced through distributed shared memory) has the shorter
    // tail -- measured 7.4 / 9.9 / 22.0 / 15.6 us on the Llama-3-8B shapes at M = 1 against 11.1 / 12.3 / 24.3 / 16.2 us for
    // the chain kernel run as a one-step chain, whose cross-CTA sums travel through L2.  The chain kernel takes over where the
    // cluster kernel has no launch plan (very long K) and for chains / tensor-parallel steps (paro_chain_forward).
    static const bool stream_only = [] { const char *v = getenv("PARO_DECODE_STREAM"); return v && *v && atoi(v) != 0; }();
    if (!stream_only) {
      const size_t head = stream_sync_bytes(L);   // the pre-rotated rows of the M >= 4 pre-pass live behind the epoch word
      const int rc = decode_forward(*shape, L, packed, x, M, bias, y, static_cast<uint8_t *>(workspace) + head,
                                    workspace_bytes > head ? workspace_bytes - head : 0, st);
      if (rc != PARO_EUNSUPPORTED) return rc;
    }
    return stream_linear_forward(*shape, L, packed, x, M, bias, y, workspace, workspace_bytes, st);
  }
  const size_t head = stream_sync_bytes(L);   // never touched by the large-M path
  if (workspace_bytes < head) { set_error("workspace too small: have %zu, need %zu", workspace_bytes, head); return PARO_EWORKSPACE; }
  return gemm_forward(*shape, L, packed, x, M, bias, y, static_cast<uint8_t *>(workspace) + head, workspace_bytes - head, st);
}

int paro_unpack_dense(const paro_linear_shape *shape, const void *packed, void *W_out, paro_stream_t stream) {
  g_launches = 0;
  Layout L;
  const char *why = "";
  if (!shape || !make_layout(*shape, L, &why)) { set_error("unpack_dense: %s", shape ? why : "null shape"); return PARO_EINVAL; }
  if (!packed || !W_out) { set_error("unpack_dense: null pointer argument"); return PARO_EINVAL; }
  return unpack_dense_launch(*shape, L, packed, W_out, static_cast<cudaStream_t>(stream));
}

int paro_debug_decode_plan(const paro_linear_shape *shape, int64_t M, int32_t sets, int32_t sms, const int32_t *resident_clusters,
                            int32_t *out) {
  Layout L;
  const char *why = "";
  if (!shape || !make_layout(*shape, L, &why)) { set_error("debug_decode_plan: %s", shape ? why : "null shape"); return PARO_EINVAL; }
  if (!resident_clusters || !out || sms < 1) { set_error("debug_decode_plan: bad arguments"); return PARO_EINVAL; }
  return decode_debug_plan(L, M, sets, sms, resident_clusters, out);
}

static int chain_to_host(const paro_chain_step *steps, int32_t n, int64_t M, HostStep *hs) {
  if (!steps || n < 1 || n > PARO_CHAIN_MAX_STEPS) { set_error("chain: 1..%d steps supported", PARO_CHAIN_MAX_STEPS); return PARO_EINVAL; }
  if (M < 1 || M > 16) { set_error("chain: 1 <= M <= 16 rows supported, got %lld", static_cast<long long>(M)); return PARO_EUNSUPPORTED; }
  for (int i = 0; i < n; ++i) {
    const paro_chain_step &c = steps[i];
    const char *why = "";
    if (!c.shape || !make_layout(*c.shape, hs[i].L, &why)) { set_error("chain: step %d: %s", i, c.shape ? why : "null shape"); return PARO_EINVAL; }
    if (!c.packed) { set_error("chain: step %d: null packed buffer", i); return PARO_EINVAL; }
    if (c.x_op < PARO_XOP_NONE || c.x_op > PARO_XOP_RMSNORM || c.epilogue < PARO_EPI_STORE || c.epilogue > PARO_EPI_ADD_RESIDUAL) {
      set_error("chain: step %d: unknown x_op / epilogue", i);
      return PARO_EINVAL;
    }
    if (!aligned(c.packed, 128) || !aligned(c.x, 16) || !aligned(c.y, 16) || !aligned(c.residual_in, 16) || !aligned(c.residual_out, 16)) {
      set_error("chain: step %d: packed must be 128-byte, activations 16-byte aligned", i);
      return PARO_EINVAL;
    }
    if (c.epilogue == PARO_EPI_ADD_RESIDUAL && c.residual_in == c.residual_out && c.residual_in) {
      set_error("chain: step %d: residual_out must not alias residual_in", i);
      return PARO_EINVAL;
    }
    hs[i].shape = c.shape;
    hs[i].packed = c.packed; hs[i].bias = c.bias; hs[i].x = c.x; hs[i].y = c.y;
    hs[i].x_op = c.x_op; hs[i].epi_op = c.epilogue;
    hs[i].res_in = c.residual_in; hs[i].res_out = c.residual_out; hs[i].norm_w = c.norm_weight; hs[i].eps = c.eps;
    hs[i].tp = c.tp;
    if (c.tp) {
      if (c.tp->world < 1 || c.tp->world > PARO_TP_MAX_RANKS || c.tp->rank < 0 || c.tp->rank >= c.tp->world) {
        set_error("chain: step %d: tensor-parallel world must be 1..%d and 0 <= rank < world", i, PARO_TP_MAX_RANKS);
        return PARO_EINVAL;
      }
      for (int r = 0; r < c.tp->world; ++r)
        if (c.tp->world > 1 && (!c.tp->peer_slots[r] || !aligned(c.tp->peer_slots[r], 256))) {
          set_error("chain: step %d: peer_slots[%d] must be a 256-byte aligned device address", i, r);
          return PARO_EINVAL;
        }
    }
  }
  return PARO_OK;
}

size_t paro_chain_workspace_bytes(const paro_chain_step *steps, int32_t n_steps, int64_t M) {
  HostStep hs[PARO_CHAIN_MAX_STEPS] = {};
  if (chain_to_host(steps, n_steps, M, hs) != PARO_OK) return 0;
  return stream_chain_workspace_bytes(hs, n_steps, M);
}

int paro_chain_forward(const paro_chain_step *steps, int32_t n_steps, int64_t M, void *workspace, size_t workspace_bytes,
                       paro_stream_t stream) {
  g_launches = 0;
  HostStep hs[PARO_CHAIN_MAX_STEPS] = {};
  const int rc = chain_to_host(steps, n_steps, M, hs);
  if (rc != PARO_OK) return rc;
  if (!workspace || !aligned(workspace, 256)) { set_error("chain: workspace must be a 256-byte aligned device buffer"); return PARO_EINVAL; }
  return stream_forward(hs, n_steps, M, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

size_t paro_tp_slot_bytes(const paro_linear_shape *shape, int64_t M, int32_t world) {
  Layout L;
  const char *why = "";
  if (!shape || !make_layout(*shape, L, &why)) { set_error("tp_slot_bytes: %s", shape ? why : "null shape"); return 0; }
  if (M < 1 || M > 16 || world < 1 || world > PARO_TP_MAX_RANKS) { set_error("tp_slot_bytes: 1 <= M <= 16, 1 <= world <= %d", PARO_TP_MAX_RANKS); return 0; }
  return static_cast<size_t>(2) * L.blocks_total * world * M * 128 * 8;
}

int paro_debug_stream_plan(const paro_linear_shape *shape, int64_t M, int32_t sets, int32_t ctas, int32_t *out) {
  Layout L;
  const char *why = "";
  if (!shape || !make_layout(*shape, L, &why)) { set_error("debug_stream_plan: %s", shape ? why : "null shape"); return PARO_EINVAL; }
  if (!out) { set_error("debug_stream_plan: bad arguments"); return PARO_EINVAL; }
  return stream_debug_plan(L, M, sets, ctas, out);
}

int paro_debug_stream_trace(unsigned long long *host_out, int32_t max_ctas) {
  if (!host_out || max_ctas <= 0) { set_error("debug_trace: bad arguments"); return PARO_EINVAL; }
  return stream_trace_read(host_out, max_ctas);
}

int paro_debug_trace(unsigned long long *host_out, int32_t max_ctas) {
  if (!host_out || max_ctas <= 0) { set_error("debug_trace: bad arguments"); return PARO_EINVAL; }
  return decode_trace_read(host_out, max_ctas);
}

}  // extern "C"
