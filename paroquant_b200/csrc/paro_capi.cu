// extern "C" surface of libparo_b200.so (declared in include/paro_b200.h): argument checks with
// the reference's error behaviour (TORCH_CHECK -> message, here: code + paro_last_error()),
// dispatch to the sm_100a kernels.  No allocation, no host synchronisation.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "paro_common.cuh"
#include "paro_layout.h"

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
int prepack_launch(const paro_linear_shape &s, const Layout &L, const int32_t *qweight, const int32_t *qzeros,
                   const void *scales, int scales_dtype, const int16_t *pairs, const void *theta, int theta_dtype,
                   const void *cscales, int cs_dtype, void *packed, cudaStream_t stream);
int unpack_dense_launch(const paro_linear_shape &s, const Layout &L, const void *packed, void *W, cudaStream_t stream);
bool decode_supported(const Layout &L, int64_t M);
int decode_forward(const paro_linear_shape &s, const Layout &L, const void *packed, const void *x, int64_t M,
                   const void *bias, void *y, cudaStream_t stream);
int decode_trace_read(unsigned long long *host, int max_ctas);
int decode_debug_plan(const Layout &L, int64_t M, int sets, int sms, const int32_t *resident4, int32_t *out20);
size_t gemm_workspace_bytes(const Layout &L, int64_t max_m);
int gemm_forward(const paro_linear_shape &s, const Layout &L, const void *packed, const void *x, int64_t M,
                 const void *bias, void *y, void *workspace, size_t workspace_bytes, cudaStream_t stream);

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
  // the small-M kernel reduces through distributed shared memory and needs no global scratch; M > 16 stages x_rot
  const size_t b = max_m > 16 ? gemm_workspace_bytes(L, max_m) : 0;
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
  if (M <= 16) return decode_forward(*shape, L, packed, x, M, bias, y, st);   // one persistent CTA per SM (paro_decode.cu)
  return gemm_forward(*shape, L, packed, x, M, bias, y, workspace, workspace_bytes, st);
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

int paro_debug_trace(unsigned long long *host_out, int32_t max_ctas) {
  if (!host_out || max_ctas <= 0) { set_error("debug_trace: bad arguments"); return PARO_EINVAL; }
  return decode_trace_read(host_out, max_ctas);
}

}  // extern "C"
