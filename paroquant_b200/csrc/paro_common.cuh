// Shared device helpers for the sm_100a ParoQuant kernels: dtype traits with the reference's
// rounding points, mbarrier / bulk-copy (TMA) / PDL PTX wrappers, error plumbing.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/paro_b200.h"

namespace paro {

// ------------------------------------------------------------------ host-side error plumbing
void set_error(const char *fmt, ...);
void note_launches(int n);
#define PARO_CUDA_OK(expr)                                                                   \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      paro::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                       \
      return PARO_ECUDA;                                                                     \
    }                                                                                        \
  } while (0)

// ------------------------------------------------------------------ dtype traits
// Every conversion below is one round-to-nearest-even, matching rotation.cuh's Traits
// (/root/reference/paroquant/kernels/cuda/rotation.cuh:177-205).
template <typename T> struct Traits;

template <> struct Traits<__half> {
  using T2 = __half2;
  static constexpr int code = PARO_F16;
  __device__ static __forceinline__ float to_float(__half v) { return __half2float(v); }
  __device__ static __forceinline__ __half from_float(float v) { return __float2half_rn(v); }
  __device__ static __forceinline__ float2 to_float2(__half2 v) { return __half22float2(v); }
  __device__ static __forceinline__ __half2 from_floats(float a, float b) { return __floats2half2_rn(a, b); }
};

template <> struct Traits<float> {   // fp32 rotate (the optimiser's dtype): no rounding anywhere
  static constexpr int code = PARO_F32;
  __device__ static __forceinline__ float to_float(float v) { return v; }
  __device__ static __forceinline__ float from_float(float v) { return v; }
};

template <> struct Traits<__nv_bfloat16> {
  using T2 = __nv_bfloat162;
  static constexpr int code = PARO_BF16;
  __device__ static __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ static __forceinline__ __nv_bfloat16 from_float(float v) { return __float2bfloat16_rn(v); }
  __device__ static __forceinline__ float2 to_float2(__nv_bfloat162 v) { return __bfloat1622float2(v); }
  __device__ static __forceinline__ __nv_bfloat162 from_floats(float a, float b) { return __floats2bfloat162_rn(a, b); }
};

template <typename T> __device__ __forceinline__ uint32_t pack2(typename Traits<T>::T2 v) {
  return *reinterpret_cast<uint32_t *>(&v);
}
template <typename T> __device__ __forceinline__ typename Traits<T>::T2 unpack2(uint32_t u) {
  return *reinterpret_cast<typename Traits<T>::T2 *>(&u);
}

// Load one element of a parameter tensor stored as `src_dtype` and return it as float AFTER
// the cast to T that rotation.cu:75-78 performs with `.to(x.dtype)`.
template <typename T> __device__ __forceinline__ float load_param_as(const void *p, int64_t i, int src_dtype) {
  float f;
  if (src_dtype == PARO_F32) f = reinterpret_cast<const float *>(p)[i];
  else if (src_dtype == PARO_F16) f = __half2float(reinterpret_cast<const __half *>(p)[i]);
  else f = __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(p)[i]);
  return Traits<T>::to_float(Traits<T>::from_float(f));
}
template <> __device__ __forceinline__ float load_param_as<float>(const void *p, int64_t i, int src_dtype) {
  if (src_dtype == PARO_F32) return reinterpret_cast<const float *>(p)[i];
  if (src_dtype == PARO_F16) return __half2float(reinterpret_cast<const __half *>(p)[i]);
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(p)[i]);
}

// The reference's Givens update (rotation.cuh:46-58,136-153; SASS of its build):
//   yi = fma(c, a, s*b)   yj = fma(c, b, s*(-a))     -- fp32, .ftz under --use_fast_math
__device__ __forceinline__ void givens(float c, float s, float a, float b, float &yi, float &yj) {
  yi = fmaf(c, a, s * b);
  yj = fmaf(c, b, s * -a);
}

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// L2 policy for data streamed exactly once (the packed weights)
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

// TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar,
                                         uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}

// generic-proxy writes to shared memory must be fenced before the async proxy (TMA) reuses it
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Programmatic dependent launch: everything before pdl_wait() may overlap the previous kernel in
// the stream; nothing written by that kernel may be read before it.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}

}  // namespace paro
