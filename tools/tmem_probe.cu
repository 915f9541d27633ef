// Micro-probe: tcgen05.st / tcgen05.ld throughput per SM (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o paroquant_b200/lib/tmem_probe tools/tmem_probe.cu
// `nw` warps (lane quarter = warp % 4) each issue `iters` x (4 x tcgen05.st.32x32b.x16 + wait::st) into
// 64 columns; reports bytes per SM clock.  Same for one x64 store per iteration and for tcgen05.ld.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

struct Args { int mode, iters; };   // 0: st x16 x4 + wait, 1: st x64 + wait, 2: ld x16 x4 + wait, 3: 4 x st x16 no wait until end, 4: STS.128 64 B x 4

__global__ void __launch_bounds__(1024, 1) probe(Args a, unsigned long long *out) {
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) uint8_t sbuf[32 * 1024];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  // warps sharing a lane quarter use different column ranges
  const uint32_t taddr = tmem + ((static_cast<uint32_t>(32 * (warp & 3))) << 16) + 64 * ((warp >> 2) & 7);
  uint32_t v = lane + warp;
  uint32_t r[16];
  const long long t0 = clock64();
  for (int i = 0; i < a.iters; ++i) {
    if (a.mode == 0 || a.mode == 3) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr + 16 * c), "r"(v) : "memory");
      if (a.mode == 0) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    } else if (a.mode == 1) {
#pragma unroll
      for (int c = 0; c < 2; ++c)
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr + 32 * c), "r"(v) : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    } else if (a.mode == 2) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr + 16 * c) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        v += r[0] ^ r[15];
      }
    } else {
      const uint32_t sa = smem_u32(sbuf) + (threadIdx.x & 255) * 64 + ((threadIdx.x >> 8) & 1) * 16384;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(sa + 16 * c), "r"(v) : "memory");
    }
  }
  if (a.mode == 3) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  const long long t1 = clock64();
  if (lane == 0) out[blockIdx.x * 32 + warp] = static_cast<unsigned long long>(t1 - t0) + (v == 0xdeadbeef);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  unsigned long long *out;
  cudaMalloc(&out, 148 * 32 * sizeof(unsigned long long));
  const char *names[] = {"st x16 x4 + wait", "st x32 x2 + wait", "ld x16 x4 (+wait each)", "st x16 x4, one wait at end", "STS.128 x4 (64 B/thread)"};
  printf("%-28s %-6s | cycles/iter/warp | bytes/clk/SM\n", "mode", "warps");
  for (int mode = 0; mode < 5; ++mode)
    for (int nw = 4; nw <= 32; nw *= 2) {
      Args a = {mode, 2000};
      probe<<<148, nw * 32>>>(a, out);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("mode %d failed: %s\n", mode, cudaGetErrorString(cudaGetLastError())); return 1; }
      unsigned long long h[148 * 32];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      double mx = 0;
      for (int b = 0; b < 148; ++b) for (int w = 0; w < nw; ++w) mx += static_cast<double>(h[b * 32 + w]);
      const double per_iter = mx / (148.0 * nw) / a.iters;
      const double bytes = 32.0 * 64 * 4 * nw;   // per iteration, all warps of the SM
      printf("%-28s %-6d | %10.1f       | %8.1f\n", names[mode], nw, per_iter, bytes / per_iter);
    }
  return 0;
}
