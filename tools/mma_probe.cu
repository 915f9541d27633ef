// Micro-probe: cost of small-N tcgen05.mma (kind::f16, M = 128, K = 16) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o paroquant_b200/lib/mma_probe tools/mma_probe.cu
//   ./mma_probe            -> table of cycles per MMA for N, A source (TMEM / smem), #accumulators, CTAs per SM
// One thread per CTA issues `rounds` x 8 MMAs, one commit per round; a second warp optionally keeps
// tcgen05.st traffic going (as the dequant workers do).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return static_cast<uint64_t>((addr >> 4) & 0x3FFF) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16) |
         (static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

struct Args { int N, ts, nacc, rounds, st_traffic, tmem_cols, m64; };

__global__ void __launch_bounds__(160, 1) probe(Args a, unsigned long long *out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_mem[2];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar = smem_u32(&bar_mem[0]);
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(a.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t fmt = 1u;  // bf16
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(a.N >> 3) << 17) | ((a.m64 ? 4u : 8u) << 24);
  const uint32_t smem0 = smem_u32(smem);
  if (warp == 0) {
    // the whole warp runs the loop, one elected lane issues (uniform registers, no divergence loop around UTCHMMA)
    const uint32_t d0 = tmem + 64;   // A operand columns 0..63, accumulators after
    const uint64_t bhi = desc_kmajor(0, a.N * 16, 128), ahi = desc_kmajor(0, 2048, 128);
    const uint32_t bstep = (a.N * 32) >> 4;
    const int nacc = a.nacc, N = a.N;
    const long long t0 = clock64();
    uint32_t elected;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
    if (elected) {
      for (int r = 0; r < a.rounds; ++r) {
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          const uint64_t bdesc = bhi | static_cast<uint64_t>(((smem0 >> 4) + s * bstep) & 0x3FFF);
          const uint32_t d = d0 + (s & (nacc - 1)) * N;
          if (a.ts) mma_ts(d, tmem + 8 * s, bdesc, idesc, 1u);
          else mma_ss(d, ahi | static_cast<uint64_t>(((smem0 + 16384 + s * 4096) >> 4) & 0x3FFF), bdesc, idesc, 1u);
        }
      }
      commit(bar);          // one commit: completes when every MMA above has
    }
    __syncwarp();
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (lane == 0) { out[blockIdx.x] = static_cast<unsigned long long>(t1 - t0); stop = 1; }
  } else if (warp >= 1 && a.st_traffic) {
    // background tcgen05.st traffic into the A columns of this warp's lane quarter
    const uint32_t taddr = tmem + ((static_cast<uint32_t>(32 * (warp & 3))) << 16);
    uint32_t v = lane;
    while (!stop) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr + 16 * c), "r"(v) : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(a.tmem_cols) : "memory");
}

int main() {
  unsigned long long *out;
  cudaMalloc(&out, 1024 * sizeof(unsigned long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d\n%-4s %-5s %-5s %-4s %-8s %-4s | cycles/MMA (per CTA)  | cycles/MMA per SM\n", sms, "N", "A", "nacc", "cta", "sttraf", "M");
  const int Ns[] = {16, 32, 64, 128, 256};
  for (int m64 = 0; m64 < 2; ++m64)
    for (int ts = 1; ts >= 0; --ts)
      for (int ni = 0; ni < 5; ++ni)
        for (int nacc = 1; nacc <= 8; nacc *= 2)
          for (int ctas = 1; ctas <= 2; ++ctas)
            for (int st = 0; st < 2; ++st) {
              const int N = Ns[ni];
              if (64 + nacc * N > 256) continue;
              if (nacc > 1 && N > 16) continue;
              if (st && !ts) continue;
              if (m64 && (N > 64 || st)) continue;
              Args a = {N, ts, nacc, 64, st, 256, m64};
              probe<<<sms * ctas, 160, 64 * 1024>>>(a, out);
              cudaError_t e = cudaDeviceSynchronize();
              if (e != cudaSuccess) { printf("N=%d ts=%d: %s\n", N, ts, cudaGetErrorString(e)); return 1; }
              unsigned long long h[1024];
              cudaMemcpy(h, out, sms * ctas * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
              double sum = 0;
              for (int i = 0; i < sms * ctas; ++i) sum += static_cast<double>(h[i]);
              const double per = sum / (sms * ctas) / (a.rounds * 8.0);
              printf("%-4d %-5s %-5d %-4d %-8d %-4d | %8.1f              | %8.1f\n", N, ts ? "tmem" : "smem", nacc, ctas, st, m64 ? 64 : 128, per, per / ctas);
            }
  return 0;
}
