// Micro-probe: the per-launch floor of a kernel shaped like the small-M kernel (paro_decode.cu) in a
// back-to-back chain: how long is one link when the kernel does (a) nothing, (b) TMEM alloc + barrier
// init + __syncthreads, (c) + one cluster barrier pair, with and without programmatic dependent launch.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o paroquant_b200/lib/launch_probe tools/launch_probe.cu
// Prints microseconds per launch (CUDA graph of 64 launches, replayed 50 times).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

struct Args { int tmem, cluster_barrier, pdl, touch; };

__global__ void __launch_bounds__(704, 1) link(Args a, float *buf) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bars[64];
  const int warp = threadIdx.x >> 5;
  if (a.tmem) {
    if (threadIdx.x < 48) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[threadIdx.x])), "r"(1));
    if (warp == 21) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (a.cluster_barrier) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  if (a.pdl) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  if (a.touch) {   // the dependency: every CTA reads what the previous link wrote, then writes its own slot
    const float v = __ldcg(buf + ((blockIdx.x + 1) % gridDim.x) * 32 + (threadIdx.x & 31));
    if (threadIdx.x < 32) __stcg(buf + blockIdx.x * 32 + threadIdx.x, v + 1.f);
  }
  if (a.cluster_barrier) {
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (a.tmem) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 21) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(512) : "memory");
  }
}

static float run(Args a, int cluster, int smem_bytes, int grid, float *buf) {
  cudaFuncSetAttribute(link, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaStream_t s;
  cudaStreamCreate(&s);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(704);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (a.pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < 64; ++i) cudaLaunchKernelEx(&cfg, link, a, buf);
  if (cudaStreamEndCapture(s, &g) != cudaSuccess || cudaGraphInstantiate(&ge, g, 0) != cudaSuccess) { printf("capture failed: %s\n", cudaGetErrorString(cudaGetLastError())); return -1.f; }
  for (int i = 0; i < 3; ++i) cudaGraphLaunch(ge, s);
  cudaStreamSynchronize(s);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, s);
  for (int i = 0; i < 50; ++i) cudaGraphLaunch(ge, s);
  cudaEventRecord(e1, s);
  cudaStreamSynchronize(s);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g); cudaStreamDestroy(s);
  return ms * 1e3f / (50 * 64);
}

int main() {
  float *buf;
  cudaMalloc(&buf, 148 * 32 * sizeof(float));
  cudaMemset(buf, 0, 148 * 32 * sizeof(float));
  printf("%-52s  us per launch\n", "configuration (704 threads, 211 KB smem, graph of 64 launches)");
  for (int pdl = 0; pdl < 2; ++pdl)
    for (int cl = 1; cl <= 4; cl *= 4)
      for (int tm = 0; tm < 2; ++tm) {
        Args a = {tm, cl > 1, pdl, 1};
        const int grid = cl > 1 ? 132 : 148;
        const float us = run(a, cl, 211 * 1024, grid, buf);
        char name[96];
        snprintf(name, sizeof name, "pdl=%d cluster=%d tmem+barriers=%d", pdl, cl, tm);
        printf("%-52s  %7.2f\n", name, us);
      }
  if (cudaDeviceSynchronize() != cudaSuccess) printf("error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
