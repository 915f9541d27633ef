// Fused small-M path (M <= 16): scaled pairwise rotation of x + INT4 group dequant + GEMV/GEMM in
// ONE launch per (merged) linear.  Replaces the reference's rotate -> Marlin kernel pairs
// (/root/reference/paroquant/inference/backends/vllm/plugin.py:281-311: 2n+1 launches for an
// n-way merged projection) on the HBM-bound side of the roofline.
//
// Work split (byte layout: paro_layout.h):
//   CTA       = one K-slice (gps groups of 128 channels) x a contiguous range of 16-column tiles
//               of ONE partition.  The `slices` CTAs that share a tile range form a thread-block
//               CLUSTER; grid = ranges x slices, sized to one resident wave.
//   warp w<W  = consumer.  Owns GPW rotation/quantisation groups of the slice for the whole
//               kernel: rotates that part of x itself (rotations are group-local, so only
//               __syncwarp is needed), keeps it as mma B fragments in registers, and multiplies
//               it with its 16 x 128 weight blocks of every tile the CTA streams.
//   warp W    = producer.  One lane streams the CTA's records with cp.async.bulk (TMA) into an
//               mbarrier ring -- all issued BEFORE griddepcontrol.wait, so under programmatic
//               dependent launch the weights of linear i+1 are in flight while linear i still
//               computes; only x is read after the wait.
//   reduction = (1) inside the CTA the W warps' partial 16 x M tiles meet in the weight bytes they
//               have just consumed (no extra shared memory); warp (tile mod W) adds them in fixed
//               order.  (2) across the K-slices the per-tile partials are PUSHED through
//               distributed shared memory to cluster rank (tile mod slices); after ONE cluster
//               barrier every CTA finishes its share of tiles locally -- no global round trip, no
//               atomics, fixed summation order (bit-reproducible).  Shapes whose K has no such
//               factorisation (or M too large for the receive buffer) use a global fp32 workspace
//               + arrival counters instead (also fixed order).
//
// Numerics (identical to the reference pipeline's operand formation):
//   x_rot : rotation.cuh:91-173 rounding points (see paro_rotate.cu)
//   W     : T((q - z) * T(s)) -- (q - z) exact, ONE rounding in the multiply, like Marlin / AWQ
//   y     : fp32 accumulate on tensor cores (mma.sync m16n8k16; W is the 16-row operand so one
//           MMA consumes 8 weights per thread), one rounding to T, bias added in T.
#include "paro_common.cuh"
#include "paro_layout.h"

namespace paro {

constexpr int kMaxStages = 8;
constexpr int kMaxConsumerWarps = 8;
constexpr int kDecodeMaxThreads = 32 * (kMaxConsumerWarps + 1);

enum ReduceMode { kDirect = 0, kCluster = 1, kWorkspace = 2 };

struct DecodeParams {
  const uint8_t *packed;
  const void *x;
  void *y;
  const void *bias;
  float *partials;
  int *counters;
  int M, K, N;
  int n_parts, slices, groups, krot, nstages, tiles_total;
  int warps, gps, rec_bytes, stage_tiles, stage_stride;
  int mode, rot_bytes, recv_tiles, payload_bytes;
  int part_tile_begin[PARO_MAX_PARTS + 1];
  int part_range_begin[PARO_MAX_PARTS + 1];
  int meta_group_bytes;
  long long meta_off, rec_off;
};

// ------------------------------------------------------------------ INT4 -> T dequant
template <typename T> struct Dequant;

// (a & mask) | magic in ONE LOP3: both constants must sit in registers (LOP3 takes one immediate)
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t mask, uint32_t magic) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(mask), "r"(magic));
  return d;
}

template <> struct Dequant<__nv_bfloat16> {
  uint32_t s_lo, s_hi, z_lo, z_hi;
  // sc = {s[g], s[g+8]} as T bits, zz = {z[g], z[g+8]} as bytes
  __device__ __forceinline__ void prep(uint32_t sc, uint32_t zz) {
    s_lo = __byte_perm(sc, sc, 0x1010);
    s_hi = __byte_perm(sc, sc, 0x3232);
    z_lo = __byte_perm(zz, 0x43434343u, 0x4040);  // bf16x2 {128 + z, 128 + z}: bytes (z, 0x43, z, 0x43)
    z_hi = __byte_perm(zz, 0x43434343u, 0x4141);
  }
  __device__ __forceinline__ uint32_t one(uint32_t w, uint32_t z, uint32_t s) const {
    // 0x4300 | q is the bf16 128 + q; (128 + q) - (128 + z) is exact; one rounding in the multiply
    const __nv_bfloat162 d = __hsub2(unpack2<__nv_bfloat16>(and_or(w, 0x000F000Fu, 0x43004300u)), unpack2<__nv_bfloat16>(z));
    return pack2<__nv_bfloat16>(__hmul2(d, unpack2<__nv_bfloat16>(s)));
  }
  __device__ __forceinline__ void run(uint32_t w, uint32_t (&a)[4]) const {
    a[0] = one(w, z_lo, s_lo);
    a[1] = one(w >> 4, z_hi, s_hi);
    a[2] = one(w >> 8, z_lo, s_lo);
    a[3] = one(w >> 12, z_hi, s_hi);
  }
};

template <> struct Dequant<__half> {
  uint32_t s_lo, s_hi, z_lo, z_hi16;
  __device__ __forceinline__ void prep(uint32_t sc, uint32_t zz) {
    s_lo = __byte_perm(sc, sc, 0x1010);
    s_hi = __byte_perm(sc, sc, 0x3232);
    z_lo = (0x6400u | (zz & 0xFFu)) * 0x00010001u;                  // {1024 + z}
    z_hi16 = (0xD400u | (((zz >> 8) & 0xFFu) << 4)) * 0x00010001u;  // {-(64 + z)}
  }
  __device__ __forceinline__ void run(uint32_t w, uint32_t (&a)[4]) const {
    const uint32_t w8 = w >> 8;
    const __half2 sixteenth = unpack2<__half>(0x2C002C00u);
    // low nibble: 0x6400 | q = 1024 + q.  High nibble in place: 0x6400 | (q << 4) = 1024 + 16 q,
    // and fma(1024 + 16 q, 1/16, -(64 + z)) = q - z exactly.
    const __half2 d0 = __hsub2(unpack2<__half>(and_or(w, 0x000F000Fu, 0x64006400u)), unpack2<__half>(z_lo));
    const __half2 d1 = __hfma2(unpack2<__half>(and_or(w, 0x00F000F0u, 0x64006400u)), sixteenth, unpack2<__half>(z_hi16));
    const __half2 d2 = __hsub2(unpack2<__half>(and_or(w8, 0x000F000Fu, 0x64006400u)), unpack2<__half>(z_lo));
    const __half2 d3 = __hfma2(unpack2<__half>(and_or(w8, 0x00F000F0u, 0x64006400u)), sixteenth, unpack2<__half>(z_hi16));
    a[0] = pack2<__half>(__hmul2(d0, unpack2<__half>(s_lo)));
    a[1] = pack2<__half>(__hmul2(d1, unpack2<__half>(s_hi)));
    a[2] = pack2<__half>(__hmul2(d2, unpack2<__half>(s_lo)));
    a[3] = pack2<__half>(__hmul2(d3, unpack2<__half>(s_hi)));
  }
};

// ------------------------------------------------------------------ cluster / DSMEM helpers
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <typename T> __device__ __forceinline__ T bits_to_T(uint32_t b) {
  const uint16_t h = static_cast<uint16_t>(b);
  return *reinterpret_cast<const T *>(&h);
}
template <typename T> __device__ __forceinline__ uint16_t T_to_bits(T v) { return *reinterpret_cast<const uint16_t *>(&v); }

// ------------------------------------------------------------------ in-warp rotation of one group
// Tile `rot` = this warp's [128 channels][ROWS] elements of T, channel-major.  Lane owns pairs
// 2*lane and 2*lane+1 of every rotation; idxw = bytes (i0, j0, i1, j1).
template <typename T, int ROWS>
__device__ __forceinline__ void rotate_stage(uint32_t rot, uint32_t idxw, float c0, float s0, float c1, float s1) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const uint32_t ci = (idxw >> (16 * q)) & 0xFFu, cj = (idxw >> (16 * q + 8)) & 0xFFu;
    const float c = q ? c1 : c0, s = q ? s1 : s0;
    if constexpr (ROWS == 1) {
      const uint32_t ai = rot + ci * 2, aj = rot + cj * 2;
      const float a = Traits<T>::to_float(bits_to_T<T>(lds16(ai)));
      const float b = Traits<T>::to_float(bits_to_T<T>(lds16(aj)));
      float yi, yj;
      givens(c, s, a, b, yi, yj);
      sts16(ai, T_to_bits<T>(Traits<T>::from_float(yi)));
      sts16(aj, T_to_bits<T>(Traits<T>::from_float(yj)));
    } else {
      constexpr int MPW = ROWS / 2;
      const uint32_t ai = rot + ci * (MPW * 4), aj = rot + cj * (MPW * 4);
      uint32_t vi[MPW], vj[MPW];
#pragma unroll
      for (int u = 0; u < MPW; ++u) {
        vi[u] = lds32(ai + 4 * u);
        vj[u] = lds32(aj + 4 * u);
      }
#pragma unroll
      for (int u = 0; u < MPW; ++u) {
        const float2 a = Traits<T>::to_float2(unpack2<T>(vi[u]));
        const float2 b = Traits<T>::to_float2(unpack2<T>(vj[u]));
        float yix, yiy, yjx, yjy;
        givens(c, s, a.x, b.x, yix, yjx);
        givens(c, s, a.y, b.y, yiy, yjy);
        sts32(ai + 4 * u, pack2<T>(Traits<T>::from_floats(yix, yiy)));
        sts32(aj + 4 * u, pack2<T>(Traits<T>::from_floats(yjx, yjy)));
      }
    }
  }
}

// load the group's rows of x, multiply by the channel scales in T (one rounding, rotation.cuh:112-113)
// and lay them out channel-major in `rot`
template <typename T, int ROWS>
__device__ __forceinline__ void load_group(const DecodeParams &p, uint32_t rot, int gk, int lane, uint2 csw) {
  using T2 = typename Traits<T>::T2;
  const T2 sc01 = unpack2<T>(csw.x), sc23 = unpack2<T>(csw.y);
  const T *xg = static_cast<const T *>(p.x) + gk * kGroup + 4 * lane;
  if constexpr (ROWS == 1) {
    const uint2 raw = __ldcg(reinterpret_cast<const uint2 *>(xg));
    const uint32_t v01 = pack2<T>(__hmul2(unpack2<T>(raw.x), sc01));
    const uint32_t v23 = pack2<T>(__hmul2(unpack2<T>(raw.y), sc23));
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(rot + 8 * lane), "r"(v01), "r"(v23) : "memory");
  } else {
    constexpr int MPW = ROWS / 2;
    uint32_t v01[ROWS], v23[ROWS];
#pragma unroll
    for (int m = 0; m < ROWS; ++m) {
      uint2 raw = make_uint2(0u, 0u);
      if (m < p.M) raw = __ldcg(reinterpret_cast<const uint2 *>(xg + static_cast<int64_t>(m) * p.K));
      v01[m] = pack2<T>(__hmul2(unpack2<T>(raw.x), sc01));
      v23[m] = pack2<T>(__hmul2(unpack2<T>(raw.y), sc23));
    }
#pragma unroll
    for (int u = 0; u < MPW; ++u) {  // word u of a channel = rows (2u, 2u+1)
      const uint32_t base = rot + (4 * lane) * (MPW * 4) + 4 * u;
      sts32(base + 0 * (MPW * 4), __byte_perm(v01[2 * u], v01[2 * u + 1], 0x5410));
      sts32(base + 1 * (MPW * 4), __byte_perm(v01[2 * u], v01[2 * u + 1], 0x7632));
      sts32(base + 2 * (MPW * 4), __byte_perm(v23[2 * u], v23[2 * u + 1], 0x5410));
      sts32(base + 3 * (MPW * 4), __byte_perm(v23[2 * u], v23[2 * u + 1], 0x7632));
    }
  }
}

// B fragments of mma.m16n8k16 for the 8 k16-steps of a group: b[kk][0] = {x[m][16kk+2t], x[m][16kk+2t+1]},
// b[kk][1] = same at +8, m = 8*mb + lane/4 (zero beyond the rows held in the tile)
template <int ROWS>
__device__ __forceinline__ void load_bfrags(uint32_t rot, int lane, int mb, uint32_t (&b)[8][2]) {
  const int m = mb * 8 + (lane >> 2), t = lane & 3;
  const bool have = m < ROWS;
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = 16 * kk + 8 * h + 2 * t;
      if constexpr (ROWS == 1) {
        const uint32_t v = lds32(rot + c * 2);
        b[kk][h] = have ? v : 0u;
      } else {
        const uint32_t off = have ? 2 * m : 0;
        const uint32_t lo = lds16(rot + c * (ROWS * 2) + off);
        const uint32_t hi = lds16(rot + (c + 1) * (ROWS * 2) + off);
        b[kk][h] = have ? (lo | (hi << 16)) : 0u;
      }
    }
  }
}

template <typename T>
__device__ __forceinline__ void sincos2(uint32_t theta_pair_bits, float &c0, float &s0, float &c1, float &s1) {
  const float2 th = Traits<T>::to_float2(unpack2<T>(theta_pair_bits));
  __sincosf(th.x, &s0, &c0);
  __sincosf(th.y, &s1, &c1);
}

struct Hoisted {  // rotation coefficients of ONE group, computed before griddepcontrol.wait
  uint32_t idxw[8];
  float c0[8], s0[8], c1[8], s1[8];
};

// Rotate the GPW groups of this warp (interleaved stage by stage for ILP) and build their B fragments.
template <typename T, int ROWS, int MB, int GPW, bool HOIST>
__device__ __forceinline__ void rotate_and_fragment(const DecodeParams &p, uint32_t rot0, const uint8_t *const (&meta)[GPW],
                                                    const int (&gk)[GPW], const bool (&valid)[GPW], int lane,
                                                    const Hoisted &h, const uint2 (&csw)[GPW],
                                                    uint32_t (&bfr)[GPW][MB][8][2]) {
#pragma unroll
  for (int u = 0; u < GPW; ++u)
    if (valid[u]) load_group<T, ROWS>(p, rot0 + u * p.rot_bytes, gk[u], lane, csw[u]);
  __syncwarp();
  const int krot = p.krot;
  if constexpr (HOIST) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (valid[0]) rotate_stage<T, ROWS>(rot0, h.idxw[r], h.c0[r], h.s0[r], h.c1[r], h.s1[r]);
      __syncwarp();
    }
  } else {
    for (int r = 0; r < krot; ++r) {
#pragma unroll
      for (int u = 0; u < GPW; ++u)
        if (valid[u]) {
          const uint32_t iw = *reinterpret_cast<const uint32_t *>(meta[u] + r * 128 + 4 * lane);
          const uint32_t tw = *reinterpret_cast<const uint32_t *>(meta[u] + krot * 128 + r * 128 + 4 * lane);
          float c0, s0, c1, s1;
          sincos2<T>(tw, c0, s0, c1, s1);
          rotate_stage<T, ROWS>(rot0 + u * p.rot_bytes, iw, c0, s0, c1, s1);
        }
      __syncwarp();
    }
  }
#pragma unroll
  for (int u = 0; u < GPW; ++u)
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) {
      if (valid[u]) {
        load_bfrags<ROWS>(rot0 + u * p.rot_bytes, lane, mb, bfr[u][mb]);
      } else {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) bfr[u][mb][kk][0] = bfr[u][mb][kk][1] = 0u;
      }
    }
}

// ------------------------------------------------------------------ epilogue helpers
template <typename T, int MB>
__device__ __forceinline__ void store_tile(const DecodeParams &p, const float (&acc)[MB][4], int tile_g, int lane) {
  const int g = lane >> 2, t = lane & 3;
  T *y = static_cast<T *>(p.y);
  const T *bias = static_cast<const T *>(p.bias);
#pragma unroll
  for (int mb = 0; mb < MB; ++mb) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int m = mb * 8 + 2 * t + (e & 1);
      const int n = tile_g * kTileN + g + ((e >> 1) << 3);
      if (m < p.M) {
        T v = Traits<T>::from_float(acc[mb][e]);
        if (bias) v = Traits<T>::from_float(Traits<T>::to_float(v) + Traits<T>::to_float(bias[n]));  // plugin.py:309-310
        y[static_cast<int64_t>(m) * p.N + n] = v;
      }
    }
  }
}

// compact per-tile payload: lanes with 2t < rows of their m8 block carry a float4; slot = float4 index
template <int MB> __device__ __forceinline__ bool payload_slot(int M, int lane, int mb, int &slot) {
  const int g = lane >> 2, t = lane & 3;
  const int nt0 = MB == 1 ? (M + 1) / 2 : 4;
  slot = (mb == 0 ? t : nt0 + t) * 8 + g;
  return 2 * t < M - 8 * mb;
}

template <typename T, int MB, int GPW>
__global__ void __launch_bounds__(kDecodeMaxThreads, (MB * GPW <= 2 ? 2 : 1)) decode_kernel(const DecodeParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr bool HOIST = (GPW == 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int W = p.warps, nst = p.nstages;
  const uint32_t smem0 = smem_u32(smem);
  const uint32_t rot_all = smem0 + nst * p.stage_stride;
  const uint32_t recv = rot_all + W * GPW * p.rot_bytes;
  const uint32_t bars = recv + p.recv_tiles * p.slices * p.payload_bytes;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * kMaxStages, bar_part = bars + 16 * kMaxStages;

  // ---- which tile range / slice / partition is mine (pure arithmetic on launch constants)
  const int range = blockIdx.x / p.slices;
  const int slice = blockIdx.x - range * p.slices;  // == %cluster_ctarank in cluster mode
  int part = 0;
  while (range >= p.part_range_begin[part + 1]) ++part;
  const int jl = range - p.part_range_begin[part];
  const int cp = p.part_range_begin[part + 1] - p.part_range_begin[part];
  const int tp = p.part_tile_begin[part + 1] - p.part_tile_begin[part];
  const int t_begin = static_cast<int>(static_cast<long long>(jl) * tp / cp);
  const int t_end = static_cast<int>(static_cast<long long>(jl + 1) * tp / cp);
  const int ntiles = t_end - t_begin;
  const int RS = p.stage_tiles;
  const int nstage_iters = (ntiles + RS - 1) / RS;
  const uint8_t *rec_src = p.packed + p.rec_off +
                           (static_cast<size_t>(p.slices) * p.part_tile_begin[part] + static_cast<size_t>(slice) * tp + t_begin) * p.rec_bytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, W);
      mbar_init(bar_part + 8 * s, W);
    }
    fence_mbar_init();
  }
  __syncthreads();
  if (p.mode == kCluster) cluster_arrive_relaxed();  // #1 "this CTA runs" -- waited on before the first DSMEM push
  pdl_launch_dependents();                           // let the next linear in the stream start prefetching its weights

  if (warp >= W) {
    // ================= producer: stream the records; nothing here depends on the previous kernel
    if (warp == W && lane == 0) {
      const uint64_t pol = policy_evict_first();
      for (int s = 0; s < nstage_iters; ++s) {
        const int slot = s % nst, it = s / nst;
        if (it > 0) mbar_wait(bar_empty + 8 * slot, (it - 1) & 1);
        const int nrec = min(RS, ntiles - s * RS);
        const uint32_t bytes = nrec * p.rec_bytes;
        mbar_arrive_expect_tx(bar_full + 8 * slot, bytes);
        bulk_g2s(smem0 + slot * p.stage_stride, rec_src + static_cast<size_t>(s) * RS * p.rec_bytes, bytes, bar_full + 8 * slot, pol);
      }
    }
    return;
  }

  // ================= consumers
  int gk[GPW];
  bool valid[GPW];
  const uint8_t *meta[GPW];
  uint2 csw[GPW];
#pragma unroll
  for (int u = 0; u < GPW; ++u) {
    gk[u] = slice * p.gps + warp * GPW + u;
    valid[u] = gk[u] < p.groups;
    meta[u] = p.packed + p.meta_off + (static_cast<size_t>(part) * p.groups + (valid[u] ? gk[u] : 0)) * p.meta_group_bytes;
    csw[u] = valid[u] ? *reinterpret_cast<const uint2 *>(meta[u] + p.krot * 256 + 8 * lane) : make_uint2(0u, 0u);
  }
  // rotation coefficients are immutable metadata: fetched and run through MUFU before the wait
  Hoisted h;
  const bool hoisted = HOIST && p.krot == 8;
  if constexpr (HOIST) {
    if (valid[0] && hoisted) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        h.idxw[r] = *reinterpret_cast<const uint32_t *>(meta[0] + r * 128 + 4 * lane);
        sincos2<T>(*reinterpret_cast<const uint32_t *>(meta[0] + 8 * 128 + r * 128 + 4 * lane), h.c0[r], h.s0[r], h.c1[r], h.s1[r]);
      }
    }
  }
  const uint32_t rot0 = rot_all + warp * GPW * p.rot_bytes;

  pdl_wait();  // x (and the workspace) may have been written by the previous kernel

  uint32_t bfr[GPW][MB][8][2];
  {
    const int M = p.M;
#define PARO_ROT(ROWS)                                                                                   \
  do {                                                                                                   \
    if (hoisted) rotate_and_fragment<T, ROWS, MB, GPW, HOIST>(p, rot0, meta, gk, valid, lane, h, csw, bfr); \
    else rotate_and_fragment<T, ROWS, MB, GPW, false>(p, rot0, meta, gk, valid, lane, h, csw, bfr);        \
  } while (0)
    if constexpr (MB == 1) {
      if (M == 1) PARO_ROT(1);
      else if (M == 2) PARO_ROT(2);
      else if (M <= 4) PARO_ROT(4);
      else PARO_ROT(8);
    } else {
      PARO_ROT(16);
    }
#undef PARO_ROT
  }

  const int g = lane >> 2, t = lane & 3;
  const uint32_t part_off = (t * 8 + g) * 16;  // this lane's float4 inside a 512-byte partial block
  const int tile_g0 = p.part_tile_begin[part] + t_begin;
  const int mode = p.mode;
  const uint32_t sc_off = p.gps * kUnitWeightBytes, z_off = p.gps * (kUnitWeightBytes + 32);
  bool cluster_ready = false;

  for (int s = 0; s < nstage_iters; ++s) {
    const int slot = s % nst, it = s / nst;
    const int nrec = min(RS, ntiles - s * RS);
    const uint32_t st = smem0 + slot * p.stage_stride;
    mbar_wait(bar_full + 8 * slot, it & 1);

    for (int r = 0; r < nrec; ++r) {
      const uint32_t rec = st + r * p.rec_bytes;
      float d[MB][4];
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) d[mb][0] = d[mb][1] = d[mb][2] = d[mb][3] = 0.f;
#pragma unroll
      for (int u = 0; u < GPW; ++u) {
        const int unit = warp * GPW + u;
        const uint4 q0 = lds128(rec + unit * 1024 + lane * 16);
        const uint4 q1 = lds128(rec + unit * 1024 + 512 + lane * 16);
        Dequant<T> dq;
        dq.prep(lds32(rec + sc_off + unit * 32 + g * 4), lds16(rec + z_off + unit * 16 + g * 2));
        const uint32_t qw[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          uint32_t a[4];
          dq.run(qw[kk], a);
#pragma unroll
          for (int mb = 0; mb < MB; ++mb) mma16816<T>(d[mb], a, bfr[u][mb][kk][0], bfr[u][mb][kk][1]);
        }
      }
      // park the partial in the weight bytes this warp has just consumed for this tile
#pragma unroll
      for (int mb = 0; mb < MB; ++mb)
        sts128(rec + (warp * GPW) * 1024 + mb * 512 + part_off, make_float4(d[mb][0], d[mb][1], d[mb][2], d[mb][3]));
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_part + 8 * slot);

    bool waited = false;
    for (int r = 0; r < nrec; ++r) {
      const int ti = s * RS + r;  // tile index inside this CTA's range
      if (ti % W != warp) continue;
      if (!waited) { mbar_wait(bar_part + 8 * slot, it & 1); waited = true; }
      const uint32_t rec = st + r * p.rec_bytes;
      float acc[MB][4];
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        acc[mb][0] = acc[mb][1] = acc[mb][2] = acc[mb][3] = 0.f;
        for (int ww = 0; ww < W; ++ww) {  // fixed order: reproducible
          const float4 v = lds128f(rec + (ww * GPW) * 1024 + mb * 512 + part_off);
          acc[mb][0] += v.x; acc[mb][1] += v.y; acc[mb][2] += v.z; acc[mb][3] += v.w;
        }
      }
      const int tile_g = tile_g0 + ti;
      if (mode == kDirect) {
        store_tile<T, MB>(p, acc, tile_g, lane);
      } else if (mode == kCluster) {
        if (!cluster_ready) { cluster_wait_acquire(); cluster_ready = true; }  // #1: every CTA of the cluster runs
        const uint32_t dst_rank = ti % p.slices;
        const uint32_t local = recv + ((ti / p.slices) * p.slices + slice) * p.payload_bytes;
        const uint32_t remote = map_to_rank(local, dst_rank);
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
          int slot_f4;
          if (payload_slot<MB>(p.M, lane, mb, slot_f4))
            st_cluster_f4(remote + slot_f4 * 16, make_float4(acc[mb][0], acc[mb][1], acc[mb][2], acc[mb][3]));
        }
      } else {
#pragma unroll
        for (int mb = 0; mb < MB; ++mb)
          if (2 * t < p.M - 8 * mb)
            stcg128f(p.partials + ((static_cast<size_t>(slice) * p.tiles_total + tile_g) * MB + mb) * 128 + (t * 8 + g) * 4,
                     make_float4(acc[mb][0], acc[mb][1], acc[mb][2], acc[mb][3]));
      }
    }
    fence_proxy_async_smem();  // our generic-proxy writes into the stage precede its TMA refill
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8 * slot);
  }

  if (mode == kDirect) return;

  if (mode == kCluster) {
    // ---- one cluster barrier, then every CTA finishes the tiles it was sent (ti % slices == slice)
    if (!cluster_ready) cluster_wait_acquire();
    cluster_arrive_release();
    cluster_wait_acquire();
    const int nmine = ntiles > slice ? (ntiles - slice + p.slices - 1) / p.slices : 0;
    for (int jj = warp; jj < nmine; jj += W) {
      float acc[MB][4];
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        acc[mb][0] = acc[mb][1] = acc[mb][2] = acc[mb][3] = 0.f;
        int slot_f4;
        const bool on = payload_slot<MB>(p.M, lane, mb, slot_f4);
        for (int src = 0; src < p.slices; ++src) {  // fixed order over K-slices
          if (on) {
            const float4 v = lds128f(recv + (jj * p.slices + src) * p.payload_bytes + slot_f4 * 16);
            acc[mb][0] += v.x; acc[mb][1] += v.y; acc[mb][2] += v.z; acc[mb][3] += v.w;
          }
        }
      }
      store_tile<T, MB>(p, acc, tile_g0 + jj * p.slices + slice, lane);
    }
    return;
  }

  // ---- workspace mode: announce my tiles; whoever completes a tile sums the slices in fixed order
  __threadfence();
  __syncwarp();
  const int nmy = ntiles > warp ? (ntiles - warp + W - 1) / W : 0;
  for (int base = 0; base < nmy; base += 32) {
    const int qi = base + lane;
    const bool has = qi < nmy;
    const int tile_g = tile_g0 + qi * W + warp;
    const int old = has ? atomicAdd(p.counters + tile_g, 1) : 0;
    unsigned done = __ballot_sync(0xFFFFFFFFu, has && old == p.slices - 1);
    if (done) __threadfence();
    while (done) {
      const int bsrc = __ffs(done) - 1;
      done &= done - 1;
      const int tg = __shfl_sync(0xFFFFFFFFu, tile_g, bsrc);
      float acc[MB][4];
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) acc[mb][0] = acc[mb][1] = acc[mb][2] = acc[mb][3] = 0.f;
      for (int sl0 = 0; sl0 < p.slices; sl0 += 8) {  // 8 independent loads in flight, then add in order
        float4 v[8][MB];
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
          for (int mb = 0; mb < MB; ++mb) {
            v[k][mb] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (sl0 + k < p.slices && 2 * t < p.M - 8 * mb)
              v[k][mb] = ldcg128f(p.partials + ((static_cast<size_t>(sl0 + k) * p.tiles_total + tg) * MB + mb) * 128 + (t * 8 + g) * 4);
          }
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
          for (int mb = 0; mb < MB; ++mb) {
            acc[mb][0] += v[k][mb].x; acc[mb][1] += v[k][mb].y; acc[mb][2] += v[k][mb].z; acc[mb][3] += v[k][mb].w;
          }
      }
      store_tile<T, MB>(p, acc, tg, lane);
      if (lane == 0) p.counters[tg] = 0;  // leave the workspace zeroed for the next launch
    }
  }
}

// ------------------------------------------------------------------ host side
struct DecodePlan {
  int ranges;
  int part_range_begin[PARO_MAX_PARTS + 1];
  int grid;
  int max_tiles_per_range;
};

// Integer number of tile ranges per partition, proportional to the partition's tile count.
static bool make_plan(const Layout &L, int ranges, DecodePlan &plan) {
  if (ranges < L.n_parts) ranges = L.n_parts;
  if (ranges > L.tiles_total) ranges = L.tiles_total;
  int alloc[PARO_MAX_PARTS];
  double frac[PARO_MAX_PARTS];
  int used = 0;
  for (int p = 0; p < L.n_parts; ++p) {
    const int tp = L.part_tile_begin[p + 1] - L.part_tile_begin[p];
    const double exact = static_cast<double>(ranges) * tp / L.tiles_total;
    alloc[p] = static_cast<int>(exact);
    if (alloc[p] < 1) alloc[p] = 1;
    if (alloc[p] > tp) alloc[p] = tp;
    frac[p] = exact - alloc[p];
    used += alloc[p];
  }
  while (used < ranges) {  // largest remainder first
    int best = -1;
    for (int p = 0; p < L.n_parts; ++p) {
      const int tp = L.part_tile_begin[p + 1] - L.part_tile_begin[p];
      if (alloc[p] < tp && (best < 0 || frac[p] > frac[best])) best = p;
    }
    if (best < 0) break;
    alloc[best]++;
    frac[best] -= 1.0;
    used++;
  }
  while (used > ranges) {
    int best = -1;
    for (int p = 0; p < L.n_parts; ++p)
      if (alloc[p] > 1 && (best < 0 || frac[p] < frac[best])) best = p;
    if (best < 0) break;
    alloc[best]--;
    frac[best] += 1.0;
    used--;
  }
  plan.part_range_begin[0] = 0;
  plan.max_tiles_per_range = 0;
  for (int p = 0; p < PARO_MAX_PARTS; ++p) {
    plan.part_range_begin[p + 1] = plan.part_range_begin[p] + (p < L.n_parts ? alloc[p] : 0);
    if (p < L.n_parts) {
      const int tp = L.part_tile_begin[p + 1] - L.part_tile_begin[p];
      const int mx = (tp + alloc[p] - 1) / alloc[p];
      if (mx > plan.max_tiles_per_range) plan.max_tiles_per_range = mx;
    }
  }
  plan.ranges = plan.part_range_begin[L.n_parts];
  plan.grid = plan.ranges * L.slices;
  return plan.grid > 0;
}

size_t decode_workspace_bytes(const Layout &L, int64_t max_m) {
  const int MB = max_m <= 8 ? 1 : 2;
  const size_t counters = (static_cast<size_t>(L.tiles_total) * 4 + 255) / 256 * 256;
  if (L.slices <= 1) return counters;
  return counters + static_cast<size_t>(L.slices) * L.tiles_total * MB * 512;
}

static int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

static int payload_bytes_for(int M) {
  const int MB = M <= 8 ? 1 : 2;
  const int nt0 = MB == 1 ? (M + 1) / 2 : 4;
  const int nt1 = MB == 2 ? (M - 8 + 1) / 2 : 0;
  return (nt0 + nt1) * 128;
}

template <typename T, int MB, int GPW>
static int launch_decode(DecodeParams &p, const Layout &L, int sms, cudaStream_t stream) {
  auto kern = decode_kernel<T, MB, GPW>;
  const int threads = 32 * (p.warps + 1);
  const int ctas_per_sm = env_int("PARO_DECODE_CTAS_PER_SM", (MB * GPW <= 2) ? 2 : 1);
  const int ranges = sms * ctas_per_sm / L.slices;
  if (ranges < 1) {
    set_error("decode: in_features=%d needs %d K-slices, more than the %d resident CTAs", L.K, L.slices, sms * ctas_per_sm);
    return PARO_EUNSUPPORTED;
  }
  const bool want_cluster = L.plan.cluster > 1 && !env_int("PARO_NO_CLUSTER", 0);
  const int fixed = p.nstages * p.stage_stride + p.warps * GPW * p.rot_bytes + 3 * 8 * kMaxStages;
  p.payload_bytes = payload_bytes_for(p.M);
  DecodePlan plan;
  for (int attempt = 0; attempt < 2; ++attempt) {
    if (!make_plan(L, ranges, plan)) { set_error("decode: empty plan"); return PARO_EINVAL; }
    p.mode = L.slices == 1 ? kDirect : kWorkspace;
    p.recv_tiles = 0;
    if (want_cluster && attempt == 0) {
      const int recv_tiles = (plan.max_tiles_per_range + L.slices - 1) / L.slices;
      if (static_cast<size_t>(recv_tiles) * L.slices * p.payload_bytes <= 40 * 1024) {
        p.mode = kCluster;
        p.recv_tiles = recv_tiles;
      }
    }
    const size_t smem = fixed + static_cast<size_t>(p.recv_tiles) * L.slices * p.payload_bytes;
    PARO_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(plan.grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (p.mode == kCluster) {
      attr[na].id = cudaLaunchAttributeClusterDimension;
      attr[na].val.clusterDim.x = L.slices;
      attr[na].val.clusterDim.y = 1;
      attr[na].val.clusterDim.z = 1;
      ++na;
      // all clusters must be co-resident (one wave): shrink the number of tile ranges if needed
      int max_clusters = 0;
      cfg.attrs = attr;
      cfg.numAttrs = na;
      const cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
      if (e != cudaSuccess || max_clusters < L.n_parts) {
        (void)cudaGetLastError();
        continue;  // no cluster launch with this footprint: workspace mode
      }
      if (max_clusters < plan.ranges) {
        if (!make_plan(L, max_clusters, plan)) { set_error("decode: empty plan"); return PARO_EINVAL; }
        if ((plan.max_tiles_per_range + L.slices - 1) / L.slices > p.recv_tiles) continue;  // receive buffer too small now
        cfg.gridDim = dim3(plan.grid);
      }
    }
    if (!env_int("PARO_NO_PDL", 0)) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    for (int i = 0; i <= PARO_MAX_PARTS; ++i) p.part_range_begin[i] = plan.part_range_begin[i];
    PARO_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p));
    note_launches(1);
    return PARO_OK;
  }
  set_error("decode: no launch configuration found");
  return PARO_EUNSUPPORTED;
}

int decode_forward(const paro_linear_shape &s, const Layout &L, const void *packed, const void *x, int64_t M,
                   const void *bias, void *y, void *workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (workspace_bytes < decode_workspace_bytes(L, M)) {
    set_error("workspace too small: have %zu, need %zu", workspace_bytes, decode_workspace_bytes(L, M));
    return PARO_EWORKSPACE;
  }
  int dev = 0, sms = 0;
  PARO_CUDA_OK(cudaGetDevice(&dev));
  PARO_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int MB = M <= 8 ? 1 : 2;
  DecodeParams p;
  p.packed = static_cast<const uint8_t *>(packed);
  p.x = x; p.y = y; p.bias = bias;
  const size_t counters = (static_cast<size_t>(L.tiles_total) * 4 + 255) / 256 * 256;
  p.counters = static_cast<int *>(workspace);
  p.partials = reinterpret_cast<float *>(static_cast<uint8_t *>(workspace) + counters);
  p.M = static_cast<int>(M); p.K = L.K; p.N = L.N;
  p.n_parts = L.n_parts; p.slices = L.slices; p.groups = L.groups; p.krot = L.krot; p.tiles_total = L.tiles_total;
  p.warps = L.plan.warps; p.gps = L.gps; p.rec_bytes = L.rec_bytes;
  int rs = (17000 + L.rec_bytes / 2) / L.rec_bytes;
  if (rs < 1) rs = 1;
  p.stage_tiles = env_int("PARO_DECODE_STAGE_TILES", rs);
  p.stage_stride = (p.stage_tiles * L.rec_bytes + 127) / 128 * 128;
  int nst = env_int("PARO_DECODE_STAGES", 4);
  if (nst < 1) nst = 1;
  if (nst > kMaxStages) nst = kMaxStages;
  p.nstages = nst;
  p.rot_bytes = kGroup * 2 * (M == 1 ? 1 : M == 2 ? 2 : M <= 4 ? 4 : M <= 8 ? 8 : 16);
  for (int i = 0; i <= PARO_MAX_PARTS; ++i) p.part_tile_begin[i] = L.part_tile_begin[i];
  p.meta_group_bytes = L.meta_group_bytes;
  p.meta_off = static_cast<long long>(L.meta_off);
  p.rec_off = static_cast<long long>(L.rec_off);

  const int gpw = L.plan.gpw;
  if (s.dtype == PARO_BF16) {
    if (MB == 1) return gpw == 1 ? launch_decode<__nv_bfloat16, 1, 1>(p, L, sms, stream) : launch_decode<__nv_bfloat16, 1, 2>(p, L, sms, stream);
    return gpw == 1 ? launch_decode<__nv_bfloat16, 2, 1>(p, L, sms, stream) : launch_decode<__nv_bfloat16, 2, 2>(p, L, sms, stream);
  }
  if (MB == 1) return gpw == 1 ? launch_decode<__half, 1, 1>(p, L, sms, stream) : launch_decode<__half, 1, 2>(p, L, sms, stream);
  return gpw == 1 ? launch_decode<__half, 2, 1>(p, L, sms, stream) : launch_decode<__half, 2, 2>(p, L, sms, stream);
}

}  // namespace paro
