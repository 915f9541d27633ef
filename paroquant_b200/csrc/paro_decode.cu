// Fused small-M path (M <= 16): scaled pairwise rotation of x + INT4 group dequant + GEMV/GEMM in
// ONE launch per (merged) linear.  Replaces the reference's rotate -> Marlin kernel pairs
// (/root/reference/paroquant/inference/backends/vllm/plugin.py:281-311: 2n+1 launches for an
// n-way merged projection) on the HBM-bound side of the roofline.
//
// Work split (see paro_layout.h for the byte layout):
//   grid     = slices x ctas_per_slice  (<= 2..3 resident CTAs per SM, one wave, no tail)
//   CTA      = one K-slice of 512 channels x a contiguous range of 16-column tiles of ONE partition
//   warp 0-3 = consumer; warp w owns rotation/quantisation group  slice*4 + w  (128 channels):
//              it rotates that part of x itself (rotations are group-local -> __syncwarp only),
//              keeps it as mma B fragments in registers, and multiplies it with the group's
//              16 x 128 weight block of every tile the CTA streams
//   warp 4   = producer; one lane streams the CTA's records with cp.async.bulk (TMA) into an
//              mbarrier ring.  All of it is issued BEFORE griddepcontrol.wait, so under
//              programmatic dependent launch the weights of linear i+1 are already in flight
//              while linear i is still computing; only x is read after the wait.
//   per stage (4 tiles): the four warps' partial 16 x M tiles are exchanged through the weight
//              bytes they have just consumed (no extra shared memory), warp r reduces tile r in
//              fixed order 0..3 and either stores y (one slice) or publishes the slice's partial
//              to the split-K workspace; the CTA that completes a tile (arrival counter) adds the
//              slices in fixed order, so results are bit-reproducible run to run.
//
// Numerics (identical to the reference pipeline's operand formation):
//   x_rot : rotation.cuh:91-173 rounding points (see paro_rotate.cu)
//   W     : T((q - z) * T(s)) -- (q - z) exact, ONE rounding in the multiply, like Marlin / AWQ
//   y     : fp32 accumulate on tensor cores (mma.sync m16n8k16; W is the 16-row operand so one
//           MMA consumes 8 weights per thread), one rounding to T, bias added in T.
#include "paro_common.cuh"
#include "paro_layout.h"

namespace paro {

constexpr int kDecodeThreads = 160;
constexpr int kMaxStages = 8;

struct DecodeParams {
  const uint8_t *packed;
  const void *x;
  void *y;
  const void *bias;
  float *partials;
  int *counters;
  int M, K, N;
  int n_parts, slices, groups, krot, nstages, tiles_total;
  int rot_bytes;  // per consumer warp: 128 channels x (padded rows) x sizeof(T)
  int part_tile_begin[PARO_MAX_PARTS + 1];
  int part_cta_begin[PARO_MAX_PARTS + 1];
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
    z_lo = (0x4300u | (zz & 0xFFu)) * 0x00010001u;          // bf16x2 {128 + z, 128 + z}
    z_hi = (0x4300u | ((zz >> 8) & 0xFFu)) * 0x00010001u;
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
    z_lo = (0x6400u | (zz & 0xFFu)) * 0x00010001u;                    // {1024 + z}
    z_hi16 = (0xD400u | (((zz >> 8) & 0xFFu) << 4)) * 0x00010001u;    // {-(64 + z)}
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

// ------------------------------------------------------------------ in-warp rotation of one group
// rot: this warp's [128 channels][MPW words] tile (word u = rows 2u, 2u+1).  Lane owns pairs
// 2*lane and 2*lane+1 of every rotation; idxw = bytes (i0, j0, i1, j1).
template <typename T, int MPW>
__device__ __forceinline__ void rotate_stage(uint32_t rot, uint32_t idxw, float c0, float s0, float c1, float s1) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const uint32_t ai = rot + ((idxw >> (16 * q)) & 0xFFu) * (MPW * 4);
    const uint32_t aj = rot + ((idxw >> (16 * q + 8)) & 0xFFu) * (MPW * 4);
    const float c = q ? c1 : c0, s = q ? s1 : s0;
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

template <typename T, int MPW, int KROT>
__device__ __forceinline__ void rotate_group(const DecodeParams &p, uint32_t rot, const uint8_t *meta, int gk, int lane,
                                             const uint32_t (&idxw)[8], const float (&cs0)[8], const float (&sn0)[8],
                                             const float (&cs1)[8], const float (&sn1)[8], uint2 csw) {
  using T2 = typename Traits<T>::T2;
  constexpr int MP = MPW * 2;
  // ---- load x rows, multiply by the channel scales in T (one rounding, rotation.cuh:112-113)
  const T2 sc01 = unpack2<T>(csw.x), sc23 = unpack2<T>(csw.y);
  uint32_t v01[MP], v23[MP];
#pragma unroll
  for (int m = 0; m < MP; ++m) {
    uint2 raw = make_uint2(0u, 0u);
    if (m < p.M)
      raw = __ldcg(reinterpret_cast<const uint2 *>(static_cast<const T *>(p.x) + static_cast<int64_t>(m) * p.K + gk * kGroup + 4 * lane));
    v01[m] = pack2<T>(__hmul2(unpack2<T>(raw.x), sc01));
    v23[m] = pack2<T>(__hmul2(unpack2<T>(raw.y), sc23));
  }
  // transpose to channel-major: channel 4*lane + c, word u = rows (2u, 2u+1)
#pragma unroll
  for (int u = 0; u < MPW; ++u) {
    const uint32_t base = rot + (4 * lane) * (MPW * 4) + 4 * u;
    sts32(base + 0 * (MPW * 4), __byte_perm(v01[2 * u], v01[2 * u + 1], 0x5410));
    sts32(base + 1 * (MPW * 4), __byte_perm(v01[2 * u], v01[2 * u + 1], 0x7632));
    sts32(base + 2 * (MPW * 4), __byte_perm(v23[2 * u], v23[2 * u + 1], 0x5410));
    sts32(base + 3 * (MPW * 4), __byte_perm(v23[2 * u], v23[2 * u + 1], 0x7632));
  }
  __syncwarp();
  if constexpr (KROT == 8) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      rotate_stage<T, MPW>(rot, idxw[r], cs0[r], sn0[r], cs1[r], sn1[r]);
      __syncwarp();
    }
  } else {
    for (int r = 0; r < p.krot; ++r) {
      const uint32_t iw = *reinterpret_cast<const uint32_t *>(meta + r * 128 + 4 * lane);
      const uint32_t tw = *reinterpret_cast<const uint32_t *>(meta + p.krot * 128 + r * 128 + 4 * lane);
      const float2 th = Traits<T>::to_float2(unpack2<T>(tw));
      float c0, s0, c1, s1;
      __sincosf(th.x, &s0, &c0);
      __sincosf(th.y, &s1, &c1);
      rotate_stage<T, MPW>(rot, iw, c0, s0, c1, s1);
      __syncwarp();
    }
  }
}

// B fragments of mma.m16n8k16 for the 8 k16-steps of the group: b[kk][0] = {x[m][16kk+2t], x[m][16kk+2t+1]},
// b[kk][1] = same at +8, m = 8*mb + lane/4 (zero beyond the rows held in the tile)
template <int MPW>
__device__ __forceinline__ void load_bfrags(uint32_t rot, int lane, int mb, uint32_t (&b)[8][2]) {
  constexpr int MP = MPW * 2;
  const int m = mb * 8 + (lane >> 2), t = lane & 3;
  const bool have = m < MP;
  const uint32_t off = have ? 2 * m : 0;
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = 16 * kk + 8 * h + 2 * t;
      const uint32_t lo = lds16(rot + c * (MPW * 4) + off);
      const uint32_t hi = lds16(rot + (c + 1) * (MPW * 4) + off);
      b[kk][h] = have ? (lo | (hi << 16)) : 0u;
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

template <typename T, int MB, int KROT>
__global__ void __launch_bounds__(kDecodeThreads, (MB == 1 ? 3 : 2)) decode_kernel(const DecodeParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nst = p.nstages;
  uint8_t *stage_base = smem;
  const uint32_t rot_all = smem_u32(smem + static_cast<size_t>(nst) * kStageBytes);
  const int kRotBytesPerWarp = p.rot_bytes;
  const uint32_t bars = rot_all + 4 * kRotBytesPerWarp;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * kMaxStages, bar_part = bars + 16 * kMaxStages;

  // ---- which slice / partition / tile range is mine (pure arithmetic on launch constants)
  const int cps = p.part_cta_begin[p.n_parts];
  const int slice = blockIdx.x / cps;
  const int j = blockIdx.x - slice * cps;
  int part = 0;
  while (j >= p.part_cta_begin[part + 1]) ++part;
  const int jl = j - p.part_cta_begin[part];
  const int cp = p.part_cta_begin[part + 1] - p.part_cta_begin[part];
  const int tp = p.part_tile_begin[part + 1] - p.part_tile_begin[part];
  const int t_begin = static_cast<int>(static_cast<long long>(jl) * tp / cp);
  const int t_end = static_cast<int>(static_cast<long long>(jl + 1) * tp / cp);
  const int ntiles = t_end - t_begin;
  const int nstage_iters = (ntiles + kStageRecs - 1) / kStageRecs;
  const uint8_t *rec_src = p.packed + p.rec_off +
                           (static_cast<size_t>(p.slices) * p.part_tile_begin[part] + static_cast<size_t>(slice) * tp + t_begin) * kRecBytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 4);
      mbar_init(bar_part + 8 * s, 4);
    }
    fence_mbar_init();
  }
  __syncthreads();
  pdl_launch_dependents();  // let the next linear in the stream start prefetching its weights

  if (warp == 4) {
    // ================= producer: stream the records; nothing here depends on the previous kernel
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();
      for (int s = 0; s < nstage_iters; ++s) {
        const int slot = s % nst, it = s / nst;
        if (it > 0) mbar_wait(bar_empty + 8 * slot, (it - 1) & 1);
        const int nrec = min(kStageRecs, ntiles - s * kStageRecs);
        const uint32_t bytes = nrec * kRecBytes;
        mbar_arrive_expect_tx(bar_full + 8 * slot, bytes);
        bulk_g2s(smem_u32(stage_base + static_cast<size_t>(slot) * kStageBytes),
                 rec_src + static_cast<size_t>(s) * kStageBytes, bytes, bar_full + 8 * slot, pol);
      }
    }
    return;
  }

  // ================= consumers
  const int gk = slice * kSliceGroups + warp;
  const bool valid = gk < p.groups;
  const uint8_t *meta = p.packed + p.meta_off + (static_cast<size_t>(part) * p.groups + (valid ? gk : 0)) * p.meta_group_bytes;
  const uint32_t rot = rot_all + warp * kRotBytesPerWarp;

  // rotation coefficients: immutable metadata, so fetched and run through MUFU before the wait
  uint32_t idxw[8];
  float cs0[8], sn0[8], cs1[8], sn1[8];
  uint2 csw = make_uint2(0u, 0u);
  if (valid) {
    csw = *reinterpret_cast<const uint2 *>(meta + p.krot * 256 + 8 * lane);
    if constexpr (KROT == 8) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        idxw[r] = *reinterpret_cast<const uint32_t *>(meta + r * 128 + 4 * lane);
        const uint32_t tw = *reinterpret_cast<const uint32_t *>(meta + 8 * 128 + r * 128 + 4 * lane);
        const float2 th = Traits<T>::to_float2(unpack2<T>(tw));
        __sincosf(th.x, &sn0[r], &cs0[r]);
        __sincosf(th.y, &sn1[r], &cs1[r]);
      }
    }
  }

  pdl_wait();  // x (and the split-K workspace) may have been written by the previous kernel

  uint32_t bfr[MB][8][2];
  if (valid) {
    const int M = p.M;
    if (MB == 1 && M <= 2) {
      rotate_group<T, 1, KROT>(p, rot, meta, gk, lane, idxw, cs0, sn0, cs1, sn1, csw);
      load_bfrags<1>(rot, lane, 0, bfr[0]);
    } else if (MB == 1 && M <= 4) {
      rotate_group<T, 2, KROT>(p, rot, meta, gk, lane, idxw, cs0, sn0, cs1, sn1, csw);
      load_bfrags<2>(rot, lane, 0, bfr[0]);
    } else if (MB == 1) {
      rotate_group<T, 4, KROT>(p, rot, meta, gk, lane, idxw, cs0, sn0, cs1, sn1, csw);
      load_bfrags<4>(rot, lane, 0, bfr[0]);
    } else {
      rotate_group<T, 4 * MB, KROT>(p, rot, meta, gk, lane, idxw, cs0, sn0, cs1, sn1, csw);
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) load_bfrags<4 * MB>(rot, lane, mb, bfr[mb]);
    }
  } else {
#pragma unroll
    for (int mb = 0; mb < MB; ++mb)
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) bfr[mb][kk][0] = bfr[mb][kk][1] = 0u;
  }

  const int g = lane >> 2, t = lane & 3;
  const uint32_t part_off = ((t * 8 + g) * 16);  // this lane's float4 inside a 512-byte partial block
  const int tile_g0 = p.part_tile_begin[part] + t_begin;
  const bool split = p.slices > 1;

  for (int s = 0; s < nstage_iters; ++s) {
    const int slot = s % nst, it = s / nst;
    const int nrec = min(kStageRecs, ntiles - s * kStageRecs);
    const uint32_t st = smem_u32(stage_base + static_cast<size_t>(slot) * kStageBytes);
    mbar_wait(bar_full + 8 * slot, it & 1);

    float own[MB][4];
#pragma unroll
    for (int mb = 0; mb < MB; ++mb) own[mb][0] = own[mb][1] = own[mb][2] = own[mb][3] = 0.f;

    for (int r = 0; r < nrec; ++r) {
      const uint32_t rec = st + r * kRecBytes;
      const uint4 q0 = lds128(rec + warp * 1024 + lane * 16);
      const uint4 q1 = lds128(rec + warp * 1024 + 512 + lane * 16);
      Dequant<T> dq;
      dq.prep(lds32(rec + kRecScaleOff + warp * 32 + g * 4), lds16(rec + kRecZeroOff + warp * 16 + g * 2));
      const uint32_t qw[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
      float d[MB][4];
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) d[mb][0] = d[mb][1] = d[mb][2] = d[mb][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        uint32_t a[4];
        dq.run(qw[kk], a);
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) mma16816<T>(d[mb], a, bfr[mb][kk][0], bfr[mb][kk][1]);
      }
      if (r == warp) {
#pragma unroll
        for (int mb = 0; mb < MB; ++mb)
#pragma unroll
          for (int e = 0; e < 4; ++e) own[mb][e] = d[mb][e];
      } else {
        // park the partial in the weight bytes this warp has just consumed for tile r
#pragma unroll
        for (int mb = 0; mb < MB; ++mb)
          sts128(rec + warp * 1024 + mb * 512 + part_off, make_float4(d[mb][0], d[mb][1], d[mb][2], d[mb][3]));
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_part + 8 * slot);

    if (warp < nrec) {
      mbar_wait(bar_part + 8 * slot, it & 1);
      const uint32_t rec = st + warp * kRecBytes;
      float acc[MB][4];
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        acc[mb][0] = acc[mb][1] = acc[mb][2] = acc[mb][3] = 0.f;
#pragma unroll
        for (int ww = 0; ww < 4; ++ww) {  // fixed order: reproducible
          float4 v;
          if (ww == warp) v = make_float4(own[mb][0], own[mb][1], own[mb][2], own[mb][3]);
          else v = lds128f(rec + ww * 1024 + mb * 512 + part_off);
          acc[mb][0] += v.x; acc[mb][1] += v.y; acc[mb][2] += v.z; acc[mb][3] += v.w;
        }
      }
      const int tile_g = tile_g0 + s * kStageRecs + warp;
      if (!split) {
        store_tile<T, MB>(p, acc, tile_g, lane);
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

  if (!split) return;

  // ---- split-K: announce my tiles; whoever completes a tile sums the slices in fixed order
  __threadfence();
  __syncwarp();
  const int nmy = ntiles > warp ? (ntiles - warp + kStageRecs - 1) / kStageRecs : 0;
  for (int base = 0; base < nmy; base += 32) {
    const int qi = base + lane;
    const bool has = qi < nmy;
    const int tile_g = tile_g0 + qi * kStageRecs + warp;
    const int old = has ? atomicAdd(p.counters + tile_g, 1) : 0;
    unsigned done = __ballot_sync(0xFFFFFFFFu, has && old == p.slices - 1);
    while (done) {
      const int bsrc = __ffs(done) - 1;
      done &= done - 1;
      const int tg = __shfl_sync(0xFFFFFFFFu, tile_g, bsrc);
      __threadfence();
      float acc[MB][4];
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) acc[mb][0] = acc[mb][1] = acc[mb][2] = acc[mb][3] = 0.f;
      for (int sl = 0; sl < p.slices; ++sl) {
#pragma unroll
        for (int mb = 0; mb < MB; ++mb)
          if (2 * t < p.M - 8 * mb) {
            const float4 v = ldcg128f(p.partials + ((static_cast<size_t>(sl) * p.tiles_total + tg) * MB + mb) * 128 + (t * 8 + g) * 4);
            acc[mb][0] += v.x; acc[mb][1] += v.y; acc[mb][2] += v.z; acc[mb][3] += v.w;
          }
      }
      store_tile<T, MB>(p, acc, tg, lane);
      if (lane == 0) p.counters[tg] = 0;  // leave the workspace zeroed for the next launch
    }
  }
}

// ------------------------------------------------------------------ host side
struct DecodePlan {
  int ctas_per_slice;
  int part_cta_begin[PARO_MAX_PARTS + 1];
  int grid;
};

// Integer number of CTAs per (partition, slice), proportional to the partition's tile count.
static bool make_plan(const Layout &L, int max_ctas, DecodePlan &plan) {
  int cps = max_ctas / L.slices;
  if (cps < L.n_parts) cps = L.n_parts;  // at least one CTA per partition and slice
  if (cps > L.tiles_total) cps = L.tiles_total;
  int alloc[PARO_MAX_PARTS];
  double frac[PARO_MAX_PARTS];
  int used = 0;
  for (int p = 0; p < L.n_parts; ++p) {
    const int tp = L.part_tile_begin[p + 1] - L.part_tile_begin[p];
    const double exact = static_cast<double>(cps) * tp / L.tiles_total;
    alloc[p] = static_cast<int>(exact);
    if (alloc[p] < 1) alloc[p] = 1;
    if (alloc[p] > tp) alloc[p] = tp;
    frac[p] = exact - alloc[p];
    used += alloc[p];
  }
  while (used < cps) {  // largest remainder first
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
  while (used > cps) {
    int best = -1;
    for (int p = 0; p < L.n_parts; ++p)
      if (alloc[p] > 1 && (best < 0 || frac[p] < frac[best])) best = p;
    if (best < 0) break;
    alloc[best]--;
    frac[best] += 1.0;
    used--;
  }
  plan.part_cta_begin[0] = 0;
  for (int p = 0; p < PARO_MAX_PARTS; ++p)
    plan.part_cta_begin[p + 1] = plan.part_cta_begin[p] + (p < L.n_parts ? alloc[p] : 0);
  plan.ctas_per_slice = plan.part_cta_begin[L.n_parts];
  plan.grid = plan.ctas_per_slice * L.slices;
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

template <typename T, int MB, int KROT>
static int launch_decode(const DecodeParams &p, const DecodePlan &plan, cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(p.nstages) * kStageBytes + 4 * p.rot_bytes + 3 * 8 * kMaxStages;
  auto kern = decode_kernel<T, MB, KROT>;
  PARO_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(plan.grid);
  cfg.blockDim = dim3(kDecodeThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = env_int("PARO_NO_PDL", 0) ? 0 : 1;
  PARO_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p));
  note_launches(1);
  return PARO_OK;
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
  const int ctas_per_sm = env_int("PARO_DECODE_CTAS_PER_SM", 2);
  DecodePlan plan;
  if (L.slices > sms * ctas_per_sm || !make_plan(L, sms * ctas_per_sm, plan)) {
    set_error("decode: in_features=%d needs %d K-slices, more than the %d resident CTAs", L.K, L.slices, sms * ctas_per_sm);
    return PARO_EUNSUPPORTED;
  }
  DecodeParams p;
  p.packed = static_cast<const uint8_t *>(packed);
  p.x = x; p.y = y; p.bias = bias;
  const size_t counters = (static_cast<size_t>(L.tiles_total) * 4 + 255) / 256 * 256;
  p.counters = static_cast<int *>(workspace);
  p.partials = reinterpret_cast<float *>(static_cast<uint8_t *>(workspace) + counters);
  p.M = static_cast<int>(M); p.K = L.K; p.N = L.N;
  p.n_parts = L.n_parts; p.slices = L.slices; p.groups = L.groups; p.krot = L.krot; p.tiles_total = L.tiles_total;
  int nst = env_int("PARO_DECODE_STAGES", MB == 1 ? 4 : 4);
  if (nst < 1) nst = 1;
  if (nst > kMaxStages) nst = kMaxStages;
  p.nstages = nst;
  p.rot_bytes = kGroup * 2 * (M <= 2 ? 2 : M <= 4 ? 4 : M <= 8 ? 8 : 16);
  for (int i = 0; i <= PARO_MAX_PARTS; ++i) {
    p.part_tile_begin[i] = L.part_tile_begin[i];
    p.part_cta_begin[i] = plan.part_cta_begin[i];
  }
  p.meta_group_bytes = L.meta_group_bytes;
  p.meta_off = static_cast<long long>(L.meta_off);
  p.rec_off = static_cast<long long>(L.rec_off);

  const bool k8 = L.krot == 8;
  if (s.dtype == PARO_BF16) {
    if (MB == 1) return k8 ? launch_decode<__nv_bfloat16, 1, 8>(p, plan, stream) : launch_decode<__nv_bfloat16, 1, 0>(p, plan, stream);
    return k8 ? launch_decode<__nv_bfloat16, 2, 8>(p, plan, stream) : launch_decode<__nv_bfloat16, 2, 0>(p, plan, stream);
  }
  if (MB == 1) return k8 ? launch_decode<__half, 1, 8>(p, plan, stream) : launch_decode<__half, 1, 0>(p, plan, stream);
  return k8 ? launch_decode<__half, 2, 8>(p, plan, stream) : launch_decode<__half, 2, 0>(p, plan, stream);
}

}  // namespace paro
