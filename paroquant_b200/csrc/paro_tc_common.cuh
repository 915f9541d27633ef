// Device helpers shared by the tensor-core kernels (paro_decode.cu, paro_gemm.cu):
// tcgen05 / TMEM wrappers, UMMA descriptors, cluster + DSMEM primitives, the INT4 -> T row dequant
// and the in-warp pairwise rotation (rounding points of /root/reference/paroquant/kernels/cuda/rotation.cuh:91-173).
#pragma once
#include "paro_common.cuh"
#include "paro_layout.h"

namespace paro {

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ------------------------------------------------------------------ tcgen05 wrappers
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem_d] (+)= A[tmem_a] * B[smem desc]; A: 128 lanes x 8 columns (16 x 16-bit k), kind::f16
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tc_ld2(uint32_t taddr, uint32_t &a, uint32_t &b) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld1(uint32_t taddr, uint32_t &a) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(a) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld4(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// K-major, no swizzle: 8-row x 16-byte core matrices; `lbo` = byte distance between the two k-halves
// of a k16 step, `sbo` = byte distance between 8-row groups (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t smem_desc_kmajor(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return static_cast<uint64_t>((addr >> 4) & 0x3FFF) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16) |
         (static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// cute::UMMA::InstrDescriptor: c_format F32 (bit 4), a/b format (bits 7, 10: 0 = F16, 1 = BF16), K-major A and B,
// N >> 3 at bit 17, M >> 4 at bit 24
template <typename T> __device__ __forceinline__ uint32_t instr_desc(int n) {
  const uint32_t fmt = Traits<T>::code == PARO_BF16 ? 1u : 0u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (8u << 24);
}

// one lane of a converged warp (CUTLASS' elect_one_sync): ptxas treats the guarded region as single-threaded
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ------------------------------------------------------------------ cluster / DSMEM helpers
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <typename T> __device__ __forceinline__ T bits_to_T(uint32_t b) {
  const uint16_t h = static_cast<uint16_t>(b);
  return *reinterpret_cast<const T *>(&h);
}
template <typename T> __device__ __forceinline__ uint16_t T_to_bits(T v) { return *reinterpret_cast<const uint16_t *>(&v); }

// ------------------------------------------------------------------ INT4 -> T dequant of one row
// (a & mask) | magic in ONE LOP3: both constants must sit in registers (LOP3 takes one immediate)
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t mask, uint32_t magic) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(mask), "r"(magic));
  return d;
}

template <typename T> struct RowDequant;

template <> struct RowDequant<__nv_bfloat16> {
  uint32_t s2, z2;
  __device__ __forceinline__ void prep(uint32_t s_bits, uint32_t z) {
    s2 = s_bits * 0x00010001u;
    z2 = (0x4300u | z) * 0x00010001u;  // bf16x2 {128 + z, 128 + z}
  }
  __device__ __forceinline__ uint32_t one(uint32_t w) const {
    // 0x4300 | q is the bf16 128 + q; (128 + q) - (128 + z) is exact; one rounding in the multiply
    const __nv_bfloat162 d = __hsub2(unpack2<__nv_bfloat16>(and_or(w, 0x000F000Fu, 0x43004300u)), unpack2<__nv_bfloat16>(z2));
    return pack2<__nv_bfloat16>(__hmul2(d, unpack2<__nv_bfloat16>(s2)));
  }
  __device__ __forceinline__ void word(uint32_t w, uint32_t *r) const {
    r[0] = one(w);
    r[1] = one(w >> 4);
    r[2] = one(w >> 8);
    r[3] = one(w >> 12);
  }
};

template <> struct RowDequant<__half> {
  uint32_t s2, z_lo, z_hi16;
  __device__ __forceinline__ void prep(uint32_t s_bits, uint32_t z) {
    s2 = s_bits * 0x00010001u;
    z_lo = (0x6400u | z) * 0x00010001u;           // {1024 + z}
    z_hi16 = (0xD400u | (z << 4)) * 0x00010001u;  // {-(64 + z)}
  }
  __device__ __forceinline__ void word(uint32_t w, uint32_t *r) const {
    const uint32_t w8 = w >> 8;
    const __half2 sixteenth = unpack2<__half>(0x2C002C00u), s = unpack2<__half>(s2);
    // low nibble: 0x6400 | q = 1024 + q.  High nibble in place: 0x6400 | (q << 4) = 1024 + 16 q,
    // and fma(1024 + 16 q, 1/16, -(64 + z)) = q - z exactly.
    r[0] = pack2<__half>(__hmul2(__hsub2(unpack2<__half>(and_or(w, 0x000F000Fu, 0x64006400u)), unpack2<__half>(z_lo)), s));
    r[1] = pack2<__half>(__hmul2(__hfma2(unpack2<__half>(and_or(w, 0x00F000F0u, 0x64006400u)), sixteenth, unpack2<__half>(z_hi16)), s));
    r[2] = pack2<__half>(__hmul2(__hsub2(unpack2<__half>(and_or(w8, 0x000F000Fu, 0x64006400u)), unpack2<__half>(z_lo)), s));
    r[3] = pack2<__half>(__hmul2(__hfma2(unpack2<__half>(and_or(w8, 0x00F000F0u, 0x64006400u)), sixteenth, unpack2<__half>(z_hi16)), s));
  }
};

// ------------------------------------------------------------------ in-warp rotation of one group
// Tile `rot` = this warp's [128 channels][ROWS] elements of T, channel-major.  Lane owns pairs
// 2*lane and 2*lane+1 of every rotation; idxw = bytes (i0, j0, i1, j1).
template <typename T, int ROWS>
__device__ __forceinline__ void rotate_stage(uint32_t rot, uint32_t idxw, float c0, float s0, float c1, float s1) {
  const uint32_t i0 = idxw & 0xFFu, j0 = (idxw >> 8) & 0xFFu, i1 = (idxw >> 16) & 0xFFu, j1 = idxw >> 24;
  if constexpr (ROWS == 1) {
    // all four loads first (the two pairs are disjoint), then the math, then the stores
    const uint32_t a0 = rot + i0 * 2, b0 = rot + j0 * 2, a1 = rot + i1 * 2, b1 = rot + j1 * 2;
    const float xa0 = Traits<T>::to_float(bits_to_T<T>(lds16(a0))), xb0 = Traits<T>::to_float(bits_to_T<T>(lds16(b0)));
    const float xa1 = Traits<T>::to_float(bits_to_T<T>(lds16(a1))), xb1 = Traits<T>::to_float(bits_to_T<T>(lds16(b1)));
    float yi0, yj0, yi1, yj1;
    givens(c0, s0, xa0, xb0, yi0, yj0);
    givens(c1, s1, xa1, xb1, yi1, yj1);
    sts16(a0, T_to_bits<T>(Traits<T>::from_float(yi0)));
    sts16(b0, T_to_bits<T>(Traits<T>::from_float(yj0)));
    sts16(a1, T_to_bits<T>(Traits<T>::from_float(yi1)));
    sts16(b1, T_to_bits<T>(Traits<T>::from_float(yj1)));
  } else {
    constexpr int MPW = ROWS / 2;
    const uint32_t ad[4] = {rot + i0 * (MPW * 4), rot + j0 * (MPW * 4), rot + i1 * (MPW * 4), rot + j1 * (MPW * 4)};
    uint32_t v[4][MPW];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int u = 0; u < MPW; ++u) v[k][u] = lds32(ad[k] + 4 * u);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const float c = q ? c1 : c0, s = q ? s1 : s0;
#pragma unroll
      for (int u = 0; u < MPW; ++u) {
        const float2 a = Traits<T>::to_float2(unpack2<T>(v[2 * q][u]));
        const float2 b = Traits<T>::to_float2(unpack2<T>(v[2 * q + 1][u]));
        float yix, yiy, yjx, yjy;
        givens(c, s, a.x, b.x, yix, yjx);
        givens(c, s, a.y, b.y, yiy, yjy);
        sts32(ad[2 * q] + 4 * u, pack2<T>(Traits<T>::from_floats(yix, yiy)));
        sts32(ad[2 * q + 1] + 4 * u, pack2<T>(Traits<T>::from_floats(yjx, yjy)));
      }
    }
  }
}

// this lane's 4 channels of the group for all rows: raw x (no dependence on the rotation metadata)
template <typename T, int ROWS, typename P>
__device__ __forceinline__ void load_x(const P &p, int gk, int lane, uint2 (&raw)[ROWS], int m0 = 0) {
  const T *xg = static_cast<const T *>(p.x) + gk * kGroup + 4 * lane;   // (p.x already points at the CTA's partition when pre-rotated)
#pragma unroll
  for (int m = 0; m < ROWS; ++m) {
    raw[m] = make_uint2(0u, 0u);
    if (m0 + m < p.M) raw[m] = __ldcg(reinterpret_cast<const uint2 *>(xg + static_cast<int64_t>(m0 + m) * p.K));
  }
}

// multiply by the channel scales in T (one rounding, rotation.cuh:112-113), lay out channel-major in `rot`
template <typename T, int ROWS>
__device__ __forceinline__ void scale_and_stage(uint32_t rot, int lane, const uint2 (&raw)[ROWS], uint2 csw) {
  using T2 = typename Traits<T>::T2;
  const T2 sc01 = unpack2<T>(csw.x), sc23 = unpack2<T>(csw.y);
  if constexpr (ROWS == 1) {
    const uint32_t v01 = pack2<T>(__hmul2(unpack2<T>(raw[0].x), sc01));
    const uint32_t v23 = pack2<T>(__hmul2(unpack2<T>(raw[0].y), sc23));
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(rot + 8 * lane), "r"(v01), "r"(v23) : "memory");
  } else {
    constexpr int MPW = ROWS / 2;
    uint32_t v01[ROWS], v23[ROWS];
#pragma unroll
    for (int m = 0; m < ROWS; ++m) {
      v01[m] = pack2<T>(__hmul2(unpack2<T>(raw[m].x), sc01));
      v23[m] = pack2<T>(__hmul2(unpack2<T>(raw[m].y), sc23));
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

template <typename T>
__device__ __forceinline__ void sincos2(uint32_t theta_pair_bits, float &c0, float &s0, float &c1, float &s1) {
  const float2 th = Traits<T>::to_float2(unpack2<T>(theta_pair_bits));
  __sincosf(th.x, &s0, &c0);
  __sincosf(th.y, &s1, &c1);
}

}  // namespace paro
