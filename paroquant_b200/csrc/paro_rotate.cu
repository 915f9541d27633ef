// Standalone scaled pairwise (Givens) rotation for sm_100a -- the torch.ops.rotation.rotate
// surface (/root/reference/paroquant/kernels/cuda/rotation.cu:10-43,62-135) and the
// activation pre-pass of the large-M GEMM path.
//
// Arithmetic is the reference's, rounding point for rounding point (rotation.cuh:91-173 for
// fp16/bf16, :16-75 for fp32): v = T(x * T(scale)); per rotation r and pair (i, j):
// (s, c) = __sincosf(float(T(theta))), vi' = T(fma(c, vi, s*vj)), vj' = T(fma(c, vj, s*-vi)).
//
// Mapping (not the reference's): one WARP owns one (block of RB rows, group) tile, so the
// krot rotations need only __syncwarp().  The tile lives in shared memory channel-major --
// rot[channel][RB rows] = one 16-byte vector per channel -- so a pair update is two 128-bit
// loads and two 128-bit stores for all RB rows at once.  The 16-byte slot of channel c is
// c ^ ((c >> 3) & 7), which makes the coalesced transpose-in / transpose-out conflict free.
#include "paro_common.cuh"

namespace paro {

template <typename T> struct RotVec;  // 16-byte vector of RB rows of one channel

template <> struct RotVec<float> {
  static constexpr int RB = 4;
};
template <> struct RotVec<__half> {
  static constexpr int RB = 8;
};
template <> struct RotVec<__nv_bfloat16> {
  static constexpr int RB = 8;
};

__device__ __forceinline__ int rot_slot(int c) { return c ^ ((c >> 3) & 7); }

// ---- half / bf16: a channel vector is 4 x T2 words, word u = rows (2u, 2u+1)
template <typename T>
__device__ __forceinline__ void pair_update(uint4 &vi, uint4 &vj, float c, float s) {
  uint32_t *pi = reinterpret_cast<uint32_t *>(&vi), *pj = reinterpret_cast<uint32_t *>(&vj);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const float2 a = Traits<T>::to_float2(unpack2<T>(pi[u]));
    const float2 b = Traits<T>::to_float2(unpack2<T>(pj[u]));
    float yix, yiy, yjx, yjy;
    givens(c, s, a.x, b.x, yix, yjx);
    givens(c, s, a.y, b.y, yiy, yjy);
    pi[u] = pack2<T>(Traits<T>::from_floats(yix, yiy));
    pj[u] = pack2<T>(Traits<T>::from_floats(yjx, yjy));
  }
}
template <>
__device__ __forceinline__ void pair_update<float>(uint4 &vi, uint4 &vj, float c, float s) {
  float *pi = reinterpret_cast<float *>(&vi), *pj = reinterpret_cast<float *>(&vj);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    float yi, yj;
    givens(c, s, pi[u], pj[u], yi, yj);
    pi[u] = yi;
    pj[u] = yj;
  }
}

template <typename T, int G>
__global__ void __launch_bounds__(128) rotate_kernel(const T *__restrict__ x, T *__restrict__ out,
                                                     const int16_t *__restrict__ idx, const void *__restrict__ theta,
                                                     int theta_dtype, const void *__restrict__ scales, int scales_dtype,
                                                     int64_t M, int K, int krot, int64_t M_store, int tiled_nt) {
  constexpr int RB = RotVec<T>::RB;
  constexpr int CPL = G / 32;  // channels per lane for the coalesced load / store
  constexpr int PPL = G / 64;  // pairs per lane
  __shared__ __align__(16) uint4 tile[4][G];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int groups = K / G;
  const int64_t task = static_cast<int64_t>(blockIdx.x) * 4 + warp;
  const int64_t row_blocks = (M_store + RB - 1) / RB;  // rows in [M, M_store) are zero padding (tiled output only)
  if (task >= row_blocks * groups) return;
  const int g = static_cast<int>(task % groups);
  const int64_t row0 = (task / groups) * RB;
  uint4 *rot = tile[warp];

  // ---- load RB rows x CPL channels per lane, scale, transpose into rot[channel]
  {
    float sc[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c)
      sc[c] = scales ? load_param_as<T>(scales, static_cast<int64_t>(g) * G + lane * CPL + c, scales_dtype) : 1.0f;
    if constexpr (sizeof(T) == 4) {
      float v[RB][CPL];
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const bool ok = row0 + r < M;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          const float xv = ok ? x[(row0 + r) * K + g * G + lane * CPL + c] : 0.0f;
          v[r][c] = scales ? xv * sc[c] : xv;  // rotation.cuh:24-32
        }
      }
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        uint4 w;
        float *pw = reinterpret_cast<float *>(&w);
#pragma unroll
        for (int r = 0; r < RB; ++r) pw[r] = v[r][c];
        rot[rot_slot(lane * CPL + c)] = w;
      }
    } else {
      T v[RB][CPL];
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const bool ok = row0 + r < M;
        if constexpr (CPL == 4) {
          uint2 raw = ok ? *reinterpret_cast<const uint2 *>(x + (row0 + r) * K + g * G + lane * 4) : make_uint2(0, 0);
          const T *pr = reinterpret_cast<const T *>(&raw);
#pragma unroll
          for (int c = 0; c < 4; ++c) v[r][c] = pr[c];
        } else {
          uint32_t raw = ok ? *reinterpret_cast<const uint32_t *>(x + (row0 + r) * K + g * G + lane * 2) : 0u;
          const T *pr = reinterpret_cast<const T *>(&raw);
          v[r][0] = pr[0];
          v[r][1] = pr[1];
        }
        if (scales) {
#pragma unroll
          for (int c = 0; c < CPL; ++c) v[r][c] = __hmul(v[r][c], Traits<T>::from_float(sc[c]));  // one rounding, cuh:112-113
        }
      }
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        uint4 w;
        T *pw = reinterpret_cast<T *>(&w);
#pragma unroll
        for (int r = 0; r < RB; ++r) pw[r] = v[r][c];
        rot[rot_slot(lane * CPL + c)] = w;
      }
    }
  }
  __syncwarp();

  // ---- krot rotations; lane owns pairs lane*PPL .. lane*PPL+PPL-1 of each rotation
  for (int r = 0; r < krot; ++r) {
    int pi[PPL], pj[PPL];
    float cs[PPL], sn[PPL];
#pragma unroll
    for (int q = 0; q < PPL; ++q) {
      const int t = lane * PPL + q;
      const int ij = *reinterpret_cast<const int *>(idx + static_cast<int64_t>(r) * K + g * G + 2 * t);
      pi[q] = ij & 0xFFFF;
      pj[q] = (ij >> 16) & 0xFFFF;
      const float th = load_param_as<T>(theta, static_cast<int64_t>(r) * (K / 2) + g * (G / 2) + t, theta_dtype);
      __sincosf(th, &sn[q], &cs[q]);
    }
#pragma unroll
    for (int q = 0; q < PPL; ++q) {
      uint4 vi = rot[rot_slot(pi[q])], vj = rot[rot_slot(pj[q])];
      pair_update<T>(vi, vj, cs[q], sn[q]);
      rot[rot_slot(pi[q])] = vi;
      rot[rot_slot(pj[q])] = vj;
    }
    __syncwarp();
  }

  // ---- transpose out
  if constexpr (sizeof(T) == 4) {
    float v[RB][CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const uint4 w = rot[rot_slot(lane * CPL + c)];
      const float *pw = reinterpret_cast<const float *>(&w);
#pragma unroll
      for (int r = 0; r < RB; ++r) v[r][c] = pw[r];
    }
#pragma unroll
    for (int r = 0; r < RB; ++r)
      if (row0 + r < M) {
#pragma unroll
        for (int c = 0; c < CPL; ++c) out[(row0 + r) * K + g * G + lane * CPL + c] = v[r][c];
      }
  } else {
    T v[RB][CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const uint4 w = rot[rot_slot(lane * CPL + c)];
      const T *pw = reinterpret_cast<const T *>(&w);
#pragma unroll
      for (int r = 0; r < RB; ++r) v[r][c] = pw[r];
    }
#pragma unroll
    for (int r = 0; r < RB; ++r)
      if (row0 + r < M_store) {
        if constexpr (CPL == 4) {
          uint2 raw;
          T *pr = reinterpret_cast<T *>(&raw);
#pragma unroll
          for (int c = 0; c < 4; ++c) pr[c] = v[r][c];
          if (tiled_nt) {
            // B-operand order of the tcgen05 GEMM (paro_gemm.cu): [token block][k16 step][k half][n/8][n%8][8 elems]
            const int64_t m = row0 + r;
            const int64_t tb = m / tiled_nt;
            const int n = static_cast<int>(m - tb * tiled_nt), k = g * G + lane * 4;
            const int64_t off = (tb * (K / 16) + (k >> 4)) * (static_cast<int64_t>(tiled_nt) * 16) + ((k >> 3) & 1) * (tiled_nt * 8) +
                                (n >> 3) * 64 + (n & 7) * 8 + (k & 7);
            *reinterpret_cast<uint2 *>(out + off) = raw;
          } else {
            *reinterpret_cast<uint2 *>(out + (row0 + r) * K + g * G + lane * 4) = raw;
          }
        } else {
          uint32_t raw;
          T *pr = reinterpret_cast<T *>(&raw);
          pr[0] = v[r][0];
          pr[1] = v[r][1];
          *reinterpret_cast<uint32_t *>(out + (row0 + r) * K + g * G + lane * 2) = raw;
        }
      }
  }
}


// ------------------------------------------------------------------ pre-pass of the tcgen05 kernels (paro_decode.cu from 5 rows, paro_gemm.cu)
// ONE launch for all partitions of a merged linear: warp = (partition, group, block of 8 rows); output in the B-operand order
// of the consumer (tiles of nt tokens, rows >= M zero).  Same arithmetic as rotate_kernel; what differs is latency: the pair
// indices and angles of all (<= 8 at a time) rotations are fetched up front, so the launch costs one global-load round trip
// instead of one per stage (rotate_kernel at 16 rows: ~7 us, almost all of it eight dependent L2 / DRAM latencies).
template <typename T>
__global__ void __launch_bounds__(128) rotate_small_kernel(const T *__restrict__ x, T *__restrict__ out, const uint8_t *__restrict__ raw_base,
                                                           long long raw_part_bytes, long long out_part_elems, int M, int M_store, int nt, int K,
                                                           int krot, int n_parts) {
  constexpr int G = 128, RB = 8;
  __shared__ __align__(16) uint4 tile[4][G];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int groups = K / G;
  const int row_blocks = M_store / RB;
  const long long task = static_cast<long long>(blockIdx.x) * 4 + warp;          // (part, row block, group)
  if (task >= static_cast<long long>(n_parts) * row_blocks * groups) return;
  const int g = static_cast<int>(task % groups), rb = static_cast<int>((task / groups) % row_blocks), part = static_cast<int>(task / (static_cast<long long>(row_blocks) * groups));
  const uint8_t *raw = raw_base + part * raw_part_bytes;
  const int16_t *idx = reinterpret_cast<const int16_t *>(raw);
  const T *theta = reinterpret_cast<const T *>(raw + static_cast<size_t>(krot) * K * 2);
  const T *scales = reinterpret_cast<const T *>(raw + static_cast<size_t>(krot) * K * 3);
  T *o = out + part * out_part_elems;
  uint4 *rot = tile[warp];
  const int row0 = rb * RB;
  // B-operand order of the tcgen05 kernels: [token block of nt][k16 step][k half][row / 8][row % 8][8 elements]
  auto tiled = [&](int m, int k) -> int64_t {
    const int tb = m / nt, n = m - tb * nt;
    return (static_cast<int64_t>(tb) * (K / 16) + (k >> 4)) * (static_cast<int64_t>(nt) * 16) + ((k >> 3) & 1) * (nt * 8) + (n >> 3) * 64 + (n & 7) * 8 +
           (k & 7);
  };
  if (row0 >= M) {   // a block of padding rows: zeros, in place
#pragma unroll
    for (int r = 0; r < RB; ++r) *reinterpret_cast<uint2 *>(o + tiled(row0 + r, g * G + lane * 4)) = make_uint2(0u, 0u);
    return;
  }
  // ---- everything this warp will need from global memory, issued together
  uint2 xr[RB];
#pragma unroll
  for (int r = 0; r < RB; ++r)
    xr[r] = row0 + r < M ? *reinterpret_cast<const uint2 *>(x + static_cast<int64_t>(row0 + r) * K + g * G + lane * 4) : make_uint2(0u, 0u);
  const uint2 scw = *reinterpret_cast<const uint2 *>(scales + g * G + lane * 4);
  for (int r0 = 0; r0 < krot; r0 += 8) {
    int ij[8][2];
    uint32_t th[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (r0 + r < krot) {
        const uint2 w = *reinterpret_cast<const uint2 *>(idx + static_cast<int64_t>(r0 + r) * K + g * G + 4 * lane);   // pairs 2 lane, 2 lane + 1
        ij[r][0] = static_cast<int>(w.x);
        ij[r][1] = static_cast<int>(w.y);
        th[r] = *reinterpret_cast<const uint32_t *>(theta + static_cast<int64_t>(r0 + r) * (K / 2) + g * (G / 2) + 2 * lane);
      }
    }
    if (r0 == 0) {
      // scale (one rounding, rotation.cuh:112-113), transpose into rot[channel]
      const T *ps = reinterpret_cast<const T *>(&scw);
      T v[RB][4];
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const T *pr = reinterpret_cast<const T *>(&xr[r]);
#pragma unroll
        for (int c = 0; c < 4; ++c) v[r][c] = __hmul(pr[c], ps[c]);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 w;
        T *pw = reinterpret_cast<T *>(&w);
#pragma unroll
        for (int r = 0; r < RB; ++r) pw[r] = v[r][c];
        rot[rot_slot(lane * 4 + c)] = w;
      }
      __syncwarp();
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (r0 + r < krot) {
        const float2 t2 = Traits<T>::to_float2(unpack2<T>(th[r]));
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float sn, cs;
          __sincosf(q ? t2.y : t2.x, &sn, &cs);
          const int pi = ij[r][q] & 0xFFFF, pj = (ij[r][q] >> 16) & 0xFFFF;
          uint4 vi = rot[rot_slot(pi)], vj = rot[rot_slot(pj)];
          pair_update<T>(vi, vj, cs, sn);
          rot[rot_slot(pi)] = vi;
          rot[rot_slot(pj)] = vj;
        }
        __syncwarp();
      }
    }
  }
  // ---- transpose out
  T v[RB][4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint4 w = rot[rot_slot(lane * 4 + c)];
    const T *pw = reinterpret_cast<const T *>(&w);
#pragma unroll
    for (int r = 0; r < RB; ++r) v[r][c] = pw[r];
  }
#pragma unroll
  for (int r = 0; r < RB; ++r) {
    const int n = row0 + r, k = g * G + lane * 4;
    uint2 w = make_uint2(0u, 0u);
    if (n < M) {
      T *pw = reinterpret_cast<T *>(&w);
#pragma unroll
      for (int c = 0; c < 4; ++c) pw[c] = v[r][c];
    }
    *reinterpret_cast<uint2 *>(o + tiled(n, k)) = w;
  }
}

// all partitions in one launch; metadata in the layout's reference-format region (theta and channel scales already of dtype T);
// output per partition: M_store rows (a multiple of 8; rows >= M zero) in tiles of `nt` tokens
int rotate_small_launch(const void *x, void *out, const void *raw_base, long long raw_part_bytes, int n_parts, int64_t M, int64_t M_store, int nt,
                        int K, int krot, int dtype, cudaStream_t stream) {
  const long long tasks = static_cast<long long>(n_parts) * (M_store / 8) * (K / 128);
  const long long blocks = (tasks + 3) / 4;
  if (blocks > 0x7FFFFFFF) { set_error("rotate: too many rows"); return PARO_EINVAL; }
  const long long part_elems = static_cast<long long>(M_store) * K;
  if (dtype == PARO_F16)
    rotate_small_kernel<__half><<<static_cast<unsigned>(blocks), 128, 0, stream>>>(static_cast<const __half *>(x), static_cast<__half *>(out),
                                                                                  static_cast<const uint8_t *>(raw_base), raw_part_bytes, part_elems,
                                                                                  static_cast<int>(M), static_cast<int>(M_store), nt, K, krot, n_parts);
  else
    rotate_small_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), 128, 0, stream>>>(
        static_cast<const __nv_bfloat16 *>(x), static_cast<__nv_bfloat16 *>(out), static_cast<const uint8_t *>(raw_base), raw_part_bytes, part_elems,
        static_cast<int>(M), static_cast<int>(M_store), nt, K, krot, n_parts);
  PARO_CUDA_OK(cudaGetLastError());
  note_launches(1);
  return PARO_OK;
}

template <typename T>
static int launch_T(const void *x, void *out, const int16_t *idx, const void *theta, int theta_dtype,
                    const void *scales, int scales_dtype, int64_t M, int K, int krot, int G, cudaStream_t stream,
                    int64_t M_store = -1, int tiled_nt = 0) {
  constexpr int RB = RotVec<T>::RB;
  if (M_store < 0) M_store = M;
  const int64_t tasks = ((M_store + RB - 1) / RB) * (K / G);
  const int64_t blocks = (tasks + 3) / 4;
  if (blocks > 0x7FFFFFFF) {
    set_error("rotate: too many rows");
    return PARO_EINVAL;
  }
  const T *xp = static_cast<const T *>(x);
  T *op = static_cast<T *>(out);
  if (G == 128)
    rotate_kernel<T, 128><<<static_cast<unsigned>(blocks), 128, 0, stream>>>(xp, op, idx, theta, theta_dtype, scales,
                                                                            scales_dtype, M, K, krot, M_store, tiled_nt);
  else
    rotate_kernel<T, 64><<<static_cast<unsigned>(blocks), 128, 0, stream>>>(xp, op, idx, theta, theta_dtype, scales,
                                                                           scales_dtype, M, K, krot, M_store, tiled_nt);
  PARO_CUDA_OK(cudaGetLastError());
  note_launches(1);
  return PARO_OK;
}

int rotate_launch(const void *x, void *out, const int16_t *idx, const void *theta, int theta_dtype, const void *scales,
                  int scales_dtype, int64_t M, int K, int krot, int G, int dtype, cudaStream_t stream) {
  if (M == 0) return PARO_OK;
  switch (dtype) {
    case PARO_F32: return launch_T<float>(x, out, idx, theta, theta_dtype, scales, scales_dtype, M, K, krot, G, stream);
    case PARO_F16: return launch_T<__half>(x, out, idx, theta, theta_dtype, scales, scales_dtype, M, K, krot, G, stream);
    case PARO_BF16:
      return launch_T<__nv_bfloat16>(x, out, idx, theta, theta_dtype, scales, scales_dtype, M, K, krot, G, stream);
  }
  set_error("rotate supports Float, Half, and BFloat16, got dtype code %d", dtype);
  return PARO_EINVAL;
}

// ------------------------------------------------------------------ backward of the rotate op (training side)
// The reference walks the krot rotations in Python, last to first: two one-rotation rotate launches (t and g with -theta), five
// gathers and a row reduction per rotation (/root/reference/paroquant/kernels/cuda/autograd.py:20-61).  Here ONE launch does the
// whole walk: a warp owns a group and a strided set of 4-row blocks; the stage output t (from y) and its gradient g sit in two
// fp32 shared-memory tiles; per rotation and pair
//     dL/dtheta += sum_rows (g_i t_j - g_j t_i)          (identity on the stage OUTPUT: dy_i/dtheta = y_j, dy_j/dtheta = -y_i)
//     (t, g) <- Givens(-theta) (t, g)                    (orthogonal: the same update back-propagates g), rounded to T per
//                                                         rotation exactly where the per-rotation launches rounded
// and after the last un-rotation grad_x = g * scale, dL/dscale += x * g.  dL/dtheta and dL/dscale are accumulated per warp over
// its row blocks and added to fp32 buffers with one atomic per (rotation, pair) / channel and warp.  A warp keeps ONE group for
// all its row blocks, so the group's pair slots and (cos, sin) are computed once into shared memory: the first version re-read
// indices and angles from L2 in every rotation of every row block and sat on those loads (13 long-scoreboard stalls per issue).
template <typename T, int G>
__global__ void __launch_bounds__(128) rotate_backward_kernel(const T *__restrict__ y, const T *__restrict__ gout, const T *__restrict__ x,
                                                              const int16_t *__restrict__ idx, const void *__restrict__ theta, int theta_dtype,
                                                              const void *__restrict__ scales, int scales_dtype, T *__restrict__ grad_x,
                                                              float *__restrict__ grad_theta, float *__restrict__ grad_scale, int64_t M, int K,
                                                              int krot, int splits, int ref_formula) {
  constexpr int RB = 4, CPL = G / 32, PPL = G / 64;
  __shared__ __align__(16) float4 tile_t[4][G], tile_g[4][G];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int groups = K / G;
  const int64_t wid = static_cast<int64_t>(blockIdx.x) * 4 + warp;   // (split, group)
  if (wid >= static_cast<int64_t>(splits) * groups) return;
  const int g = static_cast<int>(wid % groups), split = static_cast<int>(wid / groups);
  const int64_t row_blocks = (M + RB - 1) / RB;
  float4 *tt = tile_t[warp], *tg = tile_g[warp];
  // per warp, dynamic shared memory: [rotation][q][lane] {cos, sin} (float2), pair slots (i | j << 16), dL/dtheta partial -- each
  // lane touches only the entries of its own pairs (shared memory, not registers: the rotation loop stays rolled; unrolled over
  // 16 rotations the kernel needed 160-190 registers)
  extern __shared__ __align__(16) uint8_t bw_dyn[];
  constexpr int kPerRot = G / 2;   // pairs per rotation = PPL * 32
  float2 *mcs = reinterpret_cast<float2 *>(bw_dyn) + static_cast<size_t>(warp) * krot * kPerRot;
  uint32_t *mij = reinterpret_cast<uint32_t *>(bw_dyn + static_cast<size_t>(4) * krot * kPerRot * 8) + static_cast<size_t>(warp) * krot * kPerRot;
  float *acc = reinterpret_cast<float *>(bw_dyn + static_cast<size_t>(4) * krot * kPerRot * 12) + static_cast<size_t>(warp) * krot * kPerRot;
  // ref_formula: also sum_rows (g . t) per pair, for the value the reference's expression takes (autograd.py:50-52 applied after
  // g was un-rotated: cos * dL/dtheta - sin * sum(g . t)); the array exists only then
  float *acc2 = reinterpret_cast<float *>(bw_dyn + static_cast<size_t>(4) * krot * kPerRot * 16) + static_cast<size_t>(warp) * krot * kPerRot;
  for (int r = 0; r < krot; ++r) {
#pragma unroll
    for (int q = 0; q < PPL; ++q) {
      const int t = lane * PPL + q, e = r * kPerRot + q * 32 + lane;
      const int ij = *reinterpret_cast<const int *>(idx + static_cast<int64_t>(r) * K + g * G + 2 * t);
      const float th = load_param_as<T>(theta, static_cast<int64_t>(r) * (K / 2) + g * (G / 2) + t, theta_dtype);
      float sn, cs;
      __sincosf(-th, &sn, &cs);
      mcs[e] = make_float2(cs, sn);
      mij[e] = static_cast<uint32_t>(rot_slot(ij & 0xFFFF)) | (static_cast<uint32_t>(rot_slot((ij >> 16) & 0xFFFF)) << 16);
      acc[e] = 0.f;
      if (ref_formula) acc2[e] = 0.f;
    }
  }
  __syncwarp();
  float acc_s[CPL], sc[CPL];
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    acc_s[c] = 0.f;
    sc[c] = scales ? load_param_as<float>(scales, static_cast<int64_t>(g) * G + lane * CPL + c, scales_dtype) : 1.0f;   // autograd.py:55: the scale as stored
  }
  auto rnd = [](float v) { return Traits<T>::to_float(Traits<T>::from_float(v)); };   // the per-rotation launches store T
  for (int64_t rb = split; rb < row_blocks; rb += splits) {
    const int64_t row0 = rb * RB;
    // ---- stage output and its gradient, channel-major fp32 vectors of 4 rows
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      float vt[RB], vg[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const bool ok = row0 + r < M;
        const int64_t o = (row0 + r) * K + g * G + lane * CPL + c;
        vt[r] = ok ? Traits<T>::to_float(y[o]) : 0.f;
        vg[r] = ok ? Traits<T>::to_float(gout[o]) : 0.f;
      }
      tt[rot_slot(lane * CPL + c)] = make_float4(vt[0], vt[1], vt[2], vt[3]);
      tg[rot_slot(lane * CPL + c)] = make_float4(vg[0], vg[1], vg[2], vg[3]);
    }
    __syncwarp();
#pragma unroll 1
    for (int r = krot - 1; r >= 0; --r) {
      {
#pragma unroll
        for (int q = 0; q < PPL; ++q) {
          const int e = r * kPerRot + q * 32 + lane;
          const uint32_t ijw = mij[e];
          const int pi = ijw & 0xFFFF, pj = ijw >> 16;
          const float2 csn = mcs[e];
          const float cs = csn.x, sn = csn.y;
          float4 ti = tt[pi], tj = tt[pj], gi = tg[pi], gj = tg[pj];
          acc[e] += (gi.x * tj.x - gj.x * ti.x) + (gi.y * tj.y - gj.y * ti.y) + (gi.z * tj.z - gj.z * ti.z) + (gi.w * tj.w - gj.w * ti.w);
          if (ref_formula)   // g . t of the pair (rotation invariant, like the cross product above)
            acc2[e] += (gi.x * ti.x + gj.x * tj.x) + (gi.y * ti.y + gj.y * tj.y) + (gi.z * ti.z + gj.z * tj.z) + (gi.w * ti.w + gj.w * tj.w);
          float4 ni, nj, mi, mj;
          givens(cs, sn, ti.x, tj.x, ni.x, nj.x); givens(cs, sn, ti.y, tj.y, ni.y, nj.y);
          givens(cs, sn, ti.z, tj.z, ni.z, nj.z); givens(cs, sn, ti.w, tj.w, ni.w, nj.w);
          givens(cs, sn, gi.x, gj.x, mi.x, mj.x); givens(cs, sn, gi.y, gj.y, mi.y, mj.y);
          givens(cs, sn, gi.z, gj.z, mi.z, mj.z); givens(cs, sn, gi.w, gj.w, mi.w, mj.w);
          if constexpr (sizeof(T) == 2) {
            ni = make_float4(rnd(ni.x), rnd(ni.y), rnd(ni.z), rnd(ni.w)); nj = make_float4(rnd(nj.x), rnd(nj.y), rnd(nj.z), rnd(nj.w));
            mi = make_float4(rnd(mi.x), rnd(mi.y), rnd(mi.z), rnd(mi.w)); mj = make_float4(rnd(mj.x), rnd(mj.y), rnd(mj.z), rnd(mj.w));
          }
          tt[pi] = ni; tt[pj] = nj; tg[pi] = mi; tg[pj] = mj;
        }
        __syncwarp();
      }
    }
    // ---- g is now dL/d(x * scale)
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const float4 gv = tg[rot_slot(lane * CPL + c)];
      const float gr[RB] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        if (row0 + r < M) {
          const int64_t o = (row0 + r) * K + g * G + lane * CPL + c;
          grad_x[o] = Traits<T>::from_float(scales ? gr[r] * sc[c] : gr[r]);
          if (grad_scale) acc_s[c] += Traits<T>::to_float(x[o]) * gr[r];
        }
      }
    }
    __syncwarp();
  }
  for (int r = 0; r < krot; ++r) {
#pragma unroll
    for (int q = 0; q < PPL; ++q) {
      const int e = r * kPerRot + q * 32 + lane;
      float v = acc[e];
      if (ref_formula) v = mcs[e].x * v + mcs[e].y * acc2[e];   // {cos, sin}(-theta): cos(theta) * cross - sin(theta) * dot
      atomicAdd(grad_theta + static_cast<int64_t>(r) * (K / 2) + g * (G / 2) + lane * PPL + q, v);
    }
  }
  if (grad_scale) {
#pragma unroll
    for (int c = 0; c < CPL; ++c) atomicAdd(grad_scale + static_cast<int64_t>(g) * G + lane * CPL + c, acc_s[c]);
  }
}

template <typename T>
static int launch_backward_T(const void *y, const void *gout, const void *x, const int16_t *idx, const void *theta, int theta_dtype, const void *scales,
                             int scales_dtype, void *grad_x, float *grad_theta, float *grad_scale, int64_t M, int K, int krot, int G, int ref_formula,
                             cudaStream_t stream) {
  const int groups = K / G;
  const int64_t row_blocks = (M + 3) / 4;
  // per CTA: 4 warps x krot x G/2 pairs x ({cos, sin} 8 B + slots 4 B + partial 4 B) of dynamic shared memory + the two tiles
  const size_t dyn = static_cast<size_t>(4) * krot * (G / 2) * (ref_formula ? 20 : 16);   // (+ 4 B: the g . t sums of the reference's expression)
  const int resident = static_cast<int>((227 * 1024) / (dyn + static_cast<size_t>(8) * G * 16 + 1024));
  const int ctas_per_sm = resident < 1 ? 1 : resident > 6 ? 6 : resident;   // 6: the register limit
  // enough warps to fill the machine, no more splits than row blocks: every warp then walks ~row_blocks / splits blocks of its
  // group, computes the group's (cos, sin) once and pays its atomics once
  int64_t splits = (148 * ctas_per_sm * 4 + groups - 1) / groups;
  if (splits > row_blocks) splits = row_blocks;
  if (splits < 1) splits = 1;
  const int64_t blocks = (splits * groups + 3) / 4;
  if (blocks > 0x7FFFFFFF) { set_error("rotate_backward: too many groups"); return PARO_EINVAL; }
  const T *yp = static_cast<const T *>(y), *gp = static_cast<const T *>(gout), *xp = static_cast<const T *>(x);
  T *gx = static_cast<T *>(grad_x);
  if (G == 128) {
    PARO_CUDA_OK(cudaFuncSetAttribute(rotate_backward_kernel<T, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dyn)));
    rotate_backward_kernel<T, 128><<<static_cast<unsigned>(blocks), 128, dyn, stream>>>(yp, gp, xp, idx, theta, theta_dtype, scales, scales_dtype, gx, grad_theta,
                                                                                       grad_scale, M, K, krot, static_cast<int>(splits), ref_formula);
  } else {
    PARO_CUDA_OK(cudaFuncSetAttribute(rotate_backward_kernel<T, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dyn)));
    rotate_backward_kernel<T, 64><<<static_cast<unsigned>(blocks), 128, dyn, stream>>>(yp, gp, xp, idx, theta, theta_dtype, scales, scales_dtype, gx, grad_theta,
                                                                                      grad_scale, M, K, krot, static_cast<int>(splits), ref_formula);
  }
  PARO_CUDA_OK(cudaGetLastError());
  note_launches(1);
  return PARO_OK;
}

int rotate_backward_launch(const void *y, const void *gout, const void *x, const int16_t *idx, const void *theta, int theta_dtype, const void *scales,
                           int scales_dtype, void *grad_x, float *grad_theta, float *grad_scale, int64_t M, int K, int krot, int G, int dtype,
                           int ref_formula, cudaStream_t stream) {
  if (M == 0) return PARO_OK;
  switch (dtype) {
    case PARO_F32: return launch_backward_T<float>(y, gout, x, idx, theta, theta_dtype, scales, scales_dtype, grad_x, grad_theta, grad_scale, M, K, krot, G, ref_formula, stream);
    case PARO_F16: return launch_backward_T<__half>(y, gout, x, idx, theta, theta_dtype, scales, scales_dtype, grad_x, grad_theta, grad_scale, M, K, krot, G, ref_formula, stream);
    case PARO_BF16: return launch_backward_T<__nv_bfloat16>(y, gout, x, idx, theta, theta_dtype, scales, scales_dtype, grad_x, grad_theta, grad_scale, M, K, krot, G, ref_formula, stream);
  }
  set_error("rotate supports Float, Half, and BFloat16, got dtype code %d", dtype);
  return PARO_EINVAL;
}

// Rotation pre-pass of the large-M path: the raw metadata of ONE partition, output in the B-operand tile
// order of the tcgen05 GEMM, zero rows up to M_store (a multiple of tiled_nt).
int rotate_tiled_launch(const void *x, void *out, const int16_t *idx, const void *theta, int theta_dtype, const void *scales,
                        int scales_dtype, int64_t M, int64_t M_store, int tiled_nt, int K, int krot, int dtype, cudaStream_t stream) {
  if (dtype == PARO_F16)
    return launch_T<__half>(x, out, idx, theta, theta_dtype, scales, scales_dtype, M, K, krot, 128, stream, M_store, tiled_nt);
  return launch_T<__nv_bfloat16>(x, out, idx, theta, theta_dtype, scales, scales_dtype, M, K, krot, 128, stream, M_store, tiled_nt);
}

}  // namespace paro
