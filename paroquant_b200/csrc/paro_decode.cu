// Fused small-M path (M <= 16): scaled pairwise rotation of x + INT4 group dequant + GEMV/GEMM in ONE
// launch per (merged) linear, ONE persistent CTA per SM.  Replaces the reference's
// rotate -> Marlin kernel pairs (/root/reference/paroquant/inference/backends/vllm/plugin.py:281-311).
//
// Formulation (operand-swapped, as the large-M kernel): D[n, m] += W[n, k] * x_rot[m, k]
//   A = 128 output columns x 128 channels (one quantisation group) per ROUND, dequantised by CUDA cores
//       straight into TENSOR MEMORY (tcgen05.st); thread = one output column = one TMEM lane.
//   B = x_rot of the CTA's K-slice, 16 token rows (zero-padded), written once per CTA after the in-kernel
//       rotation, UMMA K-major core-matrix order.
//   D = fp32 [128 x 16] in TMEM, double buffered.  All rounds of a 128-column block accumulate into the
//       SAME D whichever dequant set produced the A operand: the tensor core does the in-CTA part of the
//       split-K reduction, and the result is read back ONCE per block (tcgen05.ld, one value per thread
//       and token) -- no shared-memory exchange, no intra-CTA barrier in the main loop.
//
// Work split: the `c` CTAs of a cluster cover K (ragged slices of whole groups); a cluster walks a range
// of 128-column blocks of one partition.  Cross-slice partials are pushed through distributed shared
// memory to rank (block mod c); one cluster barrier; fixed summation order (bit-reproducible).
//
// Roles (SETS x 4 dequant warps + 2): warps 0..4*SETS-1 workers (TMEM lane quarter = warp % 4): the
// prologue rotates the slice's groups (tasks of one group x up to 4 token rows per warp, __syncwarp only);
// the main loop takes rounds r = set, set + SETS, ...; the set that dequantised a block's last group reads
// D back.  Warp 4*SETS:
// TMA producer (ONE 8576-byte record = weights + scales + zeros of a round per cp.async.bulk, into an
// mbarrier ring, issued before griddepcontrol.wait).  Warp 4*SETS+1: TMEM allocator + tcgen05.mma issuer
// (one elected lane).
//
// Numerics: x_rot as paro_rotate.cu; W = T((q - z) * T(s)) with ONE rounding (the operand Marlin / AWQ
// form); fp32 accumulation in TMEM; one rounding to T; bias added in T (plugin.py:309-310).
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>

#include "paro_tc_common.cuh"

namespace paro {

#ifndef PARO_DEC_UNROLL
#define PARO_DEC_UNROLL 4
#endif
constexpr int kDecUnroll = PARO_DEC_UNROLL;   // chunks of the dequant loop unrolled per trip (4 = straight-line)

constexpr int kDecMaxStages = 24;
constexpr int kDecTmemCols = 512;
constexpr int kDecN = 16;            // MMA N: token rows, zero-padded (M_mma = 128 needs N % 16 == 0)
constexpr int kDecSmemLimit = 227 * 1024;
// one record per ring stage: 8576 bytes (group_size 128) or 8960 (group_size 64), both multiples of 64 -> DecParams::rec_bytes

constexpr int kDecTraceSlots = 12;
constexpr int kDecTraceCtas = 200;       // rows 200.. of the trace double as a per-round log of CTA 0 (issued / full / consumed)
constexpr int kDecTraceRows = 256;
__device__ unsigned long long g_dec_trace[kDecTraceRows * kDecTraceSlots];
#define DEC_TRACE(slot)                                                                                      \
  do {                                                                                                       \
    if (p.trace && warp == 0 && lane == 0 && blockIdx.x < kDecTraceCtas)                                      \
      g_dec_trace[blockIdx.x * kDecTraceSlots + (slot)] = static_cast<unsigned long long>(clock64() - t_entry); \
  } while (0)

struct DecParams {
  const uint8_t *packed;
  const void *x;
  void *y;
  const void *bias;
  int M, K, N;
  int n_parts, groups, krot;
  int c, c_shift;                 // cluster size (K slices of this launch), log2
  int nstages, rot_bytes, rot_warps, trace;   // rot_warps: worker warps that take rotation tasks (each owns one rot_bytes tile)
  int rec_bytes, q2;              // record (= ring stage) size; q2: group_size 64, two scale / zero sets per record (paro_layout.h)
  int pre_rotated;                // x is the pre-pass's output: [n_parts][M][K], already scaled and rotated (no stages here)
  long long x_part_stride;        // elements between partitions of a pre-rotated x (0 otherwise)
  int xb_off, rot_off, recv_off, bar_off;   // shared-memory carve-up (bytes)
  int part_col_begin[PARO_MAX_PARTS + 1];
  int part_block_begin[PARO_MAX_PARTS + 1];
  int part_range_begin[PARO_MAX_PARTS + 1];
  int meta_group_bytes;
  long long meta_off, rec_off;
};

// ---- prologue.  A rotation TASK = (group gi of the CTA's slice, block of up to ROWS <= 4 token rows); worker warp w takes
// tasks w, w + W, ...  (M = 16: four tasks per group, so all warps share the rotation instead of 16-row tiles on a few).
// B[gi][k16 step s][k half h][row m][8 k]: 16 rows x 16 bytes per core-matrix pair, 512 bytes per step.
template <typename T, int ROWS>
__device__ __forceinline__ void dec_write_b_rows(uint32_t xb_group, uint32_t rot, int m0, int nrows, int lane) {
  for (int idx = lane; idx < 16 * nrows; idx += 32) {
    const int ml = idx >> 4, s = (idx >> 1) & 7, h = idx & 1;
    const int c0 = 16 * s + 8 * h, m = m0 + ml;
    uint4 v;
    if constexpr (ROWS == 1) {
      v = lds128(rot + c0 * 2);
    } else {
      uint32_t e[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) e[k] = lds16(rot + (c0 + k) * (ROWS * 2) + 2 * ml);
      v = make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
    }
    sts128u(xb_group + s * (kDecN * 32) + h * (kDecN * 16) + (m >> 3) * 128 + (m & 7) * 16, v);
  }
}
struct DecRotMeta {
  uint32_t idxw[8], tw[8];
  uint2 csw;
  const uint8_t *meta;
};

__device__ __forceinline__ void dec_fetch_meta(const DecParams &p, int part, int gk, int lane, DecRotMeta &rm) {
  rm.meta = p.packed + p.meta_off + (static_cast<size_t>(part) * p.groups + gk) * p.meta_group_bytes;
  if (p.pre_rotated) return;   // no rotation tasks at all
  rm.csw = *reinterpret_cast<const uint2 *>(rm.meta + p.krot * 256 + 8 * lane);
  if (p.krot == 8) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      rm.idxw[r] = *reinterpret_cast<const uint32_t *>(rm.meta + r * 128 + 4 * lane);
      rm.tw[r] = *reinterpret_cast<const uint32_t *>(rm.meta + 8 * 128 + r * 128 + 4 * lane);
    }
  }
}

template <typename T, int ROWS>
__device__ __forceinline__ void dec_rotate_task(const DecParams &p, const DecRotMeta &rm, const void *xpart, int gk, int m0, int lane, uint32_t rot,
                                                uint32_t xb_group, long long t_entry) {
  uint2 raw[ROWS];
  struct { const void *x; int M, K; } px = {xpart, p.M, p.K};   // my partition's rows (a pre-rotated x has one copy per partition)
  load_x<T, ROWS>(px, gk, lane, raw, m0);
  scale_and_stage<T, ROWS>(rot, lane, raw, rm.csw);
  __syncwarp();
  if (p.trace && threadIdx.x == 0 && blockIdx.x < kDecTraceCtas) g_dec_trace[blockIdx.x * kDecTraceSlots + 0] = clock64() - t_entry;
  if (p.krot == 8) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      // (keeping all 32 coefficients live across griddepcontrol.wait was tried: it spills at 80 registers and the
      // stages get slower than with the MUFU work inline)
      float c0, s0, c1, s1;
      sincos2<T>(rm.tw[r], c0, s0, c1, s1);
      rotate_stage<T, ROWS>(rot, rm.idxw[r], c0, s0, c1, s1);
      __syncwarp();
    }
  } else {
    const int krot = p.krot;
    for (int r = 0; r < krot; ++r) {
      const uint32_t iw = *reinterpret_cast<const uint32_t *>(rm.meta + r * 128 + 4 * lane);
      const uint32_t tw = *reinterpret_cast<const uint32_t *>(rm.meta + krot * 128 + r * 128 + 4 * lane);
      float c0, s0, c1, s1;
      sincos2<T>(tw, c0, s0, c1, s1);
      rotate_stage<T, ROWS>(rot, iw, c0, s0, c1, s1);
      __syncwarp();
    }
  }
  if (p.trace && threadIdx.x == 0 && blockIdx.x < kDecTraceCtas) g_dec_trace[blockIdx.x * kDecTraceSlots + 9] = clock64() - t_entry;
  const int left = p.M - m0;
  dec_write_b_rows<T, ROWS>(xb_group, rot, m0, left < ROWS ? left : ROWS, lane);
  __syncwarp();
}

// tasks t = wi, wi + nwarps, ... of ng * nq (nq row blocks per group); the first task's metadata was fetched before the wait
template <typename T, int ROWS>
__device__ __forceinline__ void dec_prologue(const DecParams &p, DecRotMeta &rm, int part, int g_begin, int ntasks, int nq, int wi, int nwarps,
                                             int lane, uint32_t rot, uint32_t xb, long long t_entry) {
  const void *xpart = static_cast<const T *>(p.x) + static_cast<long long>(part) * p.x_part_stride;
  for (int t = wi; t < ntasks; t += nwarps) {
    const int gi = t / nq, rq = t - gi * nq;
    if (t != wi) dec_fetch_meta(p, part, g_begin + gi, lane, rm);
    dec_rotate_task<T, ROWS>(p, rm, xpart, g_begin + gi, rq * ROWS, lane, rot, xb + gi * (kDecN * 256), t_entry);
  }
}

template <typename T>
__device__ __forceinline__ void dec_store(const DecParams &p, float v, int m, int n) {
  T t = Traits<T>::from_float(v);
  if (p.bias) t = Traits<T>::from_float(Traits<T>::to_float(t) + Traits<T>::to_float(static_cast<const T *>(p.bias)[n]));  // plugin.py:309-310
  static_cast<T *>(p.y)[static_cast<int64_t>(m) * p.N + n] = t;
}

template <typename T, int SETS>
__global__ void __launch_bounds__(32 * (4 * SETS + 2), 1) decode_kernel(const DecParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int kWorkers = 4 * SETS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_entry = clock64();
  if (p.trace && threadIdx.x == 0 && blockIdx.x < kDecTraceCtas) g_dec_trace[blockIdx.x * kDecTraceSlots + 10] = globaltimer_ns();
  const int NS = p.nstages;
  const uint32_t smem0 = smem_u32(smem);
  const uint32_t xb = smem0 + p.xb_off, recv = smem0 + p.recv_off, bars = smem0 + p.bar_off;
  const uint32_t bar_wfull = bars, bar_wempty = bars + 8 * kDecMaxStages;
  const uint32_t bar_afull = bars + 16 * kDecMaxStages, bar_afree = bar_afull + 64;
  const uint32_t bar_dfull = bar_afull + 128, bar_dfree = bar_afull + 192, bar_xb = bar_afull + 208, tmem_slot = bar_afull + 216;
  const uint32_t bar_xload = bar_afull + 224;   // pre-rotated x: one bulk copy fills the whole B operand
  constexpr uint32_t d_col0 = 64 * SETS;

  // ---- my range of 128-column blocks (within one partition) and my K-slice (whole groups, ragged)
  const int range = blockIdx.x >> p.c_shift;
  const int slice = blockIdx.x & (p.c - 1);   // == %cluster_ctarank
  int part = 0;
  while (range >= p.part_range_begin[part + 1]) ++part;
  const int jl = range - p.part_range_begin[part];
  const int cp = p.part_range_begin[part + 1] - p.part_range_begin[part];
  const int cbp = p.part_block_begin[part + 1] - p.part_block_begin[part];
  const int cb_begin = static_cast<int>(static_cast<unsigned>(jl) * static_cast<unsigned>(cbp) / static_cast<unsigned>(cp));
  const int cb_end = static_cast<int>(static_cast<unsigned>(jl + 1) * static_cast<unsigned>(cbp) / static_cast<unsigned>(cp));
  const int nj = cb_end - cb_begin;
  const int g_begin = static_cast<int>(static_cast<unsigned>(slice) * static_cast<unsigned>(p.groups) >> p.c_shift);
  const int g_end = static_cast<int>(static_cast<unsigned>(slice + 1) * static_cast<unsigned>(p.groups) >> p.c_shift);
  const int ng = g_end - g_begin;
  const int nrounds = nj * ng;
  // first output column of my first block, and the end of my partition (a last partial block is masked at the stores)
  const int n_first = p.part_col_begin[part] + cb_begin * kBlockN, n_end = p.part_col_begin[part + 1];
  // record of (block j, group gi) of mine: rec0 + (j * groups + gi) * rec_bytes
  const uint32_t rec_bytes = static_cast<uint32_t>(p.rec_bytes);
  const uint8_t *rec0 = p.packed + p.rec_off + (static_cast<size_t>(p.part_block_begin[part] + cb_begin) * p.groups + g_begin) * rec_bytes;

  DecRotMeta rm;
  const int nq = p.M > 4 ? (p.M + 3) >> 2 : 1, ntasks = ng * nq;   // rotation tasks: row blocks of 4 for M > 4
  if (warp < kWorkers && warp < ntasks && warp < p.rot_warps)
    dec_fetch_meta(p, part, g_begin + warp / nq, lane, rm);   // flies during the barrier init / TMEM allocation

  if (warp == kWorkers) {
    // one barrier per lane
    if (lane < NS) {
      mbar_init(bar_wfull + 8 * lane, 1);
      mbar_init(bar_wempty + 8 * lane, 4);
    }
    // Every barrier has ONE waiting party that consumes its phases in order (a parity wait issued a phase early passes
    // on the stale phase): A-buffer and D-ready barriers are per dequant set, ring stages are always consumed by the
    // same set (stage count is a multiple of SETS), the producer / MMA lanes wait sequentially on the rest.
    if (lane < SETS) {
      mbar_init(bar_afull + 8 * lane, 4);
      mbar_init(bar_afree + 8 * lane, 1);
      mbar_init(bar_dfull + 8 * lane, 1);
    }
    if (lane < 2) mbar_init(bar_dfree + 8 * lane, 4);
    if (lane == 0) {
      mbar_init(bar_xb, kWorkers);
      mbar_init(bar_xload, 1);
    }
    fence_mbar_init();
  }
  if (warp == kWorkers + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kDecTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = lds32(tmem_slot);
  DEC_TRACE(1);
  if (p.c > 1) cluster_arrive_relaxed();   // #1 "this CTA runs" -- waited on before the first DSMEM push into a peer's receive buffer
  pdl_launch_dependents();                 // let the next linear in the stream start prefetching its weights
  bool cluster_ready = false;

  if (warp == kWorkers) {
    // ================= producer: ONE bulk copy per round (weights + scales + zeros of a (block, group), 8576 contiguous
    // bytes); nothing here depends on the previous kernel.  The whole warp runs the loop, an elected lane issues.
    const uint64_t pol = policy_evict_first();
    int st = 0, it = 0, j = 0, gi = 0;
#pragma unroll 1
    for (int r = 0; r < nrounds; ++r) {
      if (it > 0) mbar_wait(bar_wempty + 8 * st, (it - 1) & 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_wfull + 8 * st, rec_bytes);
        bulk_g2s(smem0 + st * rec_bytes, rec0 + (static_cast<size_t>(j) * p.groups + gi) * rec_bytes, rec_bytes, bar_wfull + 8 * st, pol);
        if (p.trace && blockIdx.x == 0 && r < 100) g_dec_trace[kDecTraceCtas * kDecTraceSlots + 3 * r] = clock64() - t_entry;
      }
      __syncwarp();
      if (++st == NS) { st = 0; ++it; }
      if (++gi == ng) { gi = 0; ++j; }
    }
  } else if (warp == kWorkers + 1) {
    // ================= MMA issuer.  The whole warp runs the loop (waits included) and ONE elected lane issues:
    // inside an `elect.sync` region ptxas keeps descriptors and TMEM addresses in uniform registers; under a plain
    // `lane == 0` branch every tcgen05.mma was wrapped in a divergence loop (~15 instructions, ~120 cycles per MMA,
    // tools/mma_probe.cu) and the single issuing thread, not the tensor pipe, bounded the kernel.
    const uint32_t idesc = instr_desc<T>(kDecN);
    if (p.pre_rotated) {
      // the pre-pass (paro_rotate.cu, tiled for 16 rows) wrote x_rot in exactly the B-operand order of this kernel, zero padding
      // rows included: the K-slice's groups are ONE contiguous range -> one bulk copy, no rotation tasks, no proxy fence
      pdl_wait();   // ... written by the previous kernel in the stream
      if (elect_one()) {
        const uint32_t bytes = static_cast<uint32_t>(ng) * (kDecN * 256);
        mbar_arrive_expect_tx(bar_xload, bytes);
        bulk_g2s(xb, static_cast<const uint8_t *>(p.x) + (static_cast<size_t>(part) * p.x_part_stride + static_cast<size_t>(g_begin) * (kDecN * kGroup)) * 2,
                 bytes, bar_xload, policy_evict_first());
      }
      __syncwarp();
      mbar_wait(bar_xload, 0);
    }
    mbar_wait(bar_xb, 0);   // B operand rows written (generic proxy) and fenced by the workers
    const uint64_t desc_hi = smem_desc_kmajor(0, kDecN * 16, 128);   // everything but the start address
    int set = 0, use = 0, j = 0, gi = 0;
#pragma unroll 1
    for (int r = 0; r < nrounds; ++r) {
      if (gi == 0 && j >= 2) mbar_wait(bar_dfree + 8 * (j & 1), ((j >> 1) - 1) & 1);   // D buffer read back by its previous user
      mbar_wait(bar_afull + 8 * set, use & 1);
      tc_fence_after();
      const uint32_t td = tmem + d_col0 + (j & 1) * kDecN, ta = tmem + set * 64;
      const uint64_t bdesc0 = desc_hi | static_cast<uint64_t>(((xb + gi * (kDecN * 256)) >> 4) & 0x3FFF);
      if (elect_one()) {
#pragma unroll
        for (int s = 0; s < 8; ++s) tc_mma_ts(td, ta + 8 * s, bdesc0 + s * ((kDecN * 32) >> 4), idesc, (gi | s) ? 1u : 0u);
        tc_commit(bar_afree + 8 * set);
        if (gi == ng - 1) tc_commit(bar_dfull + 8 * set);   // the set that dequantised the block's last group reads D back
      }
      __syncwarp();
      if (++set == SETS) { set = 0; ++use; }
      if (++gi == ng) { gi = 0; ++j; }
    }
  } else {
    // ================= workers
    const int wi = warp, e = wi >> 2, q = warp & 3;
    const int L128 = 32 * q + lane;               // output column inside the 128-column block = TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(32 * q) << 16;
    // rows [M, 16) of every group are the zero padding of the N = 16 MMA: all workers clear the B operand before the
    // dependency wait (it does not depend on x); the rotation tasks then write only token rows
    if (p.M < kDecN && !p.pre_rotated) {
      for (int i = threadIdx.x; i < ng * (kDecN * 16); i += 32 * kWorkers) sts128u(xb + i * 16, make_uint4(0u, 0u, 0u, 0u));
      named_bar_sync(1, 32 * kWorkers);
    }
    DEC_TRACE(2);
    if (!p.pre_rotated) pdl_wait();  // x may have been written by the previous kernel (pre-rotated: the MMA warp fetches it)
    DEC_TRACE(3);
    if (!p.pre_rotated && wi < ntasks && wi < p.rot_warps) {
      const uint32_t rot = smem0 + p.rot_off + wi * p.rot_bytes;
      const int M = p.M, rw = p.rot_warps;
      if (M == 1) dec_prologue<T, 1>(p, rm, part, g_begin, ntasks, nq, wi, rw, lane, rot, xb, t_entry);
      else if (M == 2) dec_prologue<T, 2>(p, rm, part, g_begin, ntasks, nq, wi, rw, lane, rot, xb, t_entry);
      else dec_prologue<T, 4>(p, rm, part, g_begin, ntasks, nq, wi, rw, lane, rot, xb, t_entry);
      fence_proxy_async_smem();  // B rows were written through the generic proxy, tcgen05.mma reads them through the async proxy
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_xb);
    DEC_TRACE(4);

    const int M = p.M;
    int j = 0, gi = e;
    while (gi >= ng && j < nj) { gi -= ng; ++j; }
    int st = e, use = 0;
    uint32_t par = 0, epi = 0;
    const uint32_t ta = tmem + lane_base + e * 64;
    const uint32_t col_off = (L128 >> 4) * 1024 + (L128 & 15) * 16;   // my 16-byte slots inside a record's weights
    const bool q2 = p.q2 != 0;
    const uint32_t zero_off = q2 ? block_zero_off(2) : block_zero_off(1);
    bool first = true;
#pragma unroll 1
    for (int r = e; r < nrounds; r += SETS) {
      mbar_wait(bar_wfull + 8 * st, par);
      if (first) { DEC_TRACE(5); first = false; }
      if (p.trace && q == 0 && lane == 0 && blockIdx.x == 0 && r < 100) g_dec_trace[kDecTraceCtas * kDecTraceSlots + 3 * r + 1] = clock64() - t_entry;
      const uint32_t rec = smem0 + st * rec_bytes;
      RowDequant<T> dq;
      dq.prep(lds16(rec + kBlockScaleOff + 2 * L128), lds8(rec + zero_off + L128));
      uint32_t s_hi = 0, z_hi = 0;   // group_size 64: channels 64..127 of the record have their own scale / zero
      if (q2) { s_hi = lds16(rec + kBlockScaleOff + 256 + 2 * L128); z_hi = lds8(rec + zero_off + 128 + L128); }
      if (use > 0) mbar_wait(bar_afree + 8 * e, (use - 1) & 1);  // the MMAs of my previous round have drained my A buffer
      tc_fence_after();
      const uint32_t wbase = rec + col_off;
#pragma unroll kDecUnroll
      for (int c = 0; c < 4; ++c) {
        const uint4 w4 = lds128(wbase + c * 256);
        uint32_t regs[16];
        if (c == 2 && q2) dq.prep(s_hi, z_hi);
        dq.word(w4.x, regs + 0);
        dq.word(w4.y, regs + 4);
        dq.word(w4.z, regs + 8);
        dq.word(w4.w, regs + 12);
        tc_st16(ta + 16 * c, regs);
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_afull + 8 * e);
        mbar_arrive(bar_wempty + 8 * st);
        if (p.trace && q == 0 && blockIdx.x == 0 && r < 100) g_dec_trace[kDecTraceCtas * kDecTraceSlots + 3 * r + 2] = clock64() - t_entry;
      }
      if (gi == ng - 1) {
        // ---- this set closed block j: read D back, one value per token
        const int b = j & 1;
        mbar_wait(bar_dfull + 8 * e, epi & 1);
        ++epi;
        tc_fence_after();
        uint32_t v[16];
        if (M <= 8) tc_ld8(tmem + lane_base + d_col0 + b * kDecN, v);
        else tc_ld16(tmem + lane_base + d_col0 + b * kDecN, v);
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_dfree + 8 * b);
        if (p.c == 1) {
          const int n = n_first + j * kBlockN + L128;
          if (n < n_end) {
#pragma unroll
            for (int m = 0; m < 16; ++m)
              if (m < M) dec_store<T>(p, __uint_as_float(v[m]), m, n);
          }
        } else {
          if (!cluster_ready) { cluster_wait_acquire(); cluster_ready = true; }  // #1: every peer CTA is running
          const uint32_t dst = map_to_rank(recv + (((j >> p.c_shift) << p.c_shift) + slice) * (M * 512) + L128 * 4, j & (p.c - 1));
#pragma unroll
          for (int m = 0; m < 16; ++m)
            if (m < M) st_cluster_f32(dst + m * 512, __uint_as_float(v[m]));
        }
      }
      st += SETS;
      if (st >= NS) { st -= NS; par ^= 1; }
      ++use;
      gi += SETS;
      while (gi >= ng) { gi -= ng; ++j; }
    }
    DEC_TRACE(6);
  }

  // ================= common tail (all threads; warps reconverge first)
  __syncwarp();
  if (p.c > 1) {
    if (!cluster_ready) cluster_wait_acquire();
    cluster_arrive_release();                              // #2: all my pushes are done
    cluster_wait_acquire();
    DEC_TRACE(7);
    if (warp < kWorkers) {
      // finish the blocks that were sent to me (j % c == slice); fixed order over the K-slices
      const int M = p.M;
      const int nmine = nj > slice ? ((nj - slice + p.c - 1) >> p.c_shift) : 0;
      const int per = M * 128;
      for (int idx = threadIdx.x; idx < nmine * per; idx += 32 * kWorkers) {
        const int jo = idx / per, o = idx - jo * per;
        const int m = o >> 7, col = o & 127;
        float acc = 0.f;
        for (int src = 0; src < p.c; ++src) acc += lds_f32(recv + ((jo << p.c_shift) + src) * (M * 512) + o * 4);
        const int n = n_first + ((jo << p.c_shift) + slice) * kBlockN + col;
        if (n < n_end) dec_store<T>(p, acc, m, n);
      }
    }
  }
  DEC_TRACE(8);
  if (p.trace && threadIdx.x == 0 && blockIdx.x < kDecTraceCtas) g_dec_trace[blockIdx.x * kDecTraceSlots + 11] = globaltimer_ns();
  tc_fence_before();
  __syncthreads();
  if (warp == kWorkers + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kDecTmemCols) : "memory");
  }
}

int decode_trace_read(unsigned long long *host, int max_ctas) {
  const size_t n = static_cast<size_t>(max_ctas < kDecTraceRows ? max_ctas : kDecTraceRows) * kDecTraceSlots;
  PARO_CUDA_OK(cudaMemcpyFromSymbol(host, g_dec_trace, n * sizeof(unsigned long long)));
  return PARO_OK;
}

// ------------------------------------------------------------------ host side
static int dec_env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}
// environment knobs are read ONCE per process (every launch used to pay half a dozen getenv calls)
struct DecKnobs { int stages, no_cluster, force_c, verbose, no_pdl, trace, sets; };
static const DecKnobs &dec_knobs() {
  static const DecKnobs k = {dec_env_int("PARO_DECODE_STAGES", 0), dec_env_int("PARO_NO_CLUSTER", 0), dec_env_int("PARO_DECODE_C", 0),
                             dec_env_int("PARO_DECODE_VERBOSE", 0), dec_env_int("PARO_NO_PDL", 0), dec_env_int("PARO_DECODE_TRACE", 0),
                             dec_env_int("PARO_DECODE_SETS", 5)};
  return k;
}

struct DecPlan {
  int c, c_shift, ranges, grid;
  int part_range_begin[PARO_MAX_PARTS + 1];
  int nj_max, ng_max;
};

// `ranges` block ranges over the partitions, proportional to their 128-column block counts (largest remainder)
static bool dec_split_ranges(const Layout &L, int ranges, DecPlan &plan) {
  int cb[PARO_MAX_PARTS], alloc[PARO_MAX_PARTS], total = 0;
  double frac[PARO_MAX_PARTS];
  for (int p = 0; p < L.n_parts; ++p) {
    cb[p] = L.part_block_begin[p + 1] - L.part_block_begin[p];
    total += cb[p];
  }
  if (ranges > total) ranges = total;
  if (ranges < L.n_parts) return false;
  int used = 0;
  for (int p = 0; p < L.n_parts; ++p) {
    const double exact = static_cast<double>(ranges) * cb[p] / total;
    alloc[p] = static_cast<int>(exact);
    if (alloc[p] < 1) alloc[p] = 1;
    if (alloc[p] > cb[p]) alloc[p] = cb[p];
    frac[p] = exact - alloc[p];
    used += alloc[p];
  }
  while (used < ranges) {
    int best = -1;
    for (int p = 0; p < L.n_parts; ++p)
      if (alloc[p] < cb[p] && (best < 0 || frac[p] > frac[best])) best = p;
    if (best < 0) break;
    alloc[best]++; frac[best] -= 1.0; used++;
  }
  while (used > ranges) {
    int best = -1;
    for (int p = 0; p < L.n_parts; ++p)
      if (alloc[p] > 1 && (best < 0 || frac[p] < frac[best])) best = p;
    if (best < 0) return false;
    alloc[best]--; frac[best] += 1.0; used--;
  }
  plan.part_range_begin[0] = 0;
  plan.nj_max = 0;
  for (int p = 0; p < PARO_MAX_PARTS; ++p) {
    plan.part_range_begin[p + 1] = plan.part_range_begin[p] + (p < L.n_parts ? alloc[p] : 0);
    if (p < L.n_parts) {
      const int mx = (cb[p] + alloc[p] - 1) / alloc[p];
      if (mx > plan.nj_max) plan.nj_max = mx;
    }
  }
  plan.ranges = plan.part_range_begin[L.n_parts];
  return plan.ranges > 0;
}

struct DecSmem { int xb_off, rot_off, recv_off, bar_off, nstages, rot_warps, total; };

static bool dec_carve_with(const DecPlan &plan, int M, int rot_bytes, int sets, int rot_warps, int stage_bytes, DecSmem &s) {
  const int xb_bytes = plan.ng_max * kDecN * 256;
  const int rot_total = (rot_warps * rot_bytes + 127) / 128 * 128;
  const int recv = plan.c > 1 ? ((plan.nj_max + plan.c - 1) / plan.c) * plan.c * M * 512 : 0;
  const int bar_bytes = 16 * kDecMaxStages + 256;
  const int fixed = xb_bytes + rot_total + recv + bar_bytes + 128;
  int nst = (kDecSmemLimit - fixed) / stage_bytes;
  if (nst > kDecMaxStages) nst = kDecMaxStages;
  const int want = dec_knobs().stages;
  if (want >= sets && want < nst) nst = want;
  nst = nst / sets * sets;   // a stage is always consumed by the same set
  if (nst < sets) return false;
  s.nstages = nst;
  s.rot_warps = rot_warps;
  s.xb_off = (nst * stage_bytes + 127) / 128 * 128;
  s.rot_off = s.xb_off + xb_bytes;
  s.recv_off = s.rot_off + rot_total;
  s.bar_off = s.recv_off + (recv + 127) / 128 * 128;
  s.total = s.bar_off + bar_bytes;
  return s.total <= kDecSmemLimit;
}

// every group gets its own rotating warp when the tiles fit; otherwise fewer warps take several groups each
static bool dec_carve(const DecPlan &plan, int M, int rot_bytes, int sets, int stage_bytes, DecSmem &s) {
  const int ntasks = plan.ng_max * (M > 4 ? (M + 3) / 4 : 1);
  int rw = ntasks < 4 * sets ? ntasks : 4 * sets;
  for (;;) {
    if (dec_carve_with(plan, M, rot_bytes, sets, rw, stage_bytes, s)) return true;
    if (rw <= 2) return false;
    rw = rw > 8 ? 8 : rw / 2;
  }
}

// how many clusters of `c` CTAs with this footprint the device keeps resident (GPC granularity); cached per thread
template <typename T, int SETS>
static int max_resident_clusters(int c, int smem_bytes, int sms) {
  if (c == 1) return sms;
  int slot = 0;
  while ((1 << slot) < c) ++slot;
  auto kern = decode_kernel<T, SETS>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kDecSmemLimit) != cudaSuccess) { (void)cudaGetLastError(); return 0; }   // never lowered: cached plans keep launching
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((sms / c) * c);
  cfg.blockDim = dim3(32 * (4 * SETS + 2));
  cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = c;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
  return n;   // (the caller caches the whole plan per shape, M and device)
}

// The launch plan: cluster size c (K slices), block ranges per partition, shared-memory carve-up.  Pure host logic --
// `resident(c, smem_bytes)` says how many clusters of c CTAs with that footprint the device keeps resident (all clusters
// must be co-resident: one wave) -- so tests can drive it without a GPU (paro_debug_decode_plan).
template <typename ResidentFn>
static bool dec_choose_plan(const Layout &L, int M, int rot_bytes, int sets, int sms, bool cluster_ok, int force_c, ResidentFn resident,
                            DecPlan &best, DecSmem &best_s) {
  long best_cost = -1;
  for (int cs = 0; cs <= 3; ++cs) {
    const int c = 1 << cs;
    if (c > 1 && !cluster_ok) break;
    if (c > L.groups) break;
    if (force_c && c != force_c) continue;
    DecPlan plan = {};
    plan.c = c; plan.c_shift = cs;
    plan.ng_max = (L.groups + c - 1) / c;
    if (!dec_split_ranges(L, sms / c, plan)) continue;
    DecSmem s;
    if (!dec_carve(plan, M, rot_bytes, sets, L.rec_bytes, s)) continue;
    const int res = resident(c, s.total);
    if (res < L.n_parts) continue;
    if (res < plan.ranges) {
      if (!dec_split_ranges(L, res, plan)) continue;
      if (!dec_carve(plan, M, rot_bytes, sets, L.rec_bytes, s)) continue;
    }
    plan.grid = plan.ranges * c;
    // critical path in rounds: main loop of the busiest CTA + rotation passes + a small charge per doubling of the cluster
    const int passes = (plan.ng_max * (M > 4 ? (M + 3) / 4 : 1) + 4 * sets - 1) / (4 * sets);
    long cost = static_cast<long>(plan.nj_max) * plan.ng_max * 8 + passes * 24 + cs * 2;
    if (s.nstages < 2 * sets) cost += cost / 4;   // a ring shallower than two rounds per set starves the dequant warps (large M: the receive buffer crowds it out)
    if (best_cost < 0 || cost < best_cost) { best = plan; best_s = s; best_cost = cost; }
  }
  return best_cost >= 0;
}

static int dec_rot_bytes(int64_t M) { return kGroup * 2 * (M == 1 ? 1 : M == 2 ? 2 : 4); }   // rotation tile of one task: 128 channels x up to 4 rows

// test hook: the plan for given resident-cluster counts (index log2 c), no device involved
int decode_debug_plan(const Layout &L, int64_t M, int sets, int sms, const int32_t *resident4, int32_t *out20) {
  if (M < 1 || M > 16 || sets < 4 || sets > 7) { set_error("debug_plan: bad M / sets"); return PARO_EINVAL; }
  DecPlan plan = {};
  DecSmem sm = {};
  auto res = [&](int c, int) { int i = 0; while ((1 << i) < c) ++i; return static_cast<int>(resident4[i]); };
  if (!dec_choose_plan(L, static_cast<int>(M), dec_rot_bytes(M), sets, sms, true, 0, res, plan, sm)) {
    set_error("decode: no launch configuration fits (in_features=%d, M=%d)", L.K, static_cast<int>(M));
    return PARO_EUNSUPPORTED;
  }
  const int32_t v[11] = {plan.c, plan.ranges, plan.grid, plan.nj_max, plan.ng_max, sm.nstages, sm.rot_warps, sm.total, sm.xb_off, sm.recv_off, sm.bar_off};
  for (int i = 0; i < 11; ++i) out20[i] = v[i];
  for (int i = 0; i <= PARO_MAX_PARTS; ++i) out20[11 + i] = plan.part_range_begin[i];
  return PARO_OK;
}

struct DecCached { bool found; DecPlan plan; DecSmem smem; };

template <typename T, int SETS>
static int launch_decode(DecParams &p, const Layout &L, int sms, cudaStream_t stream) {
  auto kern = decode_kernel<T, SETS>;
  const DecKnobs &kn = dec_knobs();
  // the plan (cluster size, block ranges, shared-memory carve-up) depends on the shape, M and the device only: searched once
  // (the search asks the occupancy API and sets the function attribute), then looked up -- an eager launch is a map lookup
  struct Key {
    int K, N, krot, n_parts, M, dev, qhalves, parts[PARO_MAX_PARTS];
    bool operator<(const Key &o) const { return memcmp(this, &o, sizeof(Key)) < 0; }
  };
  static std::mutex mu;
  static std::map<Key, DecCached> cache;
  Key key;
  memset(&key, 0, sizeof(key));
  key.K = L.K; key.N = L.N; key.krot = L.krot; key.n_parts = L.n_parts; key.M = p.M; key.qhalves = L.qhalves;
  if (cudaGetDevice(&key.dev) != cudaSuccess) key.dev = 0;
  for (int i = 0; i < L.n_parts; ++i) key.parts[i] = L.part_col_begin[i + 1] - L.part_col_begin[i];
  DecCached c;
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it == cache.end()) {
      DecCached fresh = {};
      fresh.found = dec_choose_plan(L, p.M, p.rot_bytes, SETS, sms, !kn.no_cluster, kn.force_c,
                                    [&](int cc, int smem) { return max_resident_clusters<T, SETS>(cc, smem, sms); }, fresh.plan, fresh.smem);
      if (fresh.found && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kDecSmemLimit) != cudaSuccess) {
        (void)cudaGetLastError();
        fresh.found = false;
      }
      it = cache.emplace(key, fresh).first;
    }
    c = it->second;
  }
  if (!c.found) { set_error("decode: no launch configuration fits (in_features=%d, M=%d)", L.K, p.M); return PARO_EUNSUPPORTED; }
  const DecPlan &best = c.plan;
  const DecSmem &best_s = c.smem;
  if (kn.verbose)
    fprintf(stderr, "[paro decode] K=%d N=%d M=%d: cluster %d x %d ranges (grid %d), blocks/CTA <= %d, groups/CTA <= %d, %d stages, smem %d\n",
            L.K, L.N, p.M, best.c, best.ranges, best.grid, best.nj_max, best.ng_max, best_s.nstages, best_s.total);

  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(best.grid);
  cfg.blockDim = dim3(32 * (4 * SETS + 2));
  cfg.dynamicSmemBytes = best_s.total;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (best.c > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = best.c;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (!kn.no_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  p.c = best.c; p.c_shift = best.c_shift;
  p.nstages = best_s.nstages;
  p.rot_warps = best_s.rot_warps;
  p.xb_off = best_s.xb_off; p.rot_off = best_s.rot_off; p.recv_off = best_s.recv_off; p.bar_off = best_s.bar_off;
  for (int i = 0; i <= PARO_MAX_PARTS; ++i) p.part_range_begin[i] = best.part_range_begin[i];
  PARO_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p));
  note_launches(1);
  return PARO_OK;
}

bool decode_supported(const Layout &L, int64_t M) { return M >= 1 && M <= 16 && L.groups >= 1; }

// From 5 rows on (3 when K >= 8192) the rotation runs ONCE, in a pre-pass (paro_rotate.cu: rotate_small_kernel, one launch for all
// partitions, ~2.5 us) that writes x_rot in this kernel's B-operand order, and the kernel fetches its K slice with one bulk copy.
// Inside the kernel a 4-row rotation task is ~1000 dependent instructions of one warp (7-14 k cycles with 20 warps contending) and
// at M = 16 a CTA has 32-56 of them: 28 k of o_proj's 40 k cycles, 48 k of down_proj's 80 k (tools/trace_decode.py).  Measured
// (us, in-kernel -> pre-pass): M = 16: o 21.7 -> 13.4, qkv 18.2 -> 16.4, gate_up 38.1 -> 29.1, down 42.3 -> 23.8; M = 8: 14.0 ->
// 11.8, 16.7 -> 15.0, 29.7 -> 27.5, 38.2 -> 20.3; M = 4: o 10.0 -> 11.4 (stays in-kernel), down 25.1 -> 18.7.
static int dec_prerot_knob() {
  static const int v = dec_env_int("PARO_DECODE_PREROT_M", 0);   // 0: the rule below; n >= 2: from n rows on; 1: never
  return v;
}
static bool decode_wants_prerot(const Layout &L, int64_t M) {
  const int k = dec_prerot_knob();
  if (k == 1) return false;
  if (k >= 2) return M >= k;
  return M >= 5 || (M >= 3 && L.groups >= 64);
}
size_t decode_scratch_bytes(const Layout &L, int64_t max_m) {
  bool any = false;
  for (int64_t m = 1; m <= (max_m < 16 ? max_m : 16); ++m) any = any || decode_wants_prerot(L, m);
  return any ? static_cast<size_t>(L.n_parts) * kDecN * L.K * 2 : 0;
}

int rotate_small_launch(const void *x, void *out, const void *raw_base, long long raw_part_bytes, int n_parts, int64_t M, int64_t M_store, int nt,
                        int K, int krot, int dtype, cudaStream_t stream);

int decode_forward(const paro_linear_shape &s, const Layout &L, const void *packed, const void *x, int64_t M,
                   const void *bias, void *y, void *scratch, size_t scratch_bytes, cudaStream_t stream) {
  int dev = 0;
  PARO_CUDA_OK(cudaGetDevice(&dev));
  static thread_local int sms_of[64] = {};   // SM count per device, asked once
  if (!sms_of[dev & 63]) PARO_CUDA_OK(cudaDeviceGetAttribute(&sms_of[dev & 63], cudaDevAttrMultiProcessorCount, dev));
  const int sms = sms_of[dev & 63];
  const size_t part_bytes = static_cast<size_t>(kDecN) * L.K * 2;   // 16 rows (zero padded) x K of T, B-operand order
  const bool prerot = decode_wants_prerot(L, M) && scratch && scratch_bytes >= L.n_parts * part_bytes;
  if (prerot) {   // one launch for all partitions (paro_rotate.cu: rotate_small_kernel)
    const int rc = rotate_small_launch(x, scratch, static_cast<const uint8_t *>(packed) + L.raw_off, static_cast<long long>(L.raw_part_bytes), L.n_parts, M,
                                       kDecN, kDecN, L.K, L.krot, s.dtype, stream);
    if (rc) return rc;
  }
  DecParams p = {};
  p.packed = static_cast<const uint8_t *>(packed);
  p.x = prerot ? scratch : x; p.y = y; p.bias = bias;
  p.pre_rotated = prerot ? 1 : 0;
  p.x_part_stride = prerot ? static_cast<long long>(kDecN) * L.K : 0;
  p.M = static_cast<int>(M); p.K = L.K; p.N = L.N;
  p.n_parts = L.n_parts; p.groups = L.groups; p.krot = L.krot;
  p.rec_bytes = L.rec_bytes; p.q2 = L.qhalves == 2;
  p.rot_bytes = dec_rot_bytes(M);
  p.trace = dec_knobs().trace;
  for (int i = 0; i <= PARO_MAX_PARTS; ++i) {
    p.part_col_begin[i] = L.part_col_begin[i];
    p.part_block_begin[i] = L.part_block_begin[i];
  }
  p.meta_group_bytes = L.meta_group_bytes;
  p.meta_off = static_cast<long long>(L.meta_off);
  p.rec_off = static_cast<long long>(L.rec_off);
  const int sets = dec_knobs().sets;   // 5 sets x 4 dequant warps measured best (tools/microbench.py); 4..7 selectable
  const bool bf16 = s.dtype == PARO_BF16;
  switch (sets) {
    case 4: return bf16 ? launch_decode<__nv_bfloat16, 4>(p, L, sms, stream) : launch_decode<__half, 4>(p, L, sms, stream);
    case 7: return bf16 ? launch_decode<__nv_bfloat16, 7>(p, L, sms, stream) : launch_decode<__half, 7>(p, L, sms, stream);
    case 6: return bf16 ? launch_decode<__nv_bfloat16, 6>(p, L, sms, stream) : launch_decode<__half, 6>(p, L, sms, stream);
    default: return bf16 ? launch_decode<__nv_bfloat16, 5>(p, L, sms, stream) : launch_decode<__half, 5>(p, L, sms, stream);
  }
}

}  // namespace paro
