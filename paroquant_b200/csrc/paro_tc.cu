// Fused small-M path (M <= 16) on 5th-generation tensor cores: scaled pairwise rotation of x +
// INT4 group dequant + GEMV/GEMM in ONE launch per (merged) linear.  Replaces the reference's
// rotate -> Marlin kernel pairs (/root/reference/paroquant/inference/backends/vllm/plugin.py:281-311:
// 2n+1 launches for an n-way merged projection).
//
// Why tcgen05 even at batch 1: measured on B200, a warp-level mma.sync.m16n8k16 occupies the legacy
// HMMA sub-pipe for 32 cycles (sm__pipe_tensor_subpipe_hmma_cycles_active, profiles/), i.e. 256
// MAC/clk/SM -- at M = 1 that alone is 1.35x the HBM time of the weight stream.  tcgen05.mma runs
// 4096 MAC/clk/SM and is issued by one thread.
//
// Formulation: D[lane, col] += A[lane, k] * B[col, k] with M_mma = 128 lanes, K = 16 per instruction.
//   lane  = (part p = 0..7, row = 0..15): output column `row` of the current 16-column tile,
//           restricted to rotation/quantisation group p of the CTA's K-slice.  A[lane, :] are
//           that row's dequantised weights for the group -- written straight from registers into
//           TENSOR MEMORY (tcgen05.st), never through shared memory.
//   col   = (token m, part p'): B[8m + p', k] = x_rot[m][group p'][k], a 16-row smem operand per
//           k16 step, written once per CTA after the in-kernel rotation.
//   so D[(p,row), 8m+p] accumulates the dot product of row `row` with token m over group p; the
//   other columns are ignored.  Eight k16 steps cover the group; the epilogue adds the 8 parts.
// One tcgen05.mma therefore retires 2048 weights whatever M <= 2 is, and N_mma = 8*Mpad.
//
// Roles (320 threads, 2 CTAs / SM, TMEM 256 columns each):
//   warp 0      producer: cp.async.bulk (TMA) of the CTA's records into an mbarrier ring; issued
//               before griddepcontrol.wait (weights do not depend on the previous kernel)
//   warp 1      allocates TMEM; one lane issues tcgen05.mma / tcgen05.commit
//   warps 2-9   workers, two sets of four (TMEM lane quarter = warp % 4).  Prologue: warp wi owns
//               group(s) wi (+8): loads x, applies the reference's rotation (rounding points of
//               rotation.cuh:91-173) with __syncwarp only, writes the B-operand rows.  Main loop:
//               set e dequantises round r = e, e+2, ...: thread = one (part, row) lane, 16 words ->
//               64 TMEM columns; then (one round later, so the MMA latency is hidden) reads its D
//               column back (tcgen05.ld), the 8 parts meet in smem and 16*M threads emit the tile.
//   K-slices    the `slices` CTAs of a cluster cover K; per-tile partials are pushed through
//               distributed shared memory to rank (tile mod slices); one cluster barrier; fixed
//               summation order everywhere (bit-reproducible).  > 8 slices: global workspace.
//
// Numerics: x_rot as paro_rotate.cu; W = T((q - z) * T(s)) with ONE rounding, exactly the operand
// Marlin / AWQ form; fp32 accumulation in TMEM; one rounding to T; bias added in T.
#include "paro_tc_common.cuh"

namespace paro {

constexpr int kMaxStages = 8;
constexpr int kTcThreads = 320;
constexpr int kTmemCols = 256;

enum ReduceMode { kDirect = 0, kCluster = 1, kWorkspace = 2 };

// Developer aid (PARO_DECODE_TRACE=1): per-CTA phase timestamps of worker warp 0, SM clock cycles
// relative to kernel entry, + %globaltimer at entry / exit.  Read back with paro_debug_trace().
constexpr int kTraceSlots = 12;
constexpr int kTraceMaxCtas = 1024;
__device__ unsigned long long g_trace[kTraceMaxCtas * kTraceSlots];
#define PARO_TRACE(slot)                                                                                  \
  do {                                                                                                    \
    if (p.trace && warp == 2 && lane == 0 && blockIdx.x < kTraceMaxCtas)                                   \
      g_trace[blockIdx.x * kTraceSlots + (slot)] = static_cast<unsigned long long>(clock64() - t_entry); \
  } while (0)

struct TcParams {
  const uint8_t *packed;
  const void *x;
  void *y;
  const void *bias;
  float *partials;
  int *counters;
  int M, K, N;
  int n_parts, slices, groups, krot, nstages, tiles_total;
  int gps, nb, rec_bytes, rec_stride;
  int mpad, nmma, nd, na;        // padded tokens (2,4,8,16), MMA N = 8*mpad, number of D / A buffers in TMEM
  int mode, rot_bytes, recv_tiles, trace;
  int xb_off, rot_off, red_off, recv_off, bar_off;  // shared-memory carve-up (bytes)
  int part_tile_begin[PARO_MAX_PARTS + 1];
  int part_range_begin[PARO_MAX_PARTS + 1];
  int meta_group_bytes;
  long long meta_off, rec_off;
};


// rows of the B operand owned by this warp: B[8m + part][k] = x_rot[m][k] (zero for m >= M / invalid group)
template <typename T, int ROWS>
__device__ __forceinline__ void write_b_rows(const TcParams &p, uint32_t xb, uint32_t rot, int part, int b, bool valid, int lane) {
  const uint32_t step_bytes = p.nmma * 32, lbo = (p.nmma >> 3) * 128;
  for (int idx = lane; idx < 16 * p.mpad; idx += 32) {
    const int m = idx >> 4, s = (idx >> 1) & 7, h = idx & 1;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (valid && m < p.M) {
      const int c0 = 16 * s + 8 * h;
      if constexpr (ROWS == 1) {
        v = lds128(rot + c0 * 2);
      } else {
        uint32_t e[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) e[k] = lds16(rot + (c0 + k) * (ROWS * 2) + 2 * m);
        v = make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
      }
    }
    const int n = 8 * m + part;
    sts128u(xb + (b * 8 + s) * step_bytes + h * lbo + (n >> 3) * 128 + (n & 7) * 16, v);
  }
}

// rotation metadata of this warp's NB groups, fetched (and for NB == 1 run through MUFU) BEFORE
// griddepcontrol.wait: it does not depend on the previous kernel
template <int NB> struct RotMeta {
  uint32_t idxw[NB][8];
  uint32_t tw[NB][8];      // theta pairs (T bits)
  float c0[8], s0[8], c1[8], s1[8];  // hoisted coefficients, NB == 1 and krot == 8 only
  uint2 csw[NB];
  bool valid[NB];
  int gk[NB];
  const uint8_t *meta[NB];
  bool hoisted;
};

template <typename T, int NB>
__device__ __forceinline__ void fetch_rot_meta(const TcParams &p, int slice, int part_idx, int wi, int lane, RotMeta<NB> &rm) {
  rm.hoisted = (NB == 1) && p.krot == 8;
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    rm.gk[b] = slice * p.gps + b * 8 + wi;
    rm.valid[b] = rm.gk[b] < p.groups;
    rm.meta[b] = p.packed + p.meta_off + (static_cast<size_t>(part_idx) * p.groups + (rm.valid[b] ? rm.gk[b] : 0)) * p.meta_group_bytes;
    rm.csw[b] = make_uint2(0u, 0u);
    if (rm.valid[b]) {
      rm.csw[b] = *reinterpret_cast<const uint2 *>(rm.meta[b] + p.krot * 256 + 8 * lane);
      if (p.krot == 8) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          rm.idxw[b][r] = *reinterpret_cast<const uint32_t *>(rm.meta[b] + r * 128 + 4 * lane);
          rm.tw[b][r] = *reinterpret_cast<const uint32_t *>(rm.meta[b] + 8 * 128 + r * 128 + 4 * lane);
        }
      }
    }
  }
  if constexpr (NB == 1) {
    if (rm.hoisted && rm.valid[0]) {
#pragma unroll
      for (int r = 0; r < 8; ++r) sincos2<T>(rm.tw[0][r], rm.c0[r], rm.s0[r], rm.c1[r], rm.s1[r]);
    }
  }
}

// x-dependent part of the prologue of one worker warp: x, rotation, B rows
template <typename T, int ROWS, int NB>
__device__ __forceinline__ void rotate_groups(const TcParams &p, const RotMeta<NB> &rm, int wi, int lane, uint32_t rot0, uint32_t xb) {
  uint2 raw[NB][ROWS];
#pragma unroll
  for (int b = 0; b < NB; ++b)
    if (rm.valid[b]) load_x<T, ROWS>(p, rm.gk[b], lane, raw[b]);
#pragma unroll
  for (int b = 0; b < NB; ++b)
    if (rm.valid[b]) scale_and_stage<T, ROWS>(rot0 + b * p.rot_bytes, lane, raw[b], rm.csw[b]);
  __syncwarp();
  if (p.krot == 8) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int b = 0; b < NB; ++b)
        if (rm.valid[b]) {
          if (NB == 1 && rm.hoisted) {
            rotate_stage<T, ROWS>(rot0, rm.idxw[0][r], rm.c0[r], rm.s0[r], rm.c1[r], rm.s1[r]);
          } else {
            float c0, s0, c1, s1;
            sincos2<T>(rm.tw[b][r], c0, s0, c1, s1);
            rotate_stage<T, ROWS>(rot0 + b * p.rot_bytes, rm.idxw[b][r], c0, s0, c1, s1);
          }
        }
      __syncwarp();
    }
  } else {
    const int krot = p.krot;
    for (int r = 0; r < krot; ++r) {
#pragma unroll
      for (int b = 0; b < NB; ++b)
        if (rm.valid[b]) {
          const uint32_t iw = *reinterpret_cast<const uint32_t *>(rm.meta[b] + r * 128 + 4 * lane);
          const uint32_t tw = *reinterpret_cast<const uint32_t *>(rm.meta[b] + krot * 128 + r * 128 + 4 * lane);
          float c0, s0, c1, s1;
          sincos2<T>(tw, c0, s0, c1, s1);
          rotate_stage<T, ROWS>(rot0 + b * p.rot_bytes, iw, c0, s0, c1, s1);
        }
      __syncwarp();
    }
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) write_b_rows<T, ROWS>(p, xb, rot0 + b * p.rot_bytes, wi, b, rm.valid[b], lane);
}

template <typename T, int NB>
__device__ __forceinline__ void worker_prologue(const TcParams &p, const RotMeta<NB> &rm, int wi, int lane, uint32_t rot0, uint32_t xb) {
  const int M = p.M;
  if (M == 1) rotate_groups<T, 1, NB>(p, rm, wi, lane, rot0, xb);
  else if (M == 2) rotate_groups<T, 2, NB>(p, rm, wi, lane, rot0, xb);
  else if (M <= 4) rotate_groups<T, 4, NB>(p, rm, wi, lane, rot0, xb);
  else if (M <= 8) rotate_groups<T, 8, NB>(p, rm, wi, lane, rot0, xb);
  else rotate_groups<T, 16, NB>(p, rm, wi, lane, rot0, xb);
}

template <typename T>
__device__ __forceinline__ void store_out(const TcParams &p, float v, int m, int n) {
  T t = Traits<T>::from_float(v);
  if (p.bias) t = Traits<T>::from_float(Traits<T>::to_float(t) + Traits<T>::to_float(static_cast<const T *>(p.bias)[n]));  // plugin.py:309-310
  static_cast<T *>(p.y)[static_cast<int64_t>(m) * p.N + n] = t;
}

template <typename T, int NB>
__global__ void __launch_bounds__(kTcThreads, 2) tc_linear_kernel(const TcParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_entry = clock64();
  if (p.trace && threadIdx.x == 64 && blockIdx.x < kTraceMaxCtas) g_trace[blockIdx.x * kTraceSlots + 10] = globaltimer_ns();
  constexpr int nb = NB;
  const int nst = p.nstages, ND = p.nd;
  const uint32_t smem0 = smem_u32(smem);
  const uint32_t xb = smem0 + p.xb_off, rot_all = smem0 + p.rot_off, red = smem0 + p.red_off, recv = smem0 + p.recv_off;
  const uint32_t bars = smem0 + p.bar_off;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * kMaxStages;
  const uint32_t bar_afull = bars + 16 * kMaxStages, bar_afree = bar_afull + 32, bar_dfull = bar_afull + 64, bar_dfree = bar_afull + 80;
  const uint32_t bar_xb = bar_afull + 96, tmem_slot = bar_afull + 104;
  const int NA = p.na;
  const uint32_t d_col0 = 64 * NA;

  // ---- which tile range / slice / partition is mine (32-bit arithmetic on launch constants)
  const int range = blockIdx.x / p.slices;
  const int slice = blockIdx.x - range * p.slices;  // == %cluster_ctarank in cluster mode
  int part = 0;
  while (range >= p.part_range_begin[part + 1]) ++part;
  const int jl = range - p.part_range_begin[part];
  const int cp = p.part_range_begin[part + 1] - p.part_range_begin[part];
  const int tp = p.part_tile_begin[part + 1] - p.part_tile_begin[part];
  const int t_begin = static_cast<int>(static_cast<unsigned>(jl) * static_cast<unsigned>(tp) / static_cast<unsigned>(cp));
  const int t_end = static_cast<int>(static_cast<unsigned>(jl + 1) * static_cast<unsigned>(tp) / static_cast<unsigned>(cp));
  const int ntiles = t_end - t_begin;
  const int nrounds = ntiles * nb;
  const int tile_g0 = p.part_tile_begin[part] + t_begin;
  const uint8_t *rec_src = p.packed + p.rec_off +
                           (static_cast<size_t>(p.slices) * p.part_tile_begin[part] + static_cast<size_t>(slice) * tp + t_begin) * p.rec_bytes;

  // rotation metadata (and, for one group per warp, the MUFU work) is independent of everything else in the
  // prologue: issue it first so the loads fly during barrier init / TMEM allocation / the CTA-wide sync
  RotMeta<NB> rm;
  if (warp >= 2) fetch_rot_meta<T, NB>(p, slice, part, warp - 2, lane, rm);

  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 4 * nb);
    }
    for (int a = 0; a < 4; ++a) {
      mbar_init(bar_afull + 8 * a, 4);
      mbar_init(bar_afree + 8 * a, 1);
    }
    for (int e = 0; e < 2; ++e) {
      mbar_init(bar_dfull + 8 * e, 1);
      mbar_init(bar_dfree + 8 * e, 4);
    }
    mbar_init(bar_xb, 8);
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = lds32(tmem_slot);
  PARO_TRACE(1);
  if (p.mode == kCluster) cluster_arrive_relaxed();  // #1 "this CTA runs" -- waited on before the first DSMEM push
  pdl_launch_dependents();                           // let the next linear in the stream start prefetching its weights
  bool cluster_ready = false;

  if (warp == 0) {
    // ================= producer: stream the records; nothing here depends on the previous kernel
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();
      for (int i = 0; i < ntiles; ++i) {
        const int st = i % nst, it = i / nst;
        if (it > 0) mbar_wait(bar_empty + 8 * st, (it - 1) & 1);
        mbar_arrive_expect_tx(bar_full + 8 * st, p.rec_bytes);
        bulk_g2s(smem0 + st * p.rec_stride, rec_src + static_cast<size_t>(i) * p.rec_bytes, p.rec_bytes, bar_full + 8 * st, pol);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (whole warp in the loop, one elected lane issues: see paro_decode.cu)
    {
      const uint32_t idesc = instr_desc<T>(p.nmma);
      const uint32_t step_bytes = p.nmma * 32, lbo = (p.nmma >> 3) * 128;
      const uint64_t desc_hi = smem_desc_kmajor(0, lbo, 128);
      mbar_wait(bar_xb, 0);  // B operand rows written (generic proxy) and fenced by the workers
      int abuf = 0, a_use = 0, dbuf = 0, d_use = 0, i = 0, b = 0;
#pragma unroll 1
      for (int r = 0; r < nrounds; ++r) {
        if (b == 0 && d_use > 0) mbar_wait(bar_dfree + 8 * dbuf, (d_use - 1) & 1);  // epilogue of the previous user of this D buffer
        mbar_wait(bar_afull + 8 * abuf, a_use & 1);
        tc_fence_after();
        const uint32_t td = tmem + d_col0 + dbuf * p.nmma, ta = tmem + abuf * 64;
        const uint64_t bdesc0 = desc_hi | static_cast<uint64_t>(((xb + b * 8 * step_bytes) >> 4) & 0x3FFF);
        if (elect_one()) {
#pragma unroll
          for (int s = 0; s < 8; ++s) tc_mma_ts(td, ta + 8 * s, bdesc0 + s * (step_bytes >> 4), idesc, (b | s) ? 1u : 0u);
          tc_commit(bar_afree + 8 * abuf);
          if (b == nb - 1) tc_commit(bar_dfull + 8 * (r & 1));  // the set that dequantised the last sub-round reads D back
        }
        __syncwarp();
        if (++abuf == NA) { abuf = 0; ++a_use; }
        if (++b == nb) {
          b = 0;
          ++i;
          if (++dbuf == ND) { dbuf = 0; ++d_use; }
        }
      }
    }
  } else {
    // ================= workers
    const int wi = warp - 2, e = wi >> 2, q = warp & 3;
    const int mypart = 2 * q + (lane >> 4), row = lane & 15;
    const uint32_t rot0 = rot_all + wi * nb * p.rot_bytes;
    PARO_TRACE(2);
    pdl_wait();  // x (and the workspace) may have been written by the previous kernel
    PARO_TRACE(3);
    worker_prologue<T, NB>(p, rm, wi, lane, rot0, xb);
    fence_proxy_async_smem();  // B rows were written through the generic proxy, tcgen05.mma reads them through the async proxy
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_xb);
    PARO_TRACE(4);

    // All loop state is carried incrementally (no integer division in the hot loop).
    // nb == 1: this set takes tiles e, e+2, ... (sub-round 0);  nb == 2: every tile, sub-round e.
    constexpr int tile_step = (NB == 1) ? 2 : 1;
    const int b = (NB == 1) ? 0 : e;
    const bool does_epilogue = (NB == 1) || e == 1;   // the set that dequantises a tile's last sub-round reads D back
    const uint32_t lane_base = static_cast<uint32_t>(32 * q) << 16;
    const int u = b * 8 + mypart;
    const uint32_t unit_off = u * kUnitWeightBytes + row * 16;
    const uint32_t sc_off = p.gps * kUnitWeightBytes + u * 32 + row * 2, z_off = p.gps * (kUnitWeightBytes + 32) + u * 16 + row;
    const int tid_set = (wi & 3) * 32 + lane;
    const int M = p.M, out_per_tile = 16 * M, mode = p.mode, nslices = p.slices;
    const uint32_t red_set = red + (e * 2) * 128 * M * 4;
    const uint32_t my_red = (mypart * 16 + row) * 4;

    int i = (NB == 1) ? e : 0;                 // current tile (local index)
    int st = i;                                // its ring stage; nst is even when NB == 1, so parity is per set
    uint32_t full_par = 0;
    int abuf = e, a_use = 0;                   // A buffer of round r = i*nb + b (r starts at e), number of earlier uses
    int i_div = 0, i_mod = i;                  // i / slices, i % slices (slices >= 1)
    while (i_mod >= nslices) { i_mod -= nslices; ++i_div; }
    int pend = -1, pend_div = 0, pend_mod = 0, pend_d = 0;
    uint32_t epi = 0;

    auto epilogue = [&]() {
      mbar_wait(bar_dfull + 8 * e, epi & 1);  // completions of dfull[e] are this set's epilogues, in order
      tc_fence_after();
      const uint32_t rb = red_set + (epi & 1) * 128 * M * 4;
      const uint32_t tcol = tmem + lane_base + d_col0 + pend_d * p.nmma + 2 * q;
#pragma unroll 1
      for (int m = 0; m < M; ++m) {
        uint32_t v0, v1;
        tc_ld2(tcol + 8 * m, v0, v1);
        tc_wait_ld();
        sts_f32(rb + m * 512 + my_red, __uint_as_float((lane >> 4) ? v1 : v0));  // red[m][part][row]
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_dfree + 8 * pend_d);
      named_bar_sync(1 + e, 128);
      if (mode == kCluster && !cluster_ready) { cluster_wait_acquire(); cluster_ready = true; }  // #1
#pragma unroll 1
      for (int o = tid_set; o < out_per_tile; o += 128) {
        float acc = 0.f;
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) acc += lds_f32(rb + (o >> 4) * 512 + (pp * 16 + (o & 15)) * 4);  // fixed order
        const int tile_g = tile_g0 + pend;
        if (mode == kDirect) {
          store_out<T>(p, acc, o >> 4, tile_g * kTileN + (o & 15));
        } else if (mode == kCluster) {
          st_cluster_f32(map_to_rank(recv + ((pend_div * nslices + slice) * out_per_tile + o) * 4, pend_mod), acc);
        } else {
          __stcg(p.partials + (static_cast<size_t>(slice) * p.tiles_total + tile_g) * out_per_tile + o, acc);
        }
      }
      ++epi;
    };

    bool first = true;
#pragma unroll 1
    for (; i < ntiles; i += tile_step) {
      mbar_wait(bar_full + 8 * st, full_par);
      if (first) { PARO_TRACE(5); first = false; }
      if (a_use > 0) mbar_wait(bar_afree + 8 * abuf, (a_use - 1) & 1);  // MMAs of the previous user have drained this A buffer
      tc_fence_after();
      const uint32_t rec = smem0 + st * p.rec_stride;
      RowDequant<T> dq;
      dq.prep(lds16(rec + sc_off), lds8(rec + z_off));
      const uint32_t wbase = rec + unit_off;
      const uint32_t ta = tmem + lane_base + abuf * 64;
      // two chunks per trip: enough independent work to hide the LDS latency, half the code of a full unroll
      // (the fully unrolled body did not fit the instruction cache next to the epilogue: stall_no_inst)
#pragma unroll 1
      for (int c = 0; c < 4; c += 2) {
        const uint4 wa = lds128(wbase + c * 256), wb = lds128(wbase + c * 256 + 256);
        uint32_t regs[16];
        dq.word(wa.x, regs + 0);
        dq.word(wa.y, regs + 4);
        dq.word(wa.z, regs + 8);
        dq.word(wa.w, regs + 12);
        tc_st16(ta + 16 * c, regs);
        dq.word(wb.x, regs + 0);
        dq.word(wb.y, regs + 4);
        dq.word(wb.z, regs + 8);
        dq.word(wb.w, regs + 12);
        tc_st16(ta + 16 * c + 16, regs);
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_afull + 8 * abuf);
        mbar_arrive(bar_empty + 8 * st);
      }
      if (pend >= 0) { epilogue(); pend = -1; }
      if (does_epilogue) { pend = i; pend_div = i_div; pend_mod = i_mod; pend_d = (ND == 2) ? (i & 1) : 0; }
      // advance the incremental state
      st += tile_step;
      if (st >= nst) { st -= nst; full_par ^= 1; }
      abuf += 2;
      if (abuf >= NA) { abuf -= NA; ++a_use; }
      i_mod += tile_step;
      while (i_mod >= nslices) { i_mod -= nslices; ++i_div; }
    }
    if (pend >= 0) epilogue();
    PARO_TRACE(6);
  }

  // ================= common tail (all threads; warps reconverge first)
  __syncwarp();
  if (p.mode == kCluster) {
    if (!cluster_ready) cluster_wait_acquire();
    cluster_arrive_release();
    cluster_wait_acquire();
    PARO_TRACE(7);
    if (warp >= 2) {
      // every CTA finishes the tiles it was sent (tile_local % slices == slice); fixed order over K-slices
      const int out_per_tile = 16 * p.M;
      const int nmine = ntiles > slice ? (ntiles - slice + p.slices - 1) / p.slices : 0;
      for (int idx = threadIdx.x - 64; idx < nmine * out_per_tile; idx += 256) {
        const int jj = idx / out_per_tile, o = idx - jj * out_per_tile;
        float acc = 0.f;
        for (int src = 0; src < p.slices; ++src) acc += lds_f32(recv + ((jj * p.slices + src) * out_per_tile + o) * 4);
        store_out<T>(p, acc, o >> 4, (tile_g0 + jj * p.slices + slice) * kTileN + (o & 15));
      }
    }
  } else if (p.mode == kWorkspace) {
    if (warp >= 2) {
      // announce my tiles; whoever completes a tile sums the slices in fixed order
      __threadfence();
      named_bar_sync(3, 256);
      const int wi = warp - 2, out_per_tile = 16 * p.M;
      for (int jj = wi; jj < ntiles; jj += 8) {
        const int tile_g = tile_g0 + jj;
        int old = 0;
        if (lane == 0) old = atomicAdd(p.counters + tile_g, 1);
        old = __shfl_sync(0xFFFFFFFFu, old, 0);
        if (old != p.slices - 1) continue;
        __threadfence();
        for (int o = lane; o < out_per_tile; o += 32) {
          float acc = 0.f;
          for (int sl0 = 0; sl0 < p.slices; sl0 += 8) {  // 8 independent loads in flight, then add in order
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
              v[k] = sl0 + k < p.slices ? __ldcg(p.partials + (static_cast<size_t>(sl0 + k) * p.tiles_total + tile_g) * out_per_tile + o) : 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) acc += v[k];
          }
          store_out<T>(p, acc, o >> 4, tile_g * kTileN + (o & 15));
        }
        if (lane == 0) p.counters[tile_g] = 0;  // leave the workspace zeroed for the next launch
      }
    }
  }
  PARO_TRACE(8);
  if (p.trace && warp == 2 && lane == 0 && blockIdx.x < kTraceMaxCtas) g_trace[blockIdx.x * kTraceSlots + 11] = globaltimer_ns();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
  }
}

int debug_trace_read(unsigned long long *host, int max_ctas) {
  const size_t n = static_cast<size_t>(max_ctas < kTraceMaxCtas ? max_ctas : kTraceMaxCtas) * kTraceSlots;
  PARO_CUDA_OK(cudaMemcpyFromSymbol(host, g_trace, n * sizeof(unsigned long long)));
  return PARO_OK;
}

// ------------------------------------------------------------------ host side
struct TcPlan {
  int ranges;
  int part_range_begin[PARO_MAX_PARTS + 1];
  int grid;
  int max_tiles_per_range;
};

// Integer number of tile ranges per partition, proportional to the partition's tile count.
static bool make_plan(const Layout &L, int ranges, TcPlan &plan) {
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
  const size_t counters = (static_cast<size_t>(L.tiles_total) * 4 + 255) / 256 * 256;
  if (L.slices <= 1) return counters;
  return counters + static_cast<size_t>(L.slices) * L.tiles_total * 16 * static_cast<size_t>(max_m) * 4;
}

static int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

template <typename T, int NB>
static int launch_tc(TcParams &p, const Layout &L, int sms, cudaStream_t stream) {
  auto kern = tc_linear_kernel<T, NB>;
  const bool want_cluster = L.plan.cluster > 1 && !env_int("PARO_NO_CLUSTER", 0);
  const int out_per_tile = 16 * p.M;
  // shared-memory carve-up: [record ring][B operand][rotation tiles, later the 8-part exchange][DSMEM receive][barriers]
  const int xb_bytes = p.nb * 8 * p.nmma * 32;
  const int rot_bytes = 8 * p.nb * p.rot_bytes, red_bytes = 4 * 128 * p.M * 4;
  const int scratch = rot_bytes > red_bytes ? rot_bytes : red_bytes;
  const int recv_budget = 40 * 1024;
  int ctas_per_sm = env_int("PARO_DECODE_CTAS_PER_SM", 2);
  int limit = ctas_per_sm >= 2 ? 110 * 1024 : 220 * 1024;
  if (xb_bytes + scratch + 2 * p.rec_stride + 4096 > limit) { ctas_per_sm = 1; limit = 220 * 1024; }
  const int ranges = sms * ctas_per_sm / L.slices;
  if (ranges < 1) {
    set_error("decode: in_features=%d needs %d K-slices, more than the %d resident CTAs", L.K, L.slices, sms * ctas_per_sm);
    return PARO_EUNSUPPORTED;
  }
  TcPlan plan;
  for (int attempt = 0; attempt < 2; ++attempt) {
    if (!make_plan(L, ranges, plan)) { set_error("decode: empty plan"); return PARO_EINVAL; }
    p.mode = L.slices == 1 ? kDirect : kWorkspace;
    p.recv_tiles = 0;
    if (want_cluster && attempt == 0) {
      const int recv_tiles = (plan.max_tiles_per_range + L.slices - 1) / L.slices;
      if (recv_tiles * L.slices * out_per_tile * 4 <= recv_budget) {
        p.mode = kCluster;
        p.recv_tiles = recv_tiles;
      }
    }
    const int recv_bytes = p.recv_tiles * L.slices * out_per_tile * 4;
    const int tail = (recv_bytes + 15) / 16 * 16 + 16 * kMaxStages + 128;
    int nst = p.nstages;
    while (nst > 2 && nst * p.rec_stride + xb_bytes + scratch + tail > limit) --nst;
    if (p.nb == 1 && (nst & 1)) --nst;  // a stage must always be consumed by the same worker set (mbarrier parity)
    int off = nst * p.rec_stride;
    p.xb_off = off;  off += xb_bytes;
    p.rot_off = off; p.red_off = off; off += scratch;   // the exchange buffers reuse the rotation tiles (dead after the prologue)
    p.recv_off = off;
    p.bar_off = (off + recv_bytes + 15) / 16 * 16;
    const size_t smem = p.bar_off + 16 * kMaxStages + 128;
    if (static_cast<int>(smem) > limit) { set_error("decode: shared-memory footprint %zu too large", smem); return PARO_EUNSUPPORTED; }
    TcParams q = p;
    q.nstages = nst;
    PARO_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(plan.grid);
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (q.mode == kCluster) {
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
        if ((plan.max_tiles_per_range + L.slices - 1) / L.slices > q.recv_tiles) continue;  // receive buffer too small now
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
    for (int i = 0; i <= PARO_MAX_PARTS; ++i) q.part_range_begin[i] = plan.part_range_begin[i];
    PARO_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, q));
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
  TcParams p;
  p.packed = static_cast<const uint8_t *>(packed);
  p.x = x; p.y = y; p.bias = bias;
  const size_t counters = (static_cast<size_t>(L.tiles_total) * 4 + 255) / 256 * 256;
  p.counters = static_cast<int *>(workspace);
  p.partials = reinterpret_cast<float *>(static_cast<uint8_t *>(workspace) + counters);
  p.M = static_cast<int>(M); p.K = L.K; p.N = L.N;
  p.n_parts = L.n_parts; p.slices = L.slices; p.groups = L.groups; p.krot = L.krot; p.tiles_total = L.tiles_total;
  p.gps = L.gps; p.nb = L.gps / 8; p.rec_bytes = L.rec_bytes;
  p.rec_stride = (L.rec_bytes + 127) / 128 * 128;
  p.mpad = M <= 2 ? 2 : M <= 4 ? 4 : M <= 8 ? 8 : 16;
  p.nmma = 8 * p.mpad;
  // TMEM (256 columns per CTA, two CTAs per SM): NA x 64 columns of A + ND x nmma columns of D
  p.na = env_int("PARO_DECODE_ABUFS", p.nmma <= 64 ? 3 : 2);
  if (p.na < 2) p.na = 2;
  if (p.na > 3) p.na = 3;
  p.nd = (kTmemCols - 64 * p.na) / p.nmma >= 2 ? 2 : 1;
  int nst = env_int("PARO_DECODE_STAGES", L.gps == 8 ? 6 : 4);
  if (nst < 2) nst = 2;
  if (nst > kMaxStages) nst = kMaxStages;
  p.nstages = nst;
  p.trace = env_int("PARO_DECODE_TRACE", 0);
  p.rot_bytes = kGroup * 2 * (M == 1 ? 1 : M == 2 ? 2 : M <= 4 ? 4 : M <= 8 ? 8 : 16);
  for (int i = 0; i <= PARO_MAX_PARTS; ++i) p.part_tile_begin[i] = L.part_tile_begin[i];
  p.meta_group_bytes = L.meta_group_bytes;
  p.meta_off = static_cast<long long>(L.meta_off);
  p.rec_off = static_cast<long long>(L.rec_off);
  if (s.dtype == PARO_BF16) return p.nb == 1 ? launch_tc<__nv_bfloat16, 1>(p, L, sms, stream) : launch_tc<__nv_bfloat16, 2>(p, L, sms, stream);
  return p.nb == 1 ? launch_tc<__half, 1>(p, L, sms, stream) : launch_tc<__half, 2>(p, L, sms, stream);
}

}  // namespace paro
