// The streaming layout paro_prepack writes and both fused kernels read.  Host + device.
//
//   packed := [ rotation metadata ][ weight records ][ rotation metadata, reference format ]
//
// Rotation metadata, one block per (partition p, group gk), groups of 128 channels:
//   [r = 0..krot-1][128] uint8   pair indices: (i, j) of pair t at bytes 2t, 2t+1
//   [r = 0..krot-1][ 64] T       theta, already cast to the activation dtype T (rotation.cu:75)
//   [128]                T       channel scales, cast to T (rotation.cu:76-78)
//   -> krot*256 + 256 bytes; lane l of the warp that owns the group reads 4 bytes of each row.
//
// Weight records.  K is cut into `slices` of `gps` groups (gps*128 channels; the plan below picks
// gps from K so that one slice is what ONE CTA of the small-M kernel owns), N into tiles of 16
// output columns.  One record = one (slice, tile); it holds gps UNITS, a unit = 16 columns x one
// group of 128 channels:
//   weights  [u = 0..gps-1][c = 0..3][row = 0..15] 16 bytes = 4 words; word j holds the 8 INT4 weights
//            W[kb .. kb+7][n0 + row], kb = (slice*gps + u)*128 + 32c + 8j, at bit positions
//            0,16,4,20,8,24,12,28 -- so (w >> 4i) & 0x000F000F is the pair (kb+2i, kb+2i+1) sitting in
//            the low mantissa bits of a bf16x2 / half2 word, i.e. the payload of one 32-bit TMEM column
//            of the tcgen05 A operand (lane = output column n, TMEM column = k/2).
//            One thread dequantises one ROW: 64 contiguous-in-k bytes per unit, read as 4 x 16 B with
//            the 16 rows of a chunk adjacent (conflict-free 128-bit shared loads).
//   scales   [u][row] T       s[slice*gps+u][n0+row]                                (gps*32 bytes)
//   zeros    [u][row] uint8   z[slice*gps+u][n0+row]                                (gps*16 bytes)
// -> gps * 1072 bytes.  Records are ordered partition-major, then slice, then tile, so the tiles
// one CTA streams are one contiguous byte range.  A last slice with fewer groups is zero-filled.
#pragma once

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/paro_b200.h"

#if defined(__CUDACC__)
#define PARO_HD __host__ __device__
#else
#define PARO_HD
#endif

namespace paro {

constexpr int kGroup = 128;
constexpr int kTileN = 16;
constexpr int kUnitWeightBytes = 1024;  // 16 columns x 128 channels of INT4
constexpr int kUnitBytes = 1072;        // + 32 bytes of scales + 16 bytes of zeros

// How the fused kernel cuts K (a function of K only, so it is part of the layout).  A CTA owns one
// slice of gps = 8 or 16 groups: 8 groups x 16 output columns fill the 128 TMEM lanes of one
// tcgen05.mma (every lane = one (column, group) pair; the group's 16 B-operand columns select it).
//   cluster = number of slices when they fit a thread-block cluster (<= 8): the slices' partial sums
//             meet through distributed shared memory.  0 = too many slices: global workspace instead.
struct SlicePlan {
  int cluster, gps, slices;
};

inline SlicePlan choose_slice_plan(int groups) {
  SlicePlan best;
  const int s8 = (groups + 7) / 8;
  if (s8 <= 8) best = {s8, 8, s8};                                                // K <= 8192
  else if (groups % 16 == 0 && groups / 16 <= 8) best = {groups / 16, 16, groups / 16};  // e.g. 12288, 14336
  else best = {0, 8, s8};                                                         // e.g. 11008 (86 groups)
  const char *e = getenv("PARO_SLICE_PLAN");  // "gps" -- experiments only
  if (e && *e) {
    const int gps = atoi(e);
    if (gps == 8 || gps == 16) {
      const int s = (groups + gps - 1) / gps;
      best = {s <= 8 ? s : 0, gps, s};
    }
  }
  return best;
}

struct Layout {
  int K, N, krot, n_parts, dtype;
  int groups;                              // K / 128
  int gps, slices, rec_bytes;              // groups per slice, #slices, gps * 1072
  SlicePlan plan;
  int tiles_total;                         // N / 16
  int part_tile_begin[PARO_MAX_PARTS + 1]; // cumulative tiles per partition
  int meta_group_bytes;                    // krot*256 + 256
  size_t meta_off, rec_off, raw_off, raw_part_bytes, total_bytes;  // raw_*: rotation metadata again, in torch.ops.rotation.rotate's own format

  PARO_HD size_t meta_offset(int part, int gk) const {
    return meta_off + (static_cast<size_t>(part) * groups + gk) * meta_group_bytes;
  }
  PARO_HD size_t record_offset(int part, int slice, int tile_in_part) const {
    const int tp = part_tile_begin[part + 1] - part_tile_begin[part];
    return rec_off + (static_cast<size_t>(slices) * part_tile_begin[part] +
                      static_cast<size_t>(slice) * tp + tile_in_part) * rec_bytes;
  }
};

// Returns false (and leaves `why` pointing at a static message) on an unsupported shape.
inline bool make_layout(const paro_linear_shape &s, Layout &L, const char **why) {
  static const char *msgs[] = {
      "group_size must be 128 for the fused kernels",
      "in_features must be a positive multiple of 128",
      "n_parts must be in 1..8",
      "every partition size must be a positive multiple of 16",
      "sum(part_sizes) != out_features",
      "krot must be in 1..16",
      "dtype must be PARO_F16 or PARO_BF16",
  };
  if (s.group_size != kGroup) { *why = msgs[0]; return false; }
  if (s.in_features <= 0 || s.in_features % kGroup) { *why = msgs[1]; return false; }
  if (s.n_parts < 1 || s.n_parts > PARO_MAX_PARTS) { *why = msgs[2]; return false; }
  if (s.krot < 1 || s.krot > 16) { *why = msgs[5]; return false; }
  if (s.dtype != PARO_F16 && s.dtype != PARO_BF16) { *why = msgs[6]; return false; }
  L.K = s.in_features; L.N = s.out_features; L.krot = s.krot; L.n_parts = s.n_parts; L.dtype = s.dtype;
  L.groups = L.K / kGroup;
  L.plan = choose_slice_plan(L.groups);
  L.gps = L.plan.gps;
  L.slices = L.plan.slices;
  L.rec_bytes = L.gps * kUnitBytes;
  int n = 0;
  L.part_tile_begin[0] = 0;
  for (int p = 0; p < s.n_parts; ++p) {
    if (s.part_sizes[p] <= 0 || s.part_sizes[p] % kTileN) { *why = msgs[3]; return false; }
    n += s.part_sizes[p];
    L.part_tile_begin[p + 1] = n / kTileN;
  }
  for (int p = s.n_parts; p < PARO_MAX_PARTS; ++p) L.part_tile_begin[p + 1] = L.part_tile_begin[s.n_parts];
  if (n != s.out_features) { *why = msgs[4]; return false; }
  L.tiles_total = n / kTileN;
  L.meta_group_bytes = s.krot * 256 + 256;
  L.meta_off = 0;
  size_t meta = static_cast<size_t>(s.n_parts) * L.groups * L.meta_group_bytes;
  L.rec_off = (meta + 127) / 128 * 128;
  L.raw_off = (L.rec_off + static_cast<size_t>(L.slices) * L.tiles_total * L.rec_bytes + 127) / 128 * 128;
  // per partition: pairs int16 [krot][K], theta T [krot][K/2], channel scales T [K] (the large-M pre-pass reads these)
  L.raw_part_bytes = (static_cast<size_t>(s.krot) * L.K * 3 + static_cast<size_t>(L.K) * 2 + 127) / 128 * 128;
  L.total_bytes = L.raw_off + L.raw_part_bytes * s.n_parts;
  return true;
}

}  // namespace paro
