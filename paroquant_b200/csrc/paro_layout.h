// The streaming layout paro_prepack writes and both fused kernels read.  Host + device.
//
//   packed := [ rotation metadata ][ weight records ]
//
// Rotation metadata, one block per (partition p, group gk), groups of 128 channels:
//   [r = 0..krot-1][128] uint8   pair indices: (i, j) of pair t at bytes 2t, 2t+1
//   [r = 0..krot-1][ 64] T       theta, already cast to the activation dtype T (rotation.cu:75)
//   [128]                T       channel scales, cast to T (rotation.cu:76-78)
//   -> krot*256 + 256 bytes; lane l of the warp that owns the group reads 4 bytes of each row.
//
// Weight records.  K is cut into `slices` of `gps` groups (gps*128 channels; the plan below picks
// gps from K so that one slice is what ONE CTA of the small-M kernel owns), N into tiles of 16
// columns.  One record = one (slice, tile) = 16 x gps*128 INT4 weights + their scales / zeros:
//   [u = 0..gps-1][kh = 0..1][lane = 0..31] 16 bytes = 4 words, word j covers the 16 k values
//        kb = (slice*gps + u)*128 + (kh*4 + j)*16 .. +15  for the two columns n0+g, n0+g+8
//        (g = lane/4, t = lane%4) in exactly the register layout of the A operand of
//        mma.m16n8k16 (rows = output columns n, cols = k):
//          bits  0..3   W[kb+2t  ][n0+g]      bits 16..19  W[kb+2t+1][n0+g]
//          bits  4..7   W[kb+2t  ][n0+g+8]    bits 20..23  W[kb+2t+1][n0+g+8]
//          bits  8..11  W[kb+2t+8][n0+g]      bits 24..27  W[kb+2t+9][n0+g]
//          bits 12..15  W[kb+2t+8][n0+g+8]    bits 28..31  W[kb+2t+9][n0+g+8]
//   [u][g = 0..7][2] T      scales  s[slice*gps+u][n0+g], s[..][n0+g+8]            (gps*32 bytes)
//   [u][g = 0..7][2] uint8  zeros   z[slice*gps+u][n0+g], z[..][n0+g+8]            (gps*16 bytes)
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

// How the small-M kernel cuts K (a function of K only, so it is part of the layout):
//   cluster = CTAs that share one range of output tiles and split K between them (thread-block
//             cluster; partial sums meet through distributed shared memory).  0 = no such
//             factorisation exists: slices are reduced through a global workspace instead.
//   warps   = consumer warps per CTA, gpw = groups per warp  ->  gps = warps * gpw
struct SlicePlan {
  int cluster, warps, gpw, gps, slices;
};

inline SlicePlan choose_slice_plan(int groups) {
  SlicePlan best = {0, 0, 0, 0, 0};
  int best_score = -1;
  const int cl[4] = {8, 4, 2, 1};
  for (int ci = 0; ci < 4; ++ci)
    for (int gpw = 1; gpw <= 2; ++gpw) {
      const int c = cl[ci];
      if (groups % (c * gpw)) continue;
      const int w = groups / (c * gpw);
      if (w < 4 || w > 8) continue;
      const int score = w * 100 + (3 - gpw) * 10 + c;  // more warps, then fewer groups per warp, then wider cluster
      if (score > best_score) { best_score = score; best = {c, w, gpw, w * gpw, c}; }
    }
  if (best_score < 0) {  // e.g. K = 11008 (86 groups): 8-group slices, ragged tail, workspace reduction
    const int gps = groups >= 8 ? 8 : (groups >= 4 ? 4 : groups);
    best = {0, gps, 1, gps, (groups + gps - 1) / gps};
  }
  const char *e = getenv("PARO_SLICE_PLAN");  // "cluster,warps,gpw" -- experiments only
  if (e && *e) {
    int c = 0, w = 0, g = 0;
    if (sscanf(e, "%d,%d,%d", &c, &w, &g) == 3 && w >= 1 && w <= 8 && g >= 1 && g <= 2) {
      const int gps = w * g, s = (groups + gps - 1) / gps;
      best = {(c >= 1 && c <= 8 && s == c && groups % gps == 0) ? c : 0, w, g, gps, s};
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
  size_t meta_off, rec_off, total_bytes;

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
  L.total_bytes = L.rec_off + static_cast<size_t>(L.slices) * L.tiles_total * L.rec_bytes;
  return true;
}

}  // namespace paro
