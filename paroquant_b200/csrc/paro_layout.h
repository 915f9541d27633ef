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
// Weight records.  N is cut into BLOCKS of 128 output columns per partition (a last partial block is
// zero-filled), K into the quantisation groups of 128 channels.  One record = one (block, group) =
// exactly what one dequant ROUND of either fused kernel consumes, and ONE cp.async.bulk (TMA) copy:
//   weights  [t = 0..7][c = 0..3][row = 0..15] 16 bytes = 4 words; word j holds the 8 INT4 weights
//            W[kb .. kb+7][n], n = 128*block + 16t + row, kb = 128*g + 32c + 8j, at bit positions
//            0,16,4,20,8,24,12,28 -- so (w >> 4i) & 0x000F000F is the pair (kb+2i, kb+2i+1) sitting in
//            the low mantissa bits of a bf16x2 / half2 word, i.e. the payload of one 32-bit TMEM column
//            of the tcgen05 A operand (lane = output column n, TMEM column = k/2).
//            One thread dequantises one COLUMN: 64 contiguous-in-k bytes per record, read as 4 x 16 B
//            with the 16 rows of a chunk adjacent (conflict-free 128-bit shared loads).   (8192 bytes)
//   scales   [128] T       s[g][n]                                                       (256 bytes)
//   zeros    [128] uint8   z[g][n]                                                       (128 bytes)
// -> 8576 bytes.  group_size 64 (the converter rotates AND quantises in groups of `group_size`, cli/convert.py:176-182):
// a record still covers 128 channels = TWO groups; it carries both groups' scales ([2][128] T at 8192) and then both
// groups' zeros ([2][128] uint8) -> 8960 bytes; the pair indices of the second group are stored + 64, which makes the two
// 64-channel rotations one 128-channel rotation of the same arithmetic (pairs never cross the halves).
// Records are ordered partition-major, then block, then group: the groups one CTA streams for a block are one
// contiguous byte range.  (Measured: a bulk copy costs ~90 SM cycles of
// issue whatever its size, so 8 x 1 KB pieces per round capped the stream at ~11 B/clk/SM.)
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
constexpr int kBlockN = 128;               // output columns per block = TMEM lanes of one tcgen05.mma
constexpr int kBlockWeightBytes = 8192;    // 128 columns x 128 channels of INT4
constexpr int kBlockScaleOff = 8192;        // [qhalves][128] T, then [qhalves][128] uint8 zeros
constexpr int kBlockBytes = 8192 + 256 + 128;      // group_size 128: one scale / zero set per record
constexpr int kBlockBytes64 = 8192 + 512 + 256;    // group_size 64: two
constexpr int kBlockBytesMax = kBlockBytes64;
PARO_HD constexpr int block_bytes(int qhalves) { return kBlockWeightBytes + qhalves * 384; }
PARO_HD constexpr int block_zero_off(int qhalves) { return kBlockScaleOff + qhalves * 256; }

struct Layout {
  int K, N, krot, n_parts, dtype;
  int groups;                               // K / 128: 128-channel record groups (each holds `qhalves` quantisation groups)
  int qhalves, rec_bytes;                   // group_size 128: 1, 8576;  group_size 64: 2, 8960
  int blocks_total;                         // sum over partitions of ceil(size / 128)
  int part_col_begin[PARO_MAX_PARTS + 1];   // first output column of each partition
  int part_block_begin[PARO_MAX_PARTS + 1]; // cumulative blocks per partition
  int meta_group_bytes;                     // krot*256 + 256
  size_t meta_off, rec_off, raw_off, raw_part_bytes, total_bytes;  // raw_*: rotation metadata again, in torch.ops.rotation.rotate's own format

  PARO_HD size_t meta_offset(int part, int gk) const {
    return meta_off + (static_cast<size_t>(part) * groups + gk) * meta_group_bytes;
  }
  // record of (block index over all partitions, group)
  PARO_HD size_t record_offset(int block, int g) const {
    return rec_off + (static_cast<size_t>(block) * groups + g) * rec_bytes;
  }
};

// Returns false (and leaves `why` pointing at a static message) on an unsupported shape.
inline bool make_layout(const paro_linear_shape &s, Layout &L, const char **why) {
  static const char *msgs[] = {
      "group_size must be 64 or 128",
      "in_features must be a positive multiple of 128",
      "n_parts must be in 1..8",
      "every partition size must be a positive multiple of 16",
      "sum(part_sizes) != out_features",
      "krot must be in 1..16",
      "dtype must be PARO_F16 or PARO_BF16",
  };
  if (s.group_size != kGroup && s.group_size != 64) { *why = msgs[0]; return false; }
  if (s.in_features <= 0 || s.in_features % kGroup) { *why = msgs[1]; return false; }
  if (s.n_parts < 1 || s.n_parts > PARO_MAX_PARTS) { *why = msgs[2]; return false; }
  if (s.krot < 1 || s.krot > 16) { *why = msgs[5]; return false; }
  if (s.dtype != PARO_F16 && s.dtype != PARO_BF16) { *why = msgs[6]; return false; }
  L.K = s.in_features; L.N = s.out_features; L.krot = s.krot; L.n_parts = s.n_parts; L.dtype = s.dtype;
  L.groups = L.K / kGroup;
  L.qhalves = kGroup / s.group_size;
  L.rec_bytes = block_bytes(L.qhalves);
  int n = 0, nb = 0;
  L.part_col_begin[0] = 0;
  L.part_block_begin[0] = 0;
  for (int p = 0; p < s.n_parts; ++p) {
    if (s.part_sizes[p] <= 0 || s.part_sizes[p] % 16) { *why = msgs[3]; return false; }
    n += s.part_sizes[p];
    nb += (s.part_sizes[p] + kBlockN - 1) / kBlockN;
    L.part_col_begin[p + 1] = n;
    L.part_block_begin[p + 1] = nb;
  }
  for (int p = s.n_parts; p < PARO_MAX_PARTS; ++p) {
    L.part_col_begin[p + 1] = n;
    L.part_block_begin[p + 1] = nb;
  }
  if (n != s.out_features) { *why = msgs[4]; return false; }
  L.blocks_total = nb;
  L.meta_group_bytes = s.krot * 256 + 256;
  L.meta_off = 0;
  size_t meta = static_cast<size_t>(s.n_parts) * L.groups * L.meta_group_bytes;
  L.rec_off = (meta + 127) / 128 * 128;
  L.raw_off = (L.rec_off + static_cast<size_t>(L.blocks_total) * L.groups * L.rec_bytes + 127) / 128 * 128;
  // per partition: pairs int16 [krot][K], theta T [krot][K/2], channel scales T [K] (the large-M pre-pass reads these)
  L.raw_part_bytes = (static_cast<size_t>(s.krot) * L.K * 3 + static_cast<size_t>(L.K) * 2 + 127) / 128 * 128;
  L.total_bytes = L.raw_off + L.raw_part_bytes * s.n_parts;
  return true;
}

}  // namespace paro
