// POD structures shared between the host plan builder and the CUDA kernels.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define FFB_HD __host__ __device__
#else
#define FFB_HD
#endif

namespace ffb {

#ifndef FFB_TPB
#define FFB_TPB 512  // CTA size the fused kernel is compiled for (registers per thread = 64K / FFB_TPB)
#endif
constexpr int kMaxGroups = 33;       // distinct electron counts inside a window
constexpr int kMaxRotPerPass = 512;  // >= 32*31/2
constexpr int kMaxSubPerPass = 96;
constexpr int kMaxLowDev = 16;       // distinct electron counts below a register block
constexpr int kOffRow = 64;          // u16 entries per row of the host block-offset table (2^6 - 2 used)
constexpr int kOffRowDev = 68;       // u32 entries per row of the device table: classes padded to 4 entries
constexpr int kMaxSeg = 5;           // classes m' = 1..w-1 of a 6-wide register block

// start of class m' inside a device row: every class starts on a 16-byte boundary so that a thread
// fetches four byte offsets with one shared-memory load
FFB_HD constexpr int dev_class_offset(int w, int mp) {
  int off = 0;
  for (int j = 1; j < mp; ++j) {
    long long c = 1;
    for (int i = 1; i <= j; ++i) c = c * (w - j + i) / i;
    off += ((int)c + 3) & ~3;
  }
  return off;
}
static_assert(dev_class_offset(6, 5) + 8 <= kOffRowDev, "device offset row too short");

// One class of register blocks inside a (group, sub-pass)
struct SegDev {
  int mp;              // electrons inside the register block
  int begin;           // first block (index into the group's block list for that sub-pass)
  int count;           // number of blocks
  unsigned inv_count;  // 0xFFFFFFFF / count + 1 (0 for count == 1)
};
struct GroupSubDev {
  uint32_t blocks_off;  // into PassParams::u32, a multiple of 4 (the list is staged with 16-byte copies)
  int n_seg;
  int n_blocks;         // length of the block list
  int pad;
  SegDev seg[kMaxSeg];
};
static_assert(sizeof(GroupSubDev) % 16 == 0, "GroupSubDev is copied in 16-byte units");
// Column stride (in elements) of the shared-memory tile, which is stored column-major.  The stride
// is chosen modulo 8 (eight 16-byte bank groups) so that the 16-byte accesses of a quarter-warp
// that walks the tile in memory order (row, then column fastest: the alpha-side load and store)
// fall in eight different bank groups: stride = cols (mod 8) for odd cols, 8 / cols for 2, 4, 8.
FFB_HD constexpr int tile_col_stride(int R, int cols) {
  const int want = (cols & 1) ? (cols & 7) : (cols == 2 ? 4 : (cols == 4 ? 2 : 1));
  return R + ((want - R) & 7);
}
// Order of the (block, column) items of a class inside its 32-item chunks.  With the column index
// fastest a quarter-warp (one shared-memory wavefront of 16-byte accesses) holds the same block in
// several tile columns -- which the column stride already places in different bank groups -- and only
// 8 / cols different blocks, so it needs far fewer blocks with distinct starting rows modulo 8 than
// with the block index fastest.  That works out for an odd column count, for 2 and 4, and for multiples
// of 8 (tile_col_stride); other even counts keep the block index fastest.
FFB_HD constexpr bool items_column_fastest(int cols) { return (cols & 1) || cols == 2 || cols == 4 || (cols & 7) == 0; }

struct GroupLaunch {
  int R;                    // tile rows
  int Rp;                   // column stride of the tile, tile_col_stride(R, cols)
  int cols;                 // tile columns (chosen at launch)
  int n_combos;
  int has_blocks;
  unsigned inv_cols;        // 0xFFFFFFFF / cols + 1
  unsigned inv_R;           // 0xFFFFFFFF / R + 1
  uint32_t tabrow_off;      // u32 [n_low][R]
  uint32_t combo_base_off;  // u32 [n_combos]
  uint32_t combo_low_off;   // u8  [n_combos]
  uint32_t gsub_off;        // GroupSubDev [n_sub]
  long long unit_begin;     // first work unit of this group
  long long n_strips;       // column strips per combo
};
struct SubMeta {
  unsigned char q0;
  unsigned char pad;
  unsigned short rot_begin, rot_end;
  unsigned short run_begin, run_end;  // dispatch units of the sub-pass (see PassParams::runcode)
};

// A run is up to kMaxRunLen consecutive rotations of a sub-pass whose pair positions descend by
// one (q, q-1, ...): the shape the Givens decomposition emits.  The kernel dispatches once per
// run to straight-line code for all of its rotations.
constexpr int kMaxRunLen = 3;
FFB_HD constexpr int run_code(int q_hi_rel, int len) { return q_hi_rel * 4 + len; }

struct PassParams {
  void *data;  // complex128 state (or transposed workspace) the tiles are read from
  long long row_stride;  // element stride between consecutive string addresses
  long long col_stride;  // element stride between consecutive batch columns
  // where the tiles are written: the same buffer and strides (in place), or -- first / last pass of a
  // beta-side rotation that works on a transposed copy -- the other layout, which folds the
  // transposition into the pass (tiles are disjoint, so out of place is as safe as in place)
  void *out;
  long long out_row_stride, out_col_stride;
  // 1 when a tile column is ONE contiguous run of the buffer it is read from / written to (string
  // stride 1 and a window that starts at orbital 0, so tile row r is string rowbase + r): such a column
  // moves as a single bulk copy of the TMA unit (cp.async.bulk + mbarrier) instead of R 16-byte copies
  int bulk_in, bulk_out;
  long long n_cols;      // batch columns
  const void *rowphase;  // complex128[dim] multiplied into every row on store, or NULL
  const uint32_t *u32;
  const uint8_t *u8;
  const GroupSubDev *gsub;
  const uint32_t *off;   // [n_sub][kMaxLowDev][kOffRowDev] byte offsets (row offset * 16)
  int n_groups, n_sub, n_rot, w;
  int blk_cap;           // u32 entries of one block-list staging buffer (longest list, multiple of 4)
  int off_rows;          // rows of a sub-pass's offset table that are staged: max q0 + 1 (<= kMaxLowDev)
  long long total_units;
  GroupLaunch g[kMaxGroups];
  SubMeta sub[kMaxSubPerPass];
  // rotation r of the pass: pair position relative to the pass window and (c, s);
  // one spare slot so that the kernel may prefetch entry r + 1
  unsigned char rq[kMaxRotPerPass + 8];
  // run k of the pass: run_code(position of its first rotation relative to the sub-window, length)
  // and the index of its first rotation
  unsigned char runcode[kMaxRotPerPass + 8];
  unsigned short runrot[kMaxRotPerPass + 8];
  double rc[kMaxRotPerPass + 1], rsr[kMaxRotPerPass + 1], rsi[kMaxRotPerPass + 1];
};

}  // namespace ffb
