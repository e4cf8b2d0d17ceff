// POD structures shared between the host plan builder and the CUDA kernels.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define FFB_HD __host__ __device__
#else
#define FFB_HD
#endif

namespace ffb {

constexpr int kMaxGroups = 33;       // distinct electron counts inside a window
constexpr int kMaxRotPerPass = 512;  // >= 32*31/2
constexpr int kMaxSubPerPass = 96;
constexpr int kMaxLowDev = 16;       // distinct electron counts below a register block
constexpr int kOffRow = 64;          // u16 entries per row of the block-offset table (2^6 - 2 used)
constexpr int kMaxSeg = 5;           // classes m' = 1..w-1 of a 6-wide register block

// One class of register blocks inside a (group, sub-pass)
struct SegDev {
  int mp;     // electrons inside the register block
  int begin;  // first block (index into the group's block list for that sub-pass)
  int count;  // number of blocks
};
struct GroupSubDev {
  uint32_t blocks_off;  // into PassParams::u32
  int n_seg;
  SegDev seg[kMaxSeg];
};
struct GroupLaunch {
  int R;                    // tile rows
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
  void *data;  // complex128 state (or transposed workspace), updated in place
  long long row_stride;  // element stride between consecutive string addresses
  long long col_stride;  // element stride between consecutive batch columns
  long long n_cols;      // batch columns
  const void *rowphase;  // complex128[dim] multiplied into every row on store, or NULL
  const uint32_t *u32;
  const uint8_t *u8;
  const GroupSubDev *gsub;
  const uint16_t *off;   // [n_sub][kMaxLow][kOffRow]
  int n_groups, n_sub, n_rot, w;
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
