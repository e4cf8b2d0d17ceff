// Host-side structures shared by the C ABI, the plan builder and the kernels.
#pragma once

#include <complex>
#include <cstdint>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ffsim_b200.h"
#include "device_structs.h"

namespace ffb {

using cplx = std::complex<double>;

constexpr int kMaxNorb = 32;      // device kernels hold strings in 32 bits
constexpr int kMaxSubWindow = 6;  // widest register block: C(6,3) = 20 amplitudes
constexpr int kMaxLow = 16;       // distinct "electrons below the sub-window" counts

void set_error(const std::string &msg);
int fail(int code, const std::string &msg);

// binom(n, k) for 0 <= n <= 64 (0 outside the triangle)
uint64_t binom(int n, int k);
// colexicographic rank of `s` among strings of the same popcount (= its address)
uint64_t rank_of(uint64_t s);
// rank-th (0-based) string with `nocc` bits set, ascending order
uint64_t unrank(uint64_t rank, int nocc);
// next larger integer with the same popcount (v != 0)
inline uint64_t next_same_popcount(uint64_t v) {
  uint64_t t = v | (v - 1);
  return (t + 1) | (((~t & -~t) - 1) >> (__builtin_ctzll(v) + 1));
}
// all strings of `nbits` bits with `nocc` set, ascending
std::vector<uint64_t> strings_of(int nbits, int nocc);

// A rotation with the orbital pair normalised to (q, q+1); "x" is the string
// with q occupied and q+1 empty:  x' = c x + s y,  y' = c y - conj(s) x.
struct NormRot {
  int q;
  double c;
  cplx s;
};
NormRot normalise(const ffb_givens_rotation &r);

// Level-2 schedule entry: rotations [rot_begin, rot_end) of the pass lie in the
// register sub-window [q0, q0 + w) (positions relative to the pass window).
struct SubPass {
  int q0, w;
  int rot_begin, rot_end;
};
// Level-1 schedule entry: one sweep over the state with orbitals [lo, lo + W)
// active inside a shared-memory tile.
struct PassSchedule {
  int lo = 0, W = 0;
  std::vector<int> rot_index;  // indices into the side's rotation list, in order
  std::vector<SubPass> subs;   // ranges refer to positions in rot_index
};
struct SideSchedule {
  int norb = 0, nocc = 0;
  std::vector<PassSchedule> passes;
};

struct PlanOptions {
  int64_t smem_bytes;  // shared-memory budget of one tile (bytes)
  int min_cols;        // smallest column strip that may define the window width
  int max_cols;        // widest column strip of a full-height tile
  int sub_window;      // register block width (2..6)
  int threads;         // CTA size of the fused pass kernel
  int bulk_copies;     // 1: contiguous tile columns move as TMA bulk copies (cp.async.bulk + mbarrier), 0: 16-byte copies
  int beta_mode;       // 0 auto, 1 native strided, 2 transposed copy (transpositions folded into the first / last
                       // pass where possible), 3 transposed copy with separate transpose kernels
};
PlanOptions current_options();

// Widest window whose tallest tile still fits `min_cols` columns.
int max_window(int norb, int nocc, const PlanOptions &opt);
// Greedy two-level tiling of an adjacent-pair rotation sequence.
SideSchedule build_schedule(int norb, int nocc, const std::vector<int> &q, const PlanOptions &opt);

// Index tables of one pass, in the form the fused kernel consumes.
struct GroupSubHost {
  int seg_mp[kMaxSubWindow];      // m' of each segment, in decreasing cost order (0 = unused)
  int seg_begin[kMaxSubWindow];   // first block of the segment inside `blocks`
  int seg_count[kMaxSubWindow];   // blocks in the segment
  int n_seg = 0;
  std::vector<uint32_t> blocks;   // base row | (l' << 24)
};
struct PassGroupHost {
  int m = 0;      // electrons inside the window
  int R = 0;      // tile rows = C(W, m)
  int l_min = 0;  // smallest count of electrons below the window
  int n_low = 0;  // number of distinct counts
  std::vector<uint32_t> tabrow;      // [n_low][R] row offset relative to the combo base
  std::vector<uint32_t> combo_base;  // [n_combos] base row of the (H, L) combination
  std::vector<uint8_t> combo_low;    // [n_combos] l - l_min
  std::vector<GroupSubHost> subs;    // one per sub-pass
  bool has_blocks = false;
};
struct PassTablesHost {
  int lo = 0, W = 0;
  std::vector<PassGroupHost> groups;          // decreasing R
  std::vector<uint16_t> off;                  // [n_sub][kMaxLow][kOffRow]
};
PassTablesHost build_pass_tables(int norb, int nocc, const PassSchedule &pass);

// The same tables in the form fused_pass_kernel reads: one u32 pool (tile row tables, combination
// bases, 16-byte aligned register-block lists), the u8 pool (combination -> row table), the segment
// descriptors of every (group, sub-pass) and the block-offset tables as byte offsets with every class
// on a 16-byte boundary.  Consumes the block lists and row tables of `T` (they are large).
struct DevicePassTables {
  struct GroupOffsets {
    uint32_t tabrow_off, combo_base_off, combo_low_off, gsub_off;
  };
  std::vector<uint32_t> u32;
  std::vector<uint8_t> u8;
  std::vector<GroupSubDev> gsub;
  std::vector<uint32_t> off32;  // [n_sub][kMaxLowDev][kOffRowDev]
  std::vector<GroupOffsets> goff;  // one per group of `T`
  int blk_cap = 4;  // longest block list of any (group, sub-pass), rounded up to 4 entries
};
DevicePassTables pack_device_tables(const PassSchedule &pass, PassTablesHost &T);

// Dispatch units of a sub-pass: maximal runs (length <= kMaxRunLen) of consecutive rotations
// whose pair positions descend by one.  rq[] holds positions relative to the pass window;
// q_hi_rel is relative to the sub-window start q0.
struct Run {
  int first;     // index of the first rotation (into the pass's rotation arrays)
  int len;
  int q_hi_rel;  // position of the first rotation's pair inside the register block
};
std::vector<Run> segment_runs(const unsigned char *rq, int rot_begin, int rot_end, int q0);

// start of the m' class inside a row of `off`: sum_{1 <= j < m'} C(w, j)
int class_offset(int w, int mprime);

}  // namespace ffb

struct ffb_tables {
  int norb = 0, nocc = 0;
  int64_t dim = 0;
  std::vector<uint64_t> strings;  // ascending
  std::mutex mu;
  int device = -1;
  uint32_t *d_strings = nullptr;  // lazy device copy
  std::vector<void *> scratch;    // per-handle device scratch (same-spin factors, matrices)
  std::vector<size_t> scratch_bytes;
};
