// Givens-rotation kernels for sm_100a.
//
//  * fused_pass_kernel<w>: one sweep over the state.  A CTA stages a tile
//    (C(W,m) string addresses x `cols` batch columns, complex128) in shared
//    memory, applies every rotation of the pass to it in sub-passes whose
//    rotations stay inside a register block of w orbitals (up to C(6,3) = 20
//    amplitudes per thread, statically unrolled), and writes the tile back.
//    HBM traffic per pass: each amplitude read once and written once (32 B),
//    however many rotations the pass fuses.
//  * givens_single_kernel / phase_shift_kernel: one launch per rotation, the
//    arithmetic of src/gates/orbital_rotation.rs:86-104 and
//    src/gates/phase_shift.rs:18-30 -- the _lib-level entry points and the
//    correctness anchor for the fused path.
//  * row_scale_kernel, transpose_kernel: helpers.
#include <cuda_runtime.h>

#include "device_structs.h"
#include "kernels.hpp"

namespace ffb {

#ifdef FFB_DEBUG_KNOBS
// what-if timing switches (developer builds only; results are wrong when any is set):
// 1 no sub-pass barriers, 2 no global traffic, 4 no rotations, 8 no register blocks, 16 no tile store, 32 no tile load
__device__ int g_debug_knobs = 0;
#define FFB_KNOB(bit) (g_debug_knobs & (bit))
#else
#define FFB_KNOB(bit) 0
#endif

#if defined(FFB_DEBUG_TIMING)
// per-phase cycle counters (developer builds only): lane 0 of every warp accumulates clock64()
// deltas in a private shared-memory row; rows are flushed to global memory when the CTA ends
__device__ unsigned long long g_phase_cycles[16];
__shared__ unsigned long long s_phase_cycles[32][8];
#define FFB_T0() long long _t0 = clock64()
#define FFB_TACC(k)                                                                  \
  do {                                                                               \
    long long _t1 = clock64();                                                       \
    if ((threadIdx.x & 31) == 0) s_phase_cycles[threadIdx.x >> 5][k] += (unsigned long long)(_t1 - _t0); \
    _t0 = clock64();                                                                 \
  } while (0)
#define FFB_TRESTART() _t0 = clock64()
#elif defined(FFB_DEBUG_TIMELINE)
// per-warp timeline (developer builds only): for the first tile of CTA 0, lane 0 of every warp logs
// (phase, begin, end) of everything it does; scripts/timeline.py turns the log into per-scheduler
// occupancy of the FP64 and LSU phases.  Phases as in scripts/phase_timing.py.
constexpr int kTlEvents = 2048;
struct TlEvent {
  unsigned long long t0, t1;
  int phase, pad;
};
__device__ TlEvent g_tl[32][kTlEvents];
__device__ int g_tl_count[32];
__shared__ int s_tl_on;
#define FFB_T0() long long _t0 = clock64()
#define FFB_TACC(k)                                                      \
  do {                                                                   \
    long long _t1 = clock64();                                           \
    if (s_tl_on && (threadIdx.x & 31) == 0) {                            \
      const int _w = threadIdx.x >> 5, _i = g_tl_count[_w];              \
      if (_i < kTlEvents) {                                              \
        g_tl[_w][_i].t0 = (unsigned long long)_t0;                       \
        g_tl[_w][_i].t1 = (unsigned long long)_t1;                       \
        g_tl[_w][_i].phase = (k);                                        \
        g_tl_count[_w] = _i + 1;                                         \
      }                                                                  \
    }                                                                    \
    _t0 = clock64();                                                     \
  } while (0)
#define FFB_TRESTART() _t0 = clock64()
#else
#define FFB_T0()
#define FFB_TACC(k)
#define FFB_TRESTART()
#endif

// ---------------------------------------------------------------- constexpr combinatorics
__host__ __device__ constexpr int cbinom(int n, int k) {
  if (k < 0 || k > n) return 0;
  long long r = 1;
  for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
  return (int)r;
}
__host__ __device__ constexpr int cpopc(unsigned s) {
  int c = 0;
  while (s) {
    c += s & 1u;
    s >>= 1;
  }
  return c;
}
// colexicographic rank of s among strings with the same popcount
__host__ __device__ constexpr int crank(unsigned s) {
  int r = 0, idx = 0, pos = 0;
  while (s) {
    if (s & 1u) r += cbinom(pos, ++idx);
    s >>= 1;
    ++pos;
  }
  return r;
}
__host__ __device__ constexpr int cclass_offset(int w, int mp) {
  int off = 0;
  for (int j = 1; j < mp; ++j) off += cbinom(w, j);
  return off;
}

__device__ __forceinline__ unsigned dev_binom(int n, int k) {
  if (k < 0 || k > n) return 0u;
  unsigned r = 1;
  for (int i = 1; i <= k; ++i) r = r * (unsigned)(n - k + i) / (unsigned)i;
  return r;
}

// ---------------------------------------------------------------- the 2x2 update
// x' = c x + s y ; y' = c y - conj(s) x      (12 DFMA-pipe operations)
__device__ __forceinline__ void zrot(double2 &x, double2 &y, double c, double sr, double si) {
  const double xr = x.x, xi = x.y, yr = y.x, yi = y.y;
  x.x = fma(-si, yi, fma(sr, yr, c * xr));
  x.y = fma(si, yr, fma(sr, yi, c * xi));
  y.x = fma(-si, xi, fma(-sr, xr, c * yr));
  y.y = fma(si, xr, fma(-sr, xi, c * yi));
}

// 16-byte shared-memory accesses by 32-bit shared address
__device__ __forceinline__ uint4 lds_u4(unsigned addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ unsigned lds_u1(unsigned addr) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ double2 lds_z(unsigned addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_z(unsigned addr, double2 v) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}

// all (x, y) pairs of a W-orbital, M-electron register block for the rotation
// on block-relative orbitals (Q, Q+1); statically unrolled
template <int W, int M, int Q, int S = 0>
__device__ __forceinline__ void rot_block(double2 (&a)[cbinom(W, M)], double c, double sr,
                                          double si) {
  if constexpr (S < (1 << W)) {
    if constexpr (cpopc(S) == M && ((S >> Q) & 3) == 1) {
      zrot(a[crank(S)], a[crank(S ^ (3 << Q))], c, sr, si);
    }
    rot_block<W, M, Q, S + 1>(a, c, sr, si);
  }
}

// A run of LEN rotations on block-relative pairs (QHI, QHI+1), (QHI-1, QHI), ...: straight-line
// code, coefficients read from the constant bank (the pass parameters).  (A shared-memory
// coefficient table with explicit prefetch was measured 35% slower: the loads compete with the
// gathers for the LSU, the constant cache does not.)
template <int W, int M, int QHI, int LEN>
__device__ __forceinline__ void run_block(double2 (&a)[cbinom(W, M)], const PassParams &p, int r) {
  if constexpr (QHI <= W - 2 && QHI - LEN + 1 >= 0) {
    // coefficients are fetched rotation by rotation (6 registers live, not 6 * LEN): the widest
    // register block leaves no room for more
    {
      const double c = p.rc[r], sr = p.rsr[r], si = p.rsi[r];
      rot_block<W, M, QHI>(a, c, sr, si);
    }
    if constexpr (LEN >= 2) {
      const double c = p.rc[r + 1], sr = p.rsr[r + 1], si = p.rsi[r + 1];
      rot_block<W, M, QHI - 1>(a, c, sr, si);
    }
    if constexpr (LEN >= 3) {
      const double c = p.rc[r + 2], sr = p.rsr[r + 2], si = p.rsi[r + 2];
      rot_block<W, M, QHI - 2>(a, c, sr, si);
    }
  }
}

#define FFB_RUN_CASE(QHI, LEN) \
  case run_code(QHI, LEN):     \
    run_block<W, M, QHI, LEN>(a, p, r); \
    break;

template <int W, int M>
__device__ __forceinline__ void run_dispatch(double2 (&a)[cbinom(W, M)], const PassParams &p, int code,
                                             int r) {
  static_assert(kMaxRunLen == 3, "run_dispatch enumerates run lengths 1..3");
  switch (code) {
    FFB_RUN_CASE(0, 1)
    FFB_RUN_CASE(1, 1) FFB_RUN_CASE(1, 2)
    FFB_RUN_CASE(2, 1) FFB_RUN_CASE(2, 2) FFB_RUN_CASE(2, 3)
    FFB_RUN_CASE(3, 1) FFB_RUN_CASE(3, 2) FFB_RUN_CASE(3, 3)
    FFB_RUN_CASE(4, 1) FFB_RUN_CASE(4, 2) FFB_RUN_CASE(4, 3)
    default: break;
  }
}
#undef FFB_RUN_CASE

// One register block: gather, rotate, scatter.  `a_addr` is the shared address of the block's base
// row in the item's tile column; `o_addr` the shared address of the block's byte-offset list (four
// offsets per 16-byte load; re-read for the scatter so that no address stays live across the math).
template <int W, int M>
__device__ __forceinline__ void process_item(unsigned a_addr, unsigned o_addr, const PassParams &p,
                                             int run0, int run1) {
  constexpr int N = cbinom(W, M), N4 = (N + 3) / 4;
  double2 a[N];
  FFB_T0();
#pragma unroll
  for (int t4 = 0; t4 < N4; ++t4) {
    const uint4 o = lds_u4(o_addr + 16 * t4);
    a[4 * t4] = lds_z(a_addr + o.x);
    if (4 * t4 + 1 < N) a[4 * t4 + 1] = lds_z(a_addr + o.y);
    if (4 * t4 + 2 < N) a[4 * t4 + 2] = lds_z(a_addr + o.z);
    if (4 * t4 + 3 < N) a[4 * t4 + 3] = lds_z(a_addr + o.w);
  }
  FFB_TACC(3);
  // one dispatch per run; the next run's descriptor is fetched while the current one executes
  int code = p.runcode[run0], r = p.runrot[run0];
  for (int run = run0; run < (FFB_KNOB(4) ? run0 : run1); ++run) {
    const int code_n = p.runcode[run + 1], r_n = p.runrot[run + 1];
    run_dispatch<W, M>(a, p, code, r);
    code = code_n;
    r = r_n;
  }
  FFB_TACC(4);
#pragma unroll
  for (int t4 = 0; t4 < N4; ++t4) {
    const uint4 o = lds_u4(o_addr + 16 * t4);
    sts_z(a_addr + o.x, a[4 * t4]);
    if (4 * t4 + 1 < N) sts_z(a_addr + o.y, a[4 * t4 + 1]);
    if (4 * t4 + 2 < N) sts_z(a_addr + o.z, a[4 * t4 + 2]);
    if (4 * t4 + 3 < N) sts_z(a_addr + o.w, a[4 * t4 + 3]);
  }
  FFB_TACC(5);
}

template <int W, int M = 1>
__device__ __forceinline__ void process_dispatch(int mp, unsigned a_addr, unsigned o_row_addr,
                                                 const PassParams &p, int run0, int run1) {
  if constexpr (M < W) {
    if (mp == M)
      process_item<W, M>(a_addr, o_row_addr + 4 * dev_class_offset(W, M), p, run0, run1);
    else
      process_dispatch<W, M + 1>(mp, a_addr, o_row_addr, p, run0, run1);
  }
}

constexpr int kOffTabEntries = kMaxLowDev * kOffRowDev;  // u32 entries of one sub-pass table
constexpr int kChunkRow = 8;  // u16 per sub-pass: chunk-prefix of <= kMaxSeg segments, pad, total
static_assert(kMaxSeg + 1 < kChunkRow, "chunk row too short");
static_assert(kOffTabEntries % 4 == 0, "offset tables are copied in 16-byte units");

// shared-memory prefix of the kernel: two block-offset tables, the group's per-sub-pass segment
// descriptors, their chunk prefixes and two block-list staging buffers
// (only the first off_rows = max q0 + 1 rows of a sub-pass's offset table are ever addressed)
__host__ __device__ inline size_t fused_smem_gsub_off(int off_rows) {
  return 2 * (size_t)off_rows * kOffRowDev * sizeof(uint32_t);
}
__host__ __device__ inline size_t fused_smem_cend_off(int n_sub, int off_rows) {
  return fused_smem_gsub_off(off_rows) + (size_t)n_sub * sizeof(GroupSubDev);
}
// per (sub-pass, warp) chunk lists of the balanced schedule: kWarpList chunk indices (0xFF = none)
constexpr int kWarpList = 8;
constexpr int kMaxWarps = FFB_TPB / 32;
__host__ __device__ inline size_t fused_smem_wl_off(int n_sub, int off_rows) {
  return fused_smem_cend_off(n_sub, off_rows) + (size_t)n_sub * kChunkRow * sizeof(uint16_t);
}
__host__ __device__ inline size_t fused_smem_blk_off(int n_sub, int off_rows) {
  // (+1 byte per sub-pass: "list valid" flag; rounded up to 16 bytes for the block-list copies)
  return (fused_smem_wl_off(n_sub, off_rows) + (size_t)n_sub * (kMaxWarps * kWarpList + 1) + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t fused_smem_tile_off(int n_sub, int blk_cap, int off_rows) {
  return fused_smem_blk_off(n_sub, off_rows) + 2 * (size_t)blk_cap * sizeof(uint32_t);
}

// 16-byte asynchronous global -> shared copy (LDGSTS): no register staging, no stall at issue
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// ---------------------------------------------------------------- bulk (TMA) copies and mbarrier
// A tile column that is one contiguous run of global memory (see PassParams::bulk_in / bulk_out) is
// moved by the copy engine of the SM: one cp.async.bulk per column, completion of the loads signalled
// on an mbarrier through its transaction count, the stores tracked as a bulk group.
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!done);
}
// generic-proxy accesses of shared memory before / after accesses of the same bytes by the bulk copies
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gmem_src, unsigned bytes,
                                          unsigned long long *bar) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
               "l"(gmem_src), "r"(bytes), "r"(b)
               : "memory");
}
__device__ __forceinline__ void bulk_store(void *gmem_dst, const void *smem_src, unsigned bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_src);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(s), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_store_commit_and_wait_read() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// n / d for n * d < 2^32 with a precomputed inv = 0xFFFFFFFF / d + 1 (which wraps to 0 for d == 1)
__device__ __forceinline__ int fast_div(int n, unsigned inv) {
  return inv ? (int)__umulhi((unsigned)n, inv) : n;
}

// One 32-item chunk of register blocks, as seen by one lane.
struct ChunkWork {
  uint32_t entry;
  int mp;   // class (electrons in the register block); 0 = nothing to do for this lane
  int col;
};

// The g-th chunk of the concatenated (heavy classes first) chunk list of a sub-pass.  `cend` is the
// sub-pass's row of the chunk-prefix table: cend[k] = chunks in segments 0..k (0xFFFF past the last
// segment), cend[kChunkRow-1] = total.  Items of a segment are (column, block) pairs, ordered as
// items_column_fastest() says.  Everything comes from
// shared memory (the block list of the sub-pass is staged there one sub-pass ahead).
__device__ __forceinline__ ChunkWork fetch_chunk(const GroupSubDev &gs, const uint16_t *cend,
                                                 const uint32_t *blk, int g, int lane, int cols, unsigned inv_cols) {
  ChunkWork w;
  w.entry = 0;
  w.mp = 0;
  w.col = 0;
  const uint4 pk = *reinterpret_cast<const uint4 *>(cend);
  const int c0 = pk.x & 0xFFFF, c1 = pk.x >> 16, c2 = pk.y & 0xFFFF, c3 = pk.y >> 16;
  static_assert(kMaxSeg == 5, "fetch_chunk unpacks five segment prefixes");
  if (g >= (int)(pk.w >> 16)) return w;
  int sg = 0, base = 0;
  if (g >= c0) sg = 1, base = c0;
  if (g >= c1) sg = 2, base = c1;
  if (g >= c2) sg = 3, base = c2;
  if (g >= c3) sg = 4, base = c3;
  const uint4 sq = *reinterpret_cast<const uint4 *>(&gs.seg[sg]);  // mp, begin, count, inv_count
  const int item = ((g - base) << 5) + lane;
  const int count = (int)sq.z;
  if (item < count * cols) {
    // (column, block) of the item: the column index runs fastest where that keeps a quarter-warp's
    // rows in different bank groups (items_column_fastest), else the block index
    int col, b;
    if (items_column_fastest(cols)) {
      b = fast_div(item, inv_cols);
      col = item - b * cols;
    } else {
      col = fast_div(item, sq.w);
      b = item - col * count;
    }
    w.col = col;
    w.mp = (int)sq.x;
    w.entry = blk[(int)sq.y + b];
  }
  return w;
}

// FFB_TPB (device_structs.h): the CTA size the kernel is compiled for; smaller CTAs may be launched
constexpr int kTilePerThread = (14336 + FFB_TPB - 1) / FFB_TPB;  // >= 220 KB / 16 B / CTA size (28 at 512 threads)
static_assert(kTilePerThread % 2 == 0, "the tile store walks a thread's elements in two halves");

template <int W>
__global__ void __launch_bounds__(FFB_TPB, 1)
    fused_pass_kernel(const __grid_constant__ PassParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int off_entries = p.off_rows * kOffRowDev;  // u32 entries of one staged offset table
  uint32_t *offbuf = reinterpret_cast<uint32_t *>(smem_raw);  // 2 x off_entries
  GroupSubDev *gsub_s = reinterpret_cast<GroupSubDev *>(smem_raw + fused_smem_gsub_off(p.off_rows));
  uint16_t *cend_s = reinterpret_cast<uint16_t *>(smem_raw + fused_smem_cend_off(p.n_sub, p.off_rows));
  uint8_t *wl_s = smem_raw + fused_smem_wl_off(p.n_sub, p.off_rows);     // [n_sub][kMaxWarps][kWarpList]
  uint8_t *wl_ok_s = wl_s + (size_t)p.n_sub * kMaxWarps * kWarpList;      // [n_sub] 1 = balanced list built
  uint32_t *blkbuf = reinterpret_cast<uint32_t *>(smem_raw + fused_smem_blk_off(p.n_sub, p.off_rows));  // 2 x blk_cap
  double2 *tile = reinterpret_cast<double2 *>(smem_raw + fused_smem_tile_off(p.n_sub, p.blk_cap, p.off_rows));
  const unsigned tile_sa = (unsigned)__cvta_generic_to_shared(tile);
  const unsigned off_sa = (unsigned)__cvta_generic_to_shared(offbuf);

  const int tid = threadIdx.x, nthr = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
  __shared__ __align__(8) unsigned long long tile_bar;  // completion of a tile's bulk loads
  unsigned tile_phase = 0;
  if (p.bulk_in) {
    if (tid == 0) {
      mbar_init(&tile_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
#ifdef FFB_DEBUG_TIMING
  if (tid < 32 * 8) s_phase_cycles[tid >> 3][tid & 7] = 0;
  __syncthreads();
#endif
#ifdef FFB_DEBUG_TIMELINE
  if (tid == 0) s_tl_on = blockIdx.x == 0;
  __syncthreads();
#endif
  int cached_group = -1;

  for (long long unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
    int gi = 0;
    while (gi + 1 < p.n_groups && unit >= p.g[gi + 1].unit_begin) ++gi;
    const GroupLaunch &G = p.g[gi];
    const int cols = G.cols;
    const int Rp = G.Rp;  // column stride of the tile in elements (see tile_col_stride)

    FFB_T0();
    {
      // ---- load the tile, column-major: element (row r, column j) at tile[j * Rp + r].  (The tile
      // geometry is decoded here and again before the store, in a scope of its own, so that none of
      // it stays in registers across the sub-passes.)
      double2 *__restrict__ data = reinterpret_cast<double2 *>(p.data);
      const long long local = unit - G.unit_begin;
      const long long combo = local / G.n_strips;
      const long long strip = local - combo * G.n_strips;
      const int R = G.R;
      const unsigned inv_R = G.inv_R, inv_cols = G.inv_cols;
      const long long col0 = strip * cols;
      const int ncv = (int)min((long long)cols, p.n_cols - col0);
      const uint32_t rowbase = p.u32[G.combo_base_off + combo];
      const uint32_t *__restrict__ tab = p.u32 + G.tabrow_off + (size_t)p.u8[G.combo_low_off + combo] * R;
      const int n_el = R * cols;
      const bool row_major = p.col_stride == 1;  // batch index contiguous in memory
      if (p.bulk_in && !FFB_KNOB(2) && !FFB_KNOB(32)) {
        // a column of the tile is the contiguous run [rowbase, rowbase + R) of its batch column: one bulk
        // copy each, issued by the first ncv threads; thread 0 posts the byte count the barrier waits for
        // (the previous tile's shared-memory reads were generic: order them before the copy engine's writes)
        fence_proxy_async();
        if (tid == 0) mbar_expect_tx(&tile_bar, (unsigned)(ncv * R * (int)sizeof(double2)));
        if (tid < ncv)
          bulk_load(tile + tid * Rp, data + (long long)rowbase + (col0 + tid) * p.col_stride,
                    (unsigned)(R * (int)sizeof(double2)), &tile_bar);
        for (int e = ncv * R + tid; e < n_el; e += nthr) {  // columns past the edge of the batch
          const int j = fast_div(e, inv_R);
          tile[j * Rp + (e - j * R)] = make_double2(0.0, 0.0);
        }
      } else
      // Each element needs its row offset from the (global) tile row table first.  All table loads
      // of a thread (a 220 KB tile is <= 28 elements per thread at 512 threads) are issued before the
      // first copy, and the copies are asynchronous (LDGSTS): the whole tile costs two memory
      // latencies, not two per batch.
      for (int e_base = 0; e_base < n_el; e_base += kTilePerThread * nthr) {
        uint32_t trow[kTilePerThread];
#pragma unroll
        for (int k = 0; k < kTilePerThread; ++k) {
          const int e = e_base + tid + k * nthr;
          const int r = row_major ? fast_div(e, inv_cols) : e - fast_div(e, inv_R) * R;
          trow[k] = e < n_el ? tab[r] : 0u;
        }
#pragma unroll
        for (int k = 0; k < kTilePerThread; ++k) {
          const int e = e_base + tid + k * nthr;
          if (e < n_el) {
            int r, j;
            if (row_major) {
              r = fast_div(e, inv_cols);
              j = e - r * cols;
            } else {
              j = fast_div(e, inv_R);
              r = e - j * R;
            }
            double2 *dst = tile + j * Rp + r;
            if (j < ncv && !FFB_KNOB(2) && !FFB_KNOB(32))
              cp_async16(dst, data + (long long)(rowbase + trow[k]) * p.row_stride + (col0 + j) * p.col_stride);
            else
              *dst = make_double2(0.0, 0.0);
          }
        }
      }
    }
    FFB_TACC(0);
    const bool work = G.has_blocks && p.n_sub > 0;
    if (work) {
      // first sub-pass: block-offset table and block list
      for (int e = tid; e < off_entries / 4; e += nthr) cp_async16(offbuf + 4 * e, p.off + 4 * e);
      {
        const GroupSubDev &g0 = p.gsub[G.gsub_off];
        const uint32_t *src = p.u32 + g0.blocks_off;
        for (int e = tid; 4 * e < g0.n_blocks; e += nthr) cp_async16(blkbuf + 4 * e, src + 4 * e);
      }
      if (cached_group != gi) {  // per-(group, sub-pass) segment descriptors: keep them in shared memory
        const uint4 *src = reinterpret_cast<const uint4 *>(p.gsub + G.gsub_off);
        const int n16 = p.n_sub * (int)(sizeof(GroupSubDev) / 16);
        for (int e = tid; e < n16; e += nthr) reinterpret_cast<uint4 *>(gsub_s)[e] = src[e];
        // chunk-prefix rows for this group's column count (read straight from global: the copy
        // above is not visible yet)
        for (int sp = tid; sp < p.n_sub; sp += nthr) {
          const GroupSubDev &gsrc = p.gsub[G.gsub_off + sp];
          uint16_t *row = cend_s + sp * kChunkRow;
          int acc = 0;
#pragma unroll
          for (int k = 0; k < kChunkRow - 1; ++k) {
            if (k < gsrc.n_seg && k < kMaxSeg) acc += (gsrc.seg[k].count * cols + 31) >> 5;
            row[k] = (k < gsrc.n_seg && k < kMaxSeg) ? (uint16_t)acc : (uint16_t)0xFFFF;
          }
          row[kChunkRow - 1] = (uint16_t)acc;
        }
        // Balanced static schedule: chunks of a sub-pass differ in cost by ~5x between classes, and a
        // sub-pass ends at a CTA barrier, so its duration is the busiest warp's.  Longest-processing-time
        // first: walk the chunk list (classes are sorted heaviest first) and give every chunk to the warp
        // with the least work so far.  One warp builds the lists of one sub-pass (lane = target warp).
        for (int sp = warp; sp < p.n_sub; sp += nwarp) {
          const GroupSubDev &gsrc = p.gsub[G.gsub_off + sp];
          const int n_rot_sp = p.sub[sp].rot_end - p.sub[sp].rot_begin;
          uint8_t *mine = wl_s + ((size_t)sp * kMaxWarps + lane) * kWarpList;
          unsigned load = 0;
          int cnt = 0, g = 0;
          bool ok = nwarp <= kMaxWarps;
          for (int k = 0; k < gsrc.n_seg && k < kMaxSeg; ++k) {
            const int mp = gsrc.seg[k].mp;
            const int n_ch = (gsrc.seg[k].count * cols + 31) >> 5;
            // cycles, roughly: 12 FP64 operations per rotation and pair at ~2.7 clk, gather + scatter, fixed part
            const unsigned cost = 32u * n_rot_sp * dev_binom(W - 2, mp - 1) + 16u * dev_binom(W, mp) + 150u;
            for (int c = 0; c < n_ch; ++c, ++g) {
              const unsigned key = lane < nwarp ? ((load << 5) | (unsigned)lane) : 0xFFFFFFFFu;
              const int winner = (int)(__reduce_min_sync(0xFFFFFFFFu, key) & 31u);
              if (lane == winner) {
                if (cnt < kWarpList && g < 255) mine[cnt] = (uint8_t)g;
                ++cnt;
                load += cost;
              }
            }
          }
          if (lane < kMaxWarps)
            for (int c = min(cnt, kWarpList); c < kWarpList; ++c) mine[c] = 0xFF;
          ok = __all_sync(0xFFFFFFFFu, cnt <= kWarpList) && g <= 255 && ok;
          if (lane == 0) wl_ok_s[sp] = ok ? 1 : 0;
        }
      }
    }
    cp_async_wait_all();  // the tile, the first offset table and the first block list
    if (p.bulk_in && !FFB_KNOB(2) && !FFB_KNOB(32)) {
      mbar_wait(&tile_bar, tile_phase);
      tile_phase ^= 1u;
    }
    __syncthreads();
    FFB_TACC(1);
    cached_group = work ? gi : cached_group;

    // ---- sub-passes.  Chunks are dealt to the warps statically, boustrophedon over the heavy-first
    // chunk list (round k: chunk k * nwarp + warp, or + nwarp - 1 - warp when k is odd), which
    // keeps the warps balanced at the barrier without any shared counter.
    if (work) {
      for (int s = 0; s < p.n_sub; ++s) {
        const unsigned offtab_sa = off_sa + (unsigned)((s & 1) * off_entries * sizeof(uint32_t));
        const uint32_t *blk = blkbuf + (s & 1) * p.blk_cap;
        if (s + 1 < p.n_sub) {  // stage the next sub-pass's tables (they land before the barrier)
          uint32_t *dst = offbuf + ((s + 1) & 1) * off_entries;
          const uint32_t *src = p.off + (size_t)(s + 1) * kOffTabEntries;
          for (int e = tid; e < off_entries / 4; e += nthr) cp_async16(dst + 4 * e, src + 4 * e);
          const GroupSubDev &gn = gsub_s[s + 1];
          uint32_t *bdst = blkbuf + ((s + 1) & 1) * p.blk_cap;
          const uint32_t *bsrc = p.u32 + gn.blocks_off;
          for (int e = tid; 4 * e < gn.n_blocks; e += nthr) cp_async16(bdst + 4 * e, bsrc + 4 * e);
        }
        const GroupSubDev &gs = gsub_s[s];
        const uint16_t *cend = cend_s + s * kChunkRow;
        const int run0 = p.sub[s].run_begin, run1 = p.sub[s].run_end;
        const int n_chunks = cend[kChunkRow - 1];
        // chunk sequence of this warp: the balanced list when it was built, else the static boustrophedon
        // deal (round k: chunk k * nwarp + warp, reversed on odd k)
        const bool balanced = wl_ok_s[s] != 0;
        const uint8_t *mylist = wl_s + ((size_t)s * kMaxWarps + warp) * kWarpList;
        auto chunk_at = [&](int k) -> int {
          if (balanced) {
            const int g = k < kWarpList ? (int)mylist[k] : 0xFF;
            return g == 0xFF ? n_chunks : g;
          }
          return k * nwarp + ((k & 1) ? nwarp - 1 - warp : warp);
        };
        int g = chunk_at(0);
        ChunkWork cur = fetch_chunk(gs, cend, blk, g, lane, cols, G.inv_cols);
        for (int k = 1; g < n_chunks; ++k) {
          const int g_next = chunk_at(k);
          ChunkWork nxt;
          nxt.entry = 0;
          nxt.mp = 0;
          nxt.col = 0;
          if (g_next < n_chunks) nxt = fetch_chunk(gs, cend, blk, g_next, lane, cols, G.inv_cols);
          FFB_TACC(2);
          if (cur.mp && !FFB_KNOB(8)) {
            const unsigned a_addr = tile_sa + ((unsigned)(cur.col * Rp + (int)(cur.entry & 0xFFFFFFu)) << 4);
            const unsigned o_row = offtab_sa + (cur.entry >> 24) * (unsigned)(kOffRowDev * sizeof(uint32_t));
            process_dispatch<W>(cur.mp, a_addr, o_row, p, run0, run1);
          }
          FFB_TRESTART();
          cur = nxt;
          g = g_next;
        }
        cp_async_wait_all();
        FFB_TACC(2);
        if (!FFB_KNOB(1)) __syncthreads();
        FFB_TACC(6);
      }
    }

    // ---- store the tile (and fold the per-row phase product in on the last pass); table and phase
    // loads are batched ahead of the stores like in the load loop
    {
      long long unit_again = unit;
      asm volatile("" : "+l"(unit_again));  // decode afresh: keep the geometry out of the sub-pass registers
      double2 *__restrict__ data = reinterpret_cast<double2 *>(p.out);
      const double2 *__restrict__ rowphase = reinterpret_cast<const double2 *>(p.rowphase);
      const long long local = unit_again - G.unit_begin;
      const long long combo = local / G.n_strips;
      const long long strip = local - combo * G.n_strips;
      const int R = G.R;
      const unsigned inv_R = G.inv_R, inv_cols = G.inv_cols;
      const long long col0 = strip * cols;
      const int ncv = (int)min((long long)cols, p.n_cols - col0);
      const uint32_t rowbase = p.u32[G.combo_base_off + combo];
      const uint32_t *__restrict__ tab = p.u32 + G.tabrow_off + (size_t)p.u8[G.combo_low_off + combo] * R;
      const int n_el = R * cols;
      const bool row_major = p.out_col_stride == 1;
      if (p.bulk_out && !FFB_KNOB(2) && !FFB_KNOB(16)) {
        // every column goes out as one bulk copy; the per-row phases are multiplied in place first
        if (rowphase) {
          for (int e = tid; e < ncv * R; e += nthr) {
            const int j = fast_div(e, inv_R), r = e - j * R;
            const double2 v = tile[j * Rp + r], f = rowphase[rowbase + r];
            tile[j * Rp + r] = make_double2(v.x * f.x - v.y * f.y, v.x * f.y + v.y * f.x);
          }
        }
        fence_proxy_async();  // this thread's writes of the tile (sub-passes, phases) before the copy engine reads it
        __syncthreads();
        if (tid < ncv) {
          bulk_store(data + (long long)rowbase + (col0 + tid) * p.out_col_stride, tile + tid * Rp,
                     (unsigned)(R * (int)sizeof(double2)));
          bulk_store_commit_and_wait_read();  // the tile may be overwritten once the engine has read it
        }
      } else
      for (int e_base = 0; e_base < n_el; e_base += kTilePerThread * nthr) {
        uint32_t grow[kTilePerThread];
#pragma unroll
        for (int k = 0; k < kTilePerThread; ++k) {
          const int e = e_base + tid + k * nthr;
          const int r = row_major ? fast_div(e, inv_cols) : e - fast_div(e, inv_R) * R;
          grow[k] = e < n_el ? rowbase + tab[r] : 0u;
        }
        constexpr int kHalf = kTilePerThread / 2;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          double2 f[kHalf];
          if (rowphase) {
#pragma unroll
            for (int k = 0; k < kHalf; ++k) {
              const int e = e_base + tid + (h * kHalf + k) * nthr;
              f[k] = e < n_el ? rowphase[grow[h * kHalf + k]] : make_double2(1.0, 0.0);
            }
          }
#pragma unroll
          for (int k = 0; k < kHalf; ++k) {
            const int e = e_base + tid + (h * kHalf + k) * nthr;
            if (e < n_el) {
              int r, j;
              if (row_major) {
                r = fast_div(e, inv_cols);
                j = e - r * cols;
              } else {
                j = fast_div(e, inv_R);
                r = e - j * R;
              }
              if (j < ncv && !FFB_KNOB(2) && !FFB_KNOB(16)) {
                double2 v = tile[j * Rp + r];
                if (rowphase) v = make_double2(v.x * f[k].x - v.y * f[k].y, v.x * f[k].y + v.y * f[k].x);
                data[(long long)grow[h * kHalf + k] * p.out_row_stride + (col0 + j) * p.out_col_stride] = v;
              }
            }
          }
        }
      }
    }
    FFB_TACC(7);
    __syncthreads();
    FFB_TACC(6);
#ifdef FFB_DEBUG_TIMELINE
    if (tid == 0) s_tl_on = 0;  // only the CTA's first tile is logged
#endif
  }
  if (p.bulk_out) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // this thread's bulk stores have landed
#ifdef FFB_DEBUG_TIMING
  __syncthreads();
  if (tid < nwarp * 8) atomicAdd(&g_phase_cycles[tid & 7], s_phase_cycles[tid >> 3][tid & 7]);
#endif
}

template <int W>
static cudaError_t launch_w(const PassParams &p, int grid, int threads, size_t smem,
                            cudaStream_t stream) {
  // the opt-in shared-memory size is an attribute of the function ON ONE DEVICE: remember it per device
  static size_t configured[64] = {0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64 || smem > configured[dev]) {
    e = cudaFuncSetAttribute(fused_pass_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = smem;
  }
  fused_pass_kernel<W><<<grid, threads, smem, stream>>>(p);
  return cudaGetLastError();
}

size_t fused_pass_smem_overhead(int n_sub, int blk_cap, int off_rows) {
  return fused_smem_tile_off(n_sub, blk_cap, off_rows);
}

template <int W>
static int occupancy_w(int threads, size_t smem) {
  int n = 0;
  cudaFuncSetAttribute(fused_pass_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fused_pass_kernel<W>, threads, smem) != cudaSuccess) {
    cudaGetLastError();
    return 1;
  }
  return n < 1 ? 1 : n;
}

// CTAs of the fused kernel that are resident on one SM at this block size and tile size
// (registers, shared memory and thread limits all counted).
int fused_pass_ctas_per_sm(int w, int threads, size_t smem) {
  switch (w) {
    case 2: return occupancy_w<2>(threads, smem);
    case 3: return occupancy_w<3>(threads, smem);
    case 4: return occupancy_w<4>(threads, smem);
    case 5: return occupancy_w<5>(threads, smem);
    case 6: return occupancy_w<6>(threads, smem);
    default: return 1;
  }
}

#ifdef FFB_DEBUG_KNOBS
void set_debug_knobs(int v) { cudaMemcpyToSymbol(g_debug_knobs, &v, sizeof(int)); }
#endif
#ifdef FFB_DEBUG_TIMELINE
// out: 32 x kTlEvents records of (t0, t1, phase); counts: 32 ints; the log is cleared afterwards
void read_timeline(void *out, int *counts) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_tl, sizeof(g_tl));
  cudaMemcpyFromSymbol(counts, g_tl_count, sizeof(g_tl_count));
  int z[32] = {0};
  cudaMemcpyToSymbol(g_tl_count, z, sizeof(z));
}
int timeline_capacity() { return kTlEvents; }
#endif
#ifdef FFB_DEBUG_TIMING
void read_phase_cycles(unsigned long long *out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof(g_phase_cycles));
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_phase_cycles, z, sizeof(z));
  }
}
#endif

cudaError_t launch_fused_pass(const PassParams &p, int grid, int threads, size_t smem,
                              cudaStream_t stream) {
  switch (p.w) {
    case 2: return launch_w<2>(p, grid, threads, smem, stream);
    case 3: return launch_w<3>(p, grid, threads, smem, stream);
    case 4: return launch_w<4>(p, grid, threads, smem, stream);
    case 5: return launch_w<5>(p, grid, threads, smem, stream);
    case 6: return launch_w<6>(p, grid, threads, smem, stream);
    default: return cudaErrorInvalidValue;
  }
}

// ---------------------------------------------------------------- one rotation per launch
__global__ void givens_single_kernel(double2 *__restrict__ vec, long long ld, long long dim_b,
                                     double c, double sr, double si,
                                     const unsigned long long *__restrict__ s1,
                                     const unsigned long long *__restrict__ s2, long long n_pairs) {
  const long long total = n_pairs * dim_b;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long k = e / dim_b, col = e - k * dim_b;
    double2 *px = vec + (long long)s1[k] * ld + col;
    double2 *py = vec + (long long)s2[k] * ld + col;
    double2 x = *px, y = *py;
    zrot(x, y, c, sr, si);
    *px = x;
    *py = y;
  }
}

__global__ void phase_shift_kernel(double2 *__restrict__ vec, long long ld, long long dim_b,
                                   double pr, double pi,
                                   const unsigned long long *__restrict__ indices, long long n) {
  const long long total = n * dim_b;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long k = e / dim_b, col = e - k * dim_b;
    double2 *px = vec + (long long)indices[k] * ld + col;
    const double2 x = *px;
    *px = make_double2(x.x * pr - x.y * pi, x.x * pi + x.y * pr);
  }
}

// vec[row, :] *= phase[row]   (row stride ld, n_cols columns; or the strided variant)
__global__ void row_scale_kernel(double2 *__restrict__ vec, long long n_rows, long long n_cols,
                                 long long row_stride, long long col_stride,
                                 const double2 *__restrict__ phase) {
  const long long total = n_rows * n_cols;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    long long row, col;
    if (col_stride == 1) {
      row = e / n_cols;
      col = e - row * n_cols;
    } else {
      col = e / n_rows;
      row = e - col * n_rows;
    }
    double2 *px = vec + row * row_stride + col * col_stride;
    const double2 x = *px, f = phase[row];
    *px = make_double2(x.x * f.x - x.y * f.y, x.x * f.y + x.y * f.x);
  }
}

// phase[row] = prod_{i in string(row)} orbital_phase[i]
__global__ void row_phase_kernel(const uint32_t *__restrict__ strings, long long dim, int norb,
                                 PhaseList ph, double2 *__restrict__ out) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < dim;
       r += (long long)gridDim.x * blockDim.x) {
    uint32_t s = strings[r];
    double ar = 1.0, ai = 0.0;
    for (int i = 0; i < norb; ++i) {
      if ((s >> i) & 1u) {
        const double br = ph.re[i], bi = ph.im[i];
        const double nr = ar * br - ai * bi, ni = ar * bi + ai * br;
        ar = nr;
        ai = ni;
      }
    }
    out[r] = make_double2(ar, ai);
  }
}

constexpr int kTT = 32;
__global__ void transpose_kernel(const double2 *__restrict__ in, double2 *__restrict__ out,
                                 long long n_rows, long long n_cols, long long ld_in,
                                 long long ld_out, long long tiles_c) {
  __shared__ double2 t[kTT][kTT + 1];
  for (long long tileid = blockIdx.x;; tileid += gridDim.x) {
    const long long tr = tileid / tiles_c, tc = tileid - tr * tiles_c;
    if (tr * kTT >= n_rows) break;
    const long long r0 = tr * kTT, c0 = tc * kTT;
    for (int i = threadIdx.y; i < kTT; i += blockDim.y) {
      const long long r = r0 + i, c = c0 + threadIdx.x;
      if (r < n_rows && c < n_cols) t[i][threadIdx.x] = in[r * ld_in + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < kTT; i += blockDim.y) {
      const long long c = c0 + i, r = r0 + threadIdx.x;
      if (r < n_rows && c < n_cols) out[c * ld_out + r] = t[threadIdx.x][i];
    }
    __syncthreads();
  }
}

static int grid_for(long long total, int threads, int sm_count) {
  long long blocks = (total + threads - 1) / threads;
  long long cap = (long long)sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

cudaError_t launch_givens_single(void *vec, long long ld, long long dim_b, double c, double sr,
                                 double si, const unsigned long long *s1,
                                 const unsigned long long *s2, long long n_pairs, int sm_count,
                                 cudaStream_t stream) {
  if (n_pairs <= 0 || dim_b <= 0) return cudaSuccess;
  givens_single_kernel<<<grid_for(n_pairs * dim_b, 256, sm_count), 256, 0, stream>>>(
      (double2 *)vec, ld, dim_b, c, sr, si, s1, s2, n_pairs);
  return cudaGetLastError();
}

cudaError_t launch_phase_shift(void *vec, long long ld, long long dim_b, double pr, double pi,
                               const unsigned long long *indices, long long n, int sm_count,
                               cudaStream_t stream) {
  if (n <= 0 || dim_b <= 0) return cudaSuccess;
  phase_shift_kernel<<<grid_for(n * dim_b, 256, sm_count), 256, 0, stream>>>(
      (double2 *)vec, ld, dim_b, pr, pi, indices, n);
  return cudaGetLastError();
}

cudaError_t launch_row_scale(void *vec, long long n_rows, long long n_cols, long long row_stride,
                             long long col_stride, const void *phase, int sm_count,
                             cudaStream_t stream) {
  if (n_rows <= 0 || n_cols <= 0) return cudaSuccess;
  row_scale_kernel<<<grid_for(n_rows * n_cols, 256, sm_count), 256, 0, stream>>>(
      (double2 *)vec, n_rows, n_cols, row_stride, col_stride, (const double2 *)phase);
  return cudaGetLastError();
}

cudaError_t launch_row_phase(const uint32_t *strings, long long dim, int norb, const PhaseList &ph,
                             void *out, int sm_count, cudaStream_t stream) {
  if (dim <= 0) return cudaSuccess;
  row_phase_kernel<<<grid_for(dim, 256, sm_count), 256, 0, stream>>>(strings, dim, norb, ph,
                                                                     (double2 *)out);
  return cudaGetLastError();
}

cudaError_t launch_transpose(const void *in, void *out, long long n_rows, long long n_cols,
                             long long ld_in, long long ld_out, int sm_count, cudaStream_t stream) {
  if (n_rows <= 0 || n_cols <= 0) return cudaSuccess;
  const long long tiles_r = (n_rows + kTT - 1) / kTT, tiles_c = (n_cols + kTT - 1) / kTT;
  long long grid = tiles_r * tiles_c;
  const long long cap = (long long)sm_count * 16;
  if (grid > cap) grid = cap;
  transpose_kernel<<<(int)grid, dim3(kTT, 8), 0, stream>>>((const double2 *)in, (double2 *)out,
                                                           n_rows, n_cols, ld_in, ld_out, tiles_c);
  return cudaGetLastError();
}

}  // namespace ffb
