// TEST INFRASTRUCTURE ONLY -- never linked into libffsim_b200.so.
//
// Host emulation of fused_pass_kernel's index logic: consumes exactly the tables
// the plan builder uploads to the device (tile row tables, combination bases,
// register-block lists, block-offset tables, rotation order) and applies the
// rotations to a host vector the way the kernel does.  It lets the CPU-only test
// suite validate the schedule and every index table without a GPU; it is not a
// product path and nothing in ffsim_b200/ can reach it.
#include <complex>
#include <stdexcept>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../ffsim_b200/csrc/host.hpp"

using namespace ffb;

static void zrot(cplx &x, cplx &y, double c, cplx s) {
  cplx nx = c * x + s * y, ny = c * y - std::conj(s) * x;
  x = nx;
  y = ny;
}

extern "C" int ffb_hostcheck_apply_side(int norb, int nocc, const ffb_givens_rotation *rots, int n_rot,
                                        const ffb_c128 *phases, void *vec /* dim x n_cols, row-major */,
                                        int64_t n_cols, int64_t smem_bytes, int min_cols, int sub_window,
                                        int *n_passes_out, int *n_subs_out) try {
  PlanOptions opt = current_options();
  if (smem_bytes > 0) opt.smem_bytes = smem_bytes;
  if (min_cols > 0) opt.min_cols = min_cols;
  if (sub_window > 0) opt.sub_window = sub_window;
  cplx *data = reinterpret_cast<cplx *>(vec);
  std::vector<NormRot> nr;
  std::vector<int> q;
  for (int k = 0; k < n_rot; ++k) {
    nr.push_back(normalise(rots[k]));
    q.push_back(nr.back().q);
  }
  if (nocc == 0 || nocc == norb) q.clear();
  SideSchedule sched = build_schedule(norb, nocc, q, opt);
  int64_t dim = (int64_t)binom(norb, nocc);
  int total_subs = 0;
  std::vector<char> seen(q.size(), 0);
  for (const PassSchedule &ps : sched.passes) {
    PassTablesHost T = build_pass_tables(norb, nocc, ps);
    total_subs += (int)ps.subs.size();
    std::vector<char> row_seen(dim, 0);
    for (const PassGroupHost &G : T.groups) {
      const int R = G.R;
      for (size_t combo = 0; combo < G.combo_base.size(); ++combo) {
        const uint32_t rowbase = G.combo_base[combo];
        const uint32_t *tab = G.tabrow.data() + (size_t)G.combo_low[combo] * R;
        for (int r = 0; r < R; ++r) {
          int64_t row = (int64_t)rowbase + tab[r];
          if (row < 0 || row >= dim || row_seen[row]) return -100;  // tiles must partition the rows
          row_seen[row] = 1;
        }
        for (int64_t col = 0; col < n_cols; ++col) {
          std::vector<cplx> tile(R);
          for (int r = 0; r < R; ++r) tile[r] = data[((int64_t)rowbase + tab[r]) * n_cols + col];
          for (size_t s = 0; s < ps.subs.size(); ++s) {
            const SubPass &sp = ps.subs[s];
            const GroupSubHost &gs = G.subs[s];
            const uint16_t *offtab = T.off.data() + s * kMaxLow * kOffRow;
            std::vector<char> used(R, 0);
            for (int sg = 0; sg < gs.n_seg; ++sg) {
              const int mp = gs.seg_mp[sg];
              std::vector<uint64_t> pats = strings_of(sp.w, mp);
              for (int b = 0; b < gs.seg_count[sg]; ++b) {
                uint32_t entry = gs.blocks[gs.seg_begin[sg] + b];
                int base = entry & 0xFFFFFF;
                const uint16_t *o = offtab + (entry >> 24) * kOffRow + class_offset(sp.w, mp);
                std::vector<int> idx(pats.size());
                for (size_t t = 0; t < pats.size(); ++t) {
                  idx[t] = base + o[t];
                  if (idx[t] < 0 || idx[t] >= R || used[idx[t]]) return -101;  // blocks are disjoint
                  used[idx[t]] = 1;
                }
                for (int rr = sp.rot_begin; rr < sp.rot_end; ++rr) {
                  const NormRot &g = nr[ps.rot_index[rr]];
                  int qq = g.q - ps.lo - sp.q0;
                  if (qq < 0 || qq + 1 >= sp.w) return -102;
                  for (size_t t = 0; t < pats.size(); ++t) {
                    uint64_t S = pats[t];
                    if (((S >> qq) & 3) == 1) {
                      uint64_t S2 = S ^ (3ull << qq);
                      zrot(tile[idx[rank_of(S)]], tile[idx[rank_of(S2)]], g.c, g.s);
                    }
                  }
                }
              }
            }
          }
          for (int r = 0; r < R; ++r) data[((int64_t)rowbase + tab[r]) * n_cols + col] = tile[r];
        }
      }
    }
    for (int64_t r = 0; r < dim; ++r)
      if (!row_seen[r]) return -103;
    // dispatch runs: partition each sub-pass's rotations into descending runs the kernel has code for
    {
      std::vector<unsigned char> rq(ps.rot_index.size() + 8, 0);
      for (size_t r = 0; r < ps.rot_index.size(); ++r) rq[r] = (unsigned char)(nr[ps.rot_index[r]].q - ps.lo);
      for (const SubPass &sp : ps.subs) {
        int expect = sp.rot_begin;
        for (const Run &run : segment_runs(rq.data(), sp.rot_begin, sp.rot_end, sp.q0)) {
          if (run.first != expect || run.len < 1 || run.len > kMaxRunLen) return -106;
          if (run.q_hi_rel > sp.w - 2 || run.q_hi_rel - run.len + 1 < 0) return -107;
          for (int i = 0; i < run.len; ++i)
            if ((int)rq[run.first + i] - sp.q0 != run.q_hi_rel - i) return -108;
          expect += run.len;
        }
        if (expect != sp.rot_end) return -109;
      }
    }
    for (int g : ps.rot_index) {
      if (seen[g]) return -104;
      seen[g] = 1;
    }
  }
  for (size_t g = 0; g < q.size(); ++g)
    if (!seen[g]) return -105;
  if (phases) {
    std::vector<uint64_t> strs = strings_of(norb, nocc);
    for (int64_t r = 0; r < dim; ++r) {
      cplx f(1.0, 0.0);
      for (int i = 0; i < norb; ++i)
        if ((strs[r] >> i) & 1) f *= cplx(phases[i].re, phases[i].im);
      for (int64_t col = 0; col < n_cols; ++col) data[r * n_cols + col] *= f;
    }
  }
  if (n_passes_out) *n_passes_out = (int)sched.passes.size();
  if (n_subs_out) *n_subs_out = total_subs;
  return 0;
} catch (const std::exception &) {
  return -200;  // a plan-builder invariant fired (plan.cpp require())
}

// Print the two-level schedule of a pair-position sequence (developer aid).
extern "C" int ffb_hostcheck_dump_schedule(int norb, int nocc, const int *q, int n, int64_t smem_bytes,
                                           int min_cols, int sub_window) try {
  PlanOptions opt = current_options();
  if (smem_bytes > 0) opt.smem_bytes = smem_bytes;
  if (min_cols > 0) opt.min_cols = min_cols;
  if (sub_window > 0) opt.sub_window = sub_window;
  std::vector<int> qq(q, q + n);
  SideSchedule sched = build_schedule(norb, nocc, qq, opt);
  for (size_t ip = 0; ip < sched.passes.size(); ++ip) {
    const PassSchedule &ps = sched.passes[ip];
    std::printf("pass %zu: lo=%d W=%d rots=%zu subs=%zu\n", ip, ps.lo, ps.W, ps.rot_index.size(), ps.subs.size());
    for (const SubPass &sp : ps.subs) {
      std::printf("  sub q0=%2d w=%d n=%2d :", sp.q0, sp.w, sp.rot_end - sp.rot_begin);
      for (int r = sp.rot_begin; r < sp.rot_end; ++r) std::printf(" %d", qq[ps.rot_index[r]] - ps.lo - sp.q0);
      std::printf("\n");
    }
  }
  return 0;
} catch (const std::exception &) {
  return -200;  // a plan-builder invariant fired (plan.cpp require())
}

// Shared-memory bank behaviour of the register-block gathers, from the block lists the kernel will
// consume: lanes of a quarter-warp (8 consecutive items of a class; the block index runs fastest)
// access rows base + o[t]; rows that agree modulo 8 fall in the same 16-byte bank group.  Counts the
// quarter-warps of every (pass, group, sub-pass, class) and the extra wavefronts their first access
// needs (0 when the eight rows differ modulo 8).  out[0] = quarter-warps, out[1] = extra wavefronts.
extern "C" int ffb_hostcheck_gather_conflicts(int norb, int nocc, const int *q, int n, int64_t *out) try {
  PlanOptions opt = current_options();
  std::vector<int> qq(q, q + n);
  SideSchedule sched = build_schedule(norb, nocc, qq, opt);
  int64_t quarters = 0, extra = 0;
  for (const PassSchedule &ps : sched.passes) {
    PassTablesHost T = build_pass_tables(norb, nocc, ps);
    for (const PassGroupHost &G : T.groups) {
      for (size_t s = 0; s < ps.subs.size(); ++s) {
        const GroupSubHost &gs = G.subs[s];
        const uint16_t *offtab = T.off.data() + s * kMaxLow * kOffRow;
        for (int sg = 0; sg < gs.n_seg; ++sg) {
          const int mp = gs.seg_mp[sg];
          const int coff = class_offset(ps.subs[s].w, mp);
          for (int b0 = 0; b0 < gs.seg_count[sg]; b0 += 8) {
            int hits[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            int lanes = std::min(8, gs.seg_count[sg] - b0), worst = 0;
            for (int l = 0; l < lanes; ++l) {
              uint32_t e = gs.blocks[gs.seg_begin[sg] + b0 + l];
              int row = (int)(e & 0xFFFFFF) + offtab[(e >> 24) * kOffRow + coff];
              worst = std::max(worst, ++hits[row & 7]);
            }
            ++quarters;
            extra += worst - 1;
          }
        }
      }
    }
  }
  out[0] = quarters;
  out[1] = extra;
  return 0;
} catch (const std::exception &) {
  return -200;  // a plan-builder invariant fired (plan.cpp require())
}

// ---------------------------------------------------------------------------------------------------
// Device view.  The same rotation, but walking exactly the data the kernel walks, the way it walks it:
// the packed tables of pack_device_tables() (u32/u8 pools, GroupSubDev segment descriptors with their
// reciprocal counts, byte-offset tables with 16-byte aligned classes, 16-byte aligned block lists), the
// column-major tile with the bank-aware column stride, the chunk-prefix rows, the boustrophedon dealing
// of 32-item chunks to `nwarp` warps, and fetch_chunk's arithmetic including the multiply-high division.
// Mirrors fused_pass_kernel / fetch_chunk / process_item of csrc/givens_kernels.cu line by line; any
// change there must be made here too (that is the point: it pins the index logic without a GPU).
static inline uint32_t umulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline int fast_div_host(int n, uint32_t inv) { return inv ? (int)umulhi32((uint32_t)n, inv) : n; }

extern "C" int ffb_hostcheck_apply_side_device(int norb, int nocc, const ffb_givens_rotation *rots, int n_rot,
                                               void *vec /* dim x n_cols, row-major */, int64_t n_cols,
                                               int64_t smem_bytes, int min_cols, int sub_window, int cols_req,
                                               int nwarp) try {
  PlanOptions opt = current_options();
  if (smem_bytes > 0) opt.smem_bytes = smem_bytes;
  if (min_cols > 0) opt.min_cols = min_cols;
  if (sub_window > 0) opt.sub_window = sub_window;
  cplx *data = reinterpret_cast<cplx *>(vec);
  std::vector<NormRot> nr;
  std::vector<int> q;
  for (int k = 0; k < n_rot; ++k) {
    nr.push_back(normalise(rots[k]));
    q.push_back(nr.back().q);
  }
  if (nocc == 0 || nocc == norb) q.clear();
  SideSchedule sched = build_schedule(norb, nocc, q, opt);
  constexpr int kChunkRow = 8;
  for (const PassSchedule &ps : sched.passes) {
    PassTablesHost T = build_pass_tables(norb, nocc, ps);
    std::vector<int> group_R, group_has;
    std::vector<size_t> group_combos;
    for (const PassGroupHost &G : T.groups) {
      group_R.push_back(G.R);
      group_has.push_back(G.has_blocks ? 1 : 0);
      group_combos.push_back(G.combo_base.size());
    }
    DevicePassTables D = pack_device_tables(ps, T);
    const int n_sub = (int)ps.subs.size();
    for (size_t gi = 0; gi < group_R.size(); ++gi) {
      if (!group_has[gi]) continue;
      const int R = group_R[gi];
      const int cols = (int)std::max<int64_t>(1, std::min<int64_t>(cols_req, n_cols));
      const int Rp = tile_col_stride(R, cols);
      if (Rp < R || (Rp - R) > 7) return -201;
      const DevicePassTables::GroupOffsets &go = D.goff[gi];
      const int64_t n_strips = (n_cols + cols - 1) / cols;
      // chunk-prefix rows (the kernel builds them when it caches a group)
      std::vector<uint16_t> cend((size_t)n_sub * kChunkRow);
      for (int sp = 0; sp < n_sub; ++sp) {
        const GroupSubDev &g = D.gsub[go.gsub_off + sp];
        int acc = 0;
        for (int k = 0; k < kChunkRow - 1; ++k) {
          if (k < g.n_seg && k < kMaxSeg) acc += (g.seg[k].count * cols + 31) >> 5;
          cend[(size_t)sp * kChunkRow + k] = (k < g.n_seg && k < kMaxSeg) ? (uint16_t)acc : (uint16_t)0xFFFF;
        }
        cend[(size_t)sp * kChunkRow + kChunkRow - 1] = (uint16_t)acc;
      }
      for (size_t combo = 0; combo < group_combos[gi]; ++combo) {
        const uint32_t rowbase = D.u32[go.combo_base_off + combo];
        const uint32_t *tab = D.u32.data() + go.tabrow_off + (size_t)D.u8[go.combo_low_off + combo] * R;
        for (int64_t strip = 0; strip < n_strips; ++strip) {
          const int64_t col0 = strip * cols;
          const int ncv = (int)std::min<int64_t>(cols, n_cols - col0);
          std::vector<cplx> tile((size_t)Rp * cols, cplx(0, 0));
          for (int r = 0; r < R; ++r)
            for (int j = 0; j < ncv; ++j) tile[(size_t)j * Rp + r] = data[((int64_t)rowbase + tab[r]) * n_cols + col0 + j];
          for (int s = 0; s < n_sub; ++s) {
            const SubPass &sp = ps.subs[s];
            const GroupSubDev &gs = D.gsub[go.gsub_off + s];
            if (gs.blocks_off % 4 || gs.n_blocks > D.blk_cap) return -202;
            const uint32_t *blk = D.u32.data() + gs.blocks_off;  // (staged copy of n_blocks entries)
            const uint32_t *offtab = D.off32.data() + (size_t)s * kMaxLowDev * kOffRowDev;
            const uint16_t *ce = cend.data() + (size_t)s * kChunkRow;
            const int n_chunks = ce[kChunkRow - 1];
            std::vector<char> used(tile.size(), 0);
            int items_done = 0;
            for (int warp = 0; warp < nwarp; ++warp) {
              for (int k = 0;; ++k) {
                const int g = k * nwarp + ((k & 1) ? nwarp - 1 - warp : warp);
                if (g >= n_chunks) break;
                // fetch_chunk
                const int c0 = ce[0], c1 = ce[1], c2 = ce[2], c3 = ce[3];
                int sg = 0, base = 0;
                if (g >= c0) sg = 1, base = c0;
                if (g >= c1) sg = 2, base = c1;
                if (g >= c2) sg = 3, base = c2;
                if (g >= c3) sg = 4, base = c3;
                const SegDev &sq = gs.seg[sg];
                const int mp = sq.mp & 0xFF, count = sq.count;
                std::vector<uint64_t> pats = strings_of(sp.w, mp);
                for (int lane = 0; lane < 32; ++lane) {
                  const int item = ((g - base) << 5) + lane;
                  if (item >= count * cols) continue;
                  int col, b;
                  if (items_column_fastest(cols)) {
                    b = fast_div_host(item, 0xFFFFFFFFu / (unsigned)cols + 1u);
                    col = item - b * cols;
                  } else {
                    col = fast_div_host(item, sq.inv_count);
                    b = item - col * count;
                  }
                  if (col < 0 || col >= cols || b < 0 || b >= count || sq.begin + b >= gs.n_blocks) return -203;
                  const uint32_t entry = blk[sq.begin + b];
                  // process_item: byte addresses relative to the tile
                  const uint32_t a_addr = (uint32_t)(col * Rp + (int)(entry & 0xFFFFFFu)) << 4;
                  const uint32_t *o = offtab + (entry >> 24) * kOffRowDev + dev_class_offset(sp.w, mp);
                  std::vector<size_t> idx(pats.size());
                  for (size_t t = 0; t < pats.size(); ++t) {
                    const uint32_t byte = a_addr + o[t];
                    if (byte % 16) return -204;
                    idx[t] = byte / 16;
                    if (idx[t] >= tile.size() || idx[t] < (size_t)col * Rp || idx[t] >= (size_t)col * Rp + R) return -205;
                    if (used[idx[t]]) return -206;  // blocks of a sub-pass are disjoint
                    used[idx[t]] = 1;
                  }
                  for (int rr = sp.rot_begin; rr < sp.rot_end; ++rr) {
                    const NormRot &gr = nr[ps.rot_index[rr]];
                    const int qq = gr.q - ps.lo - sp.q0;
                    for (size_t t = 0; t < pats.size(); ++t) {
                      const uint64_t S = pats[t];
                      if (((S >> qq) & 3) == 1) zrot(tile[idx[rank_of(S)]], tile[idx[rank_of(S ^ (3ull << qq))]], gr.c, gr.s);
                    }
                  }
                  ++items_done;
                }
              }
            }
            if (items_done != gs.n_blocks * cols) return -207;  // every (block, column) exactly once
          }
          for (int r = 0; r < R; ++r)
            for (int j = 0; j < ncv; ++j) data[((int64_t)rowbase + tab[r]) * n_cols + col0 + j] = tile[(size_t)j * Rp + r];
        }
      }
    }
  }
  return 0;
} catch (const std::exception &) {
  return -200;  // a plan-builder invariant fired (plan.cpp require())
}

// Shared-memory wavefronts of the register-block gathers (= scatters) of one tile of every group, walked
// the way the kernel walks them: 32-item chunks of (column, block) pairs with the block index fastest,
// a quarter-warp (8 lanes, 16-byte accesses) per wavefront when its rows fall in eight different bank
// groups, more when they collide.  out[0] = wavefronts, out[1] = the conflict-free count, per tile and
// summed over the groups weighted by their number of tiles per column strip (combinations).
extern "C" int ffb_hostcheck_gather_wavefronts(int norb, int nocc, const int *q, int n, int cols_req, int col_fast, int64_t *out) try {
  PlanOptions opt = current_options();
  std::vector<int> qq(q, q + n);
  SideSchedule sched = build_schedule(norb, nocc, qq, opt);
  int64_t actual = 0, ideal = 0;
  for (const PassSchedule &ps : sched.passes) {
    PassTablesHost T = build_pass_tables(norb, nocc, ps);
    std::vector<int> group_R;
    std::vector<int64_t> group_combos;
    for (const PassGroupHost &G : T.groups) {
      group_R.push_back(G.R);
      group_combos.push_back((int64_t)G.combo_base.size());
    }
    DevicePassTables D = pack_device_tables(ps, T);
    for (size_t gi = 0; gi < group_R.size(); ++gi) {
      const int R = group_R[gi];
      const int cols = cols_req > 0 ? cols_req : (int)std::max<int64_t>(1, std::min<int64_t>(8, (opt.smem_bytes / 16) / (R + 7)));
      const int Rp = tile_col_stride(R, cols);
      for (size_t s = 0; s < ps.subs.size(); ++s) {
        const GroupSubDev &gs = D.gsub[D.goff[gi].gsub_off + s];
        const uint32_t *blk = D.u32.data() + gs.blocks_off;
        const uint32_t *offtab = D.off32.data() + s * kMaxLowDev * kOffRowDev;
        for (int sg = 0; sg < gs.n_seg && sg < kMaxSeg; ++sg) {
          const SegDev &sq = gs.seg[sg];
          const int mp = sq.mp & 0xFF, count = sq.count;
          const int nt = (int)binom(ps.subs[s].w, mp);
          const int items = count * cols;
          for (int i0 = 0; i0 < items; i0 += 8) {
            for (int t = 0; t < nt; ++t) {
              int hits[8] = {0, 0, 0, 0, 0, 0, 0, 0}, worst = 0;
              for (int l = 0; l < 8 && i0 + l < items; ++l) {
                const int item = i0 + l;
                const bool cf = col_fast && items_column_fastest(cols);
                const int col = cf ? item % cols : item / count, b = cf ? item / cols : item - col * count;
                const uint32_t entry = blk[sq.begin + b];
                const uint32_t byte = ((uint32_t)(col * Rp + (int)(entry & 0xFFFFFFu)) << 4) +
                                      offtab[(entry >> 24) * kOffRowDev + dev_class_offset(ps.subs[s].w, mp) + t];
                worst = std::max(worst, ++hits[(byte >> 4) & 7]);
              }
              actual += worst * group_combos[gi];
              ideal += group_combos[gi];
            }
          }
        }
      }
    }
  }
  out[0] = actual;
  out[1] = ideal;
  return 0;
} catch (const std::exception &) {
  return -200;
}
