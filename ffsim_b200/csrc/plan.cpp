// Plan builder: turns the adjacent-pair Givens sequence of one spin sector into
// (1) passes over the state, each confining its rotations to a window of
// orbitals [lo, lo+W) whose C(W, m) strings fit a shared-memory tile, and
// (2) sub-passes inside a tile, each confining its rotations to a register
// block of w <= 6 orbitals.  Rotations on disjoint orbital pairs commute, so the
// sequence is a dependency DAG; both levels take, greedily, the largest
// dependency-closed set that fits the window.
//
// There is no counterpart in the reference: it performs one sweep per rotation
// (python/ffsim/gates/orbital_rotation.py:132-135).
#include <algorithm>
#include <stdexcept>
#include <cstdlib>
#include <cstring>

#include "device_structs.h"
#include "host.hpp"

namespace ffb {

namespace {
PlanOptions g_opt = {/*smem_bytes=*/220 * 1024, /*min_cols=*/3, /*max_cols=*/8,
                     /*sub_window=*/6,          /*threads=*/512, /*bulk_copies=*/1, /*beta_mode=*/0};
std::mutex g_opt_mu;
}  // namespace

PlanOptions current_options() {
  std::lock_guard<std::mutex> lk(g_opt_mu);
  return g_opt;
}

int class_offset(int w, int mprime) {
  int off = 0;
  for (int j = 1; j < mprime; ++j) off += (int)binom(w, j);
  return off;
}

// Invariants of the plan tables: violated only by a bug in the builder, never by user input.  They
// surface as FFB_EINTERNAL through the C ABI (capi.cu catches), not as an abort of the host process.
static void require(bool ok, const char *what) {
  if (!ok) throw std::logic_error(std::string("plan builder invariant violated: ") + what);
}

int max_window(int norb, int nocc, const PlanOptions &opt) {
  int64_t budget_amps = opt.smem_bytes / 16;
  // a register block starts at most W - w orbitals into the window, and the block-offset tables hold
  // kMaxLow distinct counts of electrons below it: cap the window accordingly (nearly-filled sectors of
  // norb >= 22 would otherwise reach counts >= kMaxLow)
  const int w_reg = std::max(2, std::min(opt.sub_window, kMaxSubWindow));
  for (int W = std::min(norb, kMaxLow - 1 + w_reg); W >= 2; --W) {
    int mlo = std::max(0, nocc - (norb - W)), mhi = std::min(W, nocc);
    uint64_t maxR = 0;
    for (int m = mlo; m <= mhi; ++m) maxR = std::max(maxR, binom(W, m));
    if (maxR >= 65536) continue;  // block offsets are 16-bit
    if ((int64_t)maxR * std::max(1, opt.min_cols) >= 65536) continue;  // tile indices are 16-bit
    if ((int64_t)maxR * std::max(1, opt.min_cols) <= budget_amps) return W;
  }
  return std::min(norb, 2);
}

// Largest dependency-closed subsequence of `remaining` inside [lo, lo+W).
static std::vector<int> admit(const std::vector<int> &remaining, const std::vector<int> &q, int lo,
                              int W, size_t cap) {
  std::vector<int> out;
  uint64_t blocked = 0;
  for (int g : remaining) {
    uint64_t mask = 3ull << q[g];
    bool inside = q[g] >= lo && q[g] + 1 < lo + W;
    if (inside && !(blocked & mask) && out.size() < cap)
      out.push_back(g);
    else
      blocked |= mask;
  }
  return out;
}

static std::vector<int> minus(const std::vector<int> &a, const std::vector<int> &taken) {
  std::vector<int> out;
  size_t t = 0;
  for (int g : a) {
    if (t < taken.size() && taken[t] == g)
      ++t;
    else
      out.push_back(g);
  }
  return out;
}

// Tile a rotation sequence (pair positions q, relative to a span of `span` orbitals) with
// windows of width `width`: returns groups in execution order; each group is a window
// start and the indices of the rotations it applies (a dependency-closed set).
// Greedy "take the fullest window" strands a few rotations in groups of their own, and a
// group here costs a sweep over the state (level 1) or over the tile (level 2), so a small
// beam search over the window choices minimises the number of groups instead.
struct Group {
  int start;
  std::vector<int> members;
};

static std::vector<Group> tile_sequence(const std::vector<int> &index, const std::vector<int> &q,
                                        int span, int width, size_t cap, size_t max_groups) {
  struct State {
    std::vector<int> remaining;
    std::vector<Group> groups;
  };
  const size_t kBeam = 12;
  std::vector<State> beam(1);
  beam[0].remaining = index;
  if (index.empty()) return {};
  for (size_t depth = 0; depth < 4 * index.size() + 4; ++depth) {
    std::vector<State> next;
    for (const State &st : beam) {
      // candidate windows, deduplicated by the set they admit
      std::vector<std::vector<int>> seen;
      for (int start = 0; start + width <= span; ++start) {
        std::vector<int> got = admit(st.remaining, q, start, width, cap);
        if (got.empty()) continue;
        bool dup = false;
        for (const auto &o : seen) dup = dup || o == got;
        if (dup) continue;
        seen.push_back(got);
        State ns;
        ns.remaining = minus(st.remaining, got);
        ns.groups = st.groups;
        ns.groups.push_back({start, std::move(got)});
        next.push_back(std::move(ns));
      }
    }
    require(!next.empty(), "no window admits a rotation");
    // finished states win; otherwise keep the states with the fewest rotations left
    std::stable_sort(next.begin(), next.end(), [](const State &a, const State &b) {
      return a.remaining.size() < b.remaining.size();
    });
    if (next[0].remaining.empty() || next[0].groups.size() >= max_groups) return next[0].groups;
    // drop duplicates (same remaining set) and cut to the beam width
    std::vector<State> cut;
    for (State &ns : next) {
      bool dup = false;
      for (const State &o : cut) dup = dup || o.remaining == ns.remaining;
      if (!dup) cut.push_back(std::move(ns));
      if (cut.size() == kBeam) break;
    }
    beam.swap(cut);
  }
  return beam[0].groups;
}

static SideSchedule build_schedule_for(int norb, int nocc, const std::vector<int> &q, const PlanOptions &opt);

// A sweep over the state is the expensive unit.  Tiles of three or more columns move more bytes
// per row access, but a two-column tile (still one full 32-byte sector per row) allows a wider
// window; take it when that saves a whole sweep (norb=20, nelec=8: 3 sweeps instead of 4).
SideSchedule build_schedule(int norb, int nocc, const std::vector<int> &q, const PlanOptions &opt) {
  SideSchedule best = build_schedule_for(norb, nocc, q, opt);
  if (opt.min_cols > 2 && best.passes.size() > 1) {
    PlanOptions narrow = opt;
    narrow.min_cols = 2;
    SideSchedule alt = build_schedule_for(norb, nocc, q, narrow);
    if (alt.passes.size() < best.passes.size()) best = std::move(alt);
  }
  return best;
}

static SideSchedule build_schedule_for(int norb, int nocc, const std::vector<int> &q, const PlanOptions &opt) {
  SideSchedule sched;
  sched.norb = norb;
  sched.nocc = nocc;
  if (q.empty() || norb < 2) return sched;
  const int W = max_window(norb, nocc, opt);
  const int w = std::max(2, std::min({opt.sub_window, kMaxSubWindow, W}));

  std::vector<int> all(q.size());
  for (size_t g = 0; g < q.size(); ++g) all[g] = (int)g;
  std::vector<int> remaining = all;
  while (!remaining.empty()) {
    // level 1: sweeps over the state.  (Re-planned from the rotations still left, so a pass
    // that had to hand rotations back -- see below -- is followed by a consistent schedule.)
    std::vector<Group> passes = tile_sequence(remaining, q, norb, W, kMaxRotPerPass, 1u << 30);
    require(!passes.empty(), "empty pass list");
    bool handed_back = false;
    for (const Group &pg : passes) {
      PassSchedule pass;
      pass.lo = pg.start;
      pass.W = W;
      // level 2: register-block sub-passes inside the tile
      std::vector<int> qrel(q.size(), 0);
      for (int g : pg.members) qrel[g] = q[g] - pg.start;
      std::vector<Group> subs = tile_sequence(pg.members, qrel, W, w, 1u << 30, kMaxSubPerPass);
      std::vector<int> done;
      for (const Group &sg : subs) {
        SubPass sp;
        sp.q0 = sg.start;
        sp.w = w;
        sp.rot_begin = (int)pass.rot_index.size();
        pass.rot_index.insert(pass.rot_index.end(), sg.members.begin(), sg.members.end());
        sp.rot_end = (int)pass.rot_index.size();
        pass.subs.push_back(sp);
        done.insert(done.end(), sg.members.begin(), sg.members.end());
      }
      std::sort(done.begin(), done.end());
      remaining = minus(remaining, done);
      sched.passes.push_back(std::move(pass));
      if (done.size() != pg.members.size()) {  // out of sub-pass slots: re-plan the rest
        handed_back = true;
        break;
      }
    }
    if (!handed_back) break;
  }
  return sched;
}

std::vector<Run> segment_runs(const unsigned char *rq, int rot_begin, int rot_end, int q0) {
  std::vector<Run> runs;
  int r = rot_begin;
  while (r < rot_end) {
    Run run{r, 1, (int)rq[r] - q0};
    while (run.len < kMaxRunLen && r + run.len < rot_end &&
           (int)rq[r + run.len] == (int)rq[r] - run.len)
      ++run.len;
    runs.push_back(run);
    r += run.len;
  }
  return runs;
}

// sum over the set bits of `pattern` (ascending) of C(shift + pos, first_index + i)
static uint64_t placed_rank(uint64_t pattern, int shift, int first_index) {
  uint64_t r = 0;
  int idx = first_index;
  while (pattern) {
    int pos = __builtin_ctzll(pattern);
    pattern &= pattern - 1;
    r += binom(shift + pos, idx++);
  }
  return r;
}

PassTablesHost build_pass_tables(int norb, int nocc, const PassSchedule &pass) {
  PassTablesHost T;
  T.lo = pass.lo;
  T.W = pass.W;
  const int lo = pass.lo, W = pass.W, hi_bits = norb - lo - W;
  const int n_sub = (int)pass.subs.size();

  // block-offset tables: one per sub-pass, shared by every group
  T.off.assign((size_t)n_sub * kMaxLow * kOffRow, 0);
  for (int s = 0; s < n_sub; ++s) {
    const SubPass &sp = pass.subs[s];
    for (int lp = 0; lp < kMaxLow && lp <= sp.q0; ++lp) {
      for (int mp = 1; mp < sp.w; ++mp) {
        std::vector<uint64_t> pats = strings_of(sp.w, mp);
        for (size_t t = 0; t < pats.size(); ++t) {
          uint64_t v = placed_rank(pats[t], sp.q0, lp + 1);
          T.off[((size_t)s * kMaxLow + lp) * kOffRow + class_offset(sp.w, mp) + t] =
              (uint16_t)std::min<uint64_t>(v, 65535);
        }
      }
    }
  }

  int mlo = std::max(0, nocc - (norb - W)), mhi = std::min(W, nocc);
  for (int m = mlo; m <= mhi; ++m) {
    PassGroupHost G;
    G.m = m;
    G.R = (int)binom(W, m);
    int outside = nocc - m;  // electrons outside the window
    G.l_min = std::max(0, outside - hi_bits);
    int l_max = std::min(lo, outside);
    G.n_low = l_max - G.l_min + 1;
    std::vector<uint64_t> pats = strings_of(W, m);
    G.tabrow.resize((size_t)G.n_low * G.R);
    for (int l = G.l_min; l <= l_max; ++l)
      for (int r = 0; r < G.R; ++r)
        G.tabrow[(size_t)(l - G.l_min) * G.R + r] = (uint32_t)placed_rank(pats[r], lo, l + 1);
    for (int l = G.l_min; l <= l_max; ++l) {
      int h = outside - l;
      std::vector<uint64_t> Hs = strings_of(hi_bits, h), Ls = strings_of(lo, l);
      for (uint64_t H : Hs) {
        uint64_t hb = placed_rank(H, lo + W, l + m + 1);
        for (uint64_t L : Ls) {
          G.combo_base.push_back((uint32_t)(hb + rank_of(L)));
          G.combo_low.push_back((uint8_t)(l - G.l_min));
        }
      }
    }
    // register blocks of every sub-pass
    G.subs.resize(n_sub);
    for (int s = 0; s < n_sub; ++s) {
      const SubPass &sp = pass.subs[s];
      GroupSubHost &gs = G.subs[s];
      std::memset(gs.seg_mp, 0, sizeof(gs.seg_mp));
      std::memset(gs.seg_begin, 0, sizeof(gs.seg_begin));
      std::memset(gs.seg_count, 0, sizeof(gs.seg_count));
      int up_bits = W - sp.q0 - sp.w;
      std::vector<int> classes;
      for (int mp = 1; mp < sp.w; ++mp) classes.push_back(mp);
      std::stable_sort(classes.begin(), classes.end(), [&](int a, int b) {
        return binom(sp.w - 2, a - 1) > binom(sp.w - 2, b - 1);
      });
      for (int mp : classes) {
        int begin = (int)gs.blocks.size();
        uint64_t n_up = 1ull << up_bits;
        for (uint64_t Hp = 0; Hp < n_up; ++Hp) {
          int lp = m - __builtin_popcountll(Hp) - mp;
          if (lp < 0 || lp > sp.q0) continue;
          require(lp < kMaxLow, "electron count below the register block exceeds the offset table");
          uint64_t hb = placed_rank(Hp, sp.q0 + sp.w, lp + mp + 1);
          uint64_t nL = binom(sp.q0, lp);
          for (uint64_t r = 0; r < nL; ++r) {
            uint64_t base = hb + r;
            require(base < (1u << 24), "block base row exceeds 24 bits");
            gs.blocks.push_back((uint32_t)base | ((uint32_t)lp << 24));
          }
        }
        int count = (int)gs.blocks.size() - begin;
        // Order the class's blocks so that any eight consecutive ones (a quarter-warp: one
        // shared-memory wavefront of 16-byte accesses) start on rows that differ modulo 8, i.e. in
        // different bank groups: blocks with the same l' share their offset list, so the whole gather
        // and scatter of such a group is conflict-free.  (Any order is valid: blocks are disjoint.)
        if (count > 8) {
          std::vector<uint32_t> seg(gs.blocks.begin() + begin, gs.blocks.end());
          std::vector<std::vector<uint32_t>> bucket((size_t)kMaxLow * 8);
          for (uint32_t e : seg) bucket[(size_t)(e >> 24) * 8 + (e & 7u)].push_back(e);
          size_t pos = (size_t)begin;
          int res = 0;  // the residue cycle runs on across l' groups
          for (int lp = 0; lp < kMaxLow; ++lp) {
            size_t idx[8] = {0, 0, 0, 0, 0, 0, 0, 0}, left = 0;
            for (int r = 0; r < 8; ++r) left += bucket[(size_t)lp * 8 + r].size();
            while (left > 0) {
              const std::vector<uint32_t> &b = bucket[(size_t)lp * 8 + res];
              if (idx[res] < b.size()) {
                gs.blocks[pos++] = b[idx[res]++];
                --left;
              }
              res = (res + 1) & 7;
            }
          }
          require(pos == gs.blocks.size(), "block reordering lost entries");
        }
        if (count > 0) {
          gs.seg_mp[gs.n_seg] = mp;
          gs.seg_begin[gs.n_seg] = begin;
          gs.seg_count[gs.n_seg] = count;
          ++gs.n_seg;
          G.has_blocks = true;
        }
      }
    }
    T.groups.push_back(std::move(G));
  }
  std::stable_sort(T.groups.begin(), T.groups.end(),
                   [](const PassGroupHost &a, const PassGroupHost &b) { return a.R > b.R; });
  return T;
}

DevicePassTables pack_device_tables(const PassSchedule &ps, PassTablesHost &T) {
  DevicePassTables D;
  std::vector<uint32_t> &u32 = D.u32;
  const int n_sub = (int)ps.subs.size();
  auto pad4 = [&]() {  // block lists are staged into shared memory with 16-byte copies
    while (u32.size() % 4) u32.push_back(0);
  };
  for (PassGroupHost &G : T.groups) {
    DevicePassTables::GroupOffsets go;
    go.tabrow_off = (uint32_t)u32.size();
    u32.insert(u32.end(), G.tabrow.begin(), G.tabrow.end());
    go.combo_base_off = (uint32_t)u32.size();
    u32.insert(u32.end(), G.combo_base.begin(), G.combo_base.end());
    go.combo_low_off = (uint32_t)D.u8.size();
    D.u8.insert(D.u8.end(), G.combo_low.begin(), G.combo_low.end());
    go.gsub_off = (uint32_t)D.gsub.size();
    for (int s = 0; s < n_sub; ++s) {
      GroupSubHost &gs = G.subs[s];
      GroupSubDev d;
      std::memset(&d, 0, sizeof(d));
      pad4();
      d.blocks_off = (uint32_t)u32.size();
      d.n_seg = gs.n_seg;
      d.n_blocks = (int)gs.blocks.size();
      D.blk_cap = std::max(D.blk_cap, (d.n_blocks + 3) & ~3);
      for (int k = 0; k < gs.n_seg && k < kMaxSeg; ++k) {
        d.seg[k].mp = gs.seg_mp[k];
        d.seg[k].begin = gs.seg_begin[k];
        d.seg[k].count = gs.seg_count[k];
        d.seg[k].inv_count = 0xFFFFFFFFu / (unsigned)std::max(1, gs.seg_count[k]) + 1u;
      }
      u32.insert(u32.end(), gs.blocks.begin(), gs.blocks.end());
      pad4();
      D.gsub.push_back(d);
      std::vector<uint32_t>().swap(gs.blocks);
    }
    D.goff.push_back(go);
    std::vector<uint32_t>().swap(G.tabrow);
  }
  // block-offset tables: byte offsets, every class starting on a 16-byte boundary
  D.off32.assign((size_t)n_sub * kMaxLowDev * kOffRowDev, 0);
  for (int s = 0; s < n_sub; ++s) {
    const int w = ps.subs[s].w;
    for (int lp = 0; lp < kMaxLow; ++lp)
      for (int mp = 1; mp < w; ++mp) {
        const int n = (int)binom(w, mp);
        for (int t = 0; t < n; ++t)
          D.off32[((size_t)s * kMaxLowDev + lp) * kOffRowDev + dev_class_offset(w, mp) + t] =
              16u * T.off[((size_t)s * kMaxLow + lp) * kOffRow + class_offset(w, mp) + t];
      }
  }
  return D;
}

}  // namespace ffb

using namespace ffb;

extern "C" {

int ffb_set_option(const char *key, int64_t value) {
  if (!key) return fail(FFB_EINVAL, "ffb_set_option: NULL key");
  std::lock_guard<std::mutex> lk(g_opt_mu);
  std::string k(key);
  if (k == "smem_bytes") {
    if (value < 4096 || value > 227 * 1024) return fail(FFB_EINVAL, "smem_bytes out of range");
    g_opt.smem_bytes = value;
  } else if (k == "min_cols") {
    if (value < 1 || value > 64) return fail(FFB_EINVAL, "min_cols out of range");
    g_opt.min_cols = (int)value;
  } else if (k == "max_cols") {
    if (value < 1 || value > 256) return fail(FFB_EINVAL, "max_cols out of range");
    g_opt.max_cols = (int)value;
  } else if (k == "sub_window") {
    if (value < 2 || value > kMaxSubWindow) return fail(FFB_EINVAL, "sub_window out of range");
    g_opt.sub_window = (int)value;
  } else if (k == "threads") {
    if (value < 32 || value > FFB_TPB || value % 32)
      return fail(FFB_EINVAL, "threads out of range (32.." + std::to_string(FFB_TPB) + ", the CTA size the kernel is built for)");
    g_opt.threads = (int)value;
  } else if (k == "bulk_copies") {
    if (value < 0 || value > 1) return fail(FFB_EINVAL, "bulk_copies out of range");
    g_opt.bulk_copies = (int)value;
  } else if (k == "beta_mode") {
    if (value < 0 || value > 3) return fail(FFB_EINVAL, "beta_mode out of range");
    g_opt.beta_mode = (int)value;
  } else {
    return fail(FFB_EINVAL, "ffb_set_option: unknown key " + k);
  }
  return FFB_OK;
}

int64_t ffb_get_option(const char *key) {
  if (!key) return -1;
  std::lock_guard<std::mutex> lk(g_opt_mu);
  std::string k(key);
  if (k == "smem_bytes") return g_opt.smem_bytes;
  if (k == "min_cols") return g_opt.min_cols;
  if (k == "max_cols") return g_opt.max_cols;
  if (k == "sub_window") return g_opt.sub_window;
  if (k == "threads") return g_opt.threads;
  if (k == "bulk_copies") return g_opt.bulk_copies;
  if (k == "beta_mode") return g_opt.beta_mode;
  return -1;
}

}  // extern "C"
