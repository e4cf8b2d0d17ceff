// C ABI: device state, plans and the entry points declared in include/ffsim_b200.h.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <sstream>

#include "device_structs.h"
#include "host.hpp"
#include "kernels.hpp"

using namespace ffb;

#ifdef FFB_DEBUG_KNOBS
namespace ffb {
void set_debug_knobs(int v);
}
#endif
#ifdef FFB_DEBUG_TIMING
namespace ffb {
void read_phase_cycles(unsigned long long *out, int reset);
}
#endif
#ifdef FFB_DEBUG_TIMELINE
namespace ffb {
void read_timeline(void *out, int *counts);
int timeline_capacity();
}
#endif

static_assert(ffb::kMaxLow == ffb::kMaxLowDev, "kMaxLow mismatch");

#define FFB_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (call);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return fail(FFB_ECUDA, std::string(#call) + ": " + cudaGetErrorString(_e));           \
  } while (0)

namespace {

// ---- optional launch profiling (bench.py): CUDA events around the hot kernels,
// recorded on the stream they are launched on.
enum { kProfFused = 0, kProfDiag = 1, kProfTranspose = 2, kProfOther = 3, kProfExchange = 4, kProfKinds = 5 };
const char *const kProfNames[kProfKinds] = {"fused_pass_kernel", "diag_kernel", "transpose_kernel", "other",
                                             "exchange_kernel"};
struct ProfState {
  bool enabled = false;
  std::mutex mu;
  struct Rec {
    cudaEvent_t a, b;
    int kind;
    double bytes, dfma;
  };
  std::vector<Rec> recs;
  long long launches[kProfKinds] = {0, 0, 0, 0, 0};
};
ProfState g_prof;

struct ProfScope {
  cudaStream_t st;
  cudaEvent_t a = nullptr, b = nullptr;
  int kind;
  double bytes, dfma;
  bool on;
  ProfScope(int kind_, double bytes_, cudaStream_t st_, double dfma_ = 0.0)
      : st(st_), kind(kind_), bytes(bytes_), dfma(dfma_) {
    g_prof.launches[kind] += 1;
    on = g_prof.enabled && kind != kProfOther;
    if (on) {
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      cudaEventRecord(a, st);
    }
  }
  ~ProfScope() {
    if (on) {
      cudaEventRecord(b, st);
      std::lock_guard<std::mutex> lk(g_prof.mu);
      g_prof.recs.push_back({a, b, kind, bytes, dfma});
    }
  }
};

struct DeviceInfo {
  int device = -1;
  int sm_count = 0;
  size_t smem_optin = 0;
};

int get_device_info(DeviceInfo *info) {
  int dev = 0;
  FFB_CUDA(cudaGetDevice(&dev));
  int sm = 0, optin = 0;
  FFB_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
  FFB_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  info->device = dev;
  info->sm_count = sm;
  info->smem_optin = (size_t)optin;
  return FFB_OK;
}

template <class T>
int upload(const std::vector<T> &host, T **dev) {
  *dev = nullptr;
  if (host.empty()) return FFB_OK;
  FFB_CUDA(cudaMalloc((void **)dev, host.size() * sizeof(T)));
  FFB_CUDA(cudaMemcpy(*dev, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  return FFB_OK;
}

// Device-resident index tables of one pass; shared between plans with the same
// (norb, nocc, pair sequence, options).
struct DevicePass {
  PassTablesHost host;  // kept for group metadata (block lists are cleared after upload)
  PassSchedule sched;
  uint32_t *d_u32 = nullptr;
  uint8_t *d_u8 = nullptr;
  GroupSubDev *d_gsub = nullptr;
  uint32_t *d_off = nullptr;
  int blk_cap = 4;  // longest block list of any (group, sub-pass), rounded up to 4 entries
  std::vector<DevicePassTables::GroupOffsets> goff;
  ~DevicePass() {
    cudaFree(d_u32);
    cudaFree(d_u8);
    cudaFree(d_gsub);
    cudaFree(d_off);
  }
};

struct SideStructure {
  SideSchedule sched;
  std::vector<std::unique_ptr<DevicePass>> passes;
};

std::mutex g_cache_mu;
std::map<std::string, std::shared_ptr<SideStructure>> g_cache;

int build_device_pass(int norb, int nocc, const PassSchedule &ps, std::unique_ptr<DevicePass> *out) {
  std::unique_ptr<DevicePass> dp(new DevicePass());
  dp->sched = ps;
  dp->host = build_pass_tables(norb, nocc, ps);
  DevicePassTables D = pack_device_tables(ps, dp->host);  // (plan.cpp; the host emulator reads the same)
  dp->goff = D.goff;
  dp->blk_cap = D.blk_cap;
  if (D.u32.size() >= (1ull << 32)) return fail(FFB_EINTERNAL, "pass tables exceed 32-bit offsets");
  int rc;
  if ((rc = upload(D.u32, &dp->d_u32)) != FFB_OK) return rc;
  if ((rc = upload(D.u8, &dp->d_u8)) != FFB_OK) return rc;
  if ((rc = upload(D.gsub, &dp->d_gsub)) != FFB_OK) return rc;
  if ((rc = upload(D.off32, &dp->d_off)) != FFB_OK) return rc;
  *out = std::move(dp);
  return FFB_OK;
}

int get_structure(int norb, int nocc, const std::vector<int> &q, const PlanOptions &opt, int device,
                  std::shared_ptr<SideStructure> *out) {
  std::ostringstream key;
  key << device << '/' << norb << '/' << nocc << '/' << opt.smem_bytes << '/' << opt.min_cols << '/'
      << opt.sub_window << ':';
  for (int v : q) key << (char)('A' + v);
  std::lock_guard<std::mutex> lk(g_cache_mu);
  auto it = g_cache.find(key.str());
  if (it != g_cache.end()) {
    *out = it->second;
    return FFB_OK;
  }
  std::shared_ptr<SideStructure> st(new SideStructure());
  try {
    st->sched = build_schedule(norb, nocc, q, opt);
    for (const PassSchedule &ps : st->sched.passes) {
      std::unique_ptr<DevicePass> dp;
      int rc = build_device_pass(norb, nocc, ps, &dp);
      if (rc != FFB_OK) return rc;
      st->passes.push_back(std::move(dp));
    }
  } catch (const std::exception &e) {  // plan-builder invariants (plan.cpp) and allocation failures
    return fail(FFB_EINTERNAL, e.what());
  }
  if (g_cache.size() > 64) g_cache.clear();  // bounded; live plans keep their own references
  g_cache[key.str()] = st;
  *out = st;
  return FFB_OK;
}

struct SidePlan {
  bool active = false;  // false: leave this spin sector alone
  ffb_tables *tables = nullptr;
  std::vector<NormRot> rots;
  bool has_phases = false;
  PhaseList phases;
  std::shared_ptr<SideStructure> structure;
  double2 *d_rowphase = nullptr;
  bool rowphase_ready = false;
};

int ensure_device_strings(ffb_tables *t) {
  std::lock_guard<std::mutex> lk(t->mu);
  int dev = 0;
  FFB_CUDA(cudaGetDevice(&dev));
  if (t->d_strings && t->device == dev) return FFB_OK;
  if (t->norb > kMaxNorb) return fail(FFB_EINVAL, "device kernels support norb <= 32");
  if (t->d_strings) {
    cudaFree(t->d_strings);
    t->d_strings = nullptr;
  }
  std::vector<uint32_t> s32(t->strings.size());
  for (size_t i = 0; i < s32.size(); ++i) s32[i] = (uint32_t)t->strings[i];
  int rc = upload(s32, &t->d_strings);
  if (rc != FFB_OK) return rc;
  t->device = dev;
  return FFB_OK;
}

}  // namespace

struct ffb_plan {
  SidePlan side[2];
  PlanOptions opt;
  DeviceInfo dev;
  int64_t dim_a = 0, dim_b = 0;
  bool beta_transposed = false;
  ~ffb_plan() {
    cudaFree(side[0].d_rowphase);
    cudaFree(side[1].d_rowphase);
  }
};

namespace {

int setup_side(ffb_plan *plan, int which, ffb_tables *t, const ffb_givens_rotation *rots, int n,
               const ffb_c128 *phases) {
  SidePlan &sp = plan->side[which];
  sp.tables = t;
  if (n < 0 || (rots == nullptr && phases == nullptr)) {
    sp.active = false;
    return FFB_OK;
  }
  sp.active = true;
  if (t->norb > kMaxNorb) return fail(FFB_EINVAL, "device kernels support norb <= 32");
  std::vector<int> q;
  for (int k = 0; k < n; ++k) {
    const ffb_givens_rotation &r = rots[k];
    if (r.i < 0 || r.j < 0 || r.i >= t->norb || r.j >= t->norb || std::abs(r.i - r.j) != 1)
      return fail(FFB_EINVAL, "ffb_plan_orbital_rotation: rotations must act on adjacent orbitals");
    sp.rots.push_back(normalise(r));
    q.push_back(sp.rots.back().q);
  }
  sp.has_phases = phases != nullptr;
  for (int i = 0; i < 32; ++i) {
    sp.phases.re[i] = 1.0;
    sp.phases.im[i] = 0.0;
  }
  if (phases)
    for (int i = 0; i < t->norb; ++i) {
      sp.phases.re[i] = phases[i].re;
      sp.phases.im[i] = phases[i].im;
    }
  // a sector with no strings that can rotate (empty or full shell) needs no schedule
  if (t->nocc == 0 || t->nocc == t->norb) q.clear(), sp.rots.clear();
  int rc = get_structure(t->norb, t->nocc, q, plan->opt, plan->dev.device, &sp.structure);
  if (rc != FFB_OK) return rc;
  if (sp.has_phases) {
    rc = ensure_device_strings(t);
    if (rc != FFB_OK) return rc;
    FFB_CUDA(cudaMalloc((void **)&sp.d_rowphase, std::max<int64_t>(t->dim, 1) * sizeof(double2)));
  }
  return FFB_OK;
}

// Rotate the string index of a (dim x n_cols) matrix: element (r, c) lives at
// data[r * row_stride + c * col_stride].
// The other layout of the same matrix, for the first / last pass of a beta-side rotation that works on
// a transposed copy: the first pass may read its tiles from `ptr` (instead of `data`) and the last
// pass may write them to `ptr`, which folds the two transpositions into those passes.
struct AltLayout {
  void *ptr = nullptr;
  int64_t row_stride = 0, col_stride = 0;
  bool read_first = false, write_last = false;
};

int apply_side(ffb_plan *plan, int which, void *data, int64_t n_cols, int64_t row_stride,
               int64_t col_stride, cudaStream_t stream, const AltLayout &alt = AltLayout()) {
  SidePlan &sp = plan->side[which];
  if (!sp.active || n_cols <= 0 || sp.tables->dim <= 0) return FFB_OK;
  const int64_t dim = sp.tables->dim;
  if (sp.has_phases && !sp.rowphase_ready) {
    ProfScope prof(kProfOther, 0.0, stream);
    FFB_CUDA(launch_row_phase(sp.tables->d_strings, dim, sp.tables->norb, sp.phases, sp.d_rowphase,
                              plan->dev.sm_count, stream));
    sp.rowphase_ready = true;
  }
  const size_t n_pass = sp.structure ? sp.structure->passes.size() : 0;
  if (n_pass == 0) {
    if (sp.has_phases) {
      ProfScope prof(kProfOther, 0.0, stream);
      FFB_CUDA(launch_row_scale(data, dim, n_cols, row_stride, col_stride, sp.d_rowphase,
                                plan->dev.sm_count, stream));
    }
    return FFB_OK;
  }
  static thread_local PassParams P;  // 16 KB: keep it off the stack
  for (size_t ip = 0; ip < n_pass; ++ip) {
    DevicePass &dp = *sp.structure->passes[ip];
    const bool last = ip + 1 == n_pass;
    int off_rows = 1;
    for (const SubPass &sub : dp.sched.subs) off_rows = std::max(off_rows, sub.q0 + 1);
    off_rows = std::min(off_rows, (int)kMaxLowDev);
    const size_t overhead = fused_pass_smem_overhead((int)dp.sched.subs.size(), dp.blk_cap, off_rows);
    if (overhead + 64 + 1024 > plan->dev.smem_optin)
      return fail(FFB_EINTERNAL, "pass tables do not fit in shared memory");
    size_t budget = std::min<size_t>((size_t)plan->opt.smem_bytes, plan->dev.smem_optin - overhead - 64);
    const int64_t budget_amps = (int64_t)(budget / 16);
    std::memset(&P, 0, sizeof(P));
    P.data = data;
    P.row_stride = row_stride;
    P.col_stride = col_stride;
    P.out = data;
    P.out_row_stride = row_stride;
    P.out_col_stride = col_stride;
    if (alt.ptr && alt.read_first && ip == 0) {
      P.data = alt.ptr;
      P.row_stride = alt.row_stride;
      P.col_stride = alt.col_stride;
    }
    if (alt.ptr && alt.write_last && last) {
      P.out = alt.ptr;
      P.out_row_stride = alt.row_stride;
      P.out_col_stride = alt.col_stride;
    }
    P.n_cols = n_cols;
    // a tile column is one contiguous run of memory when consecutive strings are adjacent and the window
    // starts at orbital 0 (tile row r is then string rowbase + r): those columns move as bulk copies
    const bool runs = plan->opt.bulk_copies != 0 && dp.sched.lo == 0;
    P.bulk_in = runs && P.row_stride == 1 && reinterpret_cast<uintptr_t>(P.data) % 16 == 0;
    P.bulk_out = runs && P.out_row_stride == 1 && reinterpret_cast<uintptr_t>(P.out) % 16 == 0;
    P.rowphase = (last && sp.has_phases) ? sp.d_rowphase : nullptr;
    P.u32 = dp.d_u32;
    P.u8 = dp.d_u8;
    P.gsub = dp.d_gsub;
    P.off = dp.d_off;
    P.n_sub = (int)dp.sched.subs.size();
    P.blk_cap = dp.blk_cap;
    P.off_rows = off_rows;
    P.n_rot = (int)dp.sched.rot_index.size();
    P.w = dp.sched.subs.empty() ? 2 : dp.sched.subs[0].w;
    if (P.n_sub > kMaxSubPerPass || P.n_rot > kMaxRotPerPass)
      return fail(FFB_EINTERNAL, "pass exceeds kernel parameter limits");
    for (int s = 0; s < P.n_sub; ++s) {
      P.sub[s].q0 = (unsigned char)dp.sched.subs[s].q0;
      P.sub[s].rot_begin = (unsigned short)dp.sched.subs[s].rot_begin;
      P.sub[s].rot_end = (unsigned short)dp.sched.subs[s].rot_end;
    }
    for (int r = 0; r < P.n_rot; ++r) {
      const NormRot &nr = sp.rots[dp.sched.rot_index[r]];
      P.rq[r] = (unsigned char)(nr.q - dp.sched.lo);
      P.rc[r] = nr.c;
      P.rsr[r] = nr.s.real();
      P.rsi[r] = nr.s.imag();
    }
    // dispatch units: descending runs inside each sub-pass
    int n_run = 0;
    for (int s = 0; s < P.n_sub; ++s) {
      const SubPass &sub = dp.sched.subs[s];
      P.sub[s].run_begin = (unsigned short)n_run;
      for (const Run &run : segment_runs(P.rq, sub.rot_begin, sub.rot_end, sub.q0)) {
        P.runcode[n_run] = (unsigned char)run_code(run.q_hi_rel, run.len);
        P.runrot[n_run] = (unsigned short)run.first;
        ++n_run;
      }
      P.sub[s].run_end = (unsigned short)n_run;
    }
    int ng = 0;
    long long units = 0;
    size_t tile_bytes = 0;
    for (size_t gi = 0; gi < dp.host.groups.size(); ++gi) {
      const PassGroupHost &G = dp.host.groups[gi];
      // nothing to do for these rows in this pass (an out-of-place pass still has to move them)
      if (!G.has_blocks && !P.rowphase && P.out == P.data) continue;
      if (ng >= kMaxGroups) return fail(FFB_EINTERNAL, "too many tile groups");
      int64_t fit = std::max<int64_t>(1, budget_amps / (G.R + 7));  // tile columns are up to R + 7 apart
      int64_t cols;
      if (col_stride == 1) {
        cols = fit >= 8 ? std::min<int64_t>(64, fit & ~7ll) : fit;
        if (fit >= 8 && cols > plan->opt.max_cols && G.R * 16ll * plan->opt.max_cols >= 32 * 1024)
          cols = std::max<int64_t>(8, plan->opt.max_cols & ~7);
      } else {
        cols = std::min<int64_t>(fit, 15);
        if (cols % 2 == 0 && cols > 1) --cols;  // odd strip: conflict-free transposing stores
      }
      cols = std::max<int64_t>(1, std::min<int64_t>(cols, n_cols));
      GroupLaunch &L = P.g[ng++];
      L.R = G.R;
      L.cols = (int)cols;
      L.Rp = tile_col_stride(G.R, (int)cols);
      L.inv_cols = 0xFFFFFFFFu / (unsigned)cols + 1u;
      L.inv_R = 0xFFFFFFFFu / (unsigned)G.R + 1u;
      L.n_combos = (int)G.combo_base.size();
      L.has_blocks = G.has_blocks ? 1 : 0;
      L.tabrow_off = dp.goff[gi].tabrow_off;
      L.combo_base_off = dp.goff[gi].combo_base_off;
      L.combo_low_off = dp.goff[gi].combo_low_off;
      L.gsub_off = dp.goff[gi].gsub_off;
      L.unit_begin = units;
      L.n_strips = (n_cols + cols - 1) / cols;
      units += L.n_strips * L.n_combos;
      tile_bytes = std::max(tile_bytes, (size_t)L.Rp * cols * 16);
    }
    P.n_groups = ng;
    P.total_units = units;
    if (units == 0) continue;
    // persistent grid: every CTA resident at once (registers, shared memory and threads counted)
    const int ctas_per_sm = fused_pass_ctas_per_sm(P.w, plan->opt.threads, tile_bytes + overhead);
    const int grid = (int)std::min<long long>(units, (long long)plan->dev.sm_count * ctas_per_sm);
    static const bool trace = std::getenv("FFB_TRACE_LAUNCH") != nullptr;  // developer aid
    if (trace)
      std::fprintf(stderr, "[ffb] fused pass %zu: w=%d subs=%d rots=%d groups=%d units=%lld grid=%d (%d CTA/SM) threads=%d smem=%zu (tile %zu + tables %zu) cols[0]=%d\n",
                   ip, P.w, P.n_sub, P.n_rot, ng, units, grid, ctas_per_sm, plan->opt.threads, tile_bytes + overhead,
                   tile_bytes, overhead, P.g[0].cols);
    {
      // FP64-pipe work of the pass: 4 DMUL + 8 DFMA per rotation and amplitude pair
      const double pairs = (double)binom(sp.tables->norb - 2, sp.tables->nocc - 1) * (double)n_cols;
      ProfScope prof(kProfFused, 32.0 * (double)dim * (double)n_cols, stream, 12.0 * P.n_rot * pairs);
      FFB_CUDA(launch_fused_pass(P, grid, plan->opt.threads, tile_bytes + overhead, stream));
    }
  }
  return FFB_OK;
}

// A beta-side rotation on a transposed copy needs no separate transposition before its first pass
// when that pass's window starts at orbital 0: its tiles are then contiguous runs of the beta index, so
// the pass reads them from the native layout (coalesced) and writes them transposed.  Likewise after
// the last pass.  (With a single pass both would hold and the rotation would simply be in place in the
// native layout, which the planner has already decided against.)
bool beta_first_fused(const ffb_plan *p) {
  const SidePlan &sb = p->side[1];
  if (!p->beta_transposed || p->opt.beta_mode == 3 || !sb.structure || sb.structure->passes.size() < 2) return false;
  return sb.structure->passes.front()->sched.lo == 0;
}
bool beta_last_fused(const ffb_plan *p) {
  const SidePlan &sb = p->side[1];
  if (!p->beta_transposed || p->opt.beta_mode == 3 || !sb.structure || sb.structure->passes.size() < 2) return false;
  return sb.structure->passes.back()->sched.lo == 0;
}

}  // namespace

extern "C" {

int ffb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

void ffb_tables_destroy(ffb_tables *t) {
  if (!t) return;
  if (t->d_strings) cudaFree(t->d_strings);
  for (void *p : t->scratch) cudaFree(p);
  delete t;
}

int ffb_plan_orbital_rotation(ffb_tables *tables_a, ffb_tables *tables_b,
                              const ffb_givens_rotation *rots_a, int n_a, const ffb_c128 *phases_a,
                              const ffb_givens_rotation *rots_b, int n_b, const ffb_c128 *phases_b,
                              ffb_plan **out) {
  if (!out) return fail(FFB_EINVAL, "ffb_plan_orbital_rotation: out is NULL");
  *out = nullptr;
  if (!tables_a || !tables_b) return fail(FFB_EINVAL, "ffb_plan_orbital_rotation: NULL tables");
  if (tables_a->norb != tables_b->norb)
    return fail(FFB_EINVAL, "ffb_plan_orbital_rotation: alpha and beta tables differ in norb");
  std::unique_ptr<ffb_plan> plan(new ffb_plan());
  plan->opt = current_options();
  int rc = get_device_info(&plan->dev);
  if (rc != FFB_OK) return rc;
  plan->dim_a = tables_a->dim;
  plan->dim_b = tables_b->dim;
  if ((rc = setup_side(plan.get(), 0, tables_a, rots_a, n_a, phases_a)) != FFB_OK) return rc;
  if ((rc = setup_side(plan.get(), 1, tables_b, rots_b, n_b, phases_b)) != FFB_OK) return rc;
  // beta side: rotate the contiguous index in place when the whole sector fits
  // one window with at least three columns, otherwise work on a transposed copy.
  const SidePlan &sb = plan->side[1];
  bool transposed = false;
  if (sb.active && sb.structure && !sb.structure->passes.empty()) {
    if (plan->opt.beta_mode >= 2) {
      transposed = true;
    } else if (plan->opt.beta_mode == 0) {
      const bool one_window = sb.structure->passes.size() == 1 &&
                              sb.structure->passes[0]->sched.W == tables_b->norb;
      int64_t fit = (plan->opt.smem_bytes / 16) / std::max<int64_t>(1, tables_b->dim);
      transposed = !(one_window && fit >= 3);
    }
  }
  plan->beta_transposed = transposed;
  *out = plan.release();
  return FFB_OK;
}

void ffb_plan_destroy(ffb_plan *p) { delete p; }

int ffb_set_device(int device) {
  FFB_CUDA(cudaSetDevice(device));
  return FFB_OK;
}

int ffb_plan_update_coefficients(ffb_plan *p, const ffb_givens_rotation *rots_a, int n_a,
                                 const ffb_c128 *phases_a, const ffb_givens_rotation *rots_b,
                                 int n_b, const ffb_c128 *phases_b) {
  if (!p) return fail(FFB_EINVAL, "ffb_plan_update_coefficients: NULL plan");
  const ffb_givens_rotation *rots[2] = {rots_a, rots_b};
  const int n[2] = {n_a, n_b};
  const ffb_c128 *ph[2] = {phases_a, phases_b};
  std::vector<NormRot> fresh[2];
  for (int s = 0; s < 2; ++s) {
    SidePlan &sp = p->side[s];
    const bool active = !(n[s] < 0 || (rots[s] == nullptr && ph[s] == nullptr));
    if (active != sp.active || (active && (ph[s] != nullptr) != sp.has_phases))
      return fail(FFB_EINVAL, "ffb_plan_update_coefficients: plan structure differs");
    if (!active) continue;
    const bool trivial = sp.tables->nocc == 0 || sp.tables->nocc == sp.tables->norb;
    if (!trivial) {
      if ((size_t)n[s] != sp.rots.size())
        return fail(FFB_EINVAL, "ffb_plan_update_coefficients: plan structure differs");
      for (int k = 0; k < n[s]; ++k) {
        if (std::abs(rots[s][k].i - rots[s][k].j) != 1)
          return fail(FFB_EINVAL, "ffb_plan_update_coefficients: non-adjacent rotation");
        NormRot nr = normalise(rots[s][k]);
        if (nr.q != sp.rots[k].q)
          return fail(FFB_EINVAL, "ffb_plan_update_coefficients: plan structure differs");
        fresh[s].push_back(nr);
      }
    }
  }
  for (int s = 0; s < 2; ++s) {
    SidePlan &sp = p->side[s];
    if (!sp.active) continue;
    sp.rots.swap(fresh[s]);
    if (ph[s])
      for (int i = 0; i < sp.tables->norb; ++i) {
        sp.phases.re[i] = ph[s][i].re;
        sp.phases.im[i] = ph[s][i].im;
      }
    sp.rowphase_ready = false;
  }
  return FFB_OK;
}

int64_t ffb_plan_workspace_bytes(const ffb_plan *p, int64_t n_rows_a) {
  if (!p || !p->beta_transposed) return 0;
  if (n_rows_a <= 0) n_rows_a = p->dim_a;
  return n_rows_a * p->dim_b * 16;
}

int ffb_plan_n_state_passes(const ffb_plan *p) {
  if (!p) return 0;
  int n = 0;
  for (int s = 0; s < 2; ++s) {
    const SidePlan &sp = p->side[s];
    if (!sp.active) continue;
    size_t np = sp.structure ? sp.structure->passes.size() : 0;
    n += np ? (int)np : (sp.has_phases ? 1 : 0);
  }
  if (p->beta_transposed) n += (beta_first_fused(p) ? 0 : 1) + (beta_last_fused(p) ? 0 : 1);
  return n;
}

int ffb_plan_describe(const ffb_plan *p, char *buf, size_t buflen) {
  if (!p || !buf || buflen == 0) return fail(FFB_EINVAL, "ffb_plan_describe: NULL argument");
  std::ostringstream os;
  for (int s = 0; s < 2; ++s) {
    const SidePlan &sp = p->side[s];
    os << (s == 0 ? "alpha" : "beta") << ":";
    if (!sp.active) {
      os << " inactive;";
      continue;
    }
    os << " rots=" << sp.rots.size() << " phases=" << (sp.has_phases ? 1 : 0);
    if (s == 1) os << " layout=" << (p->beta_transposed ? "transposed" : "native");
    if (sp.structure)
      for (const auto &dp : sp.structure->passes) {
        os << " [lo=" << dp->sched.lo << " W=" << dp->sched.W << " rots=" << dp->sched.rot_index.size()
           << " subs=" << dp->sched.subs.size() << "]";
      }
    os << ";";
  }
  std::string str = os.str();
  std::snprintf(buf, buflen, "%s", str.c_str());
  return FFB_OK;
}

int ffb_apply_orbital_rotation_rows(ffb_plan *p, int side, void *mat_dev, int64_t n_cols, int64_t ld,
                                    void *stream) {
  if (!p || (side != 0 && side != 1)) return fail(FFB_EINVAL, "ffb_apply_orbital_rotation_rows: bad argument");
  if (!mat_dev && n_cols > 0) return fail(FFB_EINVAL, "ffb_apply_orbital_rotation_rows: NULL matrix");
  if (p->side[side].active) {
    int rc = ensure_device_strings(p->side[side].tables);
    if (rc != FFB_OK) return rc;
  }
  return apply_side(p, side, mat_dev, n_cols, ld, 1, (cudaStream_t)stream);
}

int ffb_apply_orbital_rotation_strided(ffb_plan *p, int side, void *data_dev, int64_t n_batch,
                                       int64_t row_stride, int64_t col_stride, void *stream) {
  if (!p || (side != 0 && side != 1))
    return fail(FFB_EINVAL, "ffb_apply_orbital_rotation_strided: bad argument");
  if (!data_dev && n_batch > 0) return fail(FFB_EINVAL, "ffb_apply_orbital_rotation_strided: NULL data");
  if (row_stride != 1 && col_stride != 1)
    return fail(FFB_EINVAL, "ffb_apply_orbital_rotation_strided: one of the strides must be 1");
  if (p->side[side].active) {
    int rc = ensure_device_strings(p->side[side].tables);
    if (rc != FFB_OK) return rc;
  }
  return apply_side(p, side, data_dev, n_batch, row_stride, col_stride, (cudaStream_t)stream);
}

int ffb_plan_beta_in_place(const ffb_plan *p) { return p && !p->beta_transposed ? 1 : 0; }

int ffb_apply_orbital_rotation_beta_block(ffb_plan *p, void *block_dev, int64_t n_rows, int64_t ld,
                                          void *workspace_dev, void *stream) {
  if (!p) return fail(FFB_EINVAL, "ffb_apply_orbital_rotation_beta_block: NULL plan");
  if (n_rows < 0 || ld < p->dim_b) return fail(FFB_EINVAL, "ffb_apply_orbital_rotation_beta_block: bad block shape");
  if (!p->side[1].active || n_rows == 0 || p->dim_b == 0) return FFB_OK;
  if (!block_dev) return fail(FFB_EINVAL, "ffb_apply_orbital_rotation_beta_block: NULL block");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if ((rc = ensure_device_strings(p->side[1].tables)) != FFB_OK) return rc;
  if (!p->beta_transposed) return apply_side(p, 1, block_dev, n_rows, 1, ld, st);
  if (!workspace_dev) return fail(FFB_EINVAL, "ffb_apply_orbital_rotation_beta_block: this plan needs a workspace");
  const double tbytes = 32.0 * (double)n_rows * (double)p->dim_b;
  AltLayout alt;
  alt.ptr = block_dev;
  alt.row_stride = 1;  // consecutive beta strings are adjacent in the native layout
  alt.col_stride = ld;
  alt.read_first = beta_first_fused(p);
  alt.write_last = beta_last_fused(p);
  if (!alt.read_first) {
    ProfScope prof(kProfTranspose, tbytes, st);
    FFB_CUDA(launch_transpose(block_dev, workspace_dev, n_rows, p->dim_b, ld, n_rows, p->dev.sm_count, st));
  }
  if ((rc = apply_side(p, 1, workspace_dev, n_rows, n_rows, 1, st, alt)) != FFB_OK) return rc;
  if (!alt.write_last) {
    ProfScope prof(kProfTranspose, tbytes, st);
    FFB_CUDA(launch_transpose(workspace_dev, block_dev, p->dim_b, n_rows, n_rows, ld, p->dev.sm_count, st));
  }
  return FFB_OK;
}

int ffb_apply_orbital_rotation(ffb_plan *p, void *vec_dev, void *workspace_dev, void *stream) {
  if (!p) return fail(FFB_EINVAL, "ffb_apply_orbital_rotation: NULL plan");
  if (p->dim_a * p->dim_b == 0) return FFB_OK;
  if (!vec_dev) return fail(FFB_EINVAL, "ffb_apply_orbital_rotation: NULL state");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  for (int s = 0; s < 2; ++s)
    if (p->side[s].active && (rc = ensure_device_strings(p->side[s].tables)) != FFB_OK) return rc;
  if ((rc = apply_side(p, 0, vec_dev, p->dim_b, p->dim_b, 1, st)) != FFB_OK) return rc;
  return ffb_apply_orbital_rotation_beta_block(p, vec_dev, p->dim_a, p->dim_b, workspace_dev, stream);
}

// ---------------------------------------------------------------- _lib-level kernels

int ffb_apply_givens_rotation_in_place(void *vec_dev, int64_t dim_a, int64_t dim_b, int64_t ld,
                                       double c, ffb_c128 s, const uint64_t *slice1_dev,
                                       const uint64_t *slice2_dev, int64_t n_pairs, void *stream) {
  (void)dim_a;
  if (n_pairs == 0 || dim_b == 0) return FFB_OK;  // orbital_rotation.rs:27
  if (!vec_dev || !slice1_dev || !slice2_dev || n_pairs < 0)
    return fail(FFB_EINVAL, "ffb_apply_givens_rotation_in_place: bad argument");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != FFB_OK) return rc;
  FFB_CUDA(launch_givens_single(vec_dev, ld, dim_b, c, s.re, s.im,
                                (const unsigned long long *)slice1_dev,
                                (const unsigned long long *)slice2_dev, n_pairs, di.sm_count,
                                (cudaStream_t)stream));
  return FFB_OK;
}

int ffb_apply_phase_shift_in_place(void *vec_dev, int64_t dim_a, int64_t dim_b, int64_t ld,
                                   ffb_c128 phase, const uint64_t *indices_dev, int64_t n_indices,
                                   void *stream) {
  (void)dim_a;
  if (n_indices == 0 || dim_b == 0) return FFB_OK;
  if (!vec_dev || !indices_dev || n_indices < 0)
    return fail(FFB_EINVAL, "ffb_apply_phase_shift_in_place: bad argument");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != FFB_OK) return rc;
  FFB_CUDA(launch_phase_shift(vec_dev, ld, dim_b, phase.re, phase.im,
                              (const unsigned long long *)indices_dev, n_indices, di.sm_count,
                              (cudaStream_t)stream));
  return FFB_OK;
}

// ---------------------------------------------------------------- diagonal operators

}  // extern "C"

namespace {

// scratch buffer `slot` of a tables handle, grown on demand
int scratch(ffb_tables *t, int slot, size_t bytes, void **out) {
  std::lock_guard<std::mutex> lk(t->mu);
  if ((int)t->scratch.size() <= slot) {
    t->scratch.resize(slot + 1, nullptr);
    t->scratch_bytes.resize(slot + 1, 0);
  }
  if (t->scratch_bytes[slot] < bytes) {
    if (t->scratch[slot]) cudaFree(t->scratch[slot]);
    t->scratch[slot] = nullptr;
    t->scratch_bytes[slot] = 0;
    FFB_CUDA(cudaMalloc(&t->scratch[slot], bytes));
    t->scratch_bytes[slot] = bytes;
  }
  *out = t->scratch[slot];
  return FFB_OK;
}

enum { kSlotFactor = 0, kSlotMat = 1, kSlotMatAB = 2, kSlotPartial = 3 };

// shared body of the four diagonal entry points
int diag_op(bool contract, ffb_tables *ta, ffb_tables *tb, const void *m_aa, const void *m_ab,
            const void *m_bb, int zrep, const void *vec, void *out, int accumulate, int64_t row0,
            int64_t n_rows, int64_t col0, int64_t n_cols, int64_t ld, cudaStream_t st) {
  if (!ta || !tb) return fail(FFB_EINVAL, "diagonal operator: NULL tables");
  if (ta->norb != tb->norb) return fail(FFB_EINVAL, "diagonal operator: norb mismatch");
  if (row0 < 0 || n_rows < 0 || row0 + n_rows > ta->dim)
    return fail(FFB_EINVAL, "diagonal operator: row block outside the alpha sector");
  if (n_cols < 0) {  // the whole beta sector, contiguous rows
    col0 = 0;
    n_cols = tb->dim;
    ld = tb->dim;
  }
  if (col0 < 0 || col0 + n_cols > tb->dim || ld < n_cols)
    return fail(FFB_EINVAL, "diagonal operator: column block outside the beta sector");
  if (n_rows == 0 || n_cols == 0) return FFB_OK;
  if (!vec || !out) return fail(FFB_EINVAL, "diagonal operator: NULL state");
  int rc;
  if ((rc = ensure_device_strings(ta)) != FFB_OK) return rc;
  if ((rc = ensure_device_strings(tb)) != FFB_OK) return rc;
  DeviceInfo di;
  if ((rc = get_device_info(&di)) != FFB_OK) return rc;
  const int norb = ta->norb;
  if (norb == 0) m_aa = m_ab = m_bb = nullptr;  // empty matrices: identity factors
  const size_t elem = contract ? sizeof(double) : sizeof(double2);
  const size_t mat_bytes = (size_t)std::max(1, norb * norb) * elem;
  void *fa = nullptr, *fb = nullptr, *d_ab = nullptr;
  if (m_aa) {
    void *d_m;
    if ((rc = scratch(ta, kSlotMat, mat_bytes, &d_m)) != FFB_OK) return rc;
    if ((rc = scratch(ta, kSlotFactor, (size_t)ta->dim * elem, &fa)) != FFB_OK) return rc;
    FFB_CUDA(cudaMemcpyAsync(d_m, m_aa, (size_t)norb * norb * elem, cudaMemcpyHostToDevice, st));
    ProfScope prof(kProfOther, 0.0, st);
    FFB_CUDA(launch_side_factor(contract, ta->d_strings, ta->dim, norb, d_m, zrep, fa, di.sm_count, st));
  }
  if (m_bb) {
    // when both spins share one tables handle, slots must not collide with the alpha ones
    const int off = (ta == tb) ? 4 : 0;
    void *d_m;
    if ((rc = scratch(tb, kSlotMat + off, mat_bytes, &d_m)) != FFB_OK) return rc;
    if ((rc = scratch(tb, kSlotFactor + off, (size_t)tb->dim * elem, &fb)) != FFB_OK) return rc;
    FFB_CUDA(cudaMemcpyAsync(d_m, m_bb, (size_t)norb * norb * elem, cudaMemcpyHostToDevice, st));
    ProfScope prof(kProfOther, 0.0, st);
    FFB_CUDA(launch_side_factor(contract, tb->d_strings, tb->dim, norb, d_m, zrep, fb, di.sm_count, st));
  }
  if (m_ab) {
    if ((rc = scratch(ta, kSlotMatAB, mat_bytes, &d_ab)) != FFB_OK) return rc;
    FFB_CUDA(cudaMemcpyAsync(d_ab, m_ab, (size_t)norb * norb * elem, cudaMemcpyHostToDevice, st));
  }
  {
    const double amps = (double)n_rows * (double)n_cols;
    ProfScope prof(kProfDiag, (contract && accumulate ? 48.0 : 32.0) * amps, st);
    FFB_CUDA(launch_diag(contract, ta->d_strings, tb->d_strings, fa, fb, d_ab, vec, out, row0, n_rows,
                         col0, n_cols, ld, norb, zrep, accumulate, di.sm_count, st));
  }
  return FFB_OK;
}

}  // namespace

extern "C" {

int ffb_apply_diag_coulomb_evolution_block(ffb_tables *tables_a, ffb_tables *tables_b,
                                           const ffb_c128 *mat_exp_aa, const ffb_c128 *mat_exp_ab,
                                           const ffb_c128 *mat_exp_bb, int z_representation,
                                           void *vec_dev, int64_t row0, int64_t n_rows, int64_t col0,
                                           int64_t n_cols, int64_t ld, void *stream) {
  return diag_op(false, tables_a, tables_b, mat_exp_aa, mat_exp_ab, mat_exp_bb, z_representation,
                 vec_dev, vec_dev, 0, row0, n_rows, col0, n_cols, ld, (cudaStream_t)stream);
}

int ffb_apply_diag_coulomb_evolution(ffb_tables *tables_a, ffb_tables *tables_b,
                                     const ffb_c128 *mat_exp_aa, const ffb_c128 *mat_exp_ab,
                                     const ffb_c128 *mat_exp_bb, int z_representation,
                                     void *vec_dev, int64_t row0, int64_t n_rows, void *stream) {
  return ffb_apply_diag_coulomb_evolution_block(tables_a, tables_b, mat_exp_aa, mat_exp_ab, mat_exp_bb,
                                                z_representation, vec_dev, row0, n_rows, 0, -1, 0, stream);
}

int ffb_apply_num_op_sum_evolution(ffb_tables *tables_a, ffb_tables *tables_b,
                                   const ffb_c128 *phases_a, const ffb_c128 *phases_b,
                                   void *vec_dev, int64_t row0, int64_t n_rows, void *stream) {
  return ffb_apply_num_op_sum_evolution_block(tables_a, tables_b, phases_a, phases_b, vec_dev, row0, n_rows,
                                              0, -1, 0, stream);
}

int ffb_apply_num_op_sum_evolution_block(ffb_tables *tables_a, ffb_tables *tables_b,
                                         const ffb_c128 *phases_a, const ffb_c128 *phases_b,
                                         void *vec_dev, int64_t row0, int64_t n_rows, int64_t col0,
                                         int64_t n_cols, int64_t ld, void *stream) {
  if (!tables_a || !tables_b) return fail(FFB_EINVAL, "ffb_apply_num_op_sum_evolution: NULL tables");
  // prod_{i in occ} p_i == the same-spin factor of diag(p): M[j][k] = (j == k ? p_j : 1)
  const int norb = tables_a->norb;
  std::vector<ffb_c128> ma, mb;
  auto diag = [&](const ffb_c128 *p, std::vector<ffb_c128> &m) {
    m.assign((size_t)norb * norb, ffb_c128{1.0, 0.0});
    for (int i = 0; i < norb; ++i) m[(size_t)i * norb + i] = p[i];
  };
  if (phases_a) diag(phases_a, ma);
  if (phases_b) diag(phases_b, mb);
  if (!phases_a && !phases_b) return FFB_OK;
  return diag_op(false, tables_a, tables_b, phases_a ? ma.data() : nullptr, nullptr,
                 phases_b ? mb.data() : nullptr, 0, vec_dev, vec_dev, 0, row0, n_rows, col0, n_cols, ld,
                 (cudaStream_t)stream);
}

int ffb_contract_diag_coulomb(ffb_tables *tables_a, ffb_tables *tables_b, const double *mat_aa,
                              const double *mat_ab, const double *mat_bb, int z_representation,
                              const void *vec_dev, void *out_dev, int accumulate, int64_t row0,
                              int64_t n_rows, void *stream) {
  return ffb_contract_diag_coulomb_block(tables_a, tables_b, mat_aa, mat_ab, mat_bb, z_representation, vec_dev,
                                         out_dev, accumulate, row0, n_rows, 0, -1, 0, stream);
}

int ffb_contract_diag_coulomb_block(ffb_tables *tables_a, ffb_tables *tables_b, const double *mat_aa,
                                    const double *mat_ab, const double *mat_bb, int z_representation,
                                    const void *vec_dev, void *out_dev, int accumulate, int64_t row0,
                                    int64_t n_rows, int64_t col0, int64_t n_cols, int64_t ld,
                                    void *stream) {
  if (!tables_a) return fail(FFB_EINVAL, "ffb_contract_diag_coulomb: NULL tables");
  const int norb = tables_a->norb;
  std::vector<double> sa, sab, sb;
  if (z_representation) {  // the 0.25 of src/contract/diag_coulomb.rs:166-172, applied up front
    auto scaled = [&](const double *m, std::vector<double> &s) -> const double * {
      if (!m) return nullptr;
      s.assign(m, m + (size_t)norb * norb);
      for (double &v : s) v *= 0.25;
      return s.data();
    };
    mat_aa = scaled(mat_aa, sa);
    mat_ab = scaled(mat_ab, sab);
    mat_bb = scaled(mat_bb, sb);
  }
  return diag_op(true, tables_a, tables_b, mat_aa, mat_ab, mat_bb, z_representation, vec_dev,
                 out_dev, accumulate, row0, n_rows, col0, n_cols, ld, (cudaStream_t)stream);
}

int ffb_contract_num_op_sum(ffb_tables *tables_a, ffb_tables *tables_b, const double *coeffs_a,
                            const double *coeffs_b, const void *vec_dev, void *out_dev,
                            int accumulate, int64_t row0, int64_t n_rows, void *stream) {
  return ffb_contract_num_op_sum_block(tables_a, tables_b, coeffs_a, coeffs_b, vec_dev, out_dev, accumulate,
                                       row0, n_rows, 0, -1, 0, stream);
}

int ffb_contract_num_op_sum_block(ffb_tables *tables_a, ffb_tables *tables_b, const double *coeffs_a,
                                  const double *coeffs_b, const void *vec_dev, void *out_dev,
                                  int accumulate, int64_t row0, int64_t n_rows, int64_t col0,
                                  int64_t n_cols, int64_t ld, void *stream) {
  if (!tables_a || !tables_b) return fail(FFB_EINVAL, "ffb_contract_num_op_sum: NULL tables");
  const int norb = tables_a->norb;
  std::vector<double> ma, mb;
  auto diag = [&](const double *p, std::vector<double> &m) {
    m.assign((size_t)norb * norb, 0.0);
    for (int i = 0; i < norb; ++i) m[(size_t)i * norb + i] = p[i];
  };
  if (coeffs_a) diag(coeffs_a, ma);
  if (coeffs_b) diag(coeffs_b, mb);
  return diag_op(true, tables_a, tables_b, coeffs_a ? ma.data() : nullptr, nullptr,
                 coeffs_b ? mb.data() : nullptr, 0, vec_dev, out_dev, accumulate, row0, n_rows, col0, n_cols, ld,
                 (cudaStream_t)stream);
}

#ifdef FFB_DEBUG_TIMING
int ffb_debug_phase_cycles(unsigned long long *out, int reset) {
  ffb::read_phase_cycles(out, reset);
  return FFB_OK;
}
#endif

#ifdef FFB_DEBUG_TIMELINE
int ffb_debug_timeline(void *out, int *counts) {
  ffb::read_timeline(out, counts);
  return ffb::timeline_capacity();
}
#endif

#ifdef FFB_DEBUG_KNOBS
int ffb_debug_knobs(int v) {
  ffb::set_debug_knobs(v);
  return FFB_OK;
}
#endif

int ffb_profile_begin(void) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  for (auto &r : g_prof.recs) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.recs.clear();
  for (int k = 0; k < kProfKinds; ++k) g_prof.launches[k] = 0;
  g_prof.enabled = true;
  return FFB_OK;
}

int ffb_profile_end(char *buf, size_t buflen) {
  if (!buf || buflen == 0) return fail(FFB_EINVAL, "ffb_profile_end: NULL buffer");
  FFB_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_prof.mu);
  g_prof.enabled = false;
  double ms[kProfKinds] = {0, 0, 0, 0, 0}, bytes[kProfKinds] = {0, 0, 0, 0, 0}, dfma[kProfKinds] = {0, 0, 0, 0, 0};
  long long timed[kProfKinds] = {0, 0, 0, 0, 0};
  for (auto &r : g_prof.recs) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
      ms[r.kind] += t;
      bytes[r.kind] += r.bytes;
      dfma[r.kind] += r.dfma;
      timed[r.kind] += 1;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.recs.clear();
  std::ostringstream os;
  os.precision(10);
  os << "{";
  for (int k = 0; k < kProfKinds; ++k) {
    os << (k ? ", " : "") << "\"" << kProfNames[k] << "\": {\"launches\": " << g_prof.launches[k]
       << ", \"timed\": " << timed[k] << ", \"ms\": " << ms[k] << ", \"bytes\": " << bytes[k]
       << ", \"dfma_ops\": " << dfma[k] << "}";
  }
  os << "}";
  std::snprintf(buf, buflen, "%s", os.str().c_str());
  return FFB_OK;
}

int ffb_apply_num_op_prod_phase(ffb_tables *tables_a, ffb_tables *tables_b, uint32_t mask_a, uint32_t mask_b,
                                ffb_c128 phase, void *vec_dev, int64_t row0, int64_t n_rows, int64_t col0,
                                int64_t n_cols, int64_t ld, void *stream) {
  if (!tables_a || !tables_b) return fail(FFB_EINVAL, "ffb_apply_num_op_prod_phase: NULL tables");
  if (row0 < 0 || n_rows < 0 || row0 + n_rows > tables_a->dim)
    return fail(FFB_EINVAL, "ffb_apply_num_op_prod_phase: row block outside the alpha sector");
  if (n_cols < 0) {
    col0 = 0;
    n_cols = tables_b->dim;
    ld = tables_b->dim;
  }
  if (col0 < 0 || col0 + n_cols > tables_b->dim || ld < n_cols)
    return fail(FFB_EINVAL, "ffb_apply_num_op_prod_phase: column block outside the beta sector");
  if (n_rows == 0 || n_cols == 0) return FFB_OK;
  if (!vec_dev) return fail(FFB_EINVAL, "ffb_apply_num_op_prod_phase: NULL state");
  int rc;
  if ((rc = ensure_device_strings(tables_a)) != FFB_OK) return rc;
  if ((rc = ensure_device_strings(tables_b)) != FFB_OK) return rc;
  DeviceInfo di;
  if ((rc = get_device_info(&di)) != FFB_OK) return rc;
  ProfScope prof(kProfOther, 0.0, (cudaStream_t)stream);
  FFB_CUDA(launch_num_op_prod_phase(tables_a->d_strings, tables_b->d_strings, mask_a, mask_b, phase.re, phase.im,
                                    vec_dev, row0, n_rows, col0, n_cols, ld, di.sm_count, (cudaStream_t)stream));
  return FFB_OK;
}

int ffb_memcpy2d_async(void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t width_bytes,
                       size_t height, int kind, void *stream) {
  if (width_bytes == 0 || height == 0) return FFB_OK;
  if (!dst || !src || kind < 1 || kind > 3 || (height > 1 && (dst_pitch < width_bytes || src_pitch < width_bytes)))
    return fail(FFB_EINVAL, "ffb_memcpy2d_async: bad argument");
  const cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice : (kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice);
  if (height == 1)
    FFB_CUDA(cudaMemcpyAsync(dst, src, width_bytes, k, (cudaStream_t)stream));
  else
    FFB_CUDA(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, height, k, (cudaStream_t)stream));
  return FFB_OK;
}

int ffb_measure_fp64_peak(double *tflops) {
  if (!tflops) return fail(FFB_EINVAL, "ffb_measure_fp64_peak: NULL argument");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != FFB_OK) return rc;
  double best = 0.0;
  FFB_CUDA(measure_fp64_peak(di.sm_count, &best));
  *tflops = best;
  return FFB_OK;
}

int ffb_transpose(const void *in_dev, void *out_dev, int64_t n_rows, int64_t n_cols, int64_t ld_in,
                  int64_t ld_out, void *stream) {
  if (n_rows == 0 || n_cols == 0) return FFB_OK;
  if (!in_dev || !out_dev || n_rows < 0 || n_cols < 0 || in_dev == out_dev)
    return fail(FFB_EINVAL, "ffb_transpose: bad argument");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != FFB_OK) return rc;
  FFB_CUDA(launch_transpose(in_dev, out_dev, n_rows, n_cols, ld_in, ld_out, di.sm_count,
                            (cudaStream_t)stream));
  return FFB_OK;
}

int ffb_copy_blocks(const void *src_dev, int n_blocks, const int64_t *rows, const int64_t *width,
                    const int64_t *src_off, const int64_t *src_ld, void *const *dst_dev,
                    const int64_t *dst_off, const int64_t *dst_ld, void *stream) {
  if (n_blocks == 0) return FFB_OK;
  if (n_blocks < 0 || n_blocks > kMaxExchangeDst || !rows || !width || !src_off || !src_ld || !dst_dev ||
      !dst_off || !dst_ld)
    return fail(FFB_EINVAL, "ffb_copy_blocks: bad argument");
  ExchangeParams p;
  std::memset(&p, 0, sizeof(p));
  p.src = src_dev;
  p.n_dst = n_blocks;
  for (int d = 0; d < n_blocks; ++d) {
    if (rows[d] < 0 || width[d] < 0) return fail(FFB_EINVAL, "ffb_copy_blocks: negative block size");
    if (rows[d] > 0 && width[d] > 0 && (!src_dev || !dst_dev[d]))
      return fail(FFB_EINVAL, "ffb_copy_blocks: NULL buffer");
    p.rows[d] = width[d] > 0 ? rows[d] : 0;
    p.width[d] = width[d];
    p.src_off[d] = src_off[d];
    p.src_ld[d] = src_ld[d];
    p.dst[d] = dst_dev[d];
    p.dst_off[d] = dst_off[d];
    p.dst_ld[d] = dst_ld[d];
    p.max_rows = std::max<long long>(p.max_rows, p.rows[d]);
  }
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != FFB_OK) return rc;
  double bytes = 0.0;
  for (int d = 0; d < n_blocks; ++d) bytes += 32.0 * (double)p.rows[d] * (double)p.width[d];
  ProfScope prof(kProfExchange, bytes, (cudaStream_t)stream);
  FFB_CUDA(launch_exchange(p, di.sm_count, (cudaStream_t)stream));
  return FFB_OK;
}

int ffb_exchange_blocks(const void *src_dev, int64_t src_ld, int n_dst, const int64_t *rows,
                        const int64_t *width, const int64_t *src_off, void *const *dst_dev,
                        const int64_t *dst_off, const int64_t *dst_ld, void *stream) {
  if (n_dst < 0 || n_dst > kMaxExchangeDst) return fail(FFB_EINVAL, "ffb_exchange_blocks: bad argument");
  int64_t lds[kMaxExchangeDst];
  for (int d = 0; d < n_dst; ++d) lds[d] = src_ld;
  return ffb_copy_blocks(src_dev, n_dst, rows, width, src_off, lds, dst_dev, dst_off, dst_ld, stream);
}

int ffb_vdot(const void *x_dev, const void *y_dev, int64_t n, void *result_dev, void *stream) {
  if (!result_dev || n < 0 || (n > 0 && (!x_dev || !y_dev)))
    return fail(FFB_EINVAL, "ffb_vdot: bad argument");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != FFB_OK) return rc;
  static std::mutex mu;
  static void *partial[16] = {nullptr};
  const int n_partial = 2048;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (di.device < 0 || di.device >= 16) return fail(FFB_EINVAL, "ffb_vdot: device index");
    if (!partial[di.device]) FFB_CUDA(cudaMalloc(&partial[di.device], n_partial * 2 * sizeof(double)));
  }
  FFB_CUDA(launch_vdot(x_dev, y_dev, n, partial[di.device], n_partial, result_dev, di.sm_count,
                       (cudaStream_t)stream));
  return FFB_OK;
}

int ffb_axpby(ffb_c128 alpha, const void *x_dev, ffb_c128 beta, void *y_dev, int64_t n, void *stream) {
  if (n == 0) return FFB_OK;
  if (!x_dev || !y_dev || n < 0) return fail(FFB_EINVAL, "ffb_axpby: bad argument");
  DeviceInfo di;
  int rc = get_device_info(&di);
  if (rc != FFB_OK) return rc;
  ProfScope prof(kProfOther, 0.0, (cudaStream_t)stream);
  FFB_CUDA(launch_axpby(alpha.re, alpha.im, x_dev, beta.re, beta.im, y_dev, n, di.sm_count,
                        (cudaStream_t)stream));
  return FFB_OK;
}

}  // extern "C"
