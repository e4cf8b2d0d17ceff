// Occupation-string tables and the host-side Givens decomposition.
//
// Replaces python/ffsim/_cistring.py:21-42 (pyscf.fci.cistring.make_strings /
// gen_occslst, pyscf 2.14.0), the address tables of
// python/ffsim/gates/orbital_rotation.py:203-236 and _lib.givens_decomposition
// (src/linalg/givens.rs:20-149).  The tables are built directly from the
// combinatorial number system instead of by argsort: a string's address is its
// colexicographic rank, so the (i occupied, j empty) / (j occupied, i empty)
// partner lists are enumerated in closed form.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "host.hpp"

namespace ffb {

static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }
int fail(int code, const std::string &msg) {
  g_error = msg;
  return code;
}
const char *last_error() { return g_error.c_str(); }

namespace {
struct BinomTable {
  uint64_t v[65][65];
  BinomTable() {
    std::memset(v, 0, sizeof(v));
    for (int n = 0; n <= 64; ++n) {
      v[n][0] = 1;
      for (int k = 1; k <= n; ++k) {
        unsigned __int128 s = (unsigned __int128)v[n - 1][k - 1] + v[n - 1][k];
        v[n][k] = s > UINT64_MAX ? UINT64_MAX : (uint64_t)s;
      }
    }
  }
};
const BinomTable &bt() {
  static const BinomTable t;
  return t;
}
}  // namespace

uint64_t binom(int n, int k) {
  if (n < 0 || k < 0 || k > n || n > 64) return 0;
  return bt().v[n][k];
}

uint64_t rank_of(uint64_t s) {
  uint64_t r = 0;
  int idx = 0;
  while (s) {
    int pos = __builtin_ctzll(s);
    s &= s - 1;
    r += binom(pos, ++idx);
  }
  return r;
}

uint64_t unrank(uint64_t rank, int nocc) {
  uint64_t s = 0;
  for (int idx = nocc; idx >= 1; --idx) {
    int pos = idx - 1;
    while (binom(pos + 1, idx) <= rank) ++pos;
    s |= 1ull << pos;
    rank -= binom(pos, idx);
  }
  return s;
}

std::vector<uint64_t> strings_of(int nbits, int nocc) {
  std::vector<uint64_t> out;
  if (nocc < 0 || nocc > nbits) return out;
  if (nocc == 0) {
    out.push_back(0);
    return out;
  }
  uint64_t n = binom(nbits, nocc);
  out.reserve(n);
  uint64_t s = (nocc == 64) ? ~0ull : ((1ull << nocc) - 1);
  for (uint64_t i = 0; i < n; ++i) {
    out.push_back(s);
    if (i + 1 < n) s = next_same_popcount(s);
  }
  return out;
}

NormRot normalise(const ffb_givens_rotation &r) {
  // Reference call site (orbital_rotation.py:132-135) passes conj(s) and target
  // orbitals (i, j): rows with i occupied are "slice1".  For i > j swap roles.
  cplx s_applied = std::conj(cplx(r.s.re, r.s.im));
  NormRot out;
  out.c = r.c;
  if (r.i < r.j) {
    out.q = r.i;
    out.s = s_applied;
  } else {
    out.q = r.j;
    out.s = -std::conj(s_applied);
  }
  return out;
}

}  // namespace ffb

using namespace ffb;

// insert a zero bit at position `pos`
static inline uint64_t insert_zero(uint64_t s, int pos) {
  uint64_t low = s & ((1ull << pos) - 1);
  return ((s >> pos) << (pos + 1)) | low;
}

extern "C" {

int ffb_version(void) { return FFB_VERSION; }
const char *ffb_last_error(void) { return ffb::last_error(); }

int ffb_tables_create(int norb, int nocc, ffb_tables **out) {
  if (!out) return fail(FFB_EINVAL, "ffb_tables_create: out is NULL");
  *out = nullptr;
  if (norb < 0 || norb > 63) return fail(FFB_EINVAL, "ffb_tables_create: norb must be in [0, 63]");
  if (nocc < 0 || nocc > norb) return fail(FFB_EINVAL, "ffb_tables_create: nocc must be in [0, norb]");
  uint64_t dim = binom(norb, nocc);
  if (dim > (1ull << 31)) return fail(FFB_EINVAL, "ffb_tables_create: sector dimension exceeds 2^31");
  ffb_tables *t = new (std::nothrow) ffb_tables();
  if (!t) return fail(FFB_ENOMEM, "ffb_tables_create: out of memory");
  t->norb = norb;
  t->nocc = nocc;
  t->dim = (int64_t)dim;
  t->strings = strings_of(norb, nocc);
  *out = t;
  return FFB_OK;
}

int64_t ffb_tables_dim(const ffb_tables *t) { return t ? t->dim : 0; }
int ffb_tables_norb(const ffb_tables *t) { return t ? t->norb : 0; }
int ffb_tables_nocc(const ffb_tables *t) { return t ? t->nocc : 0; }

int ffb_tables_strings(const ffb_tables *t, int64_t *out) {
  if (!t || !out) return fail(FFB_EINVAL, "ffb_tables_strings: NULL argument");
  for (int64_t i = 0; i < t->dim; ++i) out[i] = (int64_t)t->strings[i];
  return FFB_OK;
}

int ffb_tables_occupations(const ffb_tables *t, uint64_t *out) {
  if (!t || (!out && t->nocc > 0)) return fail(FFB_EINVAL, "ffb_tables_occupations: NULL argument");
  for (int64_t i = 0; i < t->dim; ++i) {
    uint64_t s = t->strings[i];
    uint64_t *row = out + i * t->nocc;
    int k = 0;
    while (s) {
      row[k++] = (uint64_t)__builtin_ctzll(s);
      s &= s - 1;
    }
  }
  return FFB_OK;
}

int ffb_tables_strs2addr(const ffb_tables *t, const int64_t *strings, int64_t n, int64_t *out) {
  if (!t || (n > 0 && (!strings || !out))) return fail(FFB_EINVAL, "ffb_tables_strs2addr: NULL argument");
  for (int64_t i = 0; i < n; ++i) {
    uint64_t s = (uint64_t)strings[i];
    if (__builtin_popcountll(s) != t->nocc || (t->norb < 64 && (s >> t->norb)))
      return fail(FFB_EINVAL, "ffb_tables_strs2addr: string outside the sector");
    out[i] = (int64_t)rank_of(s);
  }
  return FFB_OK;
}

int64_t ffb_tables_n_pairs(const ffb_tables *t) {
  return t ? (int64_t)binom(t->norb - 2, t->nocc - 1) : 0;
}
int64_t ffb_tables_n_one(const ffb_tables *t) {
  return t ? (int64_t)binom(t->norb - 1, t->nocc - 1) : 0;
}

int ffb_tables_zero_one_subspace(const ffb_tables *t, int i, int j, uint64_t *out, int64_t *n_pairs) {
  if (!t) return fail(FFB_EINVAL, "ffb_tables_zero_one_subspace: NULL tables");
  if (i < 0 || j < 0 || i >= t->norb || j >= t->norb || i == j)
    return fail(FFB_EINVAL, "ffb_tables_zero_one_subspace: bad orbital pair");
  int64_t P = ffb_tables_n_pairs(t);
  if (n_pairs) *n_pairs = P;
  if (P == 0) return FFB_OK;
  if (!out) return fail(FFB_EINVAL, "ffb_tables_zero_one_subspace: out is NULL");
  int lo = std::min(i, j), hi = std::max(i, j);
  // spectators: nocc-1 electrons over the other norb-2 orbitals, ascending in
  // their compressed value; re-open the two target positions (low one first).
  std::vector<uint64_t> rest = strings_of(t->norb - 2, t->nocc - 1);
  for (int64_t k = 0; k < P; ++k) {
    uint64_t s = insert_zero(insert_zero(rest[k], lo), hi);
    out[k] = rank_of(s | (1ull << i));
    out[P + k] = rank_of(s | (1ull << j));
  }
  return FFB_OK;
}

int ffb_tables_one_subspace(const ffb_tables *t, int i, uint64_t *out, int64_t *n) {
  if (!t) return fail(FFB_EINVAL, "ffb_tables_one_subspace: NULL tables");
  if (i < 0 || i >= t->norb) return fail(FFB_EINVAL, "ffb_tables_one_subspace: bad orbital");
  int64_t K = ffb_tables_n_one(t);
  if (n) *n = K;
  if (K == 0) return FFB_OK;
  if (!out) return fail(FFB_EINVAL, "ffb_tables_one_subspace: out is NULL");
  std::vector<uint64_t> rest = strings_of(t->norb - 1, t->nocc - 1);
  for (int64_t k = 0; k < K; ++k) out[k] = rank_of(insert_zero(rest[k], i) | (1ull << i));
  return FFB_OK;
}

// ------------------------------------------------------------------ decomposition

namespace {
struct CS {
  double c;
  cplx s;
};
// BLAS zrotg with explicit handling of (near-)zero inputs, src/linalg/givens.rs:20-34
CS zrotg_safe(cplx a, cplx b, double tol) {
  double na = std::abs(a), nb = std::abs(b);
  if (nb <= tol) return {1.0, cplx(0.0, 0.0)};
  if (na <= tol) return {0.0, cplx(1.0, 0.0)};
  double r = std::hypot(na, nb);
  double c = na / r;
  cplx s = (a / na) * std::conj(b) / r;
  return {std::min(1.0, std::max(-1.0, c)), s};
}
}  // namespace

int ffb_givens_decomposition(const ffb_c128 *mat, int n, double tol, ffb_givens_rotation *rots,
                             int *n_rot, ffb_c128 *phases) {
  if (n < 0) return fail(FFB_EINVAL, "ffb_givens_decomposition: n < 0");
  if (n_rot) *n_rot = 0;
  if (n == 0) return FFB_OK;
  if (!mat || !phases || (!rots && n > 1) || !n_rot)
    return fail(FFB_EINVAL, "ffb_givens_decomposition: NULL argument");
  std::vector<cplx> m((size_t)n * n);
  for (size_t idx = 0; idx < m.size(); ++idx) m[idx] = cplx(mat[idx].re, mat[idx].im);
  auto at = [&](int r, int c) -> cplx & { return m[(size_t)r * n + c]; };

  struct Rot {
    double c;
    cplx s;
    int i, j;
  };
  std::vector<Rot> from_right, from_left;

  // Clements-style elimination: anti-diagonal sweeps, alternately by column
  // operations from the right and row operations from the left.
  for (int sweep = 0; sweep + 1 < n; ++sweep) {
    for (int step = 0; step <= sweep; ++step) {
      if (sweep % 2 == 0) {
        int col = sweep - step, row = n - 1 - step;
        if (std::abs(at(row, col)) > tol) {
          CS g = zrotg_safe(at(row, col + 1), at(row, col), tol);
          from_right.push_back({g.c, g.s, col + 1, col});
          for (int r = 0; r < n; ++r) {
            cplx x = at(r, col + 1), y = at(r, col);
            at(r, col + 1) = g.c * x + g.s * y;
            at(r, col) = g.c * y - std::conj(g.s) * x;
          }
        }
      } else {
        int row = n - 1 - sweep + step, col = step;
        if (std::abs(at(row, col)) > tol) {
          CS g = zrotg_safe(at(row - 1, col), at(row, col), tol);
          from_left.push_back({g.c, g.s, row - 1, row});
          for (int cc = 0; cc < n; ++cc) {
            cplx x = at(row - 1, cc), y = at(row, cc);
            at(row - 1, cc) = g.c * x + g.s * y;
            at(row, cc) = g.c * y - std::conj(g.s) * x;
          }
        }
      }
    }
  }
  // Move the left rotations to the right of the diagonal, last one first.
  for (auto it = from_left.rbegin(); it != from_left.rend(); ++it) {
    int i = it->i, j = it->j;
    cplx di = at(i, i), dj = at(j, j);
    CS g = zrotg_safe(it->c * dj, std::conj(it->s) * di, tol);
    from_right.push_back({g.c, -std::conj(g.s), i, j});
    cplx g00 = g.c * di, g01 = -g.s * dj, g10 = std::conj(g.s) * di, g11 = g.c * dj;
    CS h = zrotg_safe(g11, g10, tol);
    at(i, i) = g00 * h.c + g01 * (-std::conj(h.s));
    at(j, j) = g10 * h.s + g11 * h.c;
  }
  *n_rot = (int)from_right.size();
  for (size_t k = 0; k < from_right.size(); ++k) {
    rots[k].c = from_right[k].c;
    rots[k].s.re = from_right[k].s.real();
    rots[k].s.im = from_right[k].s.imag();
    rots[k].i = from_right[k].i;
    rots[k].j = from_right[k].j;
  }
  for (int d = 0; d < n; ++d) {
    phases[d].re = at(d, d).real();
    phases[d].im = at(d, d).imag();
  }
  return FFB_OK;
}

}  // extern "C"
