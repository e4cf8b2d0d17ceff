/* TEST INFRASTRUCTURE ONLY -- C restatement of the reference's Rust CPU kernels.
 *
 * Used (1) by tests as a second, independently written checker next to the numpy
 * oracle and (2) by bench.py as the CPU baseline ("port"): the reference's
 * ffsim._lib cannot be built here (no cargo/rustc), so these loops restate it
 * one-to-one, including its threading shape:
 *   - Givens: contiguous chunks of row pairs per thread
 *     (src/gates/orbital_rotation.rs:20-104; sequential below 128 pairs :51)
 *   - phase shift: sequential (src/gates/phase_shift.rs:18-30)
 *   - num-op-sum / diag-Coulomb evolution and contraction: parallel over alpha
 *     rows, three stages (src/gates/num_op_sum.rs:20-36,
 *     src/gates/diag_coulomb.rs:21-190, src/contract/diag_coulomb.rs:22-174,
 *     src/contract/num_op_sum.rs:20-39)
 * Nothing under ffsim_b200/ links or loads this file.
 */
#include <complex.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex c128;

int ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

static void rot_pair(c128 *vec, uint64_t i, uint64_t j, int64_t dim_b, double c, c128 s) {
  c128 *ri = vec + i * dim_b, *rj = vec + j * dim_b;
  const c128 sc = conj(s);
  for (int64_t k = 0; k < dim_b; ++k) {
    const c128 x = ri[k], y = rj[k];
    ri[k] = c * x + s * y;
    rj[k] = c * y - sc * x;
  }
}

void ref_apply_givens_rotation_in_place(c128 *vec, int64_t dim_b, double c, double s_re, double s_im,
                                        const uint64_t *slice1, const uint64_t *slice2,
                                        int64_t n_pairs, int n_threads) {
  if (n_pairs == 0) return;
  const c128 s = s_re + s_im * I;
  if (n_threads > n_pairs) n_threads = (int)n_pairs;
  if (n_threads <= 1 || n_pairs < 128) {
    for (int64_t k = 0; k < n_pairs; ++k) rot_pair(vec, slice1[k], slice2[k], dim_b, c, s);
    return;
  }
  const int64_t chunk = (n_pairs + n_threads - 1) / n_threads;
#pragma omp parallel for num_threads(n_threads) schedule(static, 1)
  for (int t = 0; t < n_threads; ++t) {
    int64_t start = t * chunk, end = start + chunk;
    if (end > n_pairs) end = n_pairs;
    for (int64_t k = start; k < end; ++k) rot_pair(vec, slice1[k], slice2[k], dim_b, c, s);
  }
}

void ref_apply_phase_shift_in_place(c128 *vec, int64_t dim_b, double p_re, double p_im,
                                    const uint64_t *indices, int64_t n) {
  const c128 p = p_re + p_im * I;
  for (int64_t k = 0; k < n; ++k) {
    c128 *row = vec + indices[k] * dim_b;
    for (int64_t b = 0; b < dim_b; ++b) row[b] *= p;
  }
}

/* vec may be a strided view (row_stride, col_stride in elements), as the reference
 * passes vec.T for the beta sector (python/ffsim/gates/num_op_sum.py:205-207). */
void ref_apply_num_op_sum_evolution_in_place(c128 *vec, int64_t n_rows, int64_t n_cols,
                                             int64_t row_stride, int64_t col_stride,
                                             const c128 *phases, const uint64_t *occ, int64_t nocc) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n_rows; ++r) {
    c128 p = 1.0;
    for (int64_t k = 0; k < nocc; ++k) p *= phases[occ[r * nocc + k]];
    c128 *row = vec + r * row_stride;
    for (int64_t b = 0; b < n_cols; ++b) row[b * col_stride] *= p;
  }
}

void ref_apply_diag_coulomb_evolution_in_place_num_rep(c128 *vec, int64_t dim_a, int64_t dim_b,
                                                       const c128 *aa, const c128 *ab, const c128 *bb,
                                                       int64_t norb, const uint64_t *occ_a,
                                                       int64_t n_alpha, const uint64_t *occ_b,
                                                       int64_t n_beta) {
  c128 *alpha = malloc(sizeof(c128) * (size_t)(dim_a > 0 ? dim_a : 1));
  c128 *beta = malloc(sizeof(c128) * (size_t)(dim_b > 0 ? dim_b : 1));
  c128 *pmap = malloc(sizeof(c128) * (size_t)(dim_a * norb > 0 ? dim_a * norb : 1));
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < dim_b; ++b) {
    c128 p = 1.0;
    const uint64_t *o = occ_b + b * n_beta;
    for (int64_t j = 0; j < n_beta; ++j)
      for (int64_t k = j; k < n_beta; ++k) p *= bb[o[j] * norb + o[k]];
    beta[b] = p;
  }
#pragma omp parallel for schedule(static)
  for (int64_t a = 0; a < dim_a; ++a) {
    c128 p = 1.0;
    const uint64_t *o = occ_a + a * n_alpha;
    c128 *row = pmap + a * norb;
    for (int64_t q = 0; q < norb; ++q) row[q] = 1.0;
    for (int64_t j = 0; j < n_alpha; ++j) {
      for (int64_t q = 0; q < norb; ++q) row[q] *= ab[o[j] * norb + q];
      for (int64_t k = j; k < n_alpha; ++k) p *= aa[o[j] * norb + o[k]];
    }
    alpha[a] = p;
  }
#pragma omp parallel for schedule(static)
  for (int64_t a = 0; a < dim_a; ++a) {
    const c128 *row = pmap + a * norb;
    c128 *v = vec + a * dim_b;
    for (int64_t b = 0; b < dim_b; ++b) {
      c128 p = alpha[a] * beta[b];
      const uint64_t *o = occ_b + b * n_beta;
      for (int64_t k = 0; k < n_beta; ++k) p *= row[o[k]];
      v[b] *= p;
    }
  }
  free(alpha);
  free(beta);
  free(pmap);
}

void ref_apply_diag_coulomb_evolution_in_place_z_rep(c128 *vec, int64_t dim_a, int64_t dim_b,
                                                     const c128 *aa, const c128 *ab, const c128 *bb,
                                                     int64_t norb, const int64_t *str_a,
                                                     const int64_t *str_b) {
  c128 *alpha = malloc(sizeof(c128) * (size_t)(dim_a > 0 ? dim_a : 1));
  c128 *beta = malloc(sizeof(c128) * (size_t)(dim_b > 0 ? dim_b : 1));
  c128 *pmap = malloc(sizeof(c128) * (size_t)(dim_a * norb > 0 ? dim_a * norb : 1));
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < dim_b; ++b) {
    c128 p = 1.0;
    const int64_t s = str_b[b];
    for (int64_t j = 0; j < norb; ++j)
      for (int64_t k = j + 1; k < norb; ++k) {
        const c128 m = bb[j * norb + k];
        p *= (((s >> j) ^ (s >> k)) & 1) ? conj(m) : m;
      }
    beta[b] = p;
  }
#pragma omp parallel for schedule(static)
  for (int64_t a = 0; a < dim_a; ++a) {
    c128 p = 1.0;
    const int64_t s = str_a[a];
    c128 *row = pmap + a * norb;
    for (int64_t q = 0; q < norb; ++q) row[q] = 1.0;
    for (int64_t j = 0; j < norb; ++j) {
      const int sign_j = (s >> j) & 1;
      for (int64_t q = 0; q < norb; ++q) {
        const c128 m = ab[j * norb + q];
        row[q] *= sign_j ? conj(m) : m;
      }
      for (int64_t k = j + 1; k < norb; ++k) {
        const c128 m = aa[j * norb + k];
        p *= (((s >> j) ^ (s >> k)) & 1) ? conj(m) : m;
      }
    }
    alpha[a] = p;
  }
#pragma omp parallel for schedule(static)
  for (int64_t a = 0; a < dim_a; ++a) {
    const c128 *row = pmap + a * norb;
    c128 *v = vec + a * dim_b;
    for (int64_t b = 0; b < dim_b; ++b) {
      c128 p = alpha[a] * beta[b];
      const int64_t s = str_b[b];
      for (int64_t j = 0; j < norb; ++j) p *= ((s >> j) & 1) ? conj(row[j]) : row[j];
      v[b] *= p;
    }
  }
  free(alpha);
  free(beta);
  free(pmap);
}

void ref_contract_num_op_sum_spin_into_buffer(const c128 *vec, int64_t n_rows, int64_t n_cols,
                                              int64_t row_stride, int64_t col_stride,
                                              const double *coeffs, const uint64_t *occ,
                                              int64_t nocc, c128 *out) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n_rows; ++r) {
    double c = 0.0;
    for (int64_t k = 0; k < nocc; ++k) c += coeffs[occ[r * nocc + k]];
    const c128 *src = vec + r * row_stride;
    c128 *dst = out + r * row_stride;
    for (int64_t b = 0; b < n_cols; ++b) dst[b * col_stride] += c * src[b * col_stride];
  }
}

void ref_contract_diag_coulomb_into_buffer_num_rep(const c128 *vec, int64_t dim_a, int64_t dim_b,
                                                   const double *aa, const double *ab,
                                                   const double *bb, int64_t norb,
                                                   const uint64_t *occ_a, int64_t n_alpha,
                                                   const uint64_t *occ_b, int64_t n_beta, c128 *out) {
  double *alpha = malloc(sizeof(double) * (size_t)(dim_a > 0 ? dim_a : 1));
  double *beta = malloc(sizeof(double) * (size_t)(dim_b > 0 ? dim_b : 1));
  double *cmap = malloc(sizeof(double) * (size_t)(dim_a * norb > 0 ? dim_a * norb : 1));
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < dim_b; ++b) {
    double c = 0.0;
    const uint64_t *o = occ_b + b * n_beta;
    for (int64_t j = 0; j < n_beta; ++j)
      for (int64_t k = j; k < n_beta; ++k) c += bb[o[j] * norb + o[k]];
    beta[b] = c;
  }
#pragma omp parallel for schedule(static)
  for (int64_t a = 0; a < dim_a; ++a) {
    double c = 0.0;
    const uint64_t *o = occ_a + a * n_alpha;
    double *row = cmap + a * norb;
    for (int64_t q = 0; q < norb; ++q) row[q] = 0.0;
    for (int64_t j = 0; j < n_alpha; ++j) {
      for (int64_t q = 0; q < norb; ++q) row[q] += ab[o[j] * norb + q];
      for (int64_t k = j; k < n_alpha; ++k) c += aa[o[j] * norb + o[k]];
    }
    alpha[a] = c;
  }
#pragma omp parallel for schedule(static)
  for (int64_t a = 0; a < dim_a; ++a) {
    const double *row = cmap + a * norb;
    for (int64_t b = 0; b < dim_b; ++b) {
      double c = alpha[a] + beta[b];
      const uint64_t *o = occ_b + b * n_beta;
      for (int64_t k = 0; k < n_beta; ++k) c += row[o[k]];
      out[a * dim_b + b] += c * vec[a * dim_b + b];
    }
  }
  free(alpha);
  free(beta);
  free(cmap);
}

void ref_contract_diag_coulomb_into_buffer_z_rep(const c128 *vec, int64_t dim_a, int64_t dim_b,
                                                 const double *aa, const double *ab, const double *bb,
                                                 int64_t norb, const int64_t *str_a,
                                                 const int64_t *str_b, c128 *out) {
  double *alpha = malloc(sizeof(double) * (size_t)(dim_a > 0 ? dim_a : 1));
  double *beta = malloc(sizeof(double) * (size_t)(dim_b > 0 ? dim_b : 1));
  double *cmap = malloc(sizeof(double) * (size_t)(dim_a * norb > 0 ? dim_a * norb : 1));
#define ZSIGN(s, j) ((((s) >> (j)) & 1) ? -1.0 : 1.0)
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < dim_b; ++b) {
    double c = 0.0;
    const int64_t s = str_b[b];
    for (int64_t j = 0; j < norb; ++j)
      for (int64_t k = j + 1; k < norb; ++k) c += ZSIGN(s, j) * ZSIGN(s, k) * bb[j * norb + k];
    beta[b] = c;
  }
#pragma omp parallel for schedule(static)
  for (int64_t a = 0; a < dim_a; ++a) {
    double c = 0.0;
    const int64_t s = str_a[a];
    double *row = cmap + a * norb;
    for (int64_t q = 0; q < norb; ++q) row[q] = 0.0;
    for (int64_t j = 0; j < norb; ++j) {
      for (int64_t q = 0; q < norb; ++q) row[q] += ZSIGN(s, j) * ab[j * norb + q];
      for (int64_t k = j + 1; k < norb; ++k) c += ZSIGN(s, j) * ZSIGN(s, k) * aa[j * norb + k];
    }
    alpha[a] = c;
  }
#pragma omp parallel for schedule(static)
  for (int64_t a = 0; a < dim_a; ++a) {
    const double *row = cmap + a * norb;
    for (int64_t b = 0; b < dim_b; ++b) {
      double c = alpha[a] + beta[b];
      const int64_t s = str_b[b];
      for (int64_t j = 0; j < norb; ++j) c += ZSIGN(s, j) * row[j];
      out[a * dim_b + b] += 0.25 * c * vec[a * dim_b + b];
    }
  }
#undef ZSIGN
  free(alpha);
  free(beta);
  free(cmap);
}

/* out[c, r] = in[r, c]: the np.ascontiguousarray(vec.T) copies of
 * python/ffsim/gates/orbital_rotation.py:143,154 */
void ref_transpose(const c128 *in, c128 *out, int64_t n_rows, int64_t n_cols) {
  const int64_t B = 32;
#pragma omp parallel for schedule(static) collapse(2)
  for (int64_t r0 = 0; r0 < n_rows; r0 += B)
    for (int64_t c0 = 0; c0 < n_cols; c0 += B)
      for (int64_t r = r0; r < r0 + B && r < n_rows; ++r)
        for (int64_t c = c0; c < c0 + B && c < n_cols; ++c) out[c * n_rows + r] = in[r * n_cols + c];
}
