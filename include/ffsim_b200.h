/*
 * ffsim_b200.h -- C ABI of the B200-native determinant-space statevector hot path.
 *
 * This is the drop-in boundary for the path SURVEY.md section 8 names.  Every
 * entry point says which reference interface it replaces; paths are relative to
 * the reference tree (qiskit-community/ffsim 0.0.85.dev).
 *
 * Conventions
 *   - plain C, no exceptions: every function returns FFB_OK (0) or a negative
 *     FFB_E* code; ffb_last_error() returns the thread-local message.
 *   - the state is complex128, row-major (dim_a x dim_b): row = alpha string,
 *     column = beta string (python/ffsim/states/dimensions.py:18-50).
 *   - pointers named *_dev are device pointers; everything else is host memory.
 *     The caller owns every buffer.  `stream` is a cudaStream_t passed as void*
 *     (NULL = default stream).  Calls are asynchronous with respect to the host.
 *   - a handle is safe to use from one thread / one stream at a time.
 *   - there is no CPU fallback: device entry points return FFB_ECUDA when no
 *     CUDA device is usable.
 */
#ifndef FFSIM_B200_H
#define FFSIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFB_VERSION 100 /* 0.1.0 */

enum {
  FFB_OK = 0,
  FFB_EINVAL = -1, /* bad argument (ValueError / TypeError on the reference side) */
  FFB_ECUDA = -2,  /* CUDA runtime error, or no device */
  FFB_ENOMEM = -3,
  FFB_EINTERNAL = -4
};

typedef struct ffb_c128 {
  double re, im;
} ffb_c128;

/* One Givens rotation as the reference's decomposition returns it: the tuple
 * (c, s, i, j) of src/linalg/givens.rs:16, |i-j| == 1. */
typedef struct ffb_givens_rotation {
  double c;
  ffb_c128 s;
  int32_t i, j;
} ffb_givens_rotation;

typedef struct ffb_tables ffb_tables; /* string tables of one spin sector */
typedef struct ffb_plan ffb_plan;     /* fused orbital-rotation schedule */

int ffb_version(void);
const char *ffb_last_error(void);
/* Number of usable CUDA devices (0 when there is none; never fails). */
int ffb_device_count(void);
/* Make `device` current for the calling thread (cudaSetDevice). */
int ffb_set_device(int device);

/* ------------------------------------------------------------------ tables
 * Replaces python/ffsim/_cistring.py:21-42 (pyscf.fci.cistring make_strings /
 * gen_occslst) and the cached address tables of
 * python/ffsim/gates/orbital_rotation.py:203-236.  Host side is built at
 * creation; the device copy is made on first use by a device entry point. */
/* Limits: 0 <= nocc <= norb <= 64 on the host side, norb <= 32 for every device entry point (strings
 * are held in 32 bits there) and fewer than 2^31 strings per sector; violations return FFB_EINVAL.  A
 * handle owns device buffers (strings, per-call scratch): use one handle per device and per stream. */
int ffb_tables_create(int norb, int nocc, ffb_tables **out);
void ffb_tables_destroy(ffb_tables *t);
int64_t ffb_tables_dim(const ffb_tables *t);
int ffb_tables_norb(const ffb_tables *t);
int ffb_tables_nocc(const ffb_tables *t);
/* strings: int64[dim], ascending (== cistring.make_strings(range(norb), nocc)). */
int ffb_tables_strings(const ffb_tables *t, int64_t *out);
/* occupations: uint64[dim * nocc] row-major (== gen_occslst cast to np.uint). */
int ffb_tables_occupations(const ffb_tables *t, uint64_t *out);
/* address of each string (colexicographic rank; pyscf strs2addr). */
int ffb_tables_strs2addr(const ffb_tables *t, const int64_t *strings, int64_t n, int64_t *out);
/* _zero_one_subspace_indices(norb, nocc, (i, j)) (orbital_rotation.py:203-213):
 * writes 2*P addresses, first half = orbital i occupied & j empty, second half
 * = j occupied & i empty, pair-aligned.  *n_pairs = P = C(norb-2, nocc-1). */
int64_t ffb_tables_n_pairs(const ffb_tables *t);
int ffb_tables_zero_one_subspace(const ffb_tables *t, int i, int j, uint64_t *out, int64_t *n_pairs);
/* _one_subspace_indices(norb, nocc, (i,)) (orbital_rotation.py:216-226). */
int64_t ffb_tables_n_one(const ffb_tables *t);
int ffb_tables_one_subspace(const ffb_tables *t, int i, uint64_t *out, int64_t *n);

/* ------------------------------------------------------- host decomposition
 * Replaces _lib.givens_decomposition (src/linalg/givens.rs:71-149,
 * python/ffsim/linalg/givens.py:59-156).  mat: row-major n x n.  rots must have
 * room for n(n-1)/2 entries; phases for n.  Non-square input cannot be
 * expressed here (the Python wrapper raises ValueError as givens.rs:78-80). */
int ffb_givens_decomposition(const ffb_c128 *mat, int n, double tol, ffb_givens_rotation *rots,
                             int *n_rot, ffb_c128 *phases);

/* ------------------------------------------------- _lib-level device kernels
 * One launch per call, same arithmetic as the Rust kernels.  `ld` is the row
 * stride of vec in elements (dim_b for a contiguous state). */

/* src/gates/orbital_rotation.rs:20 apply_givens_rotation_in_place */
int ffb_apply_givens_rotation_in_place(void *vec_dev, int64_t dim_a, int64_t dim_b, int64_t ld,
                                       double c, ffb_c128 s, const uint64_t *slice1_dev,
                                       const uint64_t *slice2_dev, int64_t n_pairs, void *stream);
/* src/gates/phase_shift.rs:18 apply_phase_shift_in_place */
int ffb_apply_phase_shift_in_place(void *vec_dev, int64_t dim_a, int64_t dim_b, int64_t ld,
                                   ffb_c128 phase, const uint64_t *indices_dev, int64_t n_indices,
                                   void *stream);

/* ------------------------------------------------ fused orbital rotation
 * Replaces the two loops of _apply_orbital_rotation_spinful
 * (python/ffsim/gates/orbital_rotation.py:117-154): all Givens rotations and
 * phase shifts of both spin sectors, fused into a few passes over the state.
 *
 * rots_x / phases_x are the output of ffb_givens_decomposition for that spin
 * (n_x < 0 or rots_x == NULL && phases_x == NULL means "leave this spin alone",
 * the `None` member of the reference's mat tuple).  The plan caches the pass
 * structure, which depends only on (norb, nocc, orbital pairs); it can be
 * re-used with new coefficients through ffb_plan_update_coefficients. */
int ffb_plan_orbital_rotation(ffb_tables *tables_a, ffb_tables *tables_b,
                              const ffb_givens_rotation *rots_a, int n_a, const ffb_c128 *phases_a,
                              const ffb_givens_rotation *rots_b, int n_b, const ffb_c128 *phases_b,
                              ffb_plan **out);
void ffb_plan_destroy(ffb_plan *p);
/* Replace the coefficients (c, s, phases) of an existing plan.  The orbital pairs
 * and the active/inactive state of each spin must be those the plan was built
 * with, otherwise FFB_EINVAL is returned and the plan is left unchanged. */
int ffb_plan_update_coefficients(ffb_plan *p, const ffb_givens_rotation *rots_a, int n_a,
                                 const ffb_c128 *phases_a, const ffb_givens_rotation *rots_b,
                                 int n_b, const ffb_c128 *phases_b);
/* Bytes of device workspace ffb_apply_orbital_rotation needs for this plan and
 * this many locally held alpha rows (0 = none). */
int64_t ffb_plan_workspace_bytes(const ffb_plan *p, int64_t n_rows_a);
/* Introspection for tests / DESIGN.md: number of passes per side, and the
 * number of state passes (read+write sweeps) the whole op performs. */
int ffb_plan_describe(const ffb_plan *p, char *buf, size_t buflen);
int ffb_plan_n_state_passes(const ffb_plan *p);
/* Apply to a contiguous (dim_a x dim_b) state on the device, in place. */
int ffb_apply_orbital_rotation(ffb_plan *p, void *vec_dev, void *workspace_dev, void *stream);
/* One spin side only, acting on the ROW index of an arbitrary row-major
 * (n_rows_total == tables dim) x n_cols matrix with row stride ld: the building
 * block of the row-sharded multi-GPU path (local columns = any slice of the
 * other index).  side: 0 = alpha rotations of the plan, 1 = beta rotations. */
int ffb_apply_orbital_rotation_rows(ffb_plan *p, int side, void *mat_dev, int64_t n_cols, int64_t ld,
                                    void *stream);
/* The same with general strides: element (string address r, batch index c) lives at
 * data[r * row_stride + c * col_stride].  With row_stride = 1 and col_stride = dim_b this
 * rotates the beta (contiguous) index of n_batch locally held alpha rows in place, without a
 * transposed copy; the kernel then stages whole rows and is efficient when the beta sector
 * fits one shared-memory window (ffb_plan_beta_in_place != 0). */
int ffb_apply_orbital_rotation_strided(ffb_plan *p, int side, void *data_dev, int64_t n_batch,
                                       int64_t row_stride, int64_t col_stride, void *stream);
/* 1 when the plan rotates the beta index in place, 0 when it goes through a transposed copy. */
int ffb_plan_beta_in_place(const ffb_plan *p);
/* Beta-side rotations of the plan on a block of n_rows locally held alpha rows (row-major,
 * n_rows x dim_b, row stride ld): what python/ffsim/gates/orbital_rotation.py:139-150 does with a
 * transposed view.  In place when ffb_plan_beta_in_place; otherwise through workspace_dev
 * (n_rows * dim_b elements, laid out dim_b x n_rows).  The transpositions are folded into the first
 * and the last pass when their windows start at orbital 0 (the tiles of such a pass are contiguous
 * runs of the beta index, so it can read or write the native layout directly); a separate transpose
 * kernel runs only where that does not hold (or always with option beta_mode = 3). */
int ffb_apply_orbital_rotation_beta_block(ffb_plan *p, void *block_dev, int64_t n_rows, int64_t ld,
                                          void *workspace_dev, void *stream);

/* ------------------------------------------------------ diagonal operators
 * Replaces src/gates/diag_coulomb.rs:21,95 (num / z representation),
 * src/gates/num_op_sum.rs:20, src/contract/diag_coulomb.rs:22,100 and
 * src/contract/num_op_sum.rs:20.  Matrices are host pointers, row-major
 * norb x norb; NULL means "all ones" (evolution) / "all zeros" (contraction).
 * row0 / n_rows select a contiguous block of alpha rows held locally
 * (row0 = 0, n_rows = dim_a for the whole state); vec_dev points at that block.
 */

/* vec[a,b] *= aphase(a) * bphase(b) * prod_{i in a, j in b} mat_exp_ab[i][j]
 * (z representation: conjugate selected by the string bits, all orbitals). */
int ffb_apply_diag_coulomb_evolution(ffb_tables *tables_a, ffb_tables *tables_b,
                                     const ffb_c128 *mat_exp_aa, const ffb_c128 *mat_exp_ab,
                                     const ffb_c128 *mat_exp_bb, int z_representation,
                                     void *vec_dev, int64_t row0, int64_t n_rows, void *stream);
/* vec[a,b] *= prod_{i in a} phases_a[i] * prod_{j in b} phases_b[j]; either may be NULL. */
int ffb_apply_num_op_sum_evolution(ffb_tables *tables_a, ffb_tables *tables_b,
                                   const ffb_c128 *phases_a, const ffb_c128 *phases_b,
                                   void *vec_dev, int64_t row0, int64_t n_rows, void *stream);
/* out[a,b] (+)= coeff(a,b) * vec[a,b]; accumulate != 0 keeps the old out
 * (the reference's into_buffer form), 0 overwrites (no read of out). */
int ffb_contract_diag_coulomb(ffb_tables *tables_a, ffb_tables *tables_b, const double *mat_aa,
                              const double *mat_ab, const double *mat_bb, int z_representation,
                              const void *vec_dev, void *out_dev, int accumulate, int64_t row0,
                              int64_t n_rows, void *stream);
int ffb_contract_num_op_sum(ffb_tables *tables_a, ffb_tables *tables_b, const double *coeffs_a,
                            const double *coeffs_b, const void *vec_dev, void *out_dev,
                            int accumulate, int64_t row0, int64_t n_rows, void *stream);

/* The same four operators on a rectangular block of the state: alpha rows [row0, row0 + n_rows) x
 * beta columns [col0, col0 + n_cols), element (row0 + r, col0 + c) at vec_dev[r * ld + c].  n_cols < 0
 * means the whole beta sector with contiguous rows (the calls above).  A column block is what a
 * rank holds after the distributed transpose (the state stays in that layout
 * between an alpha-side rotation and the next one, diagonal operators work on either). */
int ffb_apply_diag_coulomb_evolution_block(ffb_tables *tables_a, ffb_tables *tables_b,
                                           const ffb_c128 *mat_exp_aa, const ffb_c128 *mat_exp_ab,
                                           const ffb_c128 *mat_exp_bb, int z_representation,
                                           void *vec_dev, int64_t row0, int64_t n_rows, int64_t col0,
                                           int64_t n_cols, int64_t ld, void *stream);
int ffb_apply_num_op_sum_evolution_block(ffb_tables *tables_a, ffb_tables *tables_b,
                                         const ffb_c128 *phases_a, const ffb_c128 *phases_b,
                                         void *vec_dev, int64_t row0, int64_t n_rows, int64_t col0,
                                         int64_t n_cols, int64_t ld, void *stream);
int ffb_contract_diag_coulomb_block(ffb_tables *tables_a, ffb_tables *tables_b, const double *mat_aa,
                                    const double *mat_ab, const double *mat_bb, int z_representation,
                                    const void *vec_dev, void *out_dev, int accumulate, int64_t row0,
                                    int64_t n_rows, int64_t col0, int64_t n_cols, int64_t ld,
                                    void *stream);
int ffb_contract_num_op_sum_block(ffb_tables *tables_a, ffb_tables *tables_b, const double *coeffs_a,
                                  const double *coeffs_b, const void *vec_dev, void *out_dev,
                                  int accumulate, int64_t row0, int64_t n_rows, int64_t col0,
                                  int64_t n_cols, int64_t ld, void *stream);

/* vec[a,b] *= phase wherever string a contains every orbital of mask_a and string b every orbital of
 * mask_b (bit i = orbital i): the controlled phase shift of python/ffsim/gates/basic_gates.py:27-51
 * behind apply_num_num_interaction, apply_on_site_interaction and apply_num_op_prod_interaction.
 * Block arguments as in the _block calls above (n_cols < 0: whole beta sector). */
int ffb_apply_num_op_prod_phase(ffb_tables *tables_a, ffb_tables *tables_b, uint32_t mask_a, uint32_t mask_b,
                                ffb_c128 phase, void *vec_dev, int64_t row0, int64_t n_rows, int64_t col0,
                                int64_t n_cols, int64_t ld, void *stream);

/* ---------------------------------------------------------------- utilities */
/* out[c, r] = in[r, c]; in is n_rows x n_cols with row stride ld_in, out has row stride ld_out. */
int ffb_transpose(const void *in_dev, void *out_dev, int64_t n_rows, int64_t n_cols, int64_t ld_in,
                  int64_t ld_out, void *stream);
/* One rank's side of the distributed transpose of a row-sharded state (no reference counterpart:
 * SURVEY.md section 8e).  Block d (d < n_dst <= 16) is rows[d] x width[d] complex128 elements, read from
 * src_dev + src_off[d] with row stride src_ld and stored to dst_dev[d] + dst_off[d] with row stride
 * dst_ld[d] (all in elements).  dst_dev[d] may point into another GPU's memory mapped over NVLink
 * (CUDA IPC / symmetric memory): the pack, all-to-all and unpack steps become this one kernel.  The
 * caller orders it against the peers (a barrier before the buffers are overwritten, one after). */
int ffb_exchange_blocks(const void *src_dev, int64_t src_ld, int n_dst, const int64_t *rows,
                        const int64_t *width, const int64_t *src_off, void *const *dst_dev,
                        const int64_t *dst_off, const int64_t *dst_ld, void *stream);
/* The same kernel with one source row stride per block: the pack / unpack step of the NCCL
 * all-to-all path (all destinations local) as well as the peer-memory exchange. */
int ffb_copy_blocks(const void *src_dev, int n_blocks, const int64_t *rows, const int64_t *width,
                    const int64_t *src_off, const int64_t *src_ld, void *const *dst_dev,
                    const int64_t *dst_off, const int64_t *dst_ld, void *stream);
/* result_dev[0] = sum conj(x) * y as one complex128 (device scalar, 16 bytes). */
int ffb_vdot(const void *x_dev, const void *y_dev, int64_t n, void *result_dev, void *stream);
/* y = alpha * x + beta * y, complex scalars */
int ffb_axpby(ffb_c128 alpha, const void *x_dev, ffb_c128 beta, void *y_dev, int64_t n, void *stream);

/* Launch profiling for bench.py.  Between begin and end every fused-pass, diagonal
 * and transpose kernel launch is bracketed by CUDA events on its own stream and
 * every kernel launch of the library is counted.  ffb_profile_end synchronises the
 * device and writes a JSON object {"<kernel>": {"launches", "timed", "ms", "bytes"}}
 * (bytes = algorithmic bytes of the timed launches). */
int ffb_profile_begin(void);
int ffb_profile_end(char *buf, size_t buflen);

/* Asynchronous strided copy (cudaMemcpy2DAsync; kind 1 = host to device, 2 = device to host, 3 = device to
 * device): `height` rows of `width_bytes`, row pitches in bytes.  It moves a column strip or a row block of
 * a HOST state to / from the device while the kernels work on the strips that have already arrived
 * (ffsim_b200/pipeline.py); the host side should be page-locked. */
int ffb_memcpy2d_async(void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t width_bytes,
                       size_t height, int kind, void *stream);

/* Dense FP64 FMA throughput of the current device in TFLOP/s (a short microbenchmark, ~10 ms): the
 * second roofline denominator of the fused rotation kernel, measured where the bench runs. */
int ffb_measure_fp64_peak(double *tflops);

/* Tuning knobs (process-wide; read when a plan is built).  key/value pairs:
 *   "smem_bytes"   shared-memory budget per tile (default 220 KB)
 *   "min_cols"     smallest column strip that may define the window width (default 3)
 *   "max_cols"     largest column strip of a full-height tile (default 8)
 *   "sub_window"   register-block width, 2..6 (default 6)
 *   "threads"      CTA size of the fused pass kernel (default 512, the size it is built for)
 *   "beta_mode"    beta side: 0 auto (in place when the sector fits one window, else a transposed copy with
 *                  the transpositions folded into the first / last pass), 1 in place always, 2 transposed
 *                  copy always, 3 transposed copy with separate transpose kernels
 *   "bulk_copies"  1 (default): contiguous tile columns move as TMA bulk copies, 0: 16-byte copies only
 * Unknown keys return FFB_EINVAL. */
int ffb_set_option(const char *key, int64_t value);
int64_t ffb_get_option(const char *key);

#ifdef __cplusplus
}
#endif
#endif /* FFSIM_B200_H */
