// Launch wrappers of the CUDA kernels (implemented in the .cu files).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "device_structs.h"

namespace ffb {

struct PhaseList {  // per-orbital phases passed by value
  double re[32], im[32];
};

// shared memory the fused kernel needs besides the tile: offset tables, segment descriptors of
// `n_sub` sub-passes, two block-list staging buffers of `blk_cap` entries and two offset tables of
// `off_rows` rows (the largest register-block start position of the pass, plus one)
size_t fused_pass_smem_overhead(int n_sub, int blk_cap, int off_rows);
int fused_pass_ctas_per_sm(int w, int threads, size_t smem_bytes);
cudaError_t launch_fused_pass(const PassParams &p, int grid, int threads, size_t smem_bytes,
                              cudaStream_t stream);
cudaError_t launch_givens_single(void *vec, long long ld, long long dim_b, double c, double sr,
                                 double si, const unsigned long long *s1,
                                 const unsigned long long *s2, long long n_pairs, int sm_count,
                                 cudaStream_t stream);
cudaError_t launch_phase_shift(void *vec, long long ld, long long dim_b, double pr, double pi,
                               const unsigned long long *indices, long long n, int sm_count,
                               cudaStream_t stream);
cudaError_t launch_row_scale(void *vec, long long n_rows, long long n_cols, long long row_stride,
                             long long col_stride, const void *phase, int sm_count,
                             cudaStream_t stream);
cudaError_t launch_row_phase(const uint32_t *strings, long long dim, int norb, const PhaseList &ph,
                             void *out, int sm_count, cudaStream_t stream);
cudaError_t launch_transpose(const void *in, void *out, long long n_rows, long long n_cols,
                             long long ld_in, long long ld_out, int sm_count, cudaStream_t stream);

}  // namespace ffb

namespace ffb {
// One rank's side of the distributed transpose: block d is rows[d] x width[d] elements, read from
// src + src_off[d] (row stride src_ld[d]) and written to dst[d] + dst_off[d] (row stride dst_ld[d]).
constexpr int kMaxExchangeDst = 16;
struct ExchangeParams {
  const void *src;
  int n_dst;
  long long max_rows;
  void *dst[kMaxExchangeDst];
  long long src_off[kMaxExchangeDst], src_ld[kMaxExchangeDst], dst_off[kMaxExchangeDst], dst_ld[kMaxExchangeDst];
  long long rows[kMaxExchangeDst], width[kMaxExchangeDst];
};
cudaError_t launch_exchange(const ExchangeParams &p, int sm_count, cudaStream_t stream);

cudaError_t launch_side_factor(bool contract, const uint32_t *strings, long long dim, int norb,
                               const void *mat, int zrep, void *out, int sm_count,
                               cudaStream_t stream);
cudaError_t launch_diag(bool contract, const uint32_t *strings_a, const uint32_t *strings_b,
                        const void *rowfac, const void *colfac, const void *mab, const void *vec,
                        void *out, long long row0, long long n_rows, long long col0, long long n_cols,
                        long long ld, int norb, int zrep, int accumulate, int sm_count,
                        cudaStream_t stream);
cudaError_t launch_num_op_prod_phase(const uint32_t *strings_a, const uint32_t *strings_b, uint32_t mask_a,
                                     uint32_t mask_b, double pr, double pi, void *vec, long long row0,
                                     long long n_rows, long long col0, long long n_cols, long long ld,
                                     int sm_count, cudaStream_t stream);
// Roofline denominator of the fused Givens kernel: dense DFMA throughput of this device (TFLOP/s,
// 2 flops per DFMA), best of a few launches of an unrolled independent-chain kernel.
cudaError_t measure_fp64_peak(int sm_count, double *tflops);
cudaError_t launch_vdot(const void *x, const void *y, long long n, void *partial, int n_partial,
                        void *result, int sm_count, cudaStream_t stream);
cudaError_t launch_axpby(double ar, double ai, const void *x, double br, double bi, void *y,
                         long long n, int sm_count, cudaStream_t stream);
}  // namespace ffb
