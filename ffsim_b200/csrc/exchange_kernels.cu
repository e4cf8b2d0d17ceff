// Block exchange for the distributed transpose of a row-sharded state.
//
// Redistributing row shards [n_rows x dim_b] to column shards [dim_a x n_cols_local] (and back) is,
// seen from one rank, a set of strided 2-D block copies: one block per destination rank.  With every
// GPU of the box mapped into every other one's address space (NVLink 5 / NVSwitch peer memory) the
// pack -> all-to-all -> unpack sequence collapses into ONE kernel that reads the local shard once and
// stores each block straight into its final place in the destination rank's buffer: 16-byte stores,
// 512 contiguous bytes per warp, destinations interleaved so that all links are busy at once.
// The reference has no counterpart (single-process NumPy).
#include <cuda_runtime.h>

#include "kernels.hpp"

namespace ffb {

namespace {

constexpr int kExThreads = 256;
constexpr int kExUnroll = 4;

__global__ void __launch_bounds__(kExThreads) exchange_kernel(const __grid_constant__ ExchangeParams p) {
  const long long n_units = p.max_rows * p.n_dst;
  for (long long unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const int d = (int)(unit % p.n_dst);  // consecutive CTAs feed different links
    const long long row = unit / p.n_dst;
    if (row >= p.rows[d]) continue;
    const double2 *__restrict__ src = reinterpret_cast<const double2 *>(p.src) + p.src_off[d] + row * p.src_ld[d];
    double2 *__restrict__ dst = reinterpret_cast<double2 *>(p.dst[d]) + p.dst_off[d] + row * p.dst_ld[d];
    const long long width = p.width[d];
    for (long long c0 = threadIdx.x; c0 < width; c0 += (long long)kExThreads * kExUnroll) {
      double2 v[kExUnroll];
#pragma unroll
      for (int u = 0; u < kExUnroll; ++u) {
        const long long c = c0 + (long long)u * kExThreads;
        if (c < width) v[u] = src[c];
      }
#pragma unroll
      for (int u = 0; u < kExUnroll; ++u) {
        const long long c = c0 + (long long)u * kExThreads;
        if (c < width) dst[c] = v[u];
      }
    }
  }
}

}  // namespace

cudaError_t launch_exchange(const ExchangeParams &p, int sm_count, cudaStream_t stream) {
  const long long n_units = p.max_rows * p.n_dst;
  if (n_units <= 0) return cudaSuccess;
  long long grid = (long long)sm_count * 8;
  if (grid > n_units) grid = n_units;
  exchange_kernel<<<(int)grid, kExThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace ffb
