// Measured denominators for the rooflines bench.py reports (sy2d_measure_peaks): the bandwidth of the unit a kernel
// is bound by, measured on the device the bench runs on with a copy micro-kernel of the same access width (fp64):
//   shared memory / L1TEX data pipe : every thread copies 8-byte words inside a 96 KB shared-memory buffer
//                                     (conflict-free, consecutive lanes -> consecutive words); bytes = loads + stores
//   L2                              : a 32 MB buffer (resident in the 126 MB L2) read with 16-byte loads, written back
//                                     to a second 32 MB buffer; bytes = reads + writes
//   HBM                             : the same copy on 2 x 1 GB buffers (cross-check of MEASURED_PEAKS.json)
// The engine-2 ensemble kernel (k_problem_xline) moves 18 shared-memory and 10 L2-backed 8-byte accesses per cell and
// BiCGSTAB iteration through the L1TEX data pipe of its SM; HBM sees only the compulsory 72 B per cell and time step.
#pragma once
#include <cuda_runtime.h>

namespace sy2d {

constexpr int kPeakSmemDoubles = 12288;   // 96 KB

__global__ void __launch_bounds__(1024, 2) k_peak_smem(double* sink, int reps) {
  extern __shared__ double pk_s[];
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int n = tid; n < kPeakSmemDoubles; n += nt) pk_s[n] = (double)n;
  __syncthreads();
  // each thread owns the words tid + m * nt of both halves and copies half A -> half B and back: consecutive lanes
  // touch consecutive 8-byte words, no bank conflicts, no barrier needed (thread-private words)
  const int half = kPeakSmemDoubles / 2;
  double acc = 0.0;
  for (int r = 0; r < reps; ++r) {
#pragma unroll 6
    for (int n = tid; n < half; n += nt) pk_s[half + n] = pk_s[n] + acc;
#pragma unroll 6
    for (int n = tid; n < half; n += nt) pk_s[n] = pk_s[half + n];
    acc += 1.0e-300;
  }
  __syncthreads();
  if (tid == 0) sink[blockIdx.x] = pk_s[7] + pk_s[half + 11];
}

__global__ void __launch_bounds__(256) k_peak_copy(const double2* __restrict__ src, double2* __restrict__ dst, size_t n2) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x; n < n2; n += stride) dst[n] = src[n];
}

}  // namespace sy2d
