// FP64 pipe of a B200 SM: latency of a dependent DFMA chain and throughput of independent DFMAs at 1 .. 32 warps per SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o profiles/fp64_probe.bin profiles/fp64_probe.cu && ./profiles/fp64_probe.bin
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(double* out, long long* cycles, int n, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x * 1e-3 + k;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = fma(x[k], a, b);
  }
  const long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int ILP>
void run(int threads, int n) {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(double));
  cudaMallocManaged(&cyc, 148 * sizeof(long long));
  chain<ILP><<<148, threads>>>(out, cyc, n, 0.999999, 1e-9);
  cudaDeviceSynchronize();
  chain<ILP><<<148, threads>>>(out, cyc, n, 0.999999, 1e-9);
  cudaDeviceSynchronize();
  const double c = (double)cyc[0] / n;
  printf("ILP %d, %2d warps per SM: %.2f cycles per loop iteration, %.2f cycles per DFMA of a warp, %.1f FP64 FMA lanes per cycle per SM\n", ILP,
         threads / 32, c, c / ILP, 32.0 * ILP * (threads / 32) / c);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  const int n = 4096;
  run<1>(32, n); run<2>(32, n); run<4>(32, n); run<8>(32, n);
  run<1>(128, n); run<2>(128, n); run<4>(128, n);
  run<1>(512, n); run<2>(512, n); run<4>(512, n);
  run<1>(1024, n); run<2>(1024, n); run<4>(1024, n);
  return 0;
}
