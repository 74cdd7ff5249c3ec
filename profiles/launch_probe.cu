// Cost of a dependent kernel launch on a B200: chains of N trivial kernels on a stream and as a replayed CUDA graph, 1 CTA and 128 CTAs
// of 256 / 1024 threads, and a kernel that needs 150 KB of dynamic shared memory (carve-out switch between neighbours).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o profiles/launch_probe.bin profiles/launch_probe.cu && ./profiles/launch_probe.bin
#include <cstdio>
#include <cuda_runtime.h>

__global__ void tiny(double* p) { if (threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1.0; }
__global__ void tiny_smem(double* p) {
  extern __shared__ double s[];
  s[threadIdx.x] = p[0];
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) p[0] = s[1] + 1.0;
}

static float run(cudaStream_t st, int n, int grid, int block, bool graph, bool alternate_smem, double* d) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaGraphExec_t exec = nullptr;
  auto issue = [&]() {
    for (int k = 0; k < n; ++k) {
      if (alternate_smem && (k & 1)) tiny_smem<<<grid, block, 150 * 1024, st>>>(d);
      else tiny<<<grid, block, 0, st>>>(d);
    }
  };
  if (graph) {
    cudaGraph_t g;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
    issue();
    cudaStreamEndCapture(st, &g);
    cudaGraphInstantiate(&exec, g, 0);
    cudaGraphLaunch(exec, st);
    cudaStreamSynchronize(st);
  } else {
    issue();
    cudaStreamSynchronize(st);
  }
  cudaEventRecord(e0, st);
  for (int rep = 0; rep < 5; ++rep) { if (graph) cudaGraphLaunch(exec, st); else issue(); }
  cudaEventRecord(e1, st);
  cudaStreamSynchronize(st);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  return 1e3f * ms / (5.0f * n);
}

int main() {
  double* d;
  cudaMalloc(&d, 8);
  cudaMemset(d, 0, 8);
  cudaFuncSetAttribute(tiny_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 150 * 1024);
  cudaStream_t st;
  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  const int n = 500;
  for (int graph = 0; graph < 2; ++graph)
    for (int cfg = 0; cfg < 4; ++cfg) {
      const int grid = cfg == 0 ? 1 : 128, block = cfg == 2 ? 1024 : 256;
      const bool alt = cfg == 3;
      printf("%s, %3d CTAs x %4d threads%s: %.2f us per dependent launch\n", graph ? "graph replay" : "stream      ", grid, block,
             alt ? ", every other kernel with 150 KB of shared memory" : "", run(st, n, grid, block, graph != 0, alt, d));
    }
  return 0;
}
