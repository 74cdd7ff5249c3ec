// Stand-alone probe of the TMA box load used by k_assemble_tma (built and run on the GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe profiles/tma_probe.cu -lcuda
//   /tmp/tma_probe <fence 0 mbarrier_init|1 proxy.async|2 both> <order 0 expect first|1 copy first> <l2promo 0..3> <coord> <dtype 0 f64|1 i32>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void k(const CUtensorMap* map, double* out, int nbytes, int c0, int c1, int fence, int order) {
  extern __shared__ __align__(128) unsigned char raw[];
  double* buf = reinterpret_cast<double*>(raw);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(raw + 8192);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)) : "memory");
    if (fence == 0 || fence == 2) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (fence == 1 || fence == 2) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (order == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(nbytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(buf)),
                 "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(s32(bar)) : "memory");
    if (order == 1) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(nbytes) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(s32(bar)) : "memory");
  for (int i = threadIdx.x; i < nbytes / 8; i += blockDim.x) out[i] = buf[i];
}

int main(int argc, char** argv) {
  const int fence = argc > 1 ? atoi(argv[1]) : 0, order = argc > 2 ? atoi(argv[2]) : 0, l2 = argc > 3 ? atoi(argv[3]) : 2;
  const int coord = argc > 4 ? atoi(argv[4]) : -1, dtype = argc > 5 ? atoi(argv[5]) : 0;
  const int nx = 64, ny = 64, box0 = 34, box1 = 10;
  std::vector<double> h(nx * ny);
  for (int i = 0; i < nx * ny; ++i) h[i] = i;
  double *d, *out;
  cudaMalloc(&d, h.size() * 8);
  cudaMalloc(&out, 8192);
  cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  cuInit(0);
  CUtensorMap map;
  const int es_bytes = dtype == 0 ? 8 : 4, mul = 8 / es_bytes;
  cuuint64_t dims[2] = {(cuuint64_t)ny * mul, (cuuint64_t)nx}, strides[1] = {(cuuint64_t)ny * 8};
  cuuint32_t box[2] = {(cuuint32_t)box0 * mul, (cuuint32_t)box1}, es[2] = {1, 1};
  CUresult r = cuTensorMapEncodeTiled(&map, dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, dims, strides, box, es,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("fence %d order %d l2 %d coord %d dtype %d encode=%d: ", fence, order, l2, coord, dtype, (int)r);
  CUtensorMap* dmap;
  cudaMalloc(&dmap, sizeof map);
  cudaMemcpy(dmap, &map, sizeof map, cudaMemcpyHostToDevice);
  k<<<1, 128, 8192 + 64>>>(dmap, out, box0 * box1 * 8, coord * mul, coord, fence, order);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<double> o(box0 * box1);
    cudaMemcpy(o.data(), out, o.size() * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int a = 0; a < box1; ++a)
      for (int b = 0; b < box0; ++b) {
        const int i = a + coord, j = b + coord;
        const double want = (i < 0 || j < 0 || i >= nx || j >= ny) ? 0.0 : (double)(i * ny + j);
        if (o[a * box0 + b] != want) ++bad;
      }
    printf(", mismatches %d", bad);
  }
  printf("\n");
  return 0;
}
