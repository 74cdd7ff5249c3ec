#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int W = 64, H = 64, BW = 32, BH = 8;
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int* out) {
  __shared__ alignas(128) int smem_buffer[BH][BW];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) {
    init(&bar, blockDim.x);
    cde::fence_proxy_async_shared_cta();
  }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
  } else {
    token = bar.arrive();
  }
  bar.wait(std::move(token));
  for (int k = threadIdx.x; k < BH * BW; k += blockDim.x) out[k] = smem_buffer[k / BW][k % BW];
}
int main() {
  std::vector<int> h(W * H);
  for (int i = 0; i < W * H; ++i) h[i] = i;
  int *d, *out;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&out, BW * BH * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cuInit(0);
  CUtensorMap map;
  cuuint64_t size[2] = {W, H}, stride[1] = {W * 4};
  cuuint32_t box[2] = {BW, BH}, es[2] = {1, 1};
  CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode=%d\n", (int)r);
  kernel<<<1, 128>>>(map, 0, 0, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("sync: %s\n", cudaGetErrorString(e));
  std::vector<int> o(BW * BH);
  cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
  printf("o[0]=%d o[33]=%d (want 0, 65)\n", o[0], o[33]);
  return 0;
}
