// Does tcgen05.ld / tcgen05.st (32x32b, .x16 / .x8 / .x4) need a column address aligned to its width?  Writes column c of every
// lane with lane * 1000 + c through .x1 stores, reads 16 / 8 / 4 columns at every offset 0..31 and compares.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tmem_probe profiles/tmem_probe.cu && /tmp/tmem_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void probe(int* bad) {
  __shared__ unsigned slot;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((unsigned)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned base = slot + ((unsigned)(32 * w) << 16);
  for (int c = 0; c < 128; ++c) {
    const unsigned v = (unsigned)((32 * w + lane) * 1000 + c);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(base + c), "r"(v) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  for (int off = 0; off < 32; ++off) {
    unsigned r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                   "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(base + off)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i)
      if (r[i] != (unsigned)((32 * w + lane) * 1000 + off + i)) atomicAdd(&bad[off], 1);
    unsigned q[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]) : "r"(base + 64 + off) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 4; ++i)
      if (q[i] != (unsigned)((32 * w + lane) * 1000 + 64 + off + i)) atomicAdd(&bad[32 + off], 1);
  }
  // unaligned .x16 store, aligned read back
  {
    unsigned v[16];
    for (int i = 0; i < 16; ++i) v[i] = 7000000u + (unsigned)(32 * w + lane) * 100 + i;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(base + 202),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
                 "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) {
      unsigned x;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(x) : "r"(base + 202 + i) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (x != v[i]) atomicAdd(&bad[64], 1);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

int main() {
  int* bad;
  cudaMallocManaged(&bad, 65 * sizeof(int));
  for (int i = 0; i < 65; ++i) bad[i] = 0;
  probe<<<1, 128>>>(bad);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  printf("x16 load mismatches by column offset 0..31:");
  for (int i = 0; i < 32; ++i) printf(" %d", bad[i]);
  printf("\nx4 load mismatches by column offset 0..31:");
  for (int i = 0; i < 32; ++i) printf(" %d", bad[32 + i]);
  printf("\nx16 store at column 202 (2 mod 8), mismatches: %d\n", bad[64]);
  return 0;
}
