// Tensor memory (TMEM, sm_100a) as a thread-private scratchpad.
//
// An SM has 256 KB of tensor memory - 128 lanes x 512 columns of 32 bits - next to its 228 KB of shared memory and its
// register file.  It exists to hold tcgen05.mma accumulators, but `tcgen05.st` / `tcgen05.ld` move registers to and from it
// without any MMA: with the 32x32b shape, thread `lane` of a warp reads or writes N consecutive 32-bit columns of TMEM lane
// `32 * (warp % 4) + lane`.  That is exactly a thread-private array: warps with the same `warp % 4` share a lane quarter and
// take disjoint column ranges.  The ensemble kernel (sy2d_xline_kernel.cuh) keeps three of its per-thread arrays there
// (wS', wN', y: 60 columns per thread) - arrays that fit neither in its registers nor in its shared memory and used to live in
// an L2-backed scratch.  A load costs tens of cycles instead of an L2 round trip and does not go through the L1TEX data pipe.
#pragma once

namespace sy2d {

// One warp of the CTA: allocates all 512 columns (the kernel runs one CTA per SM) and leaves the base address in *slot.
__device__ __forceinline__ void tmem_alloc_all(unsigned* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((unsigned)__cvta_generic_to_shared(slot)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_all(unsigned base) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ND doubles (2 ND columns) of this thread's TMEM lane, starting at column address `taddr`.  Warp-collective (.sync.aligned):
// every lane of the warp must execute it.  The values are valid after tmem_wait_ld().
template <int ND>
__device__ __forceinline__ void tmem_ld(double (&d)[ND], unsigned taddr);
template <>
__device__ __forceinline__ void tmem_ld<2>(double (&d)[2], unsigned taddr) {
  unsigned r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
  tmem_wait_ld();
#pragma unroll
  for (int k = 0; k < 2; ++k) d[k] = __hiloint2double((int)r[2 * k + 1], (int)r[2 * k]);
}
template <>
__device__ __forceinline__ void tmem_ld<4>(double (&d)[4], unsigned taddr) {
  unsigned r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  tmem_wait_ld();
#pragma unroll
  for (int k = 0; k < 4; ++k) d[k] = __hiloint2double((int)r[2 * k + 1], (int)r[2 * k]);
}
template <>
__device__ __forceinline__ void tmem_ld<8>(double (&d)[8], unsigned taddr) {
  unsigned r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  tmem_wait_ld();
#pragma unroll
  for (int k = 0; k < 8; ++k) d[k] = __hiloint2double((int)r[2 * k + 1], (int)r[2 * k]);
}

// Ten doubles in two loads (8 at `ta`, 2 at `tb`) and ONE wait.
__device__ __forceinline__ void tmem_ld2(double (&a)[8], unsigned ta, double (&b)[2], unsigned tb) {
  unsigned r[16], q[4];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(ta)
      : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]) : "r"(tb) : "memory");
  tmem_wait_ld();
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = __hiloint2double((int)r[2 * k + 1], (int)r[2 * k]);
#pragma unroll
  for (int k = 0; k < 2; ++k) b[k] = __hiloint2double((int)q[2 * k + 1], (int)q[2 * k]);
}
__device__ __forceinline__ void tmem_ld2(double (&a)[1], unsigned, double (&b)[1], unsigned) { a[0] = 0.0; b[0] = 0.0; }   // (never called: keeps the non-TMEM instantiations compiling)

// The store is complete (visible to a later tmem_ld of the same warp) after tmem_wait_st().
template <int ND>
__device__ __forceinline__ void tmem_st(unsigned taddr, const double (&d)[ND]);
template <>
__device__ __forceinline__ void tmem_st<2>(unsigned taddr, const double (&d)[2]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__double2loint(d[0])), "r"(__double2hiint(d[0])),
               "r"(__double2loint(d[1])), "r"(__double2hiint(d[1]))
               : "memory");
}
template <>
__device__ __forceinline__ void tmem_st<4>(unsigned taddr, const double (&d)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(__double2loint(d[0])),
               "r"(__double2hiint(d[0])), "r"(__double2loint(d[1])), "r"(__double2hiint(d[1])), "r"(__double2loint(d[2])), "r"(__double2hiint(d[2])),
               "r"(__double2loint(d[3])), "r"(__double2hiint(d[3]))
               : "memory");
}
template <>
__device__ __forceinline__ void tmem_st<8>(unsigned taddr, const double (&d)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(__double2loint(d[0])), "r"(__double2hiint(d[0])), "r"(__double2loint(d[1])), "r"(__double2hiint(d[1])), "r"(__double2loint(d[2])),
      "r"(__double2hiint(d[2])), "r"(__double2loint(d[3])), "r"(__double2hiint(d[3])), "r"(__double2loint(d[4])), "r"(__double2hiint(d[4])),
      "r"(__double2loint(d[5])), "r"(__double2hiint(d[5])), "r"(__double2loint(d[6])), "r"(__double2hiint(d[6])), "r"(__double2loint(d[7])),
      "r"(__double2hiint(d[7]))
      : "memory");
}

// size-generic wrappers (the one-element overloads only keep the non-TMEM instantiations of the ensemble kernel compiling)
__device__ __forceinline__ void tmem_st_any(unsigned taddr, const double (&d)[8]) { tmem_st<8>(taddr, d); }
__device__ __forceinline__ void tmem_st_any(unsigned taddr, const double (&d)[2]) { tmem_st<2>(taddr, d); }
__device__ __forceinline__ void tmem_st_any(unsigned, const double (&)[1]) {}

}  // namespace sy2d
