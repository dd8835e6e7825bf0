// bulk_pipe.cuh -- producer/consumer ring of shared-memory stages filled by the TMA unit (1-D bulk copies) for the
// streaming kernels of the solvers.
//
// The downward-sweep kernels read 5-10 scratch arrays per layer and alternate between a short recurrence and a block-wide
// g-point reduction; with plain loads nothing is in flight while a CTA reduces.  Here one thread asks the TMA unit for the
// next stages ahead of time (`cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes`, SASS UBLKCP): the scratch
// arrays are [layer][g] per column, so `nl` layers of one array are ONE contiguous, 16-byte aligned block (ng*8 is a multiple
// of 16 for every spectral size: 140, 112, 96, 64, 32).  full[s] flips when the bytes of stage s have landed, empty[s] when
// every thread has read them.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ecb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) { while (!mbar_try_wait(b, parity)) {} }
// global -> shared bulk copy by the TMA unit; completion is counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Ring of NST stages; a stage holds NL layers of up to NARR arrays, [array][layer][rowlen] doubles.
template <int NST, int NL, int NARR>
struct BulkRing {
  double* buf;          // [NST][NARR][NL][rowlen]
  uint64_t* full;       // [NST]
  uint64_t* empty;      // [NST]
  int rowlen;           // doubles per layer (= ng)
  __device__ __forceinline__ static size_t bytes(int rowlen) { return sizeof(double) * NST * NARR * NL * rowlen + 2 * NST * sizeof(uint64_t); }
  __device__ __forceinline__ void carve(unsigned char* base, int rowlen_) {
    rowlen = rowlen_;
    buf = reinterpret_cast<double*>(base);
    full = reinterpret_cast<uint64_t*>(buf + (size_t)NST * NARR * NL * rowlen);
    empty = full + NST;
  }
  // block-wide: call once by all threads before use
  __device__ __forceinline__ void init(int nthreads) {
    if (threadIdx.x == 0) {
      for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], nthreads); }
      mbar_fence_init();
    }
    __syncthreads();
  }
  __device__ __forceinline__ double* stage(int s, int arr) const { return buf + ((size_t)s * NARR + arr) * NL * rowlen; }
  // producer thread: request stage number j (layers l0 .. l0+nl-1) of `narr` arrays src[k] (each [layer][rowlen])
  __device__ __forceinline__ void issue(int j, const double* const* src, int narr, int l0, int nl) {
    const int s = j % NST, use = j / NST;
    if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
    const uint32_t one = (uint32_t)(nl * rowlen * sizeof(double));
    mbar_arrive_expect_tx(&full[s], one * narr);
    for (int k = 0; k < narr; ++k) bulk_g2s(stage(s, k), src[k] + (size_t)l0 * rowlen, one, &full[s]);
  }
  __device__ __forceinline__ void wait_full(int j) { mbar_wait(&full[j % NST], (j / NST) & 1); }
  // Call AFTER the values read from the stage have been used in arithmetic (or stored): the arrive must not overtake shared-memory
  // loads that are still in flight, or the refill could land under them (seen as one wrong element in ~10^6 stages).
  __device__ __forceinline__ void release(int j) { mbar_arrive(&empty[j % NST]); }
};

}  // namespace ecb
