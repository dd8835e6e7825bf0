// solver_common.cuh -- device helpers shared by the LW and SW solver kernels.
#pragma once
#include "cloud_core.h"
#include "kernels.cuh"
#include "solver_core.h"

namespace ecb {

#define LD_IN(p, c, j) ((p)[(size_t)(j) * in.ld + (c)])

enum { LCH = 16 };   // layers between two g-point reductions

// Sum the rows of the shared-memory tile over g (4 threads per row) into sums[f][level].
// Row (f, s), s < ns, holds level lfirst + dir*s of flux f.  Ends with a barrier so the tile can be refilled.
__device__ __forceinline__ void flush_tile(const double* tile, int rs, int ng, int nf, int ns, double* const* dst, int lfirst, int dir, int lch = LCH) {
  __syncthreads();
  const int row = threadIdx.x >> 2, sub = threadIdx.x & 3;
  const int nrows = nf * ns, rows_per_round = blockDim.x >> 2;
  for (int r0 = 0; r0 < nrows; r0 += rows_per_round) {
    const int r = r0 + row;
    const bool valid = r < nrows;
    int f = 0, s = 0;
    double p = 0.0;
    if (valid) {
      f = r / ns; s = r - f * ns;
      const double* t = tile + (size_t)(f * lch + s) * rs;
      for (int g = sub; g < ng; g += 4) p += t[g];
    }
    p += __shfl_xor_sync(0xffffffffu, p, 1);
    p += __shfl_xor_sync(0xffffffffu, p, 2);
    if (valid && sub == 0) dst[f][lfirst + dir * s] = p;
  }
  __syncthreads();
}

__device__ __forceinline__ uint32_t pick4(const uint4& q, int k) { return k == 0 ? q.x : k == 1 ? q.y : k == 2 ? q.z : q.w; }

// optical-depth scaling of this (g, layer) from the generator's code word
__device__ __forceinline__ double od_scaling_from_code(const CloudMeta& C, const double* pdf_val, uint32_t code, double fsd) {
  if (!code) return 0.0;
  return pdf_sample(C, pdf_val, fsd, (double)(code & 0x3FFFFFFFu) * (1.0 / 1073741824.0));
}


}  // namespace ecb
