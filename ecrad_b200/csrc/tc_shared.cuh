// tc_shared.cuh -- per-column region data of the Tripleclouds and SPARTACUS sweeps in shared memory.
#pragma once
#include "solver_common.cuh"

namespace ecb {

struct TcShared {
  double *reg, *ods;        // shared memory: region fractions and optical-depth scalings [nlev][3]
  const double *U, *V;      // global memory (uniform, L1-resident broadcast loads): overlap matrices [nlev+1][3][3]
  int* clear;               // is_clear_sky_layer(0:nlev+1)
};

// loads the column's region data into shared memory (block-wide, ends with a barrier).  The 3x3 overlap matrices stay in
// global memory: every thread of the CTA reads the same 9 values per half-level, one L1 transaction per warp, and keeping
// their 20 KB out of shared memory doubles the resident CTAs of these kernels.
__device__ __forceinline__ TcShared tc_load_shared(unsigned char* base, const Work& w, const DevIn& in, int c, int nlev, int nthreads) {
  TcShared s;
  s.reg = reinterpret_cast<double*>(base);
  s.ods = s.reg + nlev * 3;
  s.clear = reinterpret_cast<int*>(s.ods + nlev * 3);
  s.U = w.tc_u + (size_t)c * (nlev + 1) * 9;
  s.V = w.tc_v + (size_t)c * (nlev + 1) * 9;
  const int t = threadIdx.x;
  for (int i = t; i < nlev * 3; i += nthreads) { s.reg[i] = w.tc_reg[(size_t)c * nlev * 3 + i]; s.ods[i] = w.tc_ods[(size_t)c * nlev * 3 + i]; }
  for (int i = t; i < nlev + 2; i += nthreads) s.clear[i] = (i == 0 || i == nlev + 1) ? 1 : !(LD_IN(in.frac, c, i - 1) > 0.0);
  __syncthreads();
  return s;
}
static inline size_t tc_shared_bytes(int nlev) { return sizeof(double) * (6 * nlev) + sizeof(int) * (nlev + 2) + 16; }

// out[j1] = sum_j2 A[j1][j2] * x[j2]   (singlemat_x_vec, radiation_matrix.F90:110-136)
__device__ __forceinline__ void mat3_x_vec(const double* A, double* x) {
  double o[3];
#pragma unroll
  for (int j1 = 0; j1 < 3; ++j1) {
    double acc = 0.0;
#pragma unroll
    for (int j2 = 0; j2 < 3; ++j2) acc = acc + A[j1 * 3 + j2] * x[j2];
    o[j1] = acc;
  }
  x[0] = o[0]; x[1] = o[1]; x[2] = o[2];
}


}  // namespace ecb
