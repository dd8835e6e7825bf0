// sp_coop.cuh -- warp-cooperative matrix exponential for the SPARTACUS layer kernels.
//
// Reference: radiation/radiation_matrix.F90:805-903 (expm: scaling and squaring with the degree-7 Pade approximant), :145-265
// (mat_x_mat and its shortwave sparsity pattern), :531-700 (LU factorisation / substitution without pivoting).
//
// One matrix is worked on by M lanes (M = 9 shortwave, 6 longwave): lane j keeps COLUMN j of every matrix of the algorithm
// (A^2, A^4, A^6, the two Pade polynomials, the squarings) in registers.  A product C = A B is then, per lane,
// C(:,j) = sum_k A(:,k) B(k,j): B(k,j) is a register, A(:,k) a column of the left operand read from a small shared-memory copy
// (column-major, 16-byte vector loads, every lane of the group reads the same address: a broadcast).  3 (M = 9) or 5 (M = 6)
// matrices per warp.  A lane's own matrix (the one its g-point owns) waits in thread-local memory and is touched twice: copied into
// the group's buffer when its round comes, copied back afterwards; the exponential itself runs from registers and shared memory.  The linear solve of the Pade step is a right-looking LU without pivoting
// (the reference does not pivot either): the owner of column k publishes the multipliers, every lane eliminates in its own column
// of the matrix and of the right-hand side, then substitutes backwards column by column.
//
// Same algorithm and the same operations per element as sp_expm (sp_core.h, the sequential statement that tests/ replays against
// the oracle); multiply-adds are fused and the order of the LU updates differs, so results agree to rounding (1e-15 relative), not
// bit for bit.  Shortwave pattern (x x x; x x x; 0 0 x in 3x3 blocks): the structurally zero rows 6-8 of the left operand's columns
// 0-5 are skipped by all lanes alike (63 instead of 81 multiply-adds per lane and product); structural zeros stay exact zeros.
#pragma once
#include "hd.h"

namespace ecb {

template <int M>
struct Coop {
  static constexpr int NGRP = 32 / M;                  // matrices per warp and round
  static constexpr int CS = (M % 2) ? M + 1 : M;       // column stride of the shared-memory copies (16-byte aligned columns)
  static constexpr int BUF = M * CS;                   // doubles per copy
  static constexpr int PER_WARP = (NGRP + 1) * 2 * BUF;   // two copies per group + one dummy group for the lanes left over
};

// c(:) = sum_k As(:,k) * b[k]; even and odd k accumulate separately (two independent chains per element: the warp has few
// neighbours to hide the multiply-add latency behind)
template <int M, bool SWP>
__device__ __forceinline__ void coop_mm(const double* __restrict__ As, const double (&b)[M], double (&c)[M]) {
  constexpr int CS = Coop<M>::CS, M2 = 2 * (M / 3);
  double c1[M];
#pragma unroll
  for (int i = 0; i < M; ++i) { c[i] = 0.0; c1[i] = 0.0; }
#pragma unroll
  for (int k = 0; k < M; ++k) {
    const int rows = (SWP && k < M2) ? M2 : M;
    double a[CS];
#pragma unroll
    for (int i = 0; i < CS; i += 2)
      if (i < rows) { const double2 v = *reinterpret_cast<const double2*>(As + k * CS + i); a[i] = v.x; a[i + 1] = v.y; }
#pragma unroll
    for (int i = 0; i < M; ++i)
      if (i < rows) { if (k & 1) c1[i] = fma(a[i], b[k], c1[i]); else c[i] = fma(a[i], b[k], c[i]); }
  }
#pragma unroll
  for (int i = 0; i < M; ++i) c[i] = c[i] + c1[i];
}

template <int M>
__device__ __forceinline__ void coop_store_col(double* buf, int j, const double (&v)[M]) {
  constexpr int CS = Coop<M>::CS;
#pragma unroll
  for (int i = 0; i + 1 < CS; i += 2) *reinterpret_cast<double2*>(buf + j * CS + i) = make_double2(v[i], i + 1 < M ? v[i + 1] : 0.0);
}

// a: column j of the matrix (in), column j of its exponential (out).  bufA, bufX: the group's two shared-memory copies.
// lane0: first lane of the group.  Called by all 32 lanes of the warp (groups without a matrix pass zeros and get the identity).
template <int M, bool SWP>
__device__ __forceinline__ void coop_expm(double (&a)[M], double* bufA, double* bufX, int j, int lane0) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int CS = Coop<M>::CS;
  // 1-norm = largest column sum of |a| (radiation_matrix.F90:838-846)
  double colsum = 0.0;
#pragma unroll
  for (int i = 0; i < M; ++i) colsum = colsum + fabs(a[i]);
  double normA = 0.0;
#pragma unroll
  for (int t = 0; t < M; ++t) { const double v = __shfl_sync(FULL, colsum, lane0 + t); if (v > normA) normA = v; }
  int expo = 0;
  const double frac = frexp(normA / 3.925724783138660e+00, &expo);
  if (frac == 0.5) expo = expo - 1;
  if (expo < 0) expo = 0;
  if (!(normA == normA) || expo > 1100) expo = 0;   // (a NaN matrix stays what it is; keeps the warp's loop count finite)
  const double scaling = ldexp(1.0, -expo);
#pragma unroll
  for (int i = 0; i < M; ++i) a[i] = a[i] * scaling;
  coop_store_col<M>(bufA, j, a);
  __syncwarp();
  double a2[M], a4[M], a6[M], u[M];
  coop_mm<M, SWP>(bufA, a, a2);
  coop_store_col<M>(bufX, j, a2);
  __syncwarp();
  coop_mm<M, SWP>(bufX, a2, a4);
  coop_mm<M, SWP>(bufX, a4, a6);
  // Pade polynomials (:858-872): W1 = c8 A6 + c6 A4 + c4 A2 + c2 I (left-multiplied by A below), V = c7 A6 + c5 A4 + c3 A2 + c1 I
#pragma unroll
  for (int i = 0; i < M; ++i) {
    double w1 = 1.0 * a6[i] + 1512.0 * a4[i] + 277200.0 * a2[i];
    double v = 56.0 * a6[i] + 25200.0 * a4[i] + 1995840.0 * a2[i];
    if (i == j) { w1 = w1 + 8648640.0; v = v + 17297280.0; }
    a6[i] = w1; a4[i] = v;
  }
  coop_mm<M, SWP>(bufA, a6, u);                                   // U = A W1
  double q[M], b[M];
#pragma unroll
  for (int i = 0; i < M; ++i) { q[i] = a4[i] - u[i]; b[i] = 2.0 * u[i]; }   // (V - U) X = 2 U;  exp = X + I
  // right-looking LU without pivoting; the multipliers of step k travel through bufX (A^2 is no longer needed).  The owner of
  // column k also publishes 1 / U(k,k): the back substitution multiplies by it (one division per lane in all; structurally zero
  // right-hand-side entries would otherwise send every 0 / U(k,k) through the slow path of the fp64 division)
  double* dinv = bufX + 2 * M;
  __syncwarp();
#pragma unroll
  for (int k = 0; k < M; ++k) {
    double* mult = bufX + (k & 1) * M;
    if (j == k) {
      const double inv = 1.0 / q[k];
      dinv[k] = inv;
#pragma unroll
      for (int i = k + 1; i < M; ++i) mult[i] = q[i] * inv;
    }
    if (k == M - 1) break;
    __syncwarp();
#pragma unroll
    for (int i = k + 1; i < M; ++i) {
      const double m = mult[i];
      q[i] = fma(-m, q[k], q[i]);
      b[i] = fma(-m, b[k], b[i]);
    }
  }
  // back substitution, column by column of U (published through bufA: the product U = A W1 above was its last reader, and the
  // barriers of the elimination lie in between)
  coop_store_col<M>(bufA, j, q);
  __syncwarp();
#pragma unroll
  for (int jj = M - 1; jj >= 0; --jj) {
    const double x = b[jj] * dinv[jj];
    b[jj] = x;
#pragma unroll
    for (int i = 0; i < jj; ++i) b[i] = fma(-bufA[jj * CS + i], x, b[i]);
  }
#pragma unroll
  for (int i = 0; i < M; ++i) a[i] = b[i] + (i == j ? 1.0 : 0.0);
  // repeated squaring (:889-897); the warp runs as many rounds as its slowest group
  const int rounds = __reduce_max_sync(FULL, expo);
  for (int s = 0; s < rounds; ++s) {
    __syncwarp();
    coop_store_col<M>(bufX, j, a);
    __syncwarp();
    double y[M];
    coop_mm<M, SWP>(bufX, a, y);
    if (s < expo) {
#pragma unroll
      for (int i = 0; i < M; ++i) a[i] = y[i];
    }
  }
}

// Entry numbering of a lane's own matrix Gl[]: dense row-major (M = 6), or the 63 entries of the shortwave pattern (rows 0-5
// complete, rows 6-8 columns 6-8).
template <int M, bool SWP>
__host__ __device__ constexpr int coop_entry(int i, int j) {
  constexpr int M2 = 2 * (M / 3);
  return !SWP ? i * M + j : (i < M2 ? i * M + j : M2 * M + (i - M2) * (M - M2) + (j - M2));
}
template <int M, bool SWP>
struct CoopNE { static constexpr int value = SWP ? 2 * (M / 3) * M + (M / 3) * (M / 3) : M * M; };

// Exponentials of the matrices of the lanes in `need` (a warp-uniform mask), in place.  Gl: this lane's matrix (thread-local; it is
// touched twice: copied into the group's shared-memory buffer when its round comes, and copied back).  stage: Coop<M>::PER_WARP
// doubles of this warp.  All 32 lanes call.
template <int M, bool SWP>
__device__ __forceinline__ void coop_expm_warp(double (&Gl)[CoopNE<M, SWP>::value], unsigned need, double* stage) {
  constexpr int NGRP = Coop<M>::NGRP, BUF = Coop<M>::BUF, CS = Coop<M>::CS, M2 = 2 * (M / 3);
  const int lane = threadIdx.x & 31;
  const int grp = lane / M < NGRP ? lane / M : NGRP;   // lanes left over form a dummy group
  const int lane0 = grp * M, j = lane - lane0;
  double* bufA = stage + (size_t)grp * 2 * BUF;
  double* bufX = bufA + BUF;
  unsigned rem = need;
  while (rem) {   // warp-uniform
    int src = -1, myq = -1;   // src: the lane whose matrix this group works on; myq: the group that works on this lane's matrix
#pragma unroll
    for (int qq = 0; qq < NGRP; ++qq) {
      const int t = rem ? __ffs(rem) - 1 : -1;
      if (rem) rem &= rem - 1;
      if (qq == grp) src = t;
      if (t == lane) myq = qq;
    }
    if (myq >= 0) {   // hand this lane's matrix to its group: column-major, structural zeros written out
      double* dst = stage + (size_t)myq * 2 * BUF;
#pragma unroll
      for (int k = 0; k < M; ++k)
#pragma unroll
        for (int i = 0; i < M; ++i) dst[k * CS + i] = (!SWP || i < M2 || k >= M2) ? Gl[coop_entry<M, SWP>(i, k)] : 0.0;
    }
    __syncwarp();
    double a[M];
#pragma unroll
    for (int i = 0; i < M; ++i) a[i] = src >= 0 ? bufA[j * CS + i] : 0.0;
    __syncwarp();
    coop_expm<M, SWP>(a, bufA, bufX, j, lane0);
    __syncwarp();
    coop_store_col<M>(bufX, j, a);
    __syncwarp();
    if (myq >= 0) {
      const double* srcb = stage + (size_t)myq * 2 * BUF + BUF;
#pragma unroll
      for (int k = 0; k < M; ++k)
#pragma unroll
        for (int i = 0; i < M; ++i)
          if (!SWP || i < M2 || k >= M2) Gl[coop_entry<M, SWP>(i, k)] = srcb[k * CS + i];
    }
    __syncwarp();
  }
}

// The same with the lanes' matrices in shared memory: entry e of lane t at Gs[e * SP_LD + t] (odd stride: the M lanes that fetch one
// lane's column hit different banks).  Costs SP_LD * entries doubles per warp, saves the two copies through thread-local memory:
// the longwave kernel (36 entries, register budget of 160-thread CTAs) is faster this way, the shortwave one (63 entries) the other.
enum { SP_LD = 33 };
template <int M, bool SWP>
__device__ __forceinline__ void coop_expm_warp_shared(double* Gs, unsigned need, double* stage) {
  constexpr int NGRP = Coop<M>::NGRP, BUF = Coop<M>::BUF, M2 = 2 * (M / 3);
  const int lane = threadIdx.x & 31;
  const int grp = lane / M < NGRP ? lane / M : NGRP;
  const int lane0 = grp * M, j = lane - lane0;
  double* bufA = stage + (size_t)grp * 2 * BUF;
  double* bufX = bufA + BUF;
  unsigned rem = need;
  while (rem) {   // warp-uniform
    int src = -1;
#pragma unroll
    for (int qq = 0; qq < NGRP; ++qq) {
      const int t = rem ? __ffs(rem) - 1 : -1;
      if (rem) rem &= rem - 1;
      if (qq == grp) src = t;
    }
    double a[M];
#pragma unroll
    for (int i = 0; i < M; ++i) {
      const bool inpat = !SWP || i < M2 || j >= M2;
      a[i] = (src >= 0 && inpat) ? Gs[coop_entry<M, SWP>(i, j) * SP_LD + src] : 0.0;
    }
    coop_expm<M, SWP>(a, bufA, bufX, j, lane0);
    if (src >= 0) {
#pragma unroll
      for (int i = 0; i < M; ++i)
        if (!SWP || i < M2 || j >= M2) Gs[coop_entry<M, SWP>(i, j) * SP_LD + src] = a[i];
    }
    __syncwarp();
  }
}

// Longwave: besides the exponential, the particular solution of the inhomogeneous system needs two solves with the layer matrix
// itself (radiation_spartacus_lw.F90:697-707: solution_diff = -Gamma^-1 planck_diff, solution0 = Gamma^-1 (solution_diff - planck_top)).
// Same right-looking LU, all multipliers kept (bufX, the reciprocal pivots on its diagonal), U published through bufA; the right-hand
// sides are short vectors that every lane of the group carries and updates redundantly.  q: column j of Gamma.
template <int M>
__device__ __forceinline__ void coop_lu(double (&q)[M], double* bufA, double* bufX, int j) {
  __syncwarp();
#pragma unroll
  for (int k = 0; k < M; ++k) {
    if (j == k) {
      const double inv = 1.0 / q[k];
      bufX[k * M + k] = inv;
#pragma unroll
      for (int i = k + 1; i < M; ++i) bufX[k * M + i] = q[i] * inv;
    }
    if (k == M - 1) break;
    __syncwarp();
#pragma unroll
    for (int i = k + 1; i < M; ++i) q[i] = fma(-bufX[k * M + i], q[k], q[i]);
  }
  coop_store_col<M>(bufA, j, q);
  __syncwarp();
}
template <int M>
__device__ __forceinline__ void coop_lu_solve(const double* bufA, const double* bufX, double (&r)[M]) {
  constexpr int CS = Coop<M>::CS;
#pragma unroll
  for (int k = 0; k < M - 1; ++k)
#pragma unroll
    for (int i = k + 1; i < M; ++i) r[i] = fma(-bufX[k * M + i], r[k], r[i]);
#pragma unroll
  for (int jj = M - 1; jj >= 0; --jj) {
    const double x = r[jj] * bufX[jj * M + jj];
    r[jj] = x;
#pragma unroll
    for (int i = 0; i < jj; ++i) r[i] = fma(-bufA[jj * CS + i], x, r[i]);
  }
}
// planck_top / planck_diff (in) and solution0 / solution_diff (out) belong to the lane that owns the matrix; they travel by shuffles.
template <int M>
__device__ __forceinline__ void coop_expm_warp_shared_lw(double* Gs, unsigned need, double* stage, const double (&planck_top)[M],
                                                         const double (&planck_diff)[M], double (&solution0)[M], double (&solution_diff)[M]) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int NGRP = Coop<M>::NGRP, BUF = Coop<M>::BUF;
  const int lane = threadIdx.x & 31;
  const int grp = lane / M < NGRP ? lane / M : NGRP;
  const int lane0 = grp * M, j = lane - lane0;
  double* bufA = stage + (size_t)grp * 2 * BUF;
  double* bufX = bufA + BUF;
  unsigned rem = need;
  while (rem) {   // warp-uniform
    int src = -1, myq = -1;
#pragma unroll
    for (int qq = 0; qq < NGRP; ++qq) {
      const int t = rem ? __ffs(rem) - 1 : -1;
      if (rem) rem &= rem - 1;
      if (qq == grp) src = t;
      if (t == lane) myq = qq;
    }
    const int from = src >= 0 ? src : lane;
    double a[M], q[M], sd[M], s0[M];
#pragma unroll
    for (int i = 0; i < M; ++i) {
      a[i] = src >= 0 ? Gs[(i * M + j) * SP_LD + src] : (i == j ? 1.0 : 0.0);   // (groups without a matrix factorise the identity)
      q[i] = a[i];
      sd[i] = __shfl_sync(FULL, planck_diff[i], from);
      s0[i] = __shfl_sync(FULL, planck_top[i], from);
    }
    coop_lu<M>(q, bufA, bufX, j);
    coop_lu_solve<M>(bufA, bufX, sd);
#pragma unroll
    for (int i = 0; i < M; ++i) { sd[i] = -sd[i]; s0[i] = sd[i] - s0[i]; }
    coop_lu_solve<M>(bufA, bufX, s0);
    // hand the two vectors back to the owner (every lane of the group holds them; take the group's first lane)
    const int back = myq >= 0 ? myq * M : lane;
#pragma unroll
    for (int i = 0; i < M; ++i) {
      const double v0 = __shfl_sync(FULL, s0[i], back), vd = __shfl_sync(FULL, sd[i], back);
      if (myq >= 0) { solution0[i] = v0; solution_diff[i] = vd; }
    }
    if (src < 0) {
#pragma unroll
      for (int i = 0; i < M; ++i) a[i] = 0.0;
    }
    __syncwarp();
    coop_expm<M, false>(a, bufA, bufX, j, lane0);
    if (src >= 0) {
#pragma unroll
      for (int i = 0; i < M; ++i) Gs[(i * M + j) * SP_LD + src] = a[i];
    }
    __syncwarp();
  }
}

}  // namespace ecb
