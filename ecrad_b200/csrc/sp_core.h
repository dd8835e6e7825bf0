// sp_core.h -- small-matrix arithmetic of the SPARTACUS solvers: one thread owns one g-point's matrices.
//
// Reference map (radiation/radiation_matrix.F90, which vectorises each operation over g-points; here one matrix at a time with
// the same element-wise operation order, so results agree with a non-contracting build to rounding):
//   sp_matmul            <- mat_x_mat :145-212 (dense and "shortwave" sparsity pattern), repeated_square :353-427
//   sp_lu / sp_lu_subst  <- lu_factorization :639-675, lu_substitution :681-707 (solve_mat_n :713, solve_vec :737)
//   sp_expm              <- expm :805-903 (scaling and squaring, degree-7 Pade approximant)
//   m3_*                 <- mat_x_vec :63, singlemat_x_vec :110, singlemat_x_mat :218, mat_x_singlemat :252,
//                           identity_minus_mat_x_mat :286, solve_vec_3 :484, solve_mat_3 :527, diag_mat_right_divide_3 :567
//   sp_fast_expm_exchange_3 <- fast_expm_exchange_3 :952-1028
//   sp_step_migrations   <- radiation_spartacus_sw.F90:1606-1721
// Matrices are row-major arrays, M[j1 * m + j2] = M(jg, j1, j2).  Big (6x6, 9x9) matrices live in thread-local memory;
// sp_matmul is register-blocked over three rows of the left operand (each element of the right operand is read once per row
// block) and is deliberately not inlined: one unrolled copy per (size, pattern).
#pragma once
#include <float.h>

#include "hd.h"

#if defined(__CUDACC__)
#define HDN static __host__ __device__ __noinline__
#else
#define HDN static
#endif

namespace ecb {

// C = A * B (C must not alias A or B).  SWP: operands and result have the block pattern (x x x; x x x; 0 0 x) of blocks M/3.
template <int M, bool SWP>
HDN void sp_matmul(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C) {
  constexpr int M2 = 2 * (M / 3);
#pragma unroll
  for (int r0 = 0; r0 < M; r0 += 3) {
    const bool low = SWP && r0 >= M2;
    double a[3][M];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int k = 0; k < M; ++k) a[i][k] = A[(r0 + i) * M + k];
#pragma unroll
    for (int j2 = 0; j2 < M; ++j2) {
      if (low && j2 < M2) {
        C[(r0 + 0) * M + j2] = 0.0; C[(r0 + 1) * M + j2] = 0.0; C[(r0 + 2) * M + j2] = 0.0;
      } else {
        const int k0 = low ? M2 : 0;
        const int k1 = (SWP && !low && j2 < M2) ? M2 : M;
        double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
#pragma unroll
        for (int k = k0; k < k1; ++k) {
          const double b = B[k * M + j2];
          acc0 = acc0 + a[0][k] * b;
          acc1 = acc1 + a[1][k] * b;
          acc2 = acc2 + a[2][k] * b;
        }
        C[(r0 + 0) * M + j2] = acc0; C[(r0 + 1) * M + j2] = acc1; C[(r0 + 2) * M + j2] = acc2;
      }
    }
  }
}

// In-place LU factorisation without pivoting (L below the diagonal with unit diagonal, U on and above).  Column j2 of the
// reference's Crout loop subtracts L(j1,j3)*U(j3,j2) for j3 ascending; the same sequence of operations per element here.
template <int M>
HDN void sp_lu(double* LU) {
#pragma unroll 1
  for (int j2 = 0; j2 < M; ++j2) {
#pragma unroll 1
    for (int j1 = 0; j1 < M; ++j1) {
      double s = LU[j1 * M + j2];
      const int n = j1 < j2 ? j1 : j2;
      for (int j3 = 0; j3 < n; ++j3) s = s - LU[j1 * M + j3] * LU[j3 * M + j2];
      LU[j1 * M + j2] = s;
    }
    if (j2 != M - 1) {
      const double s = 1.0 / LU[j2 * M + j2];
      for (int j1 = j2 + 1; j1 < M; ++j1) LU[j1 * M + j2] = LU[j1 * M + j2] * s;
    }
  }
}
// Solve LU x = b in place for `ncols` right-hand sides stored as the columns of X (row-major, leading dimension ldx).
template <int M>
HDN void sp_lu_subst(const double* LU, double* X, int ncols, int ldx) {
#pragma unroll 1
  for (int j = 0; j < ncols; ++j) {
    double x[M];
#pragma unroll
    for (int i = 0; i < M; ++i) x[i] = X[i * ldx + j];
#pragma unroll
    for (int j2 = 1; j2 < M; ++j2)
#pragma unroll
      for (int j1 = 0; j1 < j2; ++j1) x[j2] = x[j2] - x[j1] * LU[j2 * M + j1];
#pragma unroll
    for (int j2 = M - 1; j2 >= 0; --j2) {
#pragma unroll
      for (int j1 = j2 + 1; j1 < M; ++j1) x[j2] = x[j2] - x[j1] * LU[j2 * M + j1];
      x[j2] = x[j2] / LU[j2 * M + j2];
    }
#pragma unroll
    for (int i = 0; i < M; ++i) X[i * ldx + j] = x[i];
  }
}

// Matrix exponential of A (M x M, M = 6 or 9) in place; W: 3 * M * M doubles of thread-local work space.  Four matrices are
// live at any time (A and the three of W): the two Pade polynomials overwrite A^6 and A^4 element by element, the product
// U = A * (b7 A^6 + b5 A^4 + b3 A^2 + b1 I) overwrites A^2, and the squarings ping-pong between A and that buffer.
template <int M, bool SWP>
HD void sp_expm(double* A, double* W) {
  constexpr int MM = M * M;
  double *A2 = W, *A4 = W + MM, *A6 = W + 2 * MM;
  double normA = 0.0;
  for (int j3 = 0; j3 < M; ++j3) {
    double sum_column = 0.0;
    for (int j2 = 0; j2 < M; ++j2) sum_column = sum_column + fabs(A[j2 * M + j3]);
    if (sum_column > normA) normA = sum_column;
  }
  int expo = 0;
  const double frac = frexp(normA / 3.925724783138660e+00, &expo);   // fraction() / exponent()
  if (frac == 0.5) expo = expo - 1;
  if (expo < 0) expo = 0;
  const double scaling = ldexp(1.0, -expo);
  for (int i = 0; i < MM; ++i) A[i] = A[i] * scaling;
  sp_matmul<M, SWP>(A, A, A2);
  sp_matmul<M, SWP>(A2, A2, A4);
  sp_matmul<M, SWP>(A2, A4, A6);
  for (int i = 0; i < MM; ++i) {
    const double a2 = A2[i], a4 = A4[i], a6 = A6[i];
    double w1 = 1.0 * a6 + 1512.0 * a4 + 277200.0 * a2;        // c(8) A6 + c(6) A4 + c(4) A2
    double v = 56.0 * a6 + 25200.0 * a4 + 1995840.0 * a2;      // c(7) A6 + c(5) A4 + c(3) A2
    if (i % (M + 1) == 0) { w1 = w1 + 8648640.0; v = v + 17297280.0; }
    A6[i] = w1; A4[i] = v;
  }
  sp_matmul<M, SWP>(A, A6, A2);                                  // U
  for (int i = 0; i < MM; ++i) { A4[i] = A4[i] - A2[i]; A2[i] = 2.0 * A2[i]; }   // V - U, 2 U
  sp_lu<M>(A4);
  sp_lu_subst<M>(A4, A2, M, M);
  for (int j = 0; j < M; ++j) A2[j * M + j] = A2[j * M + j] + 1.0;
  double *src = A2, *dst = A;
  for (int k = 0; k < expo; ++k) {
    sp_matmul<M, SWP>(src, src, dst);
    double* t = src; src = dst; dst = t;
  }
  if (src != A) for (int i = 0; i < MM; ++i) A[i] = src[i];
}

// ---------------------------------------------------------------------------------------------------------
// 3 x 3
// ---------------------------------------------------------------------------------------------------------
HD void m3_x_m3(const double* A, const double* B, double* C) {   // C may alias A or B
  double R[9];
#pragma unroll
  for (int j2 = 0; j2 < 3; ++j2)
#pragma unroll
    for (int j1 = 0; j1 < 3; ++j1) {
      double acc = 0.0;
#pragma unroll
      for (int j3 = 0; j3 < 3; ++j3) acc = acc + A[j1 * 3 + j3] * B[j3 * 3 + j2];
      R[j1 * 3 + j2] = acc;
    }
#pragma unroll
  for (int i = 0; i < 9; ++i) C[i] = R[i];
}
HD void m3_x_vec(const double* A, const double* b, double* x) {   // x may alias b
  double r[3];
#pragma unroll
  for (int j1 = 0; j1 < 3; ++j1) {
    double acc = 0.0;
#pragma unroll
    for (int j2 = 0; j2 < 3; ++j2) acc = acc + A[j1 * 3 + j2] * b[j2];
    r[j1] = acc;
  }
  x[0] = r[0]; x[1] = r[1]; x[2] = r[2];
}
HD void m3_identity_minus_product(const double* A, const double* B, double* C) {
  m3_x_m3(A, B, C);
#pragma unroll
  for (int i = 0; i < 9; ++i) C[i] = -C[i];
  C[0] = 1.0 + C[0]; C[4] = 1.0 + C[4]; C[8] = 1.0 + C[8];
}
// Division by a value that is used as a divisor several times.  On the device: one correctly rounded reciprocal, then
// multiplications (within 1 ulp of the quotient; and a zero numerator -- the empty regions of cloud-free layers -- no longer sends
// the fp64 division through its slow path, which was a fifth of the sweep kernels' instructions).  On the host: the plain division,
// so the replay of this header stays bit-identical to the oracle (tests/test_core_hostcheck.py).
struct SpRcp { double b, r; };
HD SpRcp sp_rcp(double b) {
  SpRcp x; x.b = b;
#ifdef __CUDA_ARCH__
  x.r = __drcp_rn(b);
#else
  x.r = 0.0;
#endif
  return x;
}
HD double operator/(double a, const SpRcp& d) {
#ifdef __CUDA_ARCH__
  return a * d.r;
#else
  return a / d.b;
#endif
}
struct M3LU { double L21, L31, L32, U23; SpRcp A0, U22, U33; };
HD M3LU m3_lu(const double* A) {
  M3LU f;
  f.A0 = sp_rcp(A[0]);
  f.L21 = A[3] / f.A0; f.L31 = A[6] / f.A0;
  const double u22 = A[4] - f.L21 * A[1];
  f.U22 = sp_rcp(u22); f.U23 = A[5] - f.L21 * A[2];
  f.L32 = (A[7] - f.L31 * A[1]) / f.U22;
  f.U33 = sp_rcp(A[8] - f.L31 * A[2] - f.L32 * f.U23);
  return f;
}
HD void m3_solve_vec(const double* A, const double* b, double* x) {   // x may alias b
  const M3LU f = m3_lu(A);
  const double y2 = b[1] - f.L21 * b[0], y3 = b[2] - f.L31 * b[0] - f.L32 * y2;
  const double x3 = y3 / f.U33, x2 = (y2 - f.U23 * x3) / f.U22;
  const double x1 = (b[0] - A[1] * x2 - A[2] * x3) / f.A0;
  x[0] = x1; x[1] = x2; x[2] = x3;
}
HD void m3_solve_mat(const double* A, const double* B, double* X) {   // X may alias B
  const M3LU f = m3_lu(A);
  double R[9];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double y2 = B[3 + j] - f.L21 * B[j], y3 = B[6 + j] - f.L31 * B[j] - f.L32 * y2;
    R[6 + j] = y3 / f.U33;
    R[3 + j] = (y2 - f.U23 * R[6 + j]) / f.U22;
    R[j] = (B[j] - A[1] * R[3 + j] - A[2] * R[6 + j]) / f.A0;
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) X[i] = R[i];
}
// singlemat_x_mat(U, mat_x_singlemat(A, V))
HD void m3_u_a_v(const double* U, const double* A, const double* V, double* out) {
  double T[9];
  m3_x_m3(A, V, T);
  m3_x_m3(U, T, out);
}

HD void sp_diag_mat_right_divide_3(const double* A, const double* B, double* X) {
  const SpRcp a0 = sp_rcp(A[0]);
  const double L21 = A[1] / a0, L31 = A[2] / a0;
  const double u22 = A[4] - L21 * A[3], U23 = A[7] - L21 * A[6];
  const SpRcp U22 = sp_rcp(u22);
  const double L32 = (A[5] - L31 * A[3]) / U22;
  const SpRcp U33 = sp_rcp(A[8] - L31 * A[6] - L32 * U23);
  double y2 = -L21 * B[0], y3 = -L31 * B[0] - L32 * y2;
  X[2] = y3 / U33;
  X[1] = (y2 - U23 * X[2]) / U22;
  X[0] = (B[0] - A[3] * X[1] - A[6] * X[2]) / a0;
  y3 = -L32 * B[1];
  X[5] = y3 / U33;
  X[4] = (B[1] - U23 * X[5]) / U22;
  X[3] = (-A[3] * X[4] - A[6] * X[5]) / a0;
  X[8] = B[2] / U33;
  X[7] = -U23 * X[8] / U22;
  X[6] = (-A[3] * X[7] - A[6] * X[8]) / a0;
}
// exp of the two-region exchange matrix [[-a, b], [a, -b]] (fast_expm_exchange_2, radiation_matrix.F90:905-925), embedded in the
// 3x3 form of a two-region run (the empty third region maps to itself)
HD void sp_fast_expm_exchange_2(double a, double b, double* R) {
  const double factor = (1.0 - exp(-(a + b))) / dmax(1.0e-12, a + b);
  R[0] = 1.0 - factor * a; R[3] = factor * a; R[1] = factor * b; R[4] = 1.0 - factor * b;
  R[2] = 0.0; R[5] = 0.0; R[6] = 0.0; R[7] = 0.0; R[8] = 1.0;
}
// exp of the exchange matrix [[-a, b, 0], [a, -b-c, d], [0, c, -d]]
HD void sp_fast_expm_exchange_3(double a, double b, double c, double d, double* R) {
  const double my_epsilon = 1.0e-12;
  const double tmp1 = 0.5 * (a + b + c + d);
  double tmp2 = sqrt(dmax(0.0, tmp1 * tmp1 - (a * c + a * d + b * d)));
  tmp2 = dmax(tmp2, DBL_EPSILON * tmp1);
  const double lambda1 = -tmp1 + tmp2, lambda2 = -tmp1 - tmp2;
  double V[9], DV[9];
  V[0] = dmax(my_epsilon, b) / sp_rcp(copysign(dmax(my_epsilon, fabs(a + lambda1)), a + lambda1));
  V[1] = b / sp_rcp(copysign(dmax(my_epsilon, fabs(a + lambda2)), a + lambda2));   // (b, c = 0 between empty regions: see SpRcp)
  V[2] = b / sp_rcp(dmax(my_epsilon, a));
  V[3] = 1.0; V[4] = 1.0; V[5] = 1.0;
  V[6] = c / sp_rcp(copysign(dmax(my_epsilon, fabs(d + lambda1)), d + lambda1));
  V[7] = c / sp_rcp(copysign(dmax(my_epsilon, fabs(d + lambda2)), d + lambda2));
  V[8] = dmax(my_epsilon, c) / sp_rcp(dmax(my_epsilon, d));
  const double diag[3] = {exp(lambda1), exp(lambda2), 1.0};
  sp_diag_mat_right_divide_3(V, diag, DV);
#pragma unroll
  for (int j1 = 0; j1 < 3; ++j1)
#pragma unroll
    for (int j2 = 0; j2 < 3; ++j2) R[j2 * 3 + j1] = V[j2 * 3] * DV[j1] + V[j2 * 3 + 1] * DV[3 + j1] + V[j2 * 3 + 2] * DV[6 + j1];
}

// Scalars of config_type the SPARTACUS kernels read.
struct SpCfg {
  int do_3d_effects, entrapment, do_3d_lw_multilayer_effects, do_lw_side_emissivity, use_expm_everywhere;
  int two_regions;   // config%nregions == 2: one homogeneous cloudy region; carried as three regions with an empty third one
  double max_gas_od_3d, max_cloud_od, max_3d_transfer_rate, min_cloud_effective_size, overhead_sun_factor, overhang_factor,
      clear_to_thick_fraction;
};
enum { SP_ENTR_ZERO = 0, SP_ENTR_EDGE_ONLY = 1, SP_ENTR_EXPLICIT = 2, SP_ENTR_NON_FRACTAL = 3, SP_ENTR_MAXIMUM = 4 };

#define SP_PI 3.14159265358979323846
#define SP_R_OVER_G (287.058 / 9.80665)

// dz = dp R T / (p g), radiation_spartacus_sw.F90:434-441
HD double sp_layer_depth(double p_top, double p_bot, double t_top, double t_bot) {
  return SP_R_OVER_G * (p_bot - p_top) * (t_top + t_bot) / (p_top + p_bot);
}
// cloud edge lengths per unit gridbox area (radiation_spartacus_sw.F90:495-547); returns whether the layer has 3D transfer
HD bool sp_edge_lengths(const SpCfg& c, const double* reg, double inv_cloud_size, bool have_inhom_size, double inv_inhom_size, double* edge) {
  edge[0] = 0.0; edge[1] = 0.0; edge[2] = 0.0;
  if (!c.do_3d_effects || !(inv_cloud_size > 0.0)) return false;
  const double four_over_pi = 4.0 / SP_PI, max_inv = 1.0 / c.min_cloud_effective_size;
  edge[0] = four_over_pi * reg[0] * (1.0 - reg[0]) * dmin(inv_cloud_size, max_inv);
  edge[1] = four_over_pi * reg[2] * (1.0 - reg[2]) * dmin(have_inhom_size ? inv_inhom_size : inv_cloud_size, max_inv);
  if (c.clear_to_thick_fraction > 0.0) {
    edge[2] = c.clear_to_thick_fraction * dmin(edge[0], edge[1]);
    edge[0] = edge[0] - edge[2];
    edge[1] = edge[1] - edge[2];
  }
  return true;
}
// transfer_rate(i,j): rate of lateral transfer from region i to region j times the layer depth (:549-604); rate[i*3+j]
HD void sp_transfer_rates(const SpCfg& c, double dz, const double* edge, const double* reg, double tan_angle, double* rate) {
#pragma unroll
  for (int i = 0; i < 9; ++i) rate[i] = 0.0;
#pragma unroll
  for (int jreg = 0; jreg < 2; ++jreg) {
    if (reg[jreg] > DBL_EPSILON) rate[jreg * 3 + jreg + 1] = dz * edge[jreg] * tan_angle / reg[jreg];
    if (reg[jreg + 1] > DBL_EPSILON) rate[(jreg + 1) * 3 + jreg] = dz * edge[jreg] * tan_angle / reg[jreg + 1];
  }
  if (edge[2] > 0.0) {
    if (reg[0] > DBL_EPSILON) rate[2] = dz * edge[2] * tan_angle / reg[0];
    if (reg[2] > DBL_EPSILON) rate[6] = dz * edge[2] * tan_angle / reg[2];
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) if (rate[i] > c.max_3d_transfer_rate) rate[i] = c.max_3d_transfer_rate;
}

// horizontal migration distances from the base of a layer to its top (step_migrations); matrices 3x3 row-major
HD void sp_step_migrations(double cloud_frac, double layer_depth, double tan_diffuse_angle_3d, double tan_sza, const double* reflectance,
                           const double* transmittance, const double* ref_dir, const double* trans_dir_dir, const double* trans_dir_diff,
                           const double* ta_diff, const double* ta_dir, double* x_diffuse, double* x_direct) {
  int istartreg = 0, iendreg = 3;
  if (cloud_frac <= 0.0) iendreg = 1;
  else if (cloud_frac >= 1.0) istartreg = 1;
  const double x_layer_diffuse = layer_depth * tan_diffuse_angle_3d / sqrt(2.0);
  const double x_layer_direct = layer_depth * sqrt(tan_sza * tan_sza + tan_diffuse_angle_3d * tan_diffuse_angle_3d) * 0.5;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    if (r < istartreg || r >= iendreg) continue;
    const int d = r * 4;
    const double ms_enhancement = transmittance[d] / (1.0 - reflectance[d] * ta_diff[d]);
    const double x_enhancement = pow(1.0 - reflectance[d] * ta_diff[d], -1.5);
    double top_albedo = dmax(1.0e-8, ref_dir[d] + ms_enhancement * (trans_dir_diff[d] * ta_diff[d] + trans_dir_dir[d] * ta_dir[d]));
    x_direct[r] = dmax(0.0, x_layer_direct + ((trans_dir_diff[d] * ta_diff[d] * x_enhancement + trans_dir_dir[d] * ta_dir[d] * (x_enhancement - 1.0)) *
                                                  (x_diffuse[r] + x_layer_diffuse) +
                                              trans_dir_dir[d] * ta_dir[d] * (x_direct[r] + x_layer_direct)) *
                                                 transmittance[d] / top_albedo);
    top_albedo = dmax(1.0e-8, reflectance[d] + ms_enhancement * transmittance[d] * ta_diff[d]);
    x_diffuse[r] = x_layer_diffuse + x_enhancement * ta_diff[d] * (transmittance[d] * transmittance[d]) * (x_diffuse[r] + x_layer_diffuse) / top_albedo;
  }
  if (iendreg < 3) { x_diffuse[1] = 0.0; x_diffuse[2] = 0.0; x_direct[1] = 0.0; x_direct[2] = 0.0; }
  else if (istartreg == 1) { x_diffuse[0] = 0.0; x_direct[0] = 0.0; }
}

// exchange between the sub-regions of one lower region during entrapment (radiation_spartacus_sw.F90:1124-1180); rate[i*3+j]
HD void sp_entrapment_part(const SpCfg& c, const double* rate, double x, double inv_effective_size, double* part) {
  double e[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) e[i] = 0.0;
#pragma unroll
  for (int jreg = 0; jreg < 2; ++jreg) {
    if (c.entrapment == SP_ENTR_EXPLICIT) {
      const double fractal_factor = 1.0 / sqrt(dmax(1.0, 2.5 * x * inv_effective_size));
      e[(jreg + 1) * 3 + jreg] = e[(jreg + 1) * 3 + jreg] + rate[jreg * 3 + jreg + 1] * x * fractal_factor;
      e[jreg * 3 + jreg + 1] = e[jreg * 3 + jreg + 1] + rate[(jreg + 1) * 3 + jreg] * x * fractal_factor;
    } else {
      e[(jreg + 1) * 3 + jreg] = e[(jreg + 1) * 3 + jreg] + rate[jreg * 3 + jreg + 1] * x;
      e[jreg * 3 + jreg + 1] = e[jreg * 3 + jreg + 1] + rate[(jreg + 1) * 3 + jreg] * x;
    }
    e[jreg * 4] = e[jreg * 4] - e[(jreg + 1) * 3 + jreg];
    e[(jreg + 1) * 4] = e[(jreg + 1) * 4] - e[jreg * 3 + jreg + 1];
  }
  const double max_entr = -dmin(e[0], e[4]);
  if (max_entr > c.max_cloud_od) {
    const double s = c.max_cloud_od / max_entr;
#pragma unroll
    for (int i = 0; i < 9; ++i) e[i] = e[i] * s;
  }
  if (c.two_regions) sp_fast_expm_exchange_2(e[3], e[1], part);   // radiation_spartacus_sw.F90:1184-1186
  else sp_fast_expm_exchange_3(e[3], e[1], e[7], e[5], part);
}

}  // namespace ecb
