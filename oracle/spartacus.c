/* spartacus.c -- oracle restatement of the SPARTACUS solvers (3 regions, 3D effects).  TEST INFRASTRUCTURE.
 * Follows radiation/radiation_spartacus_sw.F90:64-1600 (solver_spartacus_sw) and :1606-1721 (step_migrations),
 * radiation_spartacus_lw.F90:50-1085 (solver_spartacus_lw), radiation_matrix.F90 (mat_x_vec :63, singlemat_x_vec :110, mat_x_mat :145,
 * singlemat_x_mat :218, mat_x_singlemat :252, identity_minus_mat_x_mat :286, repeated_square :353, solve_vec_3 :484, solve_mat_3 :527,
 * diag_mat_right_divide_3 :567, lu_factorization :639, lu_substitution :681, solve_mat_n :713, expm :805-903,
 * fast_expm_exchange_3 :952-1028), radiation_two_stream.F90 (calc_two_stream_gammas_lw :51, _sw :96,
 * calc_reflectance_transmittance_lw :148, _sw :421), radiation_lw_derivatives.F90:138-193 (calc_lw_derivatives_matrix).
 * The reference vectorises every matrix operation over g-points (first array index); here each g-point's small matrices are
 * handled one at a time with the same element-wise operation order.  nregions = 3.
 *
 * PARITY UNPINNED: the reference ships no SPARTACUS output (ctest `spartacus*` are XFAIL_VALIDATION without a reference file,
 * SURVEY 8c gap iv) and cannot be built here.  What pins this file: the matrix routines against scipy (expm, solve) in
 * tests/test_oracle_spartacus.py, and the solver against the golden-pinned Tripleclouds restatement in the limit
 * do_3d_effects = false, where the two schemes solve the same equations.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#undef NG_LW
#undef NG_SW
#undef NB_LW
#undef NB_SW
#define NG_LW (cfg->n_g_lw)
#define NG_SW (cfg->n_g_sw)
#define NB_LW (cfg->n_bands_lw)
#define NB_SW (cfg->n_bands_sw)

#define NREG 3
#define MMAX 9
static inline double dmin(double a, double b) { return a < b ? a : b; }
static inline double dmax(double a, double b) { return a > b ? a : b; }

void orc_region_properties(int nlev, const double* frac, const double* fsd, double frac_threshold, int lognormal, double (*reg_fracs)[NREG],
                           double (*od_scaling)[NREG]);
void orc_overlap_matrices(int nlev, double (*reg_fracs)[NREG], const double* overlap_param, double decorrelation_scaling,
                          double frac_threshold, int use_beta_overlap, double (*U)[NREG][NREG], double (*V)[NREG][NREG], double* cloud_cover);

static const double Pi = 3.14159265358979323846, GasConstantDryAir = 287.058, AccelDueToGravity = 9.80665;
static const double LwDiffusivity = 1.66;

/* ---------------------------------------------------------------------------------------------------------
 * radiation_matrix.F90 for ONE matrix; matrices are m x m in a [MMAX][MMAX] frame, M[j1][j2] = M(jg,j1,j2)
 * --------------------------------------------------------------------------------------------------------- */
typedef double mat[MMAX][MMAX];
typedef double m3[NREG][NREG];

/* mat_x_mat :145-212 (dense or shortwave sparsity pattern) */
static void mat_x_mat(int m, mat A, mat B, mat C, int sw_pattern) {
  mat R;
  for (int j1 = 0; j1 < m; ++j1) for (int j2 = 0; j2 < m; ++j2) R[j1][j2] = 0.0;
  if (sw_pattern) {
    const int mblock = m / 3, m2block = 2 * mblock;
    for (int j2 = 0; j2 < m2block; ++j2)
      for (int j1 = 0; j1 < m2block; ++j1)
        for (int j3 = 0; j3 < m2block; ++j3) R[j1][j2] = R[j1][j2] + A[j1][j3] * B[j3][j2];
    for (int j2 = m2block; j2 < m; ++j2) {
      for (int j1 = 0; j1 < m2block; ++j1)
        for (int j3 = 0; j3 < m; ++j3) R[j1][j2] = R[j1][j2] + A[j1][j3] * B[j3][j2];
      for (int j1 = m2block; j1 < m; ++j1)
        for (int j3 = m2block; j3 < m; ++j3) R[j1][j2] = R[j1][j2] + A[j1][j3] * B[j3][j2];
    }
  } else {
    for (int j2 = 0; j2 < m; ++j2)
      for (int j1 = 0; j1 < m; ++j1)
        for (int j3 = 0; j3 < m; ++j3) R[j1][j2] = R[j1][j2] + A[j1][j3] * B[j3][j2];
  }
  for (int j1 = 0; j1 < m; ++j1) for (int j2 = 0; j2 < m; ++j2) C[j1][j2] = R[j1][j2];
}

/* lu_factorization :639-675 (no pivoting) */
static void lu_factorization(int m, mat A, mat LU) {
  for (int j1 = 0; j1 < m; ++j1) for (int j2 = 0; j2 < m; ++j2) LU[j1][j2] = A[j1][j2];
  for (int j2 = 0; j2 < m; ++j2) {
    for (int j1 = 0; j1 < j2; ++j1) {
      double s = LU[j1][j2];
      for (int j3 = 0; j3 < j1; ++j3) s = s - LU[j1][j3] * LU[j3][j2];
      LU[j1][j2] = s;
    }
    for (int j1 = j2; j1 < m; ++j1) {
      double s = LU[j1][j2];
      for (int j3 = 0; j3 < j2; ++j3) s = s - LU[j1][j3] * LU[j3][j2];
      LU[j1][j2] = s;
    }
    if (j2 != m - 1) {
      const double s = 1.0 / LU[j2][j2];
      for (int j1 = j2 + 1; j1 < m; ++j1) LU[j1][j2] = LU[j1][j2] * s;
    }
  }
}
/* lu_substitution :681-707 */
static void lu_substitution(int m, mat LU, const double* b, double* x) {
  for (int j = 0; j < m; ++j) x[j] = b[j];
  for (int j2 = 1; j2 < m; ++j2)
    for (int j1 = 0; j1 < j2; ++j1) x[j2] = x[j2] - x[j1] * LU[j2][j1];
  for (int j2 = m - 1; j2 >= 0; --j2) {
    for (int j1 = j2 + 1; j1 < m; ++j1) x[j2] = x[j2] - x[j1] * LU[j2][j1];
    x[j2] = x[j2] / LU[j2][j2];
  }
}
/* solve_vec_3 :484-521 */
static void solve_vec_3(m3 A, const double* b, double* x) {
  const double L21 = A[1][0] / A[0][0], L31 = A[2][0] / A[0][0];
  const double U22 = A[1][1] - L21 * A[0][1], U23 = A[1][2] - L21 * A[0][2];
  const double L32 = (A[2][1] - L31 * A[0][1]) / U22;
  const double U33 = A[2][2] - L31 * A[0][2] - L32 * U23;
  const double y2 = b[1] - L21 * b[0], y3 = b[2] - L31 * b[0] - L32 * y2;
  const double x3 = y3 / U33, x2 = (y2 - U23 * x3) / U22;
  const double x1 = (b[0] - A[0][1] * x2 - A[0][2] * x3) / A[0][0];
  x[0] = x1; x[1] = x2; x[2] = x3;
}
/* solve_mat_3 :527-561 */
static void solve_mat_3(m3 A, m3 B, m3 X) {
  const double L21 = A[1][0] / A[0][0], L31 = A[2][0] / A[0][0];
  const double U22 = A[1][1] - L21 * A[0][1], U23 = A[1][2] - L21 * A[0][2];
  const double L32 = (A[2][1] - L31 * A[0][1]) / U22;
  const double U33 = A[2][2] - L31 * A[0][2] - L32 * U23;
  m3 R;
  for (int j = 0; j < 3; ++j) {
    const double y2 = B[1][j] - L21 * B[0][j], y3 = B[2][j] - L31 * B[0][j] - L32 * y2;
    R[2][j] = y3 / U33;
    R[1][j] = (y2 - U23 * R[2][j]) / U22;
    R[0][j] = (B[0][j] - A[0][1] * R[1][j] - A[0][2] * R[2][j]) / A[0][0];
  }
  memcpy(X, R, sizeof(m3));
}
/* solve_mat :769-797 for m > 3 (solve_mat_n :713-730) */
static void solve_mat_n(int m, mat A, mat B, mat X) {
  mat LU, R;
  lu_factorization(m, A, LU);
  for (int j = 0; j < m; ++j) {
    double b[MMAX], x[MMAX];
    for (int i = 0; i < m; ++i) b[i] = B[i][j];
    lu_substitution(m, LU, b, x);
    for (int i = 0; i < m; ++i) R[i][j] = x[i];
  }
  for (int j1 = 0; j1 < m; ++j1) for (int j2 = 0; j2 < m; ++j2) X[j1][j2] = R[j1][j2];
}
/* solve_vec :737-762 for m > 3 */
static void solve_vec_n(int m, mat A, const double* b, double* x) {
  mat LU;
  lu_factorization(m, A, LU);
  lu_substitution(m, LU, b, x);
}
/* repeated_square :353-427 */
static void repeated_square(int m, mat A, int nrepeat, int sw_pattern) {
  for (int j4 = 0; j4 < nrepeat; ++j4) mat_x_mat(m, A, A, A, sw_pattern);   /* mat_x_mat works on a copy, then stores */
}
/* expm :805-903: scaling and squaring with the degree-7 Pade approximant; in place */
static void expm(int m, mat A, int sw_pattern) {
  static const double theta3 = 3.925724783138660e+00;
  static const double c[8] = {17297280.0, 8648640.0, 1995840.0, 277200.0, 25200.0, 1512.0, 56.0, 1.0};
  double normA = 0.0;
  for (int j3 = 0; j3 < m; ++j3) {
    double sum_column = 0.0;
    for (int j2 = 0; j2 < m; ++j2) sum_column = sum_column + fabs(A[j2][j3]);
    if (sum_column > normA) normA = sum_column;
  }
  int expo = 0;
  const double frac = frexp(normA / theta3, &expo);   /* Fortran fraction()/exponent(): x = frac * 2**expo, 0.5 <= frac < 1 */
  if (frac == 0.5) expo = expo - 1;
  if (expo < 0) expo = 0;
  const double scaling = ldexp(1.0, -expo);           /* 2.0**(-expo), exact */
  for (int j3 = 0; j3 < m; ++j3) for (int j2 = 0; j2 < m; ++j2) A[j2][j3] = A[j2][j3] * scaling;
  mat A2, A4, A6, U, V;
  mat_x_mat(m, A, A, A2, sw_pattern);
  mat_x_mat(m, A2, A2, A4, sw_pattern);
  mat_x_mat(m, A2, A4, A6, sw_pattern);
  for (int j1 = 0; j1 < m; ++j1) for (int j2 = 0; j2 < m; ++j2) V[j1][j2] = c[7] * A6[j1][j2] + c[5] * A4[j1][j2] + c[3] * A2[j1][j2];
  for (int j3 = 0; j3 < m; ++j3) V[j3][j3] = V[j3][j3] + c[1];
  mat_x_mat(m, A, V, U, sw_pattern);
  for (int j1 = 0; j1 < m; ++j1) for (int j2 = 0; j2 < m; ++j2) V[j1][j2] = c[6] * A6[j1][j2] + c[4] * A4[j1][j2] + c[2] * A2[j1][j2];
  for (int j3 = 0; j3 < m; ++j3) V[j3][j3] = V[j3][j3] + c[0];
  for (int j1 = 0; j1 < m; ++j1) for (int j2 = 0; j2 < m; ++j2) { V[j1][j2] = V[j1][j2] - U[j1][j2]; U[j1][j2] = 2.0 * U[j1][j2]; }
  if (m == 3) {   /* solve_mat dispatches on m (:787-793) */
    m3 a, b, x;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { a[i][j] = V[i][j]; b[i][j] = U[i][j]; }
    solve_mat_3(a, b, x);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i][j] = x[i][j];
  } else {
    solve_mat_n(m, V, U, A);
  }
  for (int j3 = 0; j3 < m; ++j3) A[j3][j3] = A[j3][j3] + 1.0;
  if (expo > 0) repeated_square(m, A, expo, sw_pattern);
}
/* test hook: expm of one m x m matrix given in C row-major order */
void orc_expm(int m, double* a, int sw_pattern) {
  mat A;
  for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) A[i][j] = a[i * m + j];
  expm(m, A, sw_pattern);
  for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) a[i * m + j] = A[i][j];
}

/* diag_mat_right_divide_3 :567-632: X = diag(B) A^-1 */
static void diag_mat_right_divide_3(m3 A, const double* B, m3 X) {
  const double L21 = A[0][1] / A[0][0], L31 = A[0][2] / A[0][0];
  const double U22 = A[1][1] - L21 * A[1][0], U23 = A[2][1] - L21 * A[2][0];
  const double L32 = (A[1][2] - L31 * A[1][0]) / U22;
  const double U33 = A[2][2] - L31 * A[2][0] - L32 * U23;
  double y2 = -L21 * B[0], y3 = -L31 * B[0] - L32 * y2;
  X[0][2] = y3 / U33;
  X[0][1] = (y2 - U23 * X[0][2]) / U22;
  X[0][0] = (B[0] - A[1][0] * X[0][1] - A[2][0] * X[0][2]) / A[0][0];
  y3 = -L32 * B[1];
  X[1][2] = y3 / U33;
  X[1][1] = (B[1] - U23 * X[1][2]) / U22;
  X[1][0] = (-A[1][0] * X[1][1] - A[2][0] * X[1][2]) / A[0][0];
  X[2][2] = B[2] / U33;
  X[2][1] = -U23 * X[2][2] / U22;
  X[2][0] = (-A[1][0] * X[2][1] - A[2][0] * X[2][2]) / A[0][0];
}
/* fast_expm_exchange_3 :952-1028: exp of [[-a, b, 0], [a, -b-c, d], [0, c, -d]] */
static void fast_expm_exchange_3(double a, double b, double c, double d, m3 R) {
  const double my_epsilon = 1.0e-12;
  const double tmp1 = 0.5 * (a + b + c + d);
  double tmp2 = sqrt(dmax(0.0, tmp1 * tmp1 - (a * c + a * d + b * d)));
  tmp2 = dmax(tmp2, DBL_EPSILON * tmp1);
  const double lambda1 = -tmp1 + tmp2, lambda2 = -tmp1 - tmp2;
  m3 V, DV;
  V[0][0] = dmax(my_epsilon, b) / copysign(dmax(my_epsilon, fabs(a + lambda1)), a + lambda1);
  V[0][1] = b / copysign(dmax(my_epsilon, fabs(a + lambda2)), a + lambda2);
  V[0][2] = b / dmax(my_epsilon, a);
  V[1][0] = 1.0; V[1][1] = 1.0; V[1][2] = 1.0;
  V[2][0] = c / copysign(dmax(my_epsilon, fabs(d + lambda1)), d + lambda1);
  V[2][1] = c / copysign(dmax(my_epsilon, fabs(d + lambda2)), d + lambda2);
  V[2][2] = dmax(my_epsilon, c) / dmax(my_epsilon, d);
  const double diag[3] = {exp(lambda1), exp(lambda2), 1.0};
  diag_mat_right_divide_3(V, diag, DV);
  for (int j1 = 0; j1 < 3; ++j1)
    for (int j2 = 0; j2 < 3; ++j2) R[j2][j1] = V[j2][0] * DV[0][j1] + V[j2][1] * DV[1][j1] + V[j2][2] * DV[2][j1];
}
void orc_fast_expm_exchange_3(double a, double b, double c, double d, double* r) {
  m3 R;
  fast_expm_exchange_3(a, b, c, d, R);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r[i * 3 + j] = R[i][j];
}

/* 3x3 helpers with the reference's accumulation order (acc = 0; acc = acc + a*b for j3 = 1..3) */
static void m3_x_m3(m3 A, m3 B, m3 C) {
  m3 R;
  for (int j2 = 0; j2 < 3; ++j2)
    for (int j1 = 0; j1 < 3; ++j1) { double acc = 0.0; for (int j3 = 0; j3 < 3; ++j3) acc = acc + A[j1][j3] * B[j3][j2]; R[j1][j2] = acc; }
  memcpy(C, R, sizeof(m3));
}
static void m3_x_vec(m3 A, const double* b, double* x) {   /* mat_x_vec / singlemat_x_vec */
  double r[3];
  for (int j1 = 0; j1 < 3; ++j1) { double acc = 0.0; for (int j2 = 0; j2 < 3; ++j2) acc = acc + A[j1][j2] * b[j2]; r[j1] = acc; }
  x[0] = r[0]; x[1] = r[1]; x[2] = r[2];
}
static void identity_minus_m3_x_m3(m3 A, m3 B, m3 C) {   /* :286-315 */
  m3_x_m3(A, B, C);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C[i][j] = -C[i][j];
  for (int j = 0; j < 3; ++j) C[j][j] = 1.0 + C[j][j];
}
static double clamp(double x, double lo, double hi) { return dmin(hi, dmax(lo, x)); }

/* calc_two_stream_gammas_sw radiation_two_stream.F90:96-140 */
static void gammas_sw(double mu0, double ssa, double g, double* g1, double* g2, double* g3) {
  const double factor = 0.75 * g;
  *g1 = 2.0 - ssa * (1.25 + factor);
  *g2 = ssa * (0.75 - factor);
  *g3 = 0.5 - mu0 * factor;
}
/* calc_two_stream_gammas_lw :51-90 */
static void gammas_lw(double ssa, double g, double* g1, double* g2) {
  const double factor = (LwDiffusivity * 0.5) * ssa;
  *g1 = LwDiffusivity - factor * (1.0 + g);
  *g2 = factor * (1.0 - g);
}
/* calc_reflectance_transmittance_sw :421-556 for one g-point */
static void ref_trans_sw(double mu0, double od, double ssa, double gamma1, double gamma2, double gamma3, double* ref_diff,
                         double* trans_diff, double* ref_dir, double* trans_dir_diff, double* trans_dir_dir) {
  const double gamma4 = 1.0 - gamma3;
  const double alpha1 = gamma1 * gamma4 + gamma2 * gamma3, alpha2 = gamma1 * gamma3 + gamma2 * gamma4;
  const double k_exponent = sqrt(dmax((gamma1 - gamma2) * (gamma1 + gamma2), 1.0e-12));
  double mu0_local = mu0;
  if (fabs(1.0 - k_exponent * mu0) < 1000.0 * DBL_EPSILON) mu0_local = mu0 * (1.0 - 10.0 * DBL_EPSILON);
  const double od_over_mu0 = dmax(od / mu0_local, 0.0);
  const double k_mu0 = k_exponent * mu0_local, k_gamma3 = k_exponent * gamma3, k_gamma4 = k_exponent * gamma4;
  const double exponential0 = exp(-od_over_mu0);
  *trans_dir_dir = exponential0;
  const double exponential = exp(-k_exponent * od);
  const double exponential2 = exponential * exponential, k_2_exponential = 2.0 * k_exponent * exponential;
  double reftrans_factor = 1.0 / (k_exponent + gamma1 + (k_exponent - gamma1) * exponential2);
  *ref_diff = gamma2 * (1.0 - exponential2) * reftrans_factor;
  *trans_diff = k_2_exponential * reftrans_factor;
  reftrans_factor = mu0_local * ssa * reftrans_factor / (1.0 - k_mu0 * k_mu0);
  double rd = reftrans_factor * ((1.0 - k_mu0) * (alpha2 + k_gamma3) - (1.0 + k_mu0) * (alpha2 - k_gamma3) * exponential2 -
                                 k_2_exponential * (gamma3 - alpha2 * mu0_local) * exponential0);
  double td = reftrans_factor * (k_2_exponential * (gamma4 + alpha1 * mu0_local) -
                                 exponential0 * ((1.0 + k_mu0) * (alpha1 + k_gamma4) - (1.0 - k_mu0) * (alpha1 - k_gamma4) * exponential2));
  rd = dmax(0.0, dmin(rd, 1.0));
  td = dmax(0.0, dmin(td, 1.0 - rd));
  *ref_dir = rd; *trans_dir_diff = td;
}
/* calc_reflectance_transmittance_lw :148-235 for one g-point */
static void ref_trans_lw(double od, double gamma1, double gamma2, double planck_top, double planck_bot, double* reflectance,
                         double* transmittance, double* source_up, double* source_dn) {
  const double k_exponent = sqrt(dmax((gamma1 - gamma2) * (gamma1 + gamma2), 1.0e-12));
  if (od > 1.0e-3) {
    const double exponential = exp(-k_exponent * od), exponential2 = exponential * exponential;
    const double reftrans_factor = 1.0 / (k_exponent + gamma1 + (k_exponent - gamma1) * exponential2);
    *reflectance = gamma2 * (1.0 - exponential2) * reftrans_factor;
    *transmittance = 2.0 * k_exponent * exponential * reftrans_factor;
    const double coeff = (planck_bot - planck_top) / (od * (gamma1 + gamma2));
    const double coeff_up_top = coeff + planck_top, coeff_up_bot = coeff + planck_bot;
    const double coeff_dn_top = -coeff + planck_top, coeff_dn_bot = -coeff + planck_bot;
    *source_up = coeff_up_top - *reflectance * coeff_dn_top - *transmittance * coeff_up_bot;
    *source_dn = coeff_dn_bot - *reflectance * coeff_up_bot - *transmittance * coeff_dn_top;
  } else {
    *reflectance = gamma2 * od;
    *transmittance = (1.0 - k_exponent * od) / (1.0 + od * (gamma1 - k_exponent));
    *source_up = (1.0 - *reflectance - *transmittance) * 0.5 * (planck_top + planck_bot);
    *source_dn = *source_up;
  }
}

/* config%nregions = 2 (radiation_regions.F90:105-110): clear sky + one homogeneous cloudy region, od_scaling = 1.  Restated as three
 * regions with an empty third one: the three-region formulae of the overlap matrices (radiation_overlap.F90:169-209), the edge
 * lengths and the exchange terms then reduce to the two-region ones; the third region exchanges nothing and carries no flux.
 * (Not a separate two-region code path: checked against the three-region solver with two identical cloudy regions, fractional_std = 0,
 * tests/test_oracle_spartacus.py.) */
static void two_region_properties(int nlev, const double* frac, double (*reg_fracs)[NREG], double (*od_scaling)[NREG]) {
  for (int jl = 0; jl < nlev; ++jl) {
    reg_fracs[jl][1] = frac[jl]; reg_fracs[jl][0] = 1.0 - reg_fracs[jl][1]; reg_fracs[jl][2] = 0.0;
    od_scaling[jl][0] = 0.0; od_scaling[jl][1] = 1.0; od_scaling[jl][2] = 1.0;
  }
}

/* lateral transfer rates of one cloudy layer (radiation_spartacus_sw.F90:495-604 / _lw.F90:415-520).  tan_angle: tan_sza for the
 * direct beam, tan_diffuse_angle_3d for diffuse radiation.  Returns 1 if 3D effects are represented in this layer. */
static int edge_lengths(const ecrad_b200_config* cfg, double frac, const double* reg, const double* inv_cloud_size,
                        const double* inv_inhom_size, int l, double* edge_length) {
  edge_length[0] = 0.0; edge_length[1] = 0.0; edge_length[2] = 0.0;
  if (!(cfg->do_3d_effects && inv_cloud_size)) return 0;
  if (cfg->n_regions == 2 && frac > 1.0 - cfg->cloud_fraction_threshold) return 0;   /* radiation_spartacus_sw.F90:497-500, _lw.F90:423-426 */
  if (!(inv_cloud_size[l] > 0.0)) return 0;
  const double four_over_pi = 4.0 / Pi;
  edge_length[0] = four_over_pi * reg[0] * (1.0 - reg[0]) * dmin(inv_cloud_size[l], 1.0 / cfg->min_cloud_effective_size);
  if (inv_inhom_size)
    edge_length[1] = four_over_pi * reg[2] * (1.0 - reg[2]) * dmin(inv_inhom_size[l], 1.0 / cfg->min_cloud_effective_size);
  else
    edge_length[1] = four_over_pi * reg[2] * (1.0 - reg[2]) * dmin(inv_cloud_size[l], 1.0 / cfg->min_cloud_effective_size);
  if (cfg->clear_to_thick_fraction > 0.0) {
    edge_length[2] = cfg->clear_to_thick_fraction * dmin(edge_length[0], edge_length[1]);
    edge_length[0] = edge_length[0] - edge_length[2];
    edge_length[1] = edge_length[1] - edge_length[2];
  } else {
    edge_length[2] = 0.0;
  }
  return 1;
}
static void transfer_rates(const ecrad_b200_config* cfg, double dz, const double* edge_length, const double* reg, double tan_angle, m3 rate) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) rate[i][j] = 0.0;
  for (int jreg = 0; jreg < NREG - 1; ++jreg) {
    if (reg[jreg] > DBL_EPSILON) rate[jreg][jreg + 1] = dz * edge_length[jreg] * tan_angle / reg[jreg];
    if (reg[jreg + 1] > DBL_EPSILON) rate[jreg + 1][jreg] = dz * edge_length[jreg] * tan_angle / reg[jreg + 1];
  }
  if (edge_length[2] > 0.0) {
    if (reg[0] > DBL_EPSILON) rate[0][2] = dz * edge_length[2] * tan_angle / reg[0];
    if (reg[2] > DBL_EPSILON) rate[2][0] = dz * edge_length[2] * tan_angle / reg[2];
  }
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) if (rate[i][j] > cfg->max_3d_transfer_rate) rate[i][j] = cfg->max_3d_transfer_rate;
}

/* step_migrations radiation_spartacus_sw.F90:1606-1721 for one g-point */
static void step_migrations(double cloud_frac, double layer_depth, double tan_diffuse_angle_3d, double tan_sza, m3 reflectance,
                            m3 transmittance, m3 ref_dir, m3 trans_dir_dir, m3 trans_dir_diff,
                            m3 total_albedo_diff, m3 total_albedo_dir, double* x_diffuse, double* x_direct) {
  int istartreg = 0, iendreg = NREG;   /* [istartreg, iendreg) */
  if (cloud_frac <= 0.0) iendreg = 1;
  else if (cloud_frac >= 1.0) istartreg = 1;
  const double x_layer_diffuse = layer_depth * tan_diffuse_angle_3d / sqrt(2.0);
  const double x_layer_direct = layer_depth * sqrt(tan_sza * tan_sza + tan_diffuse_angle_3d * tan_diffuse_angle_3d) * 0.5;
  for (int r = istartreg; r < iendreg; ++r) {
    const double ms_enhancement = transmittance[r][r] / (1.0 - reflectance[r][r] * total_albedo_diff[r][r]);
    const double x_enhancement = pow(1.0 - reflectance[r][r] * total_albedo_diff[r][r], -1.5);
    double top_albedo = dmax(1.0e-8, ref_dir[r][r] + ms_enhancement * (trans_dir_diff[r][r] * total_albedo_diff[r][r] +
                                                                     trans_dir_dir[r][r] * total_albedo_dir[r][r]));
    x_direct[r] = dmax(0.0, x_layer_direct +
                                ((trans_dir_diff[r][r] * total_albedo_diff[r][r] * x_enhancement +
                                  trans_dir_dir[r][r] * total_albedo_dir[r][r] * (x_enhancement - 1.0)) * (x_diffuse[r] + x_layer_diffuse) +
                                 trans_dir_dir[r][r] * total_albedo_dir[r][r] * (x_direct[r] + x_layer_direct)) *
                                    transmittance[r][r] / top_albedo);
    top_albedo = dmax(1.0e-8, reflectance[r][r] + ms_enhancement * transmittance[r][r] * total_albedo_diff[r][r]);
    x_diffuse[r] = x_layer_diffuse + x_enhancement * total_albedo_diff[r][r] * (transmittance[r][r] * transmittance[r][r]) *
                                         (x_diffuse[r] + x_layer_diffuse) / top_albedo;
  }
  if (iendreg < NREG) { for (int r = iendreg; r < NREG; ++r) { x_diffuse[r] = 0.0; x_direct[r] = 0.0; } }
  else if (istartreg == 1) { x_diffuse[0] = 0.0; x_direct[0] = 0.0; }
}

/* U (singlemat) x A x V (singlemat): singlemat_x_mat(u, mat_x_singlemat(A, v)) :218-280 */
static void u_x_mat_x_v(double U[NREG][NREG], m3 A, double V[NREG][NREG], m3 out) {
  m3 T;
  m3_x_m3(A, V, T);
  m3_x_m3(U, T, out);
}

/* the explicit-entrapment increment of one lower region jreg2 for the diffuse or the direct albedo
 * (radiation_spartacus_sw.F90:1124-1180 / :1213-1263) */
static void entrapment_part(const ecrad_b200_config* cfg, m3 rate, const double x, double inv_effective_size, m3 part) {
  m3 e;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) e[i][j] = 0.0;
  for (int jreg = 0; jreg < NREG - 1; ++jreg) {
    if (cfg->i_3d_sw_entrapment == ECRAD_ENTRAPMENT_EXPLICIT) {
      const double fractal_factor = 1.0 / sqrt(dmax(1.0, 2.5 * x * inv_effective_size));
      e[jreg + 1][jreg] = e[jreg + 1][jreg] + rate[jreg][jreg + 1] * x * fractal_factor;
      e[jreg][jreg + 1] = e[jreg][jreg + 1] + rate[jreg + 1][jreg] * x * fractal_factor;
    } else {
      e[jreg + 1][jreg] = e[jreg + 1][jreg] + rate[jreg][jreg + 1] * x;
      e[jreg][jreg + 1] = e[jreg][jreg + 1] + rate[jreg + 1][jreg] * x;
    }
    e[jreg][jreg] = e[jreg][jreg] - e[jreg + 1][jreg];
    e[jreg + 1][jreg + 1] = e[jreg + 1][jreg + 1] - e[jreg][jreg + 1];
  }
  const double max_entr = -dmin(e[0][0], e[1][1]);
  if (max_entr > cfg->max_cloud_od) {
    const double s = cfg->max_cloud_od / max_entr;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) e[i][j] = e[i][j] * s;
  }
  if (cfg->n_regions == 2) {   /* fast_expm_exchange_2, radiation_matrix.F90:905-925 (radiation_spartacus_sw.F90:1184-1186) */
    const double a = e[1][0], b = e[0][1];
    const double factor = (1.0 - exp(-(a + b))) / dmax(1.0e-12, a + b);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) part[i][j] = 0.0;
    part[0][0] = 1.0 - factor * a; part[1][0] = factor * a; part[0][1] = factor * b; part[1][1] = 1.0 - factor * b; part[2][2] = 1.0;
  } else fast_expm_exchange_3(e[1][0], e[0][1], e[2][1], e[1][2], part);
}

#define L3(p, l, g) ((p) + ((size_t)(l) * ng + (g)) * 9)       /* [lev][g][3][3] */
#define AS_M3(p) ((double (*)[NREG])(p))

/* =========================================================================================================
 * solver_spartacus_sw for one sunlit column (mu0 >= 1e-10); inputs [lev][g] / [lev][band], outputs as tripleclouds.c
 * ========================================================================================================= */
void orc_spartacus_sw(const orc_tables* t, const ecrad_b200_config* cfg, int nlev, double mu0, const double* p_hl, const double* t_hl,
                      const double* frac, const double* fsd, const double* overlap_param, const double* inv_cloud_size,
                      const double* inv_inhom_size, const double* od, const double* ssa, const double* g, const double* od_cloud,
                      const double* ssa_cloud, const double* g_cloud, const double* incoming, const double* alb_diff,
                      const double* alb_dir, orc_tc_out* o) {
  const int ng = NG_SW, nreg = NREG;
  const double R_over_g = GasConstantDryAir / AccelDueToGravity;
  const double tan_diffuse_angle_3d = Pi * 0.5, min_mu0_3d = 0.004625;
  double (*reg)[NREG] = malloc(sizeof(double[NREG]) * nlev), (*ods)[NREG] = malloc(sizeof(double[NREG]) * nlev);
  double (*U)[NREG][NREG] = malloc(sizeof(double[NREG][NREG]) * (nlev + 1)), (*V)[NREG][NREG] = malloc(sizeof(double[NREG][NREG]) * (nlev + 1));
  orc_region_properties(nlev, frac, fsd, cfg->cloud_fraction_threshold, cfg->i_cloud_pdf_shape == ECRAD_PDF_LOGNORMAL, reg, ods);
  if (cfg->n_regions == 2) two_region_properties(nlev, frac, reg, ods);
  orc_overlap_matrices(nlev, reg, overlap_param, cfg->cloud_inhom_decorr_scaling, cfg->cloud_fraction_threshold, cfg->use_beta_overlap, U, V, &o->cloud_cover);

  const double one_over_mu0 = 1.0 / mu0;
  double tan_sza;
  if (mu0 < min_mu0_3d) tan_sza = sqrt(1.0 / (min_mu0_3d * min_mu0_3d) - 1.0);
  else if (one_over_mu0 > 1.0) tan_sza = sqrt(one_over_mu0 * one_over_mu0 - 1.0 + cfg->overhead_sun_factor);
  else tan_sza = sqrt(cfg->overhead_sun_factor);

  int* clear = calloc(nlev + 2, sizeof(int));   /* is_clear_sky_layer(0:nlev+1) */
  for (int i = 0; i < nlev + 2; ++i) clear[i] = 1;
  int i_cloud_top = nlev + 1;
  for (int jl = nlev; jl >= 1; --jl) if (frac[jl - 1] > 0.0) { clear[jl] = 0; i_cloud_top = jl; }

  const size_t nl = (size_t)nlev * ng, nl1 = (size_t)(nlev + 1) * ng;
  double* buf = calloc(5 * nl * 9 + 5 * nl + 2 * nl1 * 9 + 2 * nl1 + (size_t)nlev * 4, sizeof(double));
  double *reflectance = buf, *transmittance = reflectance + nl * 9, *ref_dir = transmittance + nl * 9, *trans_dir_diff = ref_dir + nl * 9,
         *trans_dir_dir = trans_dir_diff + nl * 9;
  double *ref_clear = trans_dir_dir + nl * 9, *trans_clear = ref_clear + nl, *ref_dir_clear = trans_clear + nl,
         *trans_dir_diff_clear = ref_dir_clear + nl, *trans_dir_dir_clear = trans_dir_diff_clear + nl;
  double *total_albedo = trans_dir_dir_clear + nl, *total_albedo_direct = total_albedo + nl1 * 9;
  double *total_albedo_clear = total_albedo_direct + nl1 * 9, *total_albedo_clear_direct = total_albedo_clear + nl1;
  double* layer_depth = total_albedo_clear_direct + nl1;
  double (*edge_length)[3] = (double (*)[3])(layer_depth + nlev);
  double (*od_region)[NREG] = malloc(sizeof(double[NREG]) * ng), (*ssa_region)[NREG] = malloc(sizeof(double[NREG]) * ng);
  double (*gamma1)[NREG] = malloc(sizeof(double[NREG]) * ng), (*gamma2)[NREG] = malloc(sizeof(double[NREG]) * ng),
         (*gamma3)[NREG] = malloc(sizeof(double[NREG]) * ng);

  /* ---- Section 3: reflectance, transmittance and sources of each layer (:420-835) ---- */
  for (int jlev = 1; jlev <= nlev; ++jlev) {
    const int l = jlev - 1;
    m3 transfer_rate_direct, transfer_rate_diffuse;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { transfer_rate_direct[i][j] = 0.0; transfer_rate_diffuse[i][j] = 0.0; }
    edge_length[l][0] = 0.0; edge_length[l][1] = 0.0; edge_length[l][2] = 0.0;
    layer_depth[l] = R_over_g * (p_hl[l + 1] - p_hl[l]) * (t_hl[l] + t_hl[l + 1]) / (p_hl[l] + p_hl[l + 1]);
    int nregactive, ng3D;
    for (int jg = 0; jg < ng; ++jg)
      for (int r = 0; r < NREG; ++r) { od_region[jg][r] = 0.0; ssa_region[jg][r] = 0.0; gamma1[jg][r] = 0.0; gamma2[jg][r] = 0.0; gamma3[jg][r] = 0.0; }
    if (clear[jlev]) {
      nregactive = 1;
      for (int jg = 0; jg < ng; ++jg) {
        od_region[jg][0] = od[(size_t)l * ng + jg];
        ssa_region[jg][0] = ssa[(size_t)l * ng + jg];
        gammas_sw(mu0, ssa[(size_t)l * ng + jg], g[(size_t)l * ng + jg], &gamma1[jg][0], &gamma2[jg][0], &gamma3[jg][0]);
      }
      if (cfg->use_expm_everywhere) {
        ng3D = ng;
        for (int jg = 0; jg < ng; ++jg) if (od_region[jg][0] > cfg->max_gas_od_3d) { ng3D = jg; break; }
      } else {
        ng3D = 0;
      }
    } else {
      ng3D = cfg->use_expm_everywhere ? ng : 0;
      if (edge_lengths(cfg, frac[l], reg[l], inv_cloud_size, inv_inhom_size, l, edge_length[l])) {
        ng3D = ng;
        const double dz = layer_depth[l];
        transfer_rates(cfg, dz, edge_length[l], reg[l], tan_sza, transfer_rate_direct);
        transfer_rates(cfg, dz, edge_length[l], reg[l], tan_diffuse_angle_3d, transfer_rate_diffuse);
      }
      nregactive = nreg;
      for (int jg = 0; jg < ng; ++jg) {
        const size_t i = (size_t)l * ng + jg;
        const int iband = t->band_sw[jg];
        const double scat_od = od[i] * ssa[i];
        double g_region[NREG];
        od_region[jg][0] = od[i]; ssa_region[jg][0] = ssa[i]; g_region[0] = g[i];
        for (int jreg = 1; jreg < nreg; ++jreg) {
          const double scat_od_cloud = od_cloud[l * NB_SW + iband] * ssa_cloud[l * NB_SW + iband] * ods[l][jreg];
          od_region[jg][jreg] = od[i] + od_cloud[l * NB_SW + iband] * ods[l][jreg];
          ssa_region[jg][jreg] = (scat_od + scat_od_cloud) / od_region[jg][jreg];
          g_region[jreg] = (scat_od * g[i] + scat_od_cloud * g_cloud[l * NB_SW + iband]) / (scat_od + scat_od_cloud);
          if (od_region[jg][jreg] > cfg->max_cloud_od) od_region[jg][jreg] = cfg->max_cloud_od;
        }
        for (int r = 0; r < nreg; ++r) gammas_sw(mu0, ssa_region[jg][r], g_region[r], &gamma1[jg][r], &gamma2[jg][r], &gamma3[jg][r]);
        if (ng3D == ng && od_region[jg][0] > cfg->max_gas_od_3d) ng3D = jg;
      }
    }
    /* 3.3a: g-points with 3D effects: 9x9 matrix exponential (:658-770) */
    for (int jg = 0; jg < ng3D; ++jg) {
      mat G;
      for (int i = 0; i < 9; ++i) for (int j = 0; j < 9; ++j) G[i][j] = 0.0;
      for (int jreg = 0; jreg < nregactive; ++jreg) {
        G[jreg][jreg] = od_region[jg][jreg] * gamma1[jg][jreg];
        G[jreg + nreg][jreg] = od_region[jg][jreg] * gamma2[jg][jreg];
        G[jreg][jreg + 2 * nreg] = -od_region[jg][jreg] * ssa_region[jg][jreg] * gamma3[jg][jreg];
        G[jreg + nreg][jreg + 2 * nreg] = od_region[jg][jreg] * ssa_region[jg][jreg] * (1.0 - gamma3[jg][jreg]);
        G[jreg + 2 * nreg][jreg + 2 * nreg] = -od_region[jg][jreg] * one_over_mu0;
      }
      for (int jreg = 0; jreg < nregactive - 1; ++jreg) {
        G[jreg][jreg] = G[jreg][jreg] + transfer_rate_diffuse[jreg][jreg + 1];
        G[jreg + 1][jreg + 1] = G[jreg + 1][jreg + 1] + transfer_rate_diffuse[jreg + 1][jreg];
        G[jreg + 1][jreg] = -transfer_rate_diffuse[jreg][jreg + 1];
        G[jreg][jreg + 1] = -transfer_rate_diffuse[jreg + 1][jreg];
        G[jreg + 2 * nreg][jreg + 2 * nreg] = G[jreg + 2 * nreg][jreg + 2 * nreg] - transfer_rate_direct[jreg][jreg + 1];
        G[jreg + 2 * nreg + 1][jreg + 2 * nreg + 1] = G[jreg + 2 * nreg + 1][jreg + 2 * nreg + 1] - transfer_rate_direct[jreg + 1][jreg];
        G[jreg + 2 * nreg + 1][jreg + 2 * nreg] = transfer_rate_direct[jreg][jreg + 1];
        G[jreg + 2 * nreg][jreg + 2 * nreg + 1] = transfer_rate_direct[jreg + 1][jreg];
      }
      if (edge_length[l][2] > 0.0) {
        G[0][0] = G[0][0] + transfer_rate_diffuse[0][2];
        G[2][2] = G[2][2] + transfer_rate_diffuse[2][0];
        G[2][0] = -transfer_rate_diffuse[0][2];
        G[0][2] = -transfer_rate_diffuse[2][0];
        G[2 * nreg][2 * nreg] = G[2 * nreg][2 * nreg] - transfer_rate_direct[0][2];
        G[2 + 2 * nreg][2 + 2 * nreg] = G[2 + 2 * nreg][2 + 2 * nreg] - transfer_rate_direct[2][0];
        G[2 + 2 * nreg][2 * nreg] = transfer_rate_direct[0][2];
        G[2 * nreg][2 + 2 * nreg] = transfer_rate_direct[2][0];
      }
      for (int i = 0; i < nregactive; ++i)
        for (int j = 0; j < nregactive; ++j) G[nreg + i][nreg + j] = -G[i][j];
      for (int i = 0; i < nregactive; ++i)
        for (int j = 0; j < nregactive; ++j) G[i][nreg + j] = -G[nreg + i][j];
      expm(3 * nreg, G, 1);
      m3 E11, E12, E13, E21, E22, E23, X;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          E11[i][j] = G[i][j]; E12[i][j] = G[i][3 + j]; E13[i][j] = G[i][6 + j];
          E21[i][j] = G[3 + i][j]; E22[i][j] = G[3 + i][3 + j]; E23[i][j] = G[3 + i][6 + j];
        }
      double (*R)[NREG] = AS_M3(L3(reflectance, l, jg)), (*T)[NREG] = AS_M3(L3(transmittance, l, jg));
      double (*RD)[NREG] = AS_M3(L3(ref_dir, l, jg)), (*TDD)[NREG] = AS_M3(L3(trans_dir_diff, l, jg)), (*TD)[NREG] = AS_M3(L3(trans_dir_dir, l, jg));
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) TD[i][j] = clamp(G[6 + i][6 + j], 0.0, 1.0);
      solve_mat_3(E11, E12, X);
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R[i][j] = clamp(-X[i][j], 0.0, 1.0);
      m3_x_m3(E21, R, X);
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) T[i][j] = clamp(X[i][j] + E22[i][j], 0.0, 1.0);
      solve_mat_3(E11, E13, X);
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) RD[i][j] = dmin(mu0, dmax(0.0, -X[i][j]));
      m3_x_m3(E21, RD, X);
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) TDD[i][j] = dmin(mu0, dmax(0.0, X[i][j] + E23[i][j]));
    }
    /* 3.3b: Meador-Weaver for the clear-sky arrays (all g) and for the g-points without 3D effects (:772-833) */
    for (int jg = 0; jg < ng; ++jg) {
      const size_t i = (size_t)l * ng + jg;
      ref_trans_sw(mu0, od_region[jg][0], ssa_region[jg][0], gamma1[jg][0], gamma2[jg][0], gamma3[jg][0], &ref_clear[i], &trans_clear[i],
                   &ref_dir_clear[i], &trans_dir_diff_clear[i], &trans_dir_dir_clear[i]);
    }
    for (int jg = ng3D; jg < ng; ++jg) {
      const size_t i = (size_t)l * ng + jg;
      double (*R)[NREG] = AS_M3(L3(reflectance, l, jg)), (*T)[NREG] = AS_M3(L3(transmittance, l, jg));
      double (*RD)[NREG] = AS_M3(L3(ref_dir, l, jg)), (*TDD)[NREG] = AS_M3(L3(trans_dir_diff, l, jg)), (*TD)[NREG] = AS_M3(L3(trans_dir_dir, l, jg));
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) { R[a][b] = 0.0; T[a][b] = 0.0; RD[a][b] = 0.0; TDD[a][b] = 0.0; TD[a][b] = 0.0; }
      TD[0][0] = trans_dir_dir_clear[i]; R[0][0] = ref_clear[i]; T[0][0] = trans_clear[i]; RD[0][0] = ref_dir_clear[i]; TDD[0][0] = trans_dir_diff_clear[i];
      for (int jreg = 1; jreg < nregactive; ++jreg)
        ref_trans_sw(mu0, od_region[jg][jreg], ssa_region[jg][jreg], gamma1[jg][jreg], gamma2[jg][jreg], gamma3[jg][jreg], &R[jreg][jreg],
                     &T[jreg][jreg], &RD[jreg][jreg], &TDD[jreg][jreg], &TD[jreg][jreg]);
    }
  }

  /* ---- Section 4: total albedos (:837-1322) ---- */
  double (*x_diffuse)[NREG] = calloc(ng, sizeof(double[NREG])), (*x_direct)[NREG] = calloc(ng, sizeof(double[NREG]));
  for (int jg = 0; jg < ng; ++jg) {
    double (*TA)[NREG] = AS_M3(L3(total_albedo, nlev, jg)), (*TAD)[NREG] = AS_M3(L3(total_albedo_direct, nlev, jg));
    for (int jreg = 0; jreg < nreg; ++jreg) { TA[jreg][jreg] = alb_diff[jg]; TAD[jreg][jreg] = mu0 * alb_dir[jg]; }
    total_albedo_clear[(size_t)nlev * ng + jg] = TA[0][0];
    total_albedo_clear_direct[(size_t)nlev * ng + jg] = TAD[0][0];
  }
  const int explicit_entr = cfg->i_3d_sw_entrapment == ECRAD_ENTRAPMENT_EXPLICIT_NON_FRACTAL || cfg->i_3d_sw_entrapment == ECRAD_ENTRAPMENT_EXPLICIT;
  for (int jlev = nlev; jlev >= 1; --jlev) {
    const int l = jlev - 1;
    for (int jg = 0; jg < ng; ++jg) {
      const size_t i = (size_t)l * ng + jg, ib = (size_t)jlev * ng + jg;
      /* 4.1 adding method: clear-sky column */
      {
        const double inv_denom = 1.0 / (1.0 - total_albedo_clear[ib] * ref_clear[i]);
        total_albedo_clear[i] = ref_clear[i] + trans_clear[i] * trans_clear[i] * total_albedo_clear[ib] * inv_denom;
        total_albedo_clear_direct[i] = ref_dir_clear[i] + (trans_dir_dir_clear[i] * total_albedo_clear_direct[ib] +
                                                           trans_dir_diff_clear[i] * total_albedo_clear[ib]) * trans_clear[i] * inv_denom;
      }
      double (*R)[NREG] = AS_M3(L3(reflectance, l, jg)), (*T)[NREG] = AS_M3(L3(transmittance, l, jg));
      double (*RD)[NREG] = AS_M3(L3(ref_dir, l, jg)), (*TDD)[NREG] = AS_M3(L3(trans_dir_diff, l, jg)), (*TD)[NREG] = AS_M3(L3(trans_dir_dir, l, jg));
      double (*TAb)[NREG] = AS_M3(L3(total_albedo, jlev, jg)), (*TADb)[NREG] = AS_M3(L3(total_albedo_direct, jlev, jg));
      double (*TA)[NREG] = AS_M3(L3(total_albedo, l, jg)), (*TAD)[NREG] = AS_M3(L3(total_albedo_direct, l, jg));
      m3 below, below_direct;
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) { below[a][b] = 0.0; below_direct[a][b] = 0.0; }
      if (clear[jlev]) {
        const double inv_denom = 1.0 / (1.0 - TAb[0][0] * R[0][0]);
        below[0][0] = R[0][0] + T[0][0] * T[0][0] * TAb[0][0] * inv_denom;
        below_direct[0][0] = RD[0][0] + (TD[0][0] * TADb[0][0] + TDD[0][0] * TAb[0][0]) * T[0][0] * inv_denom;
      } else {
        m3 denominator, X, Y, Z;
        identity_minus_m3_x_m3(TAb, R, denominator);
        m3_x_m3(TAb, T, X);
        solve_mat_3(denominator, X, Y);
        m3_x_m3(T, Y, Z);
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) below[a][b] = R[a][b] + Z[a][b];
        m3_x_m3(TADb, TD, X);
        m3_x_m3(TAb, TDD, Y);
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) X[a][b] = X[a][b] + Y[a][b];
        solve_mat_3(denominator, X, Y);
        m3_x_m3(T, Y, Z);
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) below_direct[a][b] = RD[a][b] + Z[a][b];
      }
      /* 4.2 overlap and entrapment */
      if (explicit_entr && jlev >= i_cloud_top)
        step_migrations(frac[l], layer_depth[l], tan_diffuse_angle_3d, tan_sza, R, T, RD, TD, TDD, TAb, TADb, x_diffuse[jg], x_direct[jg]);
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) { TA[a][b] = 0.0; TAD[a][b] = 0.0; }
      if (clear[jlev] && clear[jlev - 1]) {
        TA[0][0] = below[0][0];
        TAD[0][0] = below_direct[0][0];
      } else if (cfg->i_3d_sw_entrapment == ECRAD_ENTRAPMENT_MAXIMUM || clear[jlev - 1]) {
        u_x_mat_x_v(U[l], below, V[l], TA);
        u_x_mat_x_v(U[l], below_direct, V[l], TAD);
      } else if (cfg->i_3d_sw_entrapment == ECRAD_ENTRAPMENT_ZERO) {
        for (int jreg = 0; jreg < nreg; ++jreg)
          for (int jreg2 = 0; jreg2 < nreg; ++jreg2) {
            /* sum(total_albedo_below(:,:,jreg2),2): over the first matrix index */
            double s = 0.0, sd = 0.0;
            for (int k = 0; k < nreg; ++k) { s = s + below[k][jreg2]; sd = sd + below_direct[k][jreg2]; }
            TA[jreg][jreg] = TA[jreg][jreg] + s * V[l][jreg2][jreg];
            TAD[jreg][jreg] = TAD[jreg][jreg] + sd * V[l][jreg2][jreg];
          }
      } else {
        m3 part;
        memcpy(part, below, sizeof(m3));
        for (int jreg = 0; jreg < nreg; ++jreg) part[jreg][jreg] = 0.0;
        u_x_mat_x_v(U[l], part, V[l], TA);
        memcpy(part, below_direct, sizeof(m3));
        for (int jreg = 0; jreg < nreg; ++jreg) part[jreg][jreg] = 0.0;
        u_x_mat_x_v(U[l], part, V[l], TAD);
        if (cfg->i_3d_sw_entrapment == ECRAD_ENTRAPMENT_EDGE_ONLY || !cfg->do_3d_effects) {
          for (int jreg = 0; jreg < nreg; ++jreg)
            for (int jreg2 = 0; jreg2 < nreg; ++jreg2) {
              TA[jreg][jreg] = TA[jreg][jreg] + below[jreg2][jreg2] * V[l][jreg2][jreg];
              TAD[jreg][jreg] = TAD[jreg][jreg] + below_direct[jreg2][jreg2] * V[l][jreg2][jreg];
            }
        } else {
          for (int jreg2 = 0; jreg2 < nreg; ++jreg2) {
            m3 rate;
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) rate[a][b] = 0.0;
            if (jlev > 1) {
              const double transfer_scaling = 1.0 - (1.0 - cfg->overhang_factor) * overlap_param[jlev - 2] *
                                                        dmin(reg[l][jreg2], reg[l - 1][jreg2]) / dmax(cfg->cloud_fraction_threshold, reg[l][jreg2]);
              for (int jreg = 0; jreg < nreg - 1; ++jreg) {
                rate[jreg][jreg + 1] = transfer_scaling * edge_length[l - 1][jreg] / dmax(U[l][jreg][jreg2], 1.0e-5);
                rate[jreg + 1][jreg] = transfer_scaling * edge_length[l - 1][jreg] / dmax(U[l][jreg + 1][jreg2], 1.0e-5);
              }
              /* rates between regions 1 and 3: computed by the reference but not used by the exchange matrix */
            }
            const double inv_effective_size = dmin(inv_cloud_size[l - 1], 1.0 / cfg->min_cloud_effective_size);
            entrapment_part(cfg, rate, x_diffuse[jg][jreg2], inv_effective_size, part);
            for (int jreg3 = 0; jreg3 < nreg; ++jreg3)
              for (int jreg = 0; jreg < nreg; ++jreg) part[jreg3][jreg] = part[jreg3][jreg] * V[l][jreg2][jreg] * below[jreg2][jreg2];
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) TA[a][b] = TA[a][b] + part[a][b];
            entrapment_part(cfg, rate, x_direct[jg][jreg2], inv_effective_size, part);
            for (int jreg3 = 0; jreg3 < nreg; ++jreg3)
              for (int jreg = 0; jreg < nreg; ++jreg) part[jreg3][jreg] = part[jreg3][jreg] * V[l][jreg2][jreg] * below_direct[jreg2][jreg2];
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) TAD[a][b] = TAD[a][b] + part[a][b];
          }
        }
      }
      if (explicit_entr && !(clear[jlev] && clear[jlev - 1])) {
        double xd[NREG] = {0.0, 0.0, 0.0}, xf[NREG] = {0.0, 0.0, 0.0};
        const int nra = clear[jlev] ? 1 : nreg;
        for (int jreg = 0; jreg < nreg; ++jreg)
          for (int jreg2 = 0; jreg2 < nra; ++jreg2) {
            xd[jreg] = xd[jreg] + x_direct[jg][jreg2] * V[l][jreg2][jreg];
            xf[jreg] = xf[jreg] + x_diffuse[jg][jreg2] * V[l][jreg2][jreg];
          }
        for (int r = 0; r < nreg; ++r) { x_direct[jg][r] = xd[r]; x_diffuse[jg][r] = xf[r]; }
      }
    }
  }

  /* ---- Section 5: fluxes (:1324-1590) ---- */
  double (*flux_up)[NREG] = calloc(ng, sizeof(double[NREG])), (*flux_dn)[NREG] = calloc(ng, sizeof(double[NREG])),
         (*direct_dn)[NREG] = calloc(ng, sizeof(double[NREG]));
  double *flux_up_clear = calloc(3 * (size_t)ng, sizeof(double)), *flux_dn_clear = flux_up_clear + ng, *direct_dn_clear = flux_dn_clear + ng;
  for (int jg = 0; jg < ng; ++jg) {
    for (int jreg = 0; jreg < nreg; ++jreg) { flux_dn[jg][jreg] = 0.0; direct_dn[jg][jreg] = incoming[jg] * reg[0][jreg]; }
    m3_x_vec(AS_M3(L3(total_albedo_direct, 0, jg)), direct_dn[jg], flux_up[jg]);
    flux_dn_clear[jg] = 0.0; direct_dn_clear[jg] = incoming[jg];
    flux_up_clear[jg] = direct_dn_clear[jg] * total_albedo_clear_direct[jg];
  }
  for (int hl = 0; hl <= nlev; ++hl) {
    double sum_dir = 0.0, sum_dir_clear = 0.0;   /* direct downwelling just above the interface (before the overlap rules) */
    if (hl == 0) {
      for (int jg = 0; jg < ng; ++jg) { o->up_toa_g[jg] = flux_up[jg][0] + flux_up[jg][1] + flux_up[jg][2]; o->up_toa_clear_g[jg] = flux_up_clear[jg]; }
      double s = 0.0;
      for (int jg = 0; jg < ng; ++jg) s = s + incoming[jg];
      sum_dir = s; sum_dir_clear = s;   /* flux%sw_dn(jcol,1) = mu0 * sum(incoming_sw(:,jcol)) */
    } else {
      const int l = hl - 1, jlev = hl;
      double dir_above[NREG] = {0.0, 0.0, 0.0}, dirc = 0.0;
      for (int jg = 0; jg < ng; ++jg) {
        const size_t i = (size_t)l * ng + jg, ib = (size_t)jlev * ng + jg;
        double (*R)[NREG] = AS_M3(L3(reflectance, l, jg)), (*T)[NREG] = AS_M3(L3(transmittance, l, jg));
        double (*TDD)[NREG] = AS_M3(L3(trans_dir_diff, l, jg)), (*TD)[NREG] = AS_M3(L3(trans_dir_dir, l, jg));
        double (*TAb)[NREG] = AS_M3(L3(total_albedo, jlev, jg)), (*TADb)[NREG] = AS_M3(L3(total_albedo_direct, jlev, jg));
        double source_dn[NREG], direct_dn_above[NREG], flux_dn_above[NREG], flux_up_above[NREG];
        const double source_dn_clear = trans_dir_diff_clear[i] * direct_dn_clear[jg];
        if (clear[jlev]) { source_dn[0] = TDD[0][0] * direct_dn[jg][0]; source_dn[1] = 0.0; source_dn[2] = 0.0; }
        else m3_x_vec(TDD, direct_dn[jg], source_dn);
        direct_dn_clear[jg] = trans_dir_dir_clear[i] * direct_dn_clear[jg];
        if (clear[jlev]) { direct_dn_above[0] = TD[0][0] * direct_dn[jg][0]; direct_dn_above[1] = 0.0; direct_dn_above[2] = 0.0; }
        else m3_x_vec(TD, direct_dn[jg], direct_dn_above);
        flux_dn_clear[jg] = (trans_clear[i] * flux_dn_clear[jg] + ref_clear[i] * total_albedo_clear_direct[ib] * direct_dn_clear[jg] + source_dn_clear) /
                            (1.0 - ref_clear[i] * total_albedo_clear[ib]);
        flux_up_clear[jg] = total_albedo_clear_direct[ib] * direct_dn_clear[jg] + total_albedo_clear[ib] * flux_dn_clear[jg];
        if (clear[jlev]) {
          flux_dn_above[0] = (T[0][0] * flux_dn[jg][0] + R[0][0] * TADb[0][0] * direct_dn_above[0] + source_dn[0]) / (1.0 - R[0][0] * TAb[0][0]);
          flux_dn_above[1] = 0.0; flux_dn_above[2] = 0.0;
          flux_up_above[0] = TADb[0][0] * direct_dn_above[0] + TAb[0][0] * flux_dn_above[0];
          flux_up_above[1] = 0.0; flux_up_above[2] = 0.0;
        } else {
          m3 denominator;
          double total_source[NREG], a[NREG], b[NREG], rhs[NREG];
          identity_minus_m3_x_m3(R, TAb, denominator);
          m3_x_vec(TADb, direct_dn_above, total_source);
          m3_x_vec(T, flux_dn[jg], a);
          m3_x_vec(R, total_source, b);
          for (int r = 0; r < 3; ++r) rhs[r] = a[r] + b[r] + source_dn[r];
          solve_vec_3(denominator, rhs, flux_dn_above);
          m3_x_vec(TAb, flux_dn_above, flux_up_above);
          for (int r = 0; r < 3; ++r) flux_up_above[r] = flux_up_above[r] + total_source[r];
        }
        for (int r = 0; r < 3; ++r) { flux_up[jg][r] = flux_up_above[r]; flux_dn[jg][r] = flux_dn_above[r]; direct_dn[jg][r] = direct_dn_above[r]; }
      }
      /* sum(sum(direct_dn_above,1)): over g for each region, then over regions */
      for (int r = 0; r < 3; ++r) { double s = 0.0; for (int jg = 0; jg < ng; ++jg) s = s + direct_dn[jg][r]; dir_above[r] = s; }
      sum_dir = dir_above[0] + dir_above[1] + dir_above[2];
      for (int jg = 0; jg < ng; ++jg) dirc = dirc + direct_dn_clear[jg];
      sum_dir_clear = dirc;
    }
    /* broadband sums at this half-level (fluxes just above the interface), per-g profiles */
    {
      double su[NREG], sd[NREG], suc = 0.0, sdc = 0.0;
      for (int r = 0; r < 3; ++r) { double a = 0.0, b = 0.0; for (int jg = 0; jg < ng; ++jg) { a = a + flux_up[jg][r]; b = b + flux_dn[jg][r]; } su[r] = a; sd[r] = b; }
      for (int jg = 0; jg < ng; ++jg) { suc = suc + flux_up_clear[jg]; sdc = sdc + flux_dn_clear[jg]; }
      o->up[hl] = su[0] + su[1] + su[2];
      o->dn_direct[hl] = mu0 * sum_dir;
      o->dn[hl] = hl == 0 ? o->dn_direct[hl] : o->dn_direct[hl] + (sd[0] + sd[1] + sd[2]);
      o->up_clear[hl] = suc;
      o->dn_direct_clear[hl] = mu0 * sum_dir_clear;
      o->dn_clear[hl] = hl == 0 ? o->dn_direct_clear[hl] : o->dn_direct_clear[hl] + sdc;
      if (o->up_g_prof)
        for (int jg = 0; jg < ng; ++jg) {
          o->up_g_prof[(size_t)hl * ng + jg] = flux_up[jg][0] + flux_up[jg][1] + flux_up[jg][2];
          o->dn_dir_g_prof[(size_t)hl * ng + jg] = direct_dn[jg][0] + direct_dn[jg][1] + direct_dn[jg][2];
          o->dn_dif_g_prof[(size_t)hl * ng + jg] = flux_dn[jg][0] + flux_dn[jg][1] + flux_dn[jg][2];
        }
    }
    if (hl == nlev) {
      for (int jg = 0; jg < ng; ++jg) {
        o->dn_diffuse_surf_g[jg] = flux_dn[jg][0] + flux_dn[jg][1] + flux_dn[jg][2];
        o->dn_direct_surf_g[jg] = mu0 * (direct_dn[jg][0] + direct_dn[jg][1] + direct_dn[jg][2]);
        o->dn_diffuse_surf_clear_g[jg] = flux_dn_clear[jg];
        o->dn_direct_surf_clear_g[jg] = mu0 * direct_dn_clear[jg];
      }
    } else if (hl > 0) {
      /* overlap rules: fluxes just above the interface -> just below (:1516-1530) */
      const int jlev = hl;
      if (!(clear[jlev] && clear[jlev + 1]))
        for (int jg = 0; jg < ng; ++jg) { m3_x_vec(V[jlev], flux_dn[jg], flux_dn[jg]); m3_x_vec(V[jlev], direct_dn[jg], direct_dn[jg]); }
    }
  }
  free(flux_up); free(flux_dn); free(direct_dn); free(flux_up_clear);
  free(x_diffuse); free(x_direct);
  free(od_region); free(ssa_region); free(gamma1); free(gamma2); free(gamma3);
  free(buf); free(clear); free(reg); free(ods); free(U); free(V);
}

#define L6(p, l, g) ((p) + ((size_t)(l) * ng + (g)) * 3)       /* [lev][g][3] */

/* =========================================================================================================
 * solver_spartacus_lw for one column; ssa, g: gas + aerosol scattering properties [nlev][ng] with do_lw_aerosol_scattering
 * (radiation_spartacus_lw.F90:366-371), NULL otherwise (clear-sky ssa = g = 0)
 * ========================================================================================================= */
void orc_spartacus_lw(const orc_tables* t, const ecrad_b200_config* cfg, int nlev, const double* p_hl, const double* t_hl,
                      const double* frac, const double* fsd, const double* overlap_param, const double* inv_cloud_size,
                      const double* inv_inhom_size, const double* od, const double* ssa, const double* g, const double* planck_hl,
                      const double* od_cloud, const double* ssa_cloud, const double* g_cloud, const double* emission, const double* albedo,
                      orc_tc_out* o) {
  const int ng = NG_LW, nreg = NREG;
  const double R_over_g = GasConstantDryAir / AccelDueToGravity;
  const double tan_diffuse_angle_3d = Pi * 0.5, side_emiss_thin = 1.4107;
  double (*reg)[NREG] = malloc(sizeof(double[NREG]) * nlev), (*ods)[NREG] = malloc(sizeof(double[NREG]) * nlev);
  double (*U)[NREG][NREG] = malloc(sizeof(double[NREG][NREG]) * (nlev + 1)), (*V)[NREG][NREG] = malloc(sizeof(double[NREG][NREG]) * (nlev + 1));
  orc_region_properties(nlev, frac, fsd, cfg->cloud_fraction_threshold, cfg->i_cloud_pdf_shape == ECRAD_PDF_LOGNORMAL, reg, ods);
  if (cfg->n_regions == 2) two_region_properties(nlev, frac, reg, ods);
  orc_overlap_matrices(nlev, reg, overlap_param, cfg->cloud_inhom_decorr_scaling, cfg->cloud_fraction_threshold, cfg->use_beta_overlap, U, V, &o->cloud_cover);
  int* clear = calloc(nlev + 2, sizeof(int));
  for (int i = 0; i < nlev + 2; ++i) clear[i] = 1;
  for (int jl = 1; jl <= nlev; ++jl) if (frac[jl - 1] > 0.0) clear[jl] = 0;

  const size_t nl = (size_t)nlev * ng, nl1 = (size_t)(nlev + 1) * ng;
  double* buf = calloc(2 * nl * 9 + 2 * nl * 3 + 4 * nl + nl1 * 9 + nl1 * 3 + 2 * nl1, sizeof(double));
  double *reflectance = buf, *transmittance = reflectance + nl * 9, *source_up = transmittance + nl * 9, *source_dn = source_up + nl * 3;
  double *ref_clear = source_dn + nl * 3, *trans_clear = ref_clear + nl, *source_up_clear = trans_clear + nl, *source_dn_clear = source_up_clear + nl;
  double *total_albedo = source_dn_clear + nl, *total_source = total_albedo + nl1 * 9;
  double *total_albedo_clear = total_source + nl1 * 3, *total_source_clear = total_albedo_clear + nl1;
  double (*od_region)[NREG] = malloc(sizeof(double[NREG]) * ng), (*ssa_region)[NREG] = malloc(sizeof(double[NREG]) * ng),
         (*g_region)[NREG] = malloc(sizeof(double[NREG]) * ng);
  double (*gamma1)[NREG] = malloc(sizeof(double[NREG]) * ng), (*gamma2)[NREG] = malloc(sizeof(double[NREG]) * ng);
  double dz = 1.0;

  /* ---- Section 3 (:349-780) ---- */
  for (int jlev = 1; jlev <= nlev; ++jlev) {
    const int l = jlev - 1;
    m3 transfer_rate;
    double edge_length[3] = {0.0, 0.0, 0.0};   /* NB the reference keeps edge_length across layers (a scalar work array); it is only
                                                  read in layers where it has just been set, or where transfer_rate is zero anyway */
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) transfer_rate[a][b] = 0.0;
    int nregactive, ng3D;
    for (int jg = 0; jg < ng; ++jg)
      for (int r = 0; r < NREG; ++r) { od_region[jg][r] = 0.0; ssa_region[jg][r] = 0.0; g_region[jg][r] = 0.0; gamma1[jg][r] = 0.0; gamma2[jg][r] = 0.0; }
    for (int jg = 0; jg < ng; ++jg) od_region[jg][0] = od[(size_t)l * ng + jg];
    if (cfg->do_lw_aerosol_scattering && ssa && g)
      for (int jg = 0; jg < ng; ++jg) { ssa_region[jg][0] = ssa[(size_t)l * ng + jg]; g_region[jg][0] = g[(size_t)l * ng + jg]; }
    int did_3d = 0;
    if (clear[jlev]) {
      nregactive = 1;
      for (int jg = 0; jg < ng; ++jg) gammas_lw(ssa_region[jg][0], g_region[jg][0], &gamma1[jg][0], &gamma2[jg][0]);
      if (cfg->use_expm_everywhere) {
        ng3D = ng;
        for (int jg = 0; jg < ng; ++jg) if (od_region[jg][0] > cfg->max_gas_od_3d) { ng3D = jg; break; }
      } else ng3D = 0;
    } else {
      ng3D = cfg->use_expm_everywhere ? ng : 0;
      if (edge_lengths(cfg, frac[l], reg[l], inv_cloud_size, inv_inhom_size, l, edge_length)) {
        ng3D = ng; did_3d = 1;
        dz = R_over_g * (p_hl[l + 1] - p_hl[l]) * (t_hl[l] + t_hl[l + 1]) / (p_hl[l] + p_hl[l + 1]);
        transfer_rates(cfg, dz, edge_length, reg[l], tan_diffuse_angle_3d, transfer_rate);
      }
      nregactive = nreg;
      for (int jg = 0; jg < ng; ++jg) {
        const int iband = t->band_lw[jg];
        const double scat_od = od_region[jg][0] * ssa_region[jg][0];
        for (int jreg = 1; jreg < nreg; ++jreg) {
          od_region[jg][jreg] = od_region[jg][0] + od_cloud[l * NB_LW + iband] * ods[l][jreg];
          if (cfg->do_lw_cloud_scattering) {
            const double scat_od_cloud = od_cloud[l * NB_LW + iband] * ssa_cloud[l * NB_LW + iband] * ods[l][jreg];
            ssa_region[jg][jreg] = (scat_od + scat_od_cloud) / od_region[jg][jreg];
            if (scat_od + scat_od_cloud > 0.0)
              g_region[jg][jreg] = (scat_od * g_region[jg][0] + scat_od_cloud * g_cloud[l * NB_LW + iband]) / (scat_od + scat_od_cloud);
          }
          if (od_region[jg][jreg] > cfg->max_cloud_od) od_region[jg][jreg] = cfg->max_cloud_od;
        }
        for (int r = 0; r < nreg; ++r) gammas_lw(ssa_region[jg][r], g_region[jg][r], &gamma1[jg][r], &gamma2[jg][r]);
        if (ng3D == ng && od_region[jg][0] > cfg->max_gas_od_3d) ng3D = jg;
      }
    }
    (void)did_3d;
    /* 3.3a: 6x6 matrix exponential (:596-727) */
    for (int jg = 0; jg < ng3D; ++jg) {
      mat G;
      double planck_top[6], planck_diff[6], solution0[6], solution_diff[6], rhs[6];
      for (int i = 0; i < 6; ++i) { for (int j = 0; j < 6; ++j) G[i][j] = 0.0; planck_top[i] = 0.0; planck_diff[i] = 0.0; }
      const double pt = planck_hl[(size_t)l * ng + jg], pb = planck_hl[(size_t)(l + 1) * ng + jg];
      for (int jreg = 0; jreg < nregactive; ++jreg) {
        G[jreg][jreg] = od_region[jg][jreg] * gamma1[jg][jreg];
        G[jreg + nreg][jreg] = od_region[jg][jreg] * gamma2[jg][jreg];
        planck_top[nreg + jreg] = od_region[jg][jreg] * (1.0 - ssa_region[jg][jreg]) * reg[l][jreg] * pt * LwDiffusivity;
        planck_top[jreg] = -planck_top[nreg + jreg];
        planck_diff[nreg + jreg] = od_region[jg][jreg] * (1.0 - ssa_region[jg][jreg]) * reg[l][jreg] * (pb - pt) * LwDiffusivity;
        planck_diff[jreg] = -planck_diff[nreg + jreg];
      }
      for (int jreg = nregactive; jreg < nreg; ++jreg) { G[jreg][jreg] = G[0][0]; G[nreg + jreg][jreg] = G[nreg][0]; }
      double side_emiss = 1.0;
      if (cfg->do_lw_side_emissivity && reg[l][0] > 0.0 && reg[l][1] > 0.0 && cfg->do_3d_effects && inv_cloud_size && inv_cloud_size[l] > 0.0) {
        const double aspect_ratio = 1.0 / (dmin(inv_cloud_size[l], 1.0 / cfg->min_cloud_effective_size) * reg[l][0] * dz);
        double s = 0.0;   /* sum(od_region(:,2:nreg)*(1-ssa_region(:,2:nreg)),2) */
        for (int r = 1; r < nreg; ++r) s = s + od_region[jg][r] * (1.0 - ssa_region[jg][r]);
        const double lateral_od = (aspect_ratio / (nreg - 1.0)) * s;
        const double sqrt_1_minus_ssa = sqrt(1.0 - ssa_region[jg][1]);
        const double side_emiss_thick = 2.0 * sqrt_1_minus_ssa / (sqrt_1_minus_ssa + sqrt(1.0 - ssa_region[jg][1] * g_region[jg][1]));
        side_emiss = (side_emiss_thin - side_emiss_thick) / (lateral_od + 1.0) + side_emiss_thick;
      }
      for (int jreg = 0; jreg < nregactive - 1; ++jreg) {
        G[jreg][jreg] = G[jreg][jreg] + transfer_rate[jreg][jreg + 1];
        G[jreg + 1][jreg] = -transfer_rate[jreg][jreg + 1];
        if (jreg > 0) {
          G[jreg + 1][jreg + 1] = G[jreg + 1][jreg + 1] + transfer_rate[jreg + 1][jreg];
          G[jreg][jreg + 1] = -transfer_rate[jreg + 1][jreg];
        } else {
          G[jreg + 1][jreg + 1] = G[jreg + 1][jreg + 1] + side_emiss * transfer_rate[jreg + 1][jreg];
          G[jreg][jreg + 1] = -side_emiss * transfer_rate[jreg + 1][jreg];
        }
      }
      if (edge_length[2] > 0.0) {
        G[0][0] = G[0][0] + transfer_rate[0][2];
        G[2][0] = -transfer_rate[0][2];
        G[2][2] = G[2][2] + side_emiss * transfer_rate[2][0];
        G[0][2] = -side_emiss * transfer_rate[2][0];
      }
      for (int i = 0; i < nreg; ++i) for (int j = 0; j < nreg; ++j) G[nreg + i][nreg + j] = -G[i][j];
      for (int i = 0; i < nreg; ++i) for (int j = 0; j < nreg; ++j) G[i][nreg + j] = -G[nreg + i][j];
      solve_vec_n(6, G, planck_diff, solution_diff);
      for (int i = 0; i < 6; ++i) solution_diff[i] = -solution_diff[i];
      for (int i = 0; i < 6; ++i) rhs[i] = solution_diff[i] - planck_top[i];
      solve_vec_n(6, G, rhs, solution0);
      expm(6, G, 0);
      m3 E11, E12, E21, E22, X;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { E11[i][j] = G[i][j]; E12[i][j] = G[i][3 + j]; E21[i][j] = G[3 + i][j]; E22[i][j] = G[3 + i][3 + j]; }
      double (*R)[NREG] = AS_M3(L3(reflectance, l, jg)), (*T)[NREG] = AS_M3(L3(transmittance, l, jg));
      double *SU = L6(source_up, l, jg), *SD = L6(source_dn, l, jg);
      solve_mat_3(E11, E12, X);
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R[i][j] = -X[i][j];
      m3_x_m3(E21, R, X);
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) T[i][j] = X[i][j] + E22[i][j];
      double tmp[3], a[3], b[3];
      m3_x_vec(E12, solution0 + 3, a);
      for (int r = 0; r < 3; ++r) tmp[r] = solution0[r] + solution_diff[r] - a[r];
      solve_vec_3(E11, tmp, a);
      for (int r = 0; r < 3; ++r) SU[r] = solution0[r] - a[r];
      for (int r = 0; r < 3; ++r) tmp[r] = SU[r] - solution0[r];
      m3_x_vec(E21, tmp, a);
      m3_x_vec(E22, solution0 + 3, b);
      for (int r = 0; r < 3; ++r) SD[r] = a[r] + solution0[3 + r] - b[r] + solution_diff[3 + r];
    }
    /* 3.3b (:729-778) */
    for (int jg = 0; jg < ng; ++jg) {
      const size_t i = (size_t)l * ng + jg;
      ref_trans_lw(od_region[jg][0], gamma1[jg][0], gamma2[jg][0], planck_hl[i], planck_hl[i + ng], &ref_clear[i], &trans_clear[i],
                   &source_up_clear[i], &source_dn_clear[i]);
    }
    for (int jg = ng3D; jg < ng; ++jg) {
      const size_t i = (size_t)l * ng + jg;
      double (*R)[NREG] = AS_M3(L3(reflectance, l, jg)), (*T)[NREG] = AS_M3(L3(transmittance, l, jg));
      double *SU = L6(source_up, l, jg), *SD = L6(source_dn, l, jg);
      for (int a = 0; a < 3; ++a) { for (int b = 0; b < 3; ++b) { R[a][b] = 0.0; T[a][b] = 0.0; } SU[a] = 0.0; SD[a] = 0.0; }
      R[0][0] = ref_clear[i]; T[0][0] = trans_clear[i];
      SU[0] = reg[l][0] * source_up_clear[i]; SD[0] = reg[l][0] * source_dn_clear[i];
      for (int jreg = 1; jreg < nregactive; ++jreg)
        ref_trans_lw(od_region[jg][jreg], gamma1[jg][jreg], gamma2[jg][jreg], reg[l][jreg] * planck_hl[i], reg[l][jreg] * planck_hl[i + ng],
                     &R[jreg][jreg], &T[jreg][jreg], &SU[jreg], &SD[jreg]);
    }
  }

  /* ---- Section 4: total sources and albedos (:782-905) ---- */
  const int matrix_adding = cfg->do_3d_effects || cfg->do_3d_lw_multilayer_effects;
  for (int jg = 0; jg < ng; ++jg) {
    double (*TA)[NREG] = AS_M3(L3(total_albedo, nlev, jg));
    double* TS = L6(total_source, nlev, jg);
    for (int jreg = 0; jreg < nreg; ++jreg) { TS[jreg] = reg[nlev - 1][jreg] * emission[jg]; TA[jreg][jreg] = albedo[jg]; }
    total_source_clear[(size_t)nlev * ng + jg] = emission[jg];
    total_albedo_clear[(size_t)nlev * ng + jg] = TA[0][0];
  }
  for (int jlev = nlev; jlev >= 1; --jlev) {
    const int l = jlev - 1;
    for (int jg = 0; jg < ng; ++jg) {
      const size_t i = (size_t)l * ng + jg, ib = (size_t)jlev * ng + jg;
      {
        const double inv_denom = 1.0 / (1.0 - total_albedo_clear[ib] * ref_clear[i]);
        total_albedo_clear[i] = ref_clear[i] + trans_clear[i] * trans_clear[i] * total_albedo_clear[ib] * inv_denom;
        total_source_clear[i] = source_up_clear[i] + trans_clear[i] * (total_source_clear[ib] + total_albedo_clear[ib] * source_dn_clear[i]) * inv_denom;
      }
      double (*R)[NREG] = AS_M3(L3(reflectance, l, jg)), (*T)[NREG] = AS_M3(L3(transmittance, l, jg));
      double *SU = L6(source_up, l, jg), *SD = L6(source_dn, l, jg);
      double (*TAb)[NREG] = AS_M3(L3(total_albedo, jlev, jg)), (*TA)[NREG] = AS_M3(L3(total_albedo, l, jg));
      double *TSb = L6(total_source, jlev, jg), *TS = L6(total_source, l, jg);
      m3 below; double sbelow[NREG] = {0.0, 0.0, 0.0};
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) below[a][b] = 0.0;
      if (clear[jlev]) {
        const double inv_denom = 1.0 / (1.0 - TAb[0][0] * R[0][0]);
        below[0][0] = R[0][0] + T[0][0] * T[0][0] * TAb[0][0] * inv_denom;
        sbelow[0] = SU[0] + T[0][0] * (TSb[0] + TAb[0][0] * SD[0]) * inv_denom;
      } else if (matrix_adding) {
        m3 denominator, X, Y, Z;
        double a[NREG], b[NREG];
        identity_minus_m3_x_m3(TAb, R, denominator);
        m3_x_m3(TAb, T, X);
        solve_mat_3(denominator, X, Y);
        m3_x_m3(T, Y, Z);
        for (int p = 0; p < 3; ++p) for (int q = 0; q < 3; ++q) below[p][q] = R[p][q] + Z[p][q];
        m3_x_vec(TAb, SD, a);
        for (int r = 0; r < 3; ++r) a[r] = TSb[r] + a[r];
        solve_vec_3(denominator, a, b);
        m3_x_vec(T, b, a);
        for (int r = 0; r < 3; ++r) sbelow[r] = SU[r] + a[r];
      } else {
        for (int jreg = 0; jreg < nreg; ++jreg) {
          const double inv_denom = 1.0 / (1.0 - TAb[jreg][jreg] * R[jreg][jreg]);
          below[jreg][jreg] = R[jreg][jreg] + T[jreg][jreg] * T[jreg][jreg] * TAb[jreg][jreg] * inv_denom;
          sbelow[jreg] = SU[jreg] + T[jreg][jreg] * (TSb[jreg] + TAb[jreg][jreg] * SD[jreg]) * inv_denom;
        }
      }
      for (int a = 0; a < 3; ++a) { for (int b = 0; b < 3; ++b) TA[a][b] = 0.0; TS[a] = 0.0; }
      if (clear[jlev] && clear[jlev - 1]) {
        TA[0][0] = below[0][0];
        TS[0] = sbelow[0];
      } else {
        m3_x_vec(U[l], sbelow, TS);
        if (cfg->do_3d_lw_multilayer_effects) {
          u_x_mat_x_v(U[l], below, V[l], TA);
        } else {
          for (int jreg = 0; jreg < nreg; ++jreg)
            for (int jreg2 = 0; jreg2 < nreg; ++jreg2) TA[jreg][jreg] = TA[jreg][jreg] + below[jreg2][jreg2] * V[l][jreg2][jreg];
        }
      }
    }
  }

  /* ---- Section 5: fluxes (:907-1040) ---- */
  double (*flux_up)[NREG] = calloc(ng, sizeof(double[NREG])), (*flux_dn)[NREG] = calloc(ng, sizeof(double[NREG]));
  double *flux_up_clear = calloc(2 * (size_t)ng, sizeof(double)), *flux_dn_clear = flux_up_clear + ng;
  for (int hl = 0; hl <= nlev; ++hl) {
    if (hl == 0) {
      for (int jg = 0; jg < ng; ++jg) {
        const double* TS = L6(total_source, 0, jg);
        for (int r = 0; r < 3; ++r) { flux_up[jg][r] = TS[r]; flux_dn[jg][r] = 0.0; }
        flux_up_clear[jg] = total_source_clear[jg]; flux_dn_clear[jg] = 0.0;
        o->up_toa_g[jg] = TS[0] + TS[1] + TS[2];
        o->up_toa_clear_g[jg] = total_source_clear[jg];
      }
    } else {
      const int l = hl - 1, jlev = hl;
      for (int jg = 0; jg < ng; ++jg) {
        const size_t i = (size_t)l * ng + jg, ib = (size_t)jlev * ng + jg;
        double (*R)[NREG] = AS_M3(L3(reflectance, l, jg)), (*T)[NREG] = AS_M3(L3(transmittance, l, jg));
        double* SD = L6(source_dn, l, jg);
        double (*TAb)[NREG] = AS_M3(L3(total_albedo, jlev, jg));
        double* TSb = L6(total_source, jlev, jg);
        double flux_dn_above[NREG], flux_up_above[NREG];
        flux_dn_clear[jg] = (trans_clear[i] * flux_dn_clear[jg] + ref_clear[i] * total_source_clear[ib] + source_dn_clear[i]) /
                            (1.0 - ref_clear[i] * total_albedo_clear[ib]);
        flux_up_clear[jg] = total_source_clear[ib] + total_albedo_clear[ib] * flux_dn_clear[jg];
        if (clear[jlev]) {
          flux_dn_above[0] = (T[0][0] * flux_dn[jg][0] + R[0][0] * TSb[0] + SD[0]) / (1.0 - R[0][0] * TAb[0][0]);
          flux_dn_above[1] = 0.0; flux_dn_above[2] = 0.0;
          flux_up_above[0] = TSb[0] + TAb[0][0] * flux_dn_above[0];
          flux_up_above[1] = 0.0; flux_up_above[2] = 0.0;
        } else if (matrix_adding) {
          m3 denominator;
          double a[NREG], b[NREG], rhs[NREG];
          identity_minus_m3_x_m3(R, TAb, denominator);
          m3_x_vec(T, flux_dn[jg], a);
          m3_x_vec(R, TSb, b);
          for (int r = 0; r < 3; ++r) rhs[r] = a[r] + b[r] + SD[r];
          solve_vec_3(denominator, rhs, flux_dn_above);
          m3_x_vec(TAb, flux_dn_above, flux_up_above);
          for (int r = 0; r < 3; ++r) flux_up_above[r] = flux_up_above[r] + TSb[r];
        } else {
          for (int jreg = 0; jreg < nreg; ++jreg) {
            flux_dn_above[jreg] = (T[jreg][jreg] * flux_dn[jg][jreg] + R[jreg][jreg] * TSb[jreg] + SD[jreg]) / (1.0 - R[jreg][jreg] * TAb[jreg][jreg]);
            flux_up_above[jreg] = TSb[jreg] + TAb[jreg][jreg] * flux_dn_above[jreg];
          }
        }
        for (int r = 0; r < 3; ++r) { flux_up[jg][r] = flux_up_above[r]; flux_dn[jg][r] = flux_dn_above[r]; }
      }
    }
    {
      double su[NREG], sd[NREG], suc = 0.0, sdc = 0.0;
      for (int r = 0; r < 3; ++r) { double a = 0.0, b = 0.0; for (int jg = 0; jg < ng; ++jg) { a = a + flux_up[jg][r]; b = b + flux_dn[jg][r]; } su[r] = a; sd[r] = b; }
      for (int jg = 0; jg < ng; ++jg) { suc = suc + flux_up_clear[jg]; sdc = sdc + flux_dn_clear[jg]; }
      o->up[hl] = su[0] + su[1] + su[2];
      o->dn[hl] = hl == 0 ? 0.0 : sd[0] + sd[1] + sd[2];
      o->up_clear[hl] = suc;
      o->dn_clear[hl] = hl == 0 ? 0.0 : sdc;
      if (o->up_g_prof)
        for (int jg = 0; jg < ng; ++jg) {
          o->up_g_prof[(size_t)hl * ng + jg] = flux_up[jg][0] + flux_up[jg][1] + flux_up[jg][2];
          o->dn_dif_g_prof[(size_t)hl * ng + jg] = flux_dn[jg][0] + flux_dn[jg][1] + flux_dn[jg][2];
        }
    }
    if (hl == nlev) {
      for (int jg = 0; jg < ng; ++jg) { o->dn_diffuse_surf_g[jg] = flux_dn[jg][0] + flux_dn[jg][1] + flux_dn[jg][2]; o->dn_diffuse_surf_clear_g[jg] = flux_dn_clear[jg]; }
    } else if (hl > 0) {
      const int jlev = hl;
      if (!(clear[jlev] && clear[jlev + 1]))
        for (int jg = 0; jg < ng; ++jg) m3_x_vec(V[jlev], flux_dn[jg], flux_dn[jg]);
    }
  }
  /* calc_lw_derivatives_matrix radiation_lw_derivatives.F90:138-193 */
  if (cfg->do_lw_derivatives && o->lw_deriv) {
    double (*d)[NREG] = calloc(ng, sizeof(double[NREG]));
    double tot = 0.0;
    for (int jg = 0; jg < ng; ++jg) tot = tot + (flux_up[jg][0] + flux_up[jg][1] + flux_up[jg][2]);
    for (int jg = 0; jg < ng; ++jg) { d[jg][0] = (flux_up[jg][0] + flux_up[jg][1] + flux_up[jg][2]) / tot; d[jg][1] = 0.0; d[jg][2] = 0.0; }
    o->lw_deriv[nlev] = 1.0;
    for (int jlev = nlev; jlev >= 1; --jlev) {
      double sr[NREG] = {0.0, 0.0, 0.0};
      for (int jg = 0; jg < ng; ++jg) {
        m3_x_vec(U[jlev], d[jg], d[jg]);
        m3_x_vec(AS_M3(L3(transmittance, jlev - 1, jg)), d[jg], d[jg]);
      }
      for (int r = 0; r < 3; ++r) { double s = 0.0; for (int jg = 0; jg < ng; ++jg) s = s + d[jg][r]; sr[r] = s; }
      o->lw_deriv[jlev - 1] = sr[0] + sr[1] + sr[2];
    }
    free(d);
  }
  free(flux_up); free(flux_dn); free(flux_up_clear);
  free(od_region); free(ssa_region); free(g_region); free(gamma1); free(gamma2);
  free(buf); free(clear); free(reg); free(ods); free(U); free(V);
}
