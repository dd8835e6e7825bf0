/* solvers.c -- oracle restatement of the two-stream layer solutions and the adding method.  TEST INFRASTRUCTURE.
 * Arrays are [lev][g] with g fastest, exactly the reference's (ng, nlev) Fortran arrays.
 * Follows radiation/radiation_two_stream.F90, radiation_adding_ica_sw.F90, radiation_adding_ica_lw.F90.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include "oracle.h"

static const double LwDiffusivity = 1.66;   /* radiation_two_stream.F90:36-37 */
static inline double dmin(double a, double b) { return a < b ? a : b; }
static inline double dmax(double a, double b) { return a > b ? a : b; }

/* radiation_two_stream.F90:246-333 calc_ref_trans_lw */
void orc_calc_ref_trans_lw(int ng, const double* od, const double* ssa, const double* asym, const double* planck_top,
                           const double* planck_bot, double* ref, double* trans, double* source_up, double* source_dn) {
  for (int jg = 0; jg < ng; ++jg) {
    double factor = (LwDiffusivity * 0.5) * ssa[jg];
    double gamma1 = LwDiffusivity - factor * (1.0 + asym[jg]);
    double gamma2 = factor * (1.0 - asym[jg]);
    double k_exponent = sqrt(dmax((gamma1 - gamma2) * (gamma1 + gamma2), 1.0e-12));
    if (od[jg] > 1.0e-3) {
      double exponential = exp(-k_exponent * od[jg]);
      double exponential2 = exponential * exponential;
      double reftrans_factor = 1.0 / (k_exponent + gamma1 + (k_exponent - gamma1) * exponential2);
      ref[jg] = gamma2 * (1.0 - exponential2) * reftrans_factor;
      trans[jg] = 2.0 * k_exponent * exponential * reftrans_factor;
      double coeff = (planck_bot[jg] - planck_top[jg]) / (od[jg] * (gamma1 + gamma2));
      double coeff_up_top = coeff + planck_top[jg];
      double coeff_up_bot = coeff + planck_bot[jg];
      double coeff_dn_top = -coeff + planck_top[jg];
      double coeff_dn_bot = -coeff + planck_bot[jg];
      source_up[jg] = coeff_up_top - ref[jg] * coeff_dn_top - trans[jg] * coeff_up_bot;
      source_dn[jg] = coeff_dn_bot - ref[jg] * coeff_up_bot - trans[jg] * coeff_dn_top;
    } else {
      ref[jg] = gamma2 * od[jg];
      trans[jg] = (1.0 - k_exponent * od[jg]) / (1.0 + od[jg] * (gamma1 - k_exponent));
      /* "0.5" is a default-kind literal in the source; it is exact in single precision */
      source_up[jg] = (1.0 - ref[jg] - trans[jg]) * 0.5 * (planck_top[jg] + planck_bot[jg]);
      source_dn[jg] = source_up[jg];
    }
  }
}

/* radiation_two_stream.F90:342-409 calc_no_scattering_transmittance_lw */
void orc_calc_no_scattering_transmittance_lw(int ng, const double* od, const double* planck_top,
                                             const double* planck_bot, double* trans, double* source_up,
                                             double* source_dn) {
  for (int jg = 0; jg < ng; ++jg) {
    trans[jg] = exp(-LwDiffusivity * od[jg]);
    double coeff = LwDiffusivity * od[jg];
    if (od[jg] > 1.0e-3) {
      coeff = (planck_bot[jg] - planck_top[jg]) / coeff;
      double coeff_up_top = coeff + planck_top[jg];
      double coeff_up_bot = coeff + planck_bot[jg];
      double coeff_dn_top = -coeff + planck_top[jg];
      double coeff_dn_bot = -coeff + planck_bot[jg];
      source_up[jg] = coeff_up_top - trans[jg] * coeff_up_bot;
      source_dn[jg] = coeff_dn_bot - trans[jg] * coeff_dn_top;
    } else {
      source_up[jg] = coeff * 0.5 * (planck_top[jg] + planck_bot[jg]);
      source_dn[jg] = source_up[jg];
    }
  }
}

/* radiation_two_stream.F90:563-696 calc_ref_trans_sw (non-DWD branch, double precision) */
void orc_calc_ref_trans_sw(int ng, double mu0, const double* od, const double* ssa, const double* asym,
                           double* ref_diff, double* trans_diff, double* ref_dir, double* trans_dir_diff,
                           double* trans_dir_dir) {
  const double eps = DBL_EPSILON;
  for (int jg = 0; jg < ng; ++jg) {
    double tdd = dmax(-dmax(od[jg] * (1.0 / mu0), 0.0), -1000.0);
    tdd = exp(tdd);
    trans_dir_dir[jg] = tdd;
    double factor = 0.75 * asym[jg];
    double gamma1 = 2.0 - ssa[jg] * (1.25 + factor);
    double gamma2 = ssa[jg] * (0.75 - factor);
    double gamma3 = 0.5 - mu0 * factor;
    double gamma4 = 1.0 - gamma3;
    double alpha1 = gamma1 * gamma4 + gamma2 * gamma3;
    double alpha2 = gamma1 * gamma3 + gamma2 * gamma4;
    double k_exponent = sqrt(dmax((gamma1 - gamma2) * (gamma1 + gamma2), 1.0e-12));
    double exponential = exp(-k_exponent * od[jg]);
    double k_mu0 = k_exponent * mu0;
    double one_minus_kmu0_sqr = 1.0 - k_mu0 * k_mu0;
    double k_gamma3 = k_exponent * gamma3;
    double k_gamma4 = k_exponent * gamma4;
    double exponential2 = exponential * exponential;
    double k_2_exponential = 2.0 * k_exponent * exponential;
    double reftrans_factor = 1.0 / (k_exponent + gamma1 + (k_exponent - gamma1) * exponential2);
    ref_diff[jg] = gamma2 * (1.0 - exponential2) * reftrans_factor;
    trans_diff[jg] = dmax(0.0, dmin(k_2_exponential * reftrans_factor, 1.0 - ref_diff[jg]));
    reftrans_factor = mu0 * ssa[jg] * reftrans_factor / (fabs(one_minus_kmu0_sqr) > eps ? one_minus_kmu0_sqr : eps);
    double rd = reftrans_factor * ((1.0 - k_mu0) * (alpha2 + k_gamma3) - (1.0 + k_mu0) * (alpha2 - k_gamma3) * exponential2 -
                                   k_2_exponential * (gamma3 - alpha2 * mu0) * tdd);
    double td = reftrans_factor * (k_2_exponential * (gamma4 + alpha1 * mu0) -
                                   tdd * ((1.0 + k_mu0) * (alpha1 + k_gamma4) - (1.0 - k_mu0) * (alpha1 - k_gamma4) * exponential2));
    rd = dmax(0.0, dmin(rd, mu0 * (1.0 - tdd)));
    td = dmax(0.0, dmin(td, mu0 * (1.0 - tdd) - rd));
    ref_dir[jg] = rd;
    trans_dir_diff[jg] = td;
  }
}

/* radiation_two_stream.F90:96-146 calc_two_stream_gammas_sw + :421-560 calc_reflectance_transmittance_sw
 * (the pair used by the Cloudless solver, radiation_cloudless_sw.F90:104-116) */
void orc_calc_reflectance_transmittance_sw(int ng, double mu0, const double* od, const double* ssa, const double* asym,
                                           double* ref_diff, double* trans_diff, double* ref_dir,
                                           double* trans_dir_diff, double* trans_dir_dir) {
  for (int jg = 0; jg < ng; ++jg) {
    double factor = 0.75 * asym[jg];
    double gamma1 = 2.0 - ssa[jg] * (1.25 + factor);
    double gamma2 = ssa[jg] * (0.75 - factor);
    double gamma3 = 0.5 - mu0 * factor;
    double gamma4 = 1.0 - gamma3;
    double alpha1 = gamma1 * gamma4 + gamma2 * gamma3;
    double alpha2 = gamma1 * gamma3 + gamma2 * gamma4;
    double k_exponent = sqrt(dmax((gamma1 - gamma2) * (gamma1 + gamma2), 1.0e-12));
    double mu0_local = mu0;
    if (fabs(1.0 - k_exponent * mu0) < 1000.0 * DBL_EPSILON) mu0_local = mu0 * (1.0 - 10.0 * DBL_EPSILON);
    double od_over_mu0 = dmax(od[jg] / mu0_local, 0.0);
    double k_mu0 = k_exponent * mu0_local;
    double k_gamma3 = k_exponent * gamma3;
    double k_gamma4 = k_exponent * gamma4;
    double exponential0 = exp(-od_over_mu0);
    trans_dir_dir[jg] = exponential0;
    double exponential = exp(-k_exponent * od[jg]);
    double exponential2 = exponential * exponential;
    double k_2_exponential = 2.0 * k_exponent * exponential;
    double reftrans_factor = 1.0 / (k_exponent + gamma1 + (k_exponent - gamma1) * exponential2);
    ref_diff[jg] = gamma2 * (1.0 - exponential2) * reftrans_factor;
    trans_diff[jg] = k_2_exponential * reftrans_factor;
    reftrans_factor = mu0_local * ssa[jg] * reftrans_factor / (1.0 - k_mu0 * k_mu0);
    double rd = reftrans_factor * ((1.0 - k_mu0) * (alpha2 + k_gamma3) - (1.0 + k_mu0) * (alpha2 - k_gamma3) * exponential2 -
                                   k_2_exponential * (gamma3 - alpha2 * mu0_local) * exponential0);
    double td = reftrans_factor * (k_2_exponential * (gamma4 + alpha1 * mu0_local) -
                                   exponential0 * ((1.0 + k_mu0) * (alpha1 + k_gamma4) - (1.0 - k_mu0) * (alpha1 - k_gamma4) * exponential2));
    rd = dmax(0.0, dmin(rd, 1.0));
    td = dmax(0.0, dmin(td, 1.0 - rd));
    ref_dir[jg] = rd;
    trans_dir_diff[jg] = td;
  }
}

#define IX(l, g) ((size_t)(l) * ng + (g))

/* radiation_adding_ica_sw.F90:24-151.  Layer arrays [nlev][ng]; flux arrays [nlev+1][ng]. */
void orc_adding_ica_sw(int ng, int nlev, const double* incoming, const double* alb_diff, const double* alb_dir,
                       double cos_sza, const double* ref, const double* trans, const double* ref_dir,
                       const double* trans_dir_diff, const double* trans_dir_dir,
                       double* flux_up, double* flux_dn_diffuse, double* flux_dn_direct) {
  double* albedo = (double*)malloc(sizeof(double) * (size_t)(nlev + 1) * ng);
  double* source = (double*)malloc(sizeof(double) * (size_t)(nlev + 1) * ng);
  double* inv_den = (double*)malloc(sizeof(double) * (size_t)nlev * ng);
  for (int g = 0; g < ng; ++g) flux_dn_direct[IX(0, g)] = incoming[g];
  for (int l = 0; l < nlev; ++l)
    for (int g = 0; g < ng; ++g) flux_dn_direct[IX(l + 1, g)] = flux_dn_direct[IX(l, g)] * trans_dir_dir[IX(l, g)];
  for (int g = 0; g < ng; ++g) {
    albedo[IX(nlev, g)] = alb_diff[g];
    source[IX(nlev, g)] = alb_dir[g] * flux_dn_direct[IX(nlev, g)] * cos_sza;
  }
  for (int l = nlev - 1; l >= 0; --l) {
    for (int g = 0; g < ng; ++g) {
      inv_den[IX(l, g)] = 1.0 / (1.0 - albedo[IX(l + 1, g)] * ref[IX(l, g)]);
      albedo[IX(l, g)] = ref[IX(l, g)] + trans[IX(l, g)] * trans[IX(l, g)] * albedo[IX(l + 1, g)] * inv_den[IX(l, g)];
      source[IX(l, g)] = ref_dir[IX(l, g)] * flux_dn_direct[IX(l, g)] +
                         trans[IX(l, g)] * (source[IX(l + 1, g)] + albedo[IX(l + 1, g)] * trans_dir_diff[IX(l, g)] * flux_dn_direct[IX(l, g)]) *
                             inv_den[IX(l, g)];
    }
  }
  for (int g = 0; g < ng; ++g) { flux_dn_diffuse[IX(0, g)] = 0.0; flux_up[IX(0, g)] = source[IX(0, g)]; }
  for (int l = 0; l < nlev; ++l) {
    for (int g = 0; g < ng; ++g) {
      flux_dn_diffuse[IX(l + 1, g)] = (trans[IX(l, g)] * flux_dn_diffuse[IX(l, g)] + ref[IX(l, g)] * source[IX(l + 1, g)] +
                                       trans_dir_diff[IX(l, g)] * flux_dn_direct[IX(l, g)]) * inv_den[IX(l, g)];
      flux_up[IX(l + 1, g)] = albedo[IX(l + 1, g)] * flux_dn_diffuse[IX(l + 1, g)] + source[IX(l + 1, g)];
      flux_dn_direct[IX(l, g)] = flux_dn_direct[IX(l, g)] * cos_sza;
    }
  }
  for (int g = 0; g < ng; ++g) flux_dn_direct[IX(nlev, g)] = flux_dn_direct[IX(nlev, g)] * cos_sza;
  free(albedo); free(source); free(inv_den);
}

/* radiation_adding_ica_lw.F90:272-332 calc_fluxes_no_scattering_lw */
void orc_calc_fluxes_no_scattering_lw(int ng, int nlev, const double* trans, const double* source_up,
                                      const double* source_dn, const double* emission, const double* albedo,
                                      double* flux_up, double* flux_dn) {
  for (int g = 0; g < ng; ++g) flux_dn[IX(0, g)] = 0.0;
  for (int l = 0; l < nlev; ++l)
    for (int g = 0; g < ng; ++g) flux_dn[IX(l + 1, g)] = trans[IX(l, g)] * flux_dn[IX(l, g)] + source_dn[IX(l, g)];
  for (int g = 0; g < ng; ++g) flux_up[IX(nlev, g)] = emission[g] + albedo[g] * flux_dn[IX(nlev, g)];
  for (int l = nlev - 1; l >= 0; --l)
    for (int g = 0; g < ng; ++g) flux_up[IX(l, g)] = trans[IX(l, g)] * flux_up[IX(l + 1, g)] + source_up[IX(l, g)];
}

/* radiation_adding_ica_lw.F90:137-263 fast_adding_ica_lw.  i_cloud_top is 1-based as in the source. */
void orc_fast_adding_ica_lw(int ng, int nlev, const double* ref, const double* trans, const double* source_up,
                            const double* source_dn, const double* emission, const double* albedo_surf,
                            const int* is_clear_sky_layer, int i_cloud_top, const double* flux_dn_clear,
                            double* flux_up, double* flux_dn) {
  double* albedo = (double*)malloc(sizeof(double) * (size_t)(nlev + 1) * ng);
  double* source = (double*)malloc(sizeof(double) * (size_t)(nlev + 1) * ng);
  double* inv_den = (double*)malloc(sizeof(double) * (size_t)nlev * ng);
  const int ict = i_cloud_top - 1; /* 0-based half-level index of cloud top */
  for (int l = 0; l <= ict; ++l)
    for (int g = 0; g < ng; ++g) flux_dn[IX(l, g)] = flux_dn_clear[IX(l, g)];
  for (int g = 0; g < ng; ++g) { albedo[IX(nlev, g)] = albedo_surf[g]; source[IX(nlev, g)] = emission[g]; }
  for (int l = nlev - 1; l >= ict; --l) {
    if (is_clear_sky_layer[l]) {
      for (int g = 0; g < ng; ++g) {
        albedo[IX(l, g)] = trans[IX(l, g)] * trans[IX(l, g)] * albedo[IX(l + 1, g)];
        source[IX(l, g)] = source_up[IX(l, g)] + trans[IX(l, g)] * (source[IX(l + 1, g)] + albedo[IX(l + 1, g)] * source_dn[IX(l, g)]);
      }
    } else {
      for (int g = 0; g < ng; ++g) {
        inv_den[IX(l, g)] = 1.0 / (1.0 - albedo[IX(l + 1, g)] * ref[IX(l, g)]);
        albedo[IX(l, g)] = ref[IX(l, g)] + trans[IX(l, g)] * trans[IX(l, g)] * albedo[IX(l + 1, g)] * inv_den[IX(l, g)];
        source[IX(l, g)] = source_up[IX(l, g)] +
                           trans[IX(l, g)] * (source[IX(l + 1, g)] + albedo[IX(l + 1, g)] * source_dn[IX(l, g)]) * inv_den[IX(l, g)];
      }
    }
  }
  for (int g = 0; g < ng; ++g) flux_up[IX(ict, g)] = source[IX(ict, g)] + albedo[IX(ict, g)] * flux_dn[IX(ict, g)];
  for (int l = ict - 1; l >= 0; --l)
    for (int g = 0; g < ng; ++g) flux_up[IX(l, g)] = trans[IX(l, g)] * flux_up[IX(l + 1, g)] + source_up[IX(l, g)];
  for (int l = ict; l < nlev; ++l) {
    if (is_clear_sky_layer[l]) {
      for (int g = 0; g < ng; ++g) {
        flux_dn[IX(l + 1, g)] = trans[IX(l, g)] * flux_dn[IX(l, g)] + source_dn[IX(l, g)];
        flux_up[IX(l + 1, g)] = albedo[IX(l + 1, g)] * flux_dn[IX(l + 1, g)] + source[IX(l + 1, g)];
      }
    } else {
      for (int g = 0; g < ng; ++g) {
        flux_dn[IX(l + 1, g)] = (trans[IX(l, g)] * flux_dn[IX(l, g)] + ref[IX(l, g)] * source[IX(l + 1, g)] + source_dn[IX(l, g)]) * inv_den[IX(l, g)];
        flux_up[IX(l + 1, g)] = albedo[IX(l + 1, g)] * flux_dn[IX(l + 1, g)] + source[IX(l + 1, g)];
      }
    }
  }
  free(albedo); free(source); free(inv_den);
}

/* radiation_adding_ica_lw.F90:32-131 adding_ica_lw (full adding method, used by the Homogeneous solver). */
void orc_adding_ica_lw(int ng, int nlev, const double* ref, const double* trans, const double* source_up, const double* source_dn,
                       const double* emission, const double* albedo_surf, double* flux_up, double* flux_dn) {
  double* albedo = (double*)malloc(sizeof(double) * (size_t)(2 * (nlev + 1) + nlev) * ng);
  double* source = albedo + (size_t)(nlev + 1) * ng;
  double* inv_denominator = source + (size_t)(nlev + 1) * ng;
  for (int g = 0; g < ng; ++g) { albedo[IX(nlev, g)] = albedo_surf[g]; source[IX(nlev, g)] = emission[g]; }
  for (int jl = nlev - 1; jl >= 0; --jl)
    for (int g = 0; g < ng; ++g) {
      inv_denominator[IX(jl, g)] = 1.0 / (1.0 - albedo[IX(jl + 1, g)] * ref[IX(jl, g)]);
      albedo[IX(jl, g)] = ref[IX(jl, g)] + trans[IX(jl, g)] * trans[IX(jl, g)] * albedo[IX(jl + 1, g)] * inv_denominator[IX(jl, g)];
      source[IX(jl, g)] = source_up[IX(jl, g)] +
                          trans[IX(jl, g)] * (source[IX(jl + 1, g)] + albedo[IX(jl + 1, g)] * source_dn[IX(jl, g)]) * inv_denominator[IX(jl, g)];
    }
  for (int g = 0; g < ng; ++g) { flux_dn[IX(0, g)] = 0.0; flux_up[IX(0, g)] = source[IX(0, g)]; }
  for (int jl = 0; jl < nlev; ++jl)
    for (int g = 0; g < ng; ++g) {
      flux_dn[IX(jl + 1, g)] = (trans[IX(jl, g)] * flux_dn[IX(jl, g)] + ref[IX(jl, g)] * source[IX(jl + 1, g)] + source_dn[IX(jl, g)]) *
                               inv_denominator[IX(jl, g)];
      flux_up[IX(jl + 1, g)] = albedo[IX(jl + 1, g)] * flux_dn[IX(jl + 1, g)] + source[IX(jl + 1, g)];
    }
  free(albedo);
}
