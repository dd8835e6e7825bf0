// solver_core.h -- two-stream layer solutions (per g-point, per layer) used by the McICA / Cloudless sweeps.
//
// Reference map:  lw_no_scat   <- radiation/radiation_two_stream.F90:342-409 calc_no_scattering_transmittance_lw
//                 lw_ref_trans <- radiation/radiation_two_stream.F90:246-333 calc_ref_trans_lw
//                 sw_ref_trans <- radiation/radiation_two_stream.F90:563-696 calc_ref_trans_sw (double precision)
//                 sw_ref_trans_cloudless <- :96-146 calc_two_stream_gammas_sw + :421-560
//                                           calc_reflectance_transmittance_sw (pair used by radiation_cloudless_sw.F90)
#pragma once
#include <float.h>

#include "hd.h"

namespace ecb {

#define ECB_LW_DIFFUSIVITY 1.66

struct LwLayer { double ref, trans, source_up, source_dn; };

HD LwLayer lw_no_scat(double od, double planck_top, double planck_bot) {
  LwLayer r;
  r.ref = 0.0;
  r.trans = exp(-ECB_LW_DIFFUSIVITY * od);
  double coeff = ECB_LW_DIFFUSIVITY * od;
  if (od > 1.0e-3) {
    coeff = (planck_bot - planck_top) / coeff;
    double coeff_up_top = coeff + planck_top, coeff_up_bot = coeff + planck_bot;
    double coeff_dn_top = -coeff + planck_top, coeff_dn_bot = -coeff + planck_bot;
    r.source_up = coeff_up_top - r.trans * coeff_up_bot;
    r.source_dn = coeff_dn_bot - r.trans * coeff_dn_top;
  } else {
    r.source_up = coeff * 0.5 * (planck_top + planck_bot);
    r.source_dn = r.source_up;
  }
  return r;
}

HD LwLayer lw_ref_trans(double od, double ssa, double asym, double planck_top, double planck_bot) {
  LwLayer r;
  double factor = (ECB_LW_DIFFUSIVITY * 0.5) * ssa;
  double gamma1 = ECB_LW_DIFFUSIVITY - factor * (1.0 + asym);
  double gamma2 = factor * (1.0 - asym);
  double k_exponent = sqrt(dmax((gamma1 - gamma2) * (gamma1 + gamma2), 1.0e-12));
  if (od > 1.0e-3) {
    double exponential = exp(-k_exponent * od);
    double exponential2 = exponential * exponential;
    double reftrans_factor = 1.0 / (k_exponent + gamma1 + (k_exponent - gamma1) * exponential2);
    r.ref = gamma2 * (1.0 - exponential2) * reftrans_factor;
    r.trans = 2.0 * k_exponent * exponential * reftrans_factor;
    double coeff = (planck_bot - planck_top) / (od * (gamma1 + gamma2));
    double coeff_up_top = coeff + planck_top, coeff_up_bot = coeff + planck_bot;
    double coeff_dn_top = -coeff + planck_top, coeff_dn_bot = -coeff + planck_bot;
    r.source_up = coeff_up_top - r.ref * coeff_dn_top - r.trans * coeff_up_bot;
    r.source_dn = coeff_dn_bot - r.ref * coeff_up_bot - r.trans * coeff_dn_top;
  } else {
    r.ref = gamma2 * od;
    r.trans = (1.0 - k_exponent * od) / (1.0 + od * (gamma1 - k_exponent));
    r.source_up = (1.0 - r.ref - r.trans) * 0.5 * (planck_top + planck_bot);
    r.source_dn = r.source_up;
  }
  return r;
}

struct SwLayer { double ref, trans, ref_dir, trans_dir_diff, trans_dir_dir; };

// delta_eddington, radiation_delta_eddington.h:20-37: with do_sw_delta_scaling_with_gases the solvers scale the gas-aerosol(-cloud)
// mixture instead of the cloud and aerosol optics on their own
HD void sw_delta_eddington(double& od, double& ssa, double& g) {
  const double f = g * g;
  od = od * (1.0 - ssa * f);
  ssa = ssa * (1.0 - f) / (1.0 - ssa * f);
  g = g / (1.0 + g);
}

HD SwLayer sw_ref_trans(double mu0, double od, double ssa, double asym) {
  SwLayer r;
  const double eps = DBL_EPSILON;
  double tdd = dmax(-dmax(od * (1.0 / mu0), 0.0), -1000.0);
  tdd = exp(tdd);
  r.trans_dir_dir = tdd;
  double factor = 0.75 * asym;
  double gamma1 = 2.0 - ssa * (1.25 + factor);
  double gamma2 = ssa * (0.75 - factor);
  double gamma3 = 0.5 - mu0 * factor;
  double gamma4 = 1.0 - gamma3;
  double alpha1 = gamma1 * gamma4 + gamma2 * gamma3;
  double alpha2 = gamma1 * gamma3 + gamma2 * gamma4;
  double k_exponent = sqrt(dmax((gamma1 - gamma2) * (gamma1 + gamma2), 1.0e-12));
  double exponential = exp(-k_exponent * od);
  double k_mu0 = k_exponent * mu0;
  double one_minus_kmu0_sqr = 1.0 - k_mu0 * k_mu0;
  double k_gamma3 = k_exponent * gamma3;
  double k_gamma4 = k_exponent * gamma4;
  double exponential2 = exponential * exponential;
  double k_2_exponential = 2.0 * k_exponent * exponential;
  double reftrans_factor = 1.0 / (k_exponent + gamma1 + (k_exponent - gamma1) * exponential2);
  r.ref = gamma2 * (1.0 - exponential2) * reftrans_factor;
  r.trans = dmax(0.0, dmin(k_2_exponential * reftrans_factor, 1.0 - r.ref));
  reftrans_factor = mu0 * ssa * reftrans_factor / (fabs(one_minus_kmu0_sqr) > eps ? one_minus_kmu0_sqr : eps);
  double rd = reftrans_factor * ((1.0 - k_mu0) * (alpha2 + k_gamma3) - (1.0 + k_mu0) * (alpha2 - k_gamma3) * exponential2 -
                                 k_2_exponential * (gamma3 - alpha2 * mu0) * tdd);
  double td = reftrans_factor * (k_2_exponential * (gamma4 + alpha1 * mu0) -
                                 tdd * ((1.0 + k_mu0) * (alpha1 + k_gamma4) - (1.0 - k_mu0) * (alpha1 - k_gamma4) * exponential2));
  rd = dmax(0.0, dmin(rd, mu0 * (1.0 - tdd)));
  td = dmax(0.0, dmin(td, mu0 * (1.0 - tdd) - rd));
  r.ref_dir = rd;
  r.trans_dir_diff = td;
  return r;
}

HD SwLayer sw_ref_trans_cloudless(double mu0, double od, double ssa, double asym) {
  SwLayer r;
  double factor = 0.75 * asym;
  double gamma1 = 2.0 - ssa * (1.25 + factor);
  double gamma2 = ssa * (0.75 - factor);
  double gamma3 = 0.5 - mu0 * factor;
  double gamma4 = 1.0 - gamma3;
  double alpha1 = gamma1 * gamma4 + gamma2 * gamma3;
  double alpha2 = gamma1 * gamma3 + gamma2 * gamma4;
  double k_exponent = sqrt(dmax((gamma1 - gamma2) * (gamma1 + gamma2), 1.0e-12));
  double mu0_local = mu0;
  if (fabs(1.0 - k_exponent * mu0) < 1000.0 * DBL_EPSILON) mu0_local = mu0 * (1.0 - 10.0 * DBL_EPSILON);
  double od_over_mu0 = dmax(od / mu0_local, 0.0);
  double k_mu0 = k_exponent * mu0_local;
  double k_gamma3 = k_exponent * gamma3;
  double k_gamma4 = k_exponent * gamma4;
  double exponential0 = exp(-od_over_mu0);
  r.trans_dir_dir = exponential0;
  double exponential = exp(-k_exponent * od);
  double exponential2 = exponential * exponential;
  double k_2_exponential = 2.0 * k_exponent * exponential;
  double reftrans_factor = 1.0 / (k_exponent + gamma1 + (k_exponent - gamma1) * exponential2);
  r.ref = gamma2 * (1.0 - exponential2) * reftrans_factor;
  r.trans = k_2_exponential * reftrans_factor;
  reftrans_factor = mu0_local * ssa * reftrans_factor / (1.0 - k_mu0 * k_mu0);
  double rd = reftrans_factor * ((1.0 - k_mu0) * (alpha2 + k_gamma3) - (1.0 + k_mu0) * (alpha2 - k_gamma3) * exponential2 -
                                 k_2_exponential * (gamma3 - alpha2 * mu0_local) * exponential0);
  double td = reftrans_factor * (k_2_exponential * (gamma4 + alpha1 * mu0_local) -
                                 exponential0 * ((1.0 + k_mu0) * (alpha1 + k_gamma4) - (1.0 - k_mu0) * (alpha1 - k_gamma4) * exponential2));
  rd = dmax(0.0, dmin(rd, 1.0));
  td = dmax(0.0, dmin(td, 1.0 - rd));
  r.ref_dir = rd;
  r.trans_dir_diff = td;
  return r;
}

}  // namespace ecb
