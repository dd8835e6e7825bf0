// tc_core.h -- Tripleclouds: region properties and overlap matrices of one column (one thread).
//
// Reference map:  tc_region      <- radiation/radiation_regions.F90:35-199 (nreg = 3, gamma PDF)
//                 tc_overlap     <- radiation/radiation_overlap.F90:130-209 (calc_alpha_overlap_matrix), :280-457
// Same operations in the same order as the reference (the library is built without FMA contraction).
#pragma once
#include "hd.h"

namespace ecb {

enum { TC_NREG = 3 };

// region fractions reg[3] and optical-depth scalings ods[3] (ods[0] unused) of one layer
HD void tc_region(double frac, double fsd, double frac_threshold, double* reg, double* ods, bool do_gamma = true) {
  if (!do_gamma) {   // lognormal PDF: two cloudy regions of equal size (radiation_regions.F90:110-126)
    ods[0] = 0.0;
    if (frac < frac_threshold) { reg[0] = 1.0; reg[1] = 0.0; reg[2] = 0.0; ods[1] = 1.0; ods[2] = 1.0; }
    else {
      reg[0] = 1.0 - frac; reg[1] = frac * 0.5; reg[2] = frac * 0.5;
      ods[1] = exp(-sqrt(log(fsd * fsd + 1.0))) / sqrt(fsd * fsd + 1.0);
      ods[2] = 2.0 - ods[1];
    }
    return;
  }
  const double MinGammaODScaling = 0.025, MinLowerFrac = 0.5, MaxLowerFrac = 0.9, FSDAtMinLowerFrac = 1.5, FSDAtMaxLowerFrac = 3.725;
  const double LowerFracFSDGradient = (MaxLowerFrac - MinLowerFrac) / (FSDAtMaxLowerFrac - FSDAtMinLowerFrac);
  const double LowerFracFSDIntercept = MinLowerFrac - FSDAtMinLowerFrac * LowerFracFSDGradient;
  ods[0] = 0.0;
  if (frac < frac_threshold) {
    reg[0] = 1.0; reg[1] = 0.0; reg[2] = 0.0;
    ods[1] = 1.0; ods[2] = 1.0;
  } else {
    reg[0] = 1.0 - frac;
    reg[1] = frac * dmax(MinLowerFrac, dmin(MaxLowerFrac, LowerFracFSDIntercept + fsd * LowerFracFSDGradient));
    ods[1] = MinGammaODScaling + (1.0 - MinGammaODScaling) * exp(-fsd * (1.0 + 0.5 * fsd * (1.0 + 0.5 * fsd)));
    reg[2] = 1.0 - reg[0] - reg[1];
    ods[2] = (frac - reg[1] * ods[1]) / reg[2];
  }
}

// Homogeneous solvers on the Tripleclouds kernels: a cloudy layer (cropped fraction >= threshold) is entirely region 2 with the
// unscaled gridbox-mean cloud (radiation_homogeneous_sw.F90:218-260), a clear layer entirely region 1
HD void tc_region_homogeneous(double frac, double frac_threshold, double* reg, double* ods) {
  const bool cloudy = frac >= frac_threshold;
  reg[0] = cloudy ? 0.0 : 1.0; reg[1] = cloudy ? 1.0 : 0.0; reg[2] = 0.0;
  ods[0] = 0.0; ods[1] = 1.0; ods[2] = 1.0;
}

// overlap matrix M[jupper][jlower] between the region fractions above (fu) and below (fl) an interface
HD void tc_alpha_overlap_matrix(double op, double op_inhom, const double* fu, const double* fl, double M[3][3]) {
  double cf_upper = fu[1] + fu[2], cf_lower = fl[1] + fl[2];
  double pair_cloud_cover = op * dmax(cf_upper, cf_lower) + (1.0 - op) * (cf_upper + cf_lower - cf_upper * cf_lower);
  M[0][0] = 1.0 - pair_cloud_cover;
  double one_over_cf = 1.0 / dmax(cf_lower, 1.0e-6);
  M[0][1] = (pair_cloud_cover - cf_upper) * fl[1] * one_over_cf;
  M[0][2] = (pair_cloud_cover - cf_upper) * fl[2] * one_over_cf;
  one_over_cf = 1.0 / dmax(cf_upper, 1.0e-6);
  M[1][0] = (pair_cloud_cover - cf_lower) * fu[1] * one_over_cf;
  M[2][0] = (pair_cloud_cover - cf_lower) * fu[2] * one_over_cf;
  const double frac_both = cf_upper + cf_lower - pair_cloud_cover;
  cf_upper = fu[2] / dmax(cf_upper, 1.0e-6);
  cf_lower = fl[2] / dmax(cf_lower, 1.0e-6);
  pair_cloud_cover = op_inhom * dmax(cf_upper, cf_lower) + (1.0 - op_inhom) * (cf_upper + cf_lower - cf_upper * cf_lower);
  M[1][1] = frac_both * (1.0 - pair_cloud_cover);
  M[1][2] = frac_both * (pair_cloud_cover - cf_upper);
  M[2][1] = frac_both * (pair_cloud_cover - cf_lower);
  M[2][2] = frac_both * (cf_upper + cf_lower - pair_cloud_cover);
}

// calc_beta_overlap_matrix, radiation_overlap.F90:63-122: one "beta" overlap parameter per region (Shonk et al. 2010)
HD void tc_beta_overlap_matrix(const double* op, const double* fu, const double* fl, double frac_threshold, double M[3][3]) {
  double denominator = 1.0, op_x_frac_min[3];
  for (int r = 0; r < 3; ++r) {
    op_x_frac_min[r] = op[r] * dmin(fu[r], fl[r]);
    denominator = denominator - op_x_frac_min[r];
  }
  if (denominator >= frac_threshold) {
    const double factor = 1.0 / denominator;
    for (int ju = 0; ju < 3; ++ju)
      for (int jw = 0; jw < 3; ++jw) M[ju][jw] = factor * (fl[jw] - op_x_frac_min[jw]) * (fu[ju] - op_x_frac_min[ju]);
  } else {
    for (int ju = 0; ju < 3; ++ju) for (int jw = 0; jw < 3; ++jw) M[ju][jw] = 0.0;
  }
  for (int r = 0; r < 3; ++r) M[r][r] = M[r][r] + op_x_frac_min[r];
}

// One column: reg/ods [nlev][3], U/V [nlev+1][3][3] with U[hl][jupper][jlower] = u_matrix(jupper,jlower,hl+1) and
// V[hl][a][b] = v_matrix(a,b,hl+1); returns the total cloud cover 1 - prod v_matrix(1,1,:).
// frac/fsd/overlap_param are strided by `fstride` (reference layout, column fastest).
HD double tc_prepare_column(int nlev, int fstride, const double* frac, const double* fsd, const double* overlap_param,
                            double decorrelation_scaling, double frac_threshold, double* reg, double* ods, double* U, double* V) {
  double frac_upper[3] = {1.0, 0.0, 0.0}, frac_lower[3], M[3][3];
  const double expo = 1.0 / decorrelation_scaling;
  double prod = 1.0;
  for (int jlev = 1; jlev <= nlev + 1; ++jlev) {
    if (jlev > nlev) { frac_lower[0] = 1.0; frac_lower[1] = 0.0; frac_lower[2] = 0.0; }
    else {
      tc_region(frac[(size_t)(jlev - 1) * fstride], fsd[(size_t)(jlev - 1) * fstride], frac_threshold, reg + (jlev - 1) * 3, ods + (jlev - 1) * 3);
      for (int r = 0; r < 3; ++r) frac_lower[r] = reg[(jlev - 1) * 3 + r];
    }
    double op1, op2;
    if (jlev == 1 || jlev > nlev) { op1 = 1.0; op2 = 1.0; }
    else {
      op1 = overlap_param[(size_t)(jlev - 2) * fstride];
      op2 = op1 >= 0.0 ? (expo == 2.0 ? mul_rn(op1, op1) : pow(op1, expo)) : op1;
    }
    tc_alpha_overlap_matrix(op1, op2, frac_upper, frac_lower, M);
    double* u = U + (jlev - 1) * 9; double* v = V + (jlev - 1) * 9;
    for (int ju = 0; ju < 3; ++ju)
      for (int jw = 0; jw < 3; ++jw) {
        u[ju * 3 + jw] = frac_lower[jw] >= frac_threshold ? M[ju][jw] / frac_lower[jw] : 0.0;
        v[jw * 3 + ju] = frac_upper[ju] >= frac_threshold ? M[ju][jw] / frac_upper[ju] : 0.0;
      }
    prod = prod * v[0];
    for (int r = 0; r < 3; ++r) frac_upper[r] = frac_lower[r];
  }
  return 1.0 - prod;
}

}  // namespace ecb
