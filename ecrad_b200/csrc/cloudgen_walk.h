// cloudgen_walk.h -- McICA stochastic cloud generator for one column (one thread).
//
// Reference map:  gen_prepare <- radiation/radiation_cloud_cover.F90:53-69 (beta2alpha), :169-225 (Max-Ran),
//                                :231-300 (Exp-Ran); radiation_cloud_generator.F90:120-200 (ibegin/iend, overhang,
//                                overlap_param_inhom)
//                 gen_walk    <- radiation/radiation_cloud_generator.F90:202-255 (seed, cloud-top triggers) and
//                                :262-390 generate_column_exp_ran
// The generator's output is NOT the optical-depth scaling itself but, per (g-point, layer), the 30-bit random
// integer that selects it (bit 31 = "cloudy"); the consumer evaluates the PDF look-up (pdf_sample) lane-parallel.
// All arithmetic that feeds a comparison with a random number is done with the non-contracting *_rn helpers, in the
// reference's operation order, so the cloud masks are bit-identical.
#pragma once
#include <float.h>

#include "cloud_core.h"

namespace ecb {

struct GenColumn {
  int nlev, fstride, stride, ibegin, iend;  // element l: frac[l*fstride]; cum/pair/opi[l*stride]; ibegin/iend 1-based
  const double *frac, *cum, *pair, *opi;
};

HD double beta2alpha(double beta, double f1, double f2) {
  if (beta < 1.0) {
    double d = fabs(sub_rn(f1, f2));
    return add_rn(beta, div_rn(mul_rn(sub_rn(1.0, beta), d), sub_rn(add_rn(d, div_rn(1.0, beta)), 1.0)));
  }
  return 1.0;
}

// scheme: 0 = Max-Ran, 1 = Exp-Ran.  Writes cum, pair, opi (strided); returns the total cloud cover
// (0 if below the threshold, radiation_cloud_generator.F90:130-133).
HD double gen_prepare(int scheme, int nlev, int fstride, int stride, const double* frac, const double* overlap_param, bool beta,
                      double decorr_scaling, double frac_threshold, double* cum, double* pair, double* opi,
                      int* ibegin_out, int* iend_out) {
  const double MaxCloudFrac = 1.0 - DBL_EPSILON * 10.0;
  double f1 = frac[0];
  double cum_product = sub_rn(1.0, f1);
  cum[0] = f1;
  int ibegin = 0, iend = 0;
  if (f1 > 0.0) { ibegin = 1; iend = 1; }
  for (int jl = 0; jl < nlev - 1; ++jl) {
    double f2 = frac[(size_t)(jl + 1) * fstride];
    double pr;
    if (scheme == 1) {
      double op = overlap_param[(size_t)jl * fstride];
      double alpha = beta ? beta2alpha(op, f1, f2) : op;
      pr = add_rn(mul_rn(alpha, dmax(f1, f2)), mul_rn(sub_rn(1.0, alpha), sub_rn(add_rn(f1, f2), mul_rn(f1, f2))));
    } else {
      pr = dmax(f1, f2);
    }
    if (f1 >= MaxCloudFrac) cum_product = 0.0;
    else cum_product = div_rn(mul_rn(cum_product, sub_rn(1.0, pr)), sub_rn(1.0, f1));
    cum[(size_t)(jl + 1) * stride] = sub_rn(1.0, cum_product);
    pair[(size_t)jl * stride] = pr;
    if (f2 > 0.0) { if (!ibegin) ibegin = jl + 2; iend = jl + 2; }
    f1 = f2;
  }
  double tcc = cum[(size_t)(nlev - 1) * stride];
  *ibegin_out = ibegin; *iend_out = iend;
  if (tcc < frac_threshold || !ibegin) return 0.0;
  const double expo = 1.0 / decorr_scaling;
  for (int jl = 0; jl < nlev - 1; ++jl) {
    double op = overlap_param[(size_t)jl * fstride];
    if (jl + 1 >= ibegin && jl + 1 <= iend - 1 && op > 0.0)
      op = (expo == 2.0) ? mul_rn(op, op) : pow(op, expo);   // overlap_param ** (1/decorrelation_scaling)
    opi[(size_t)jl * stride] = op;
  }
  return tcc;
}

// rtop[ng], rcloud[nlev], ri1[nlev]: thread-private integer work arrays.  code: [ng][rowlen] for this column,
// pre-zeroed; entry = 0x80000000 | rand30 for cloudy (g, layer).
HD void gen_walk(const GenColumn& c, RngMix& rs, int32_t iseed, int ng, double tcc, int32_t* rtop, int32_t* rcloud,
                 int32_t* ri1, uint32_t* code, int rowlen) {
  const double RM = 1.0 / 1073741824.0;  // 2^-30
  const size_t st = (size_t)c.stride, fs = (size_t)c.fstride;
  rs.init(iseed);
  for (int g = 0; g < ng; ++g) rtop[g] = rs.next_int();
  for (int g = 0; g < ng; ++g) {
    double trigger = mul_rn((double)rtop[g] * RM, tcc);
    int jlev = c.ibegin;
    while (trigger > c.cum[(size_t)(jlev - 1) * st] && jlev < c.iend) ++jlev;
    const int itrigger = jlev;
    const int nrand = c.iend + 1 - itrigger;
    for (int i = 0; i < nrand; ++i) rcloud[i] = rs.next_int();
    int n = 1, iy = 0;
    uint32_t* out = code + (size_t)g * rowlen;
    for (jlev = itrigger + 1; jlev <= c.iend + 1; ++jlev) {
      bool fill = false;
      if (jlev <= c.iend) {
        double r = (double)rcloud[iy] * RM; ++iy;
        double f_prev = c.frac[(size_t)(jlev - 2) * fs], pr = c.pair[(size_t)(jlev - 2) * st];
        if (n > 0) {
          double f_cur = c.frac[(size_t)(jlev - 1) * fs];
          if (mul_rn(r, f_prev) < sub_rn(add_rn(f_cur, f_prev), pr)) ++n; else fill = true;
        } else {
          double cum_prev = c.cum[(size_t)(jlev - 2) * st];
          double overhang = sub_rn(c.cum[(size_t)(jlev - 1) * st], cum_prev);
          if (mul_rn(r, sub_rn(cum_prev, f_prev)) < sub_rn(sub_rn(pr, overhang), f_prev)) n = 1;
        }
      } else fill = true;
      if (fill) {
        for (int k = 0; k < n; ++k) ri1[k] = rs.next_int();
        for (int jc = 1; jc <= n; ++jc) {
          double r2 = (double)rs.next_int() * RM;
          if (jc >= 2 && r2 < c.opi[(size_t)(jlev - n + jc - 3) * st]) ri1[jc - 1] = ri1[jc - 2];
        }
        for (int k = 0; k < n; ++k) out[jlev - n + k - 1] = 0x80000000u | (uint32_t)ri1[k];
        n = 0;
      }
    }
  }
}

}  // namespace ecb
