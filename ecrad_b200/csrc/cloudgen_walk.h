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

enum { GEN_MAXLEV = 256 };

// radiation_cloud_cover.F90:339-623 cum_cloud_cover_exp_exp for one column: cloud "objects" (contiguous layers around a
// local maximum of cloud fraction) are overlapped exponentially within themselves and merged pairwise, most correlated
// pair first.  Same operations in the same order as the reference (the library is built without FMA contraction).
// frac/overlap_param: stride fstride; cum/pair: stride `stride`.  Object indices are 1-based as in the source.
HD void cum_pair_exp_exp(int nlev, int fstride, int stride, const double* frac, const double* overlap_param, bool beta,
                         double* cum, double* pair) {
  const double min_frac = 1.0e-6;
  const double MaxCloudFrac = 1.0 - DBL_EPSILON * 10.0;
  int i_top[GEN_MAXLEV / 2 + 2], i_max[GEN_MAXLEV / 2 + 2], i_base[GEN_MAXLEV / 2 + 2], i_next[GEN_MAXLEV / 2 + 2];
  double cc_obj[GEN_MAXLEV / 2 + 2], alpha_obj[GEN_MAXLEV / 2 + 2];
#define FR(l) frac[(size_t)((l) - 1) * fstride]
#define OP(l) overlap_param[(size_t)((l) - 1) * fstride]
#define CUM(l) cum[(size_t)((l) - 1) * stride]
#define PAIR(l) pair[(size_t)((l) - 1) * stride]
  int jlev = 1, nobj = 0;
  while (jlev <= nlev) {
    if (FR(jlev) > min_frac) {
      ++nobj;
      i_top[nobj] = jlev;
      ++jlev;
      while (jlev <= nlev) { if (FR(jlev) < FR(jlev - 1)) break; ++jlev; }
      i_max[nobj] = jlev - 1;
      while (jlev <= nlev) { if (FR(jlev) > FR(jlev - 1) || FR(jlev) <= min_frac) break; ++jlev; }
      i_base[nobj] = jlev - 1;
      i_next[nobj] = nobj + 1;
    } else ++jlev;
  }
  for (int l = 1; l <= nlev; ++l) CUM(l) = 0.0;
  for (int l = 1; l <= nlev - 1; ++l) PAIR(l) = 0.0;
  if (nobj == 0) return;
  for (int l = 1; l <= nlev - 1; ++l) {
    const double a = beta ? beta2alpha(OP(l), FR(l), FR(l + 1)) : OP(l);
    PAIR(l) = a * dmax(FR(l), FR(l + 1)) + (1.0 - a) * (FR(l) + FR(l + 1) - FR(l) * FR(l + 1));
  }
  for (int jobj = 1; jobj <= nobj - 1; ++jobj) {
    double prod = 1.0;
    for (int l = i_max[jobj]; l <= i_max[jobj + 1] - 1; ++l) prod = prod * (beta ? beta2alpha(OP(l), FR(l), FR(l + 1)) : OP(l));
    alpha_obj[jobj] = prod;
  }
  for (int jobj = 1; jobj <= nobj; ++jobj) {
    CUM(i_top[jobj]) = FR(i_top[jobj]);
    for (int l = i_top[jobj]; l <= i_base[jobj] - 1; ++l) {
      if (FR(l) >= MaxCloudFrac) CUM(l + 1) = 1.0;
      else CUM(l + 1) = 1.0 - (1.0 - CUM(l)) * (1.0 - PAIR(l)) / (1.0 - FR(l));
    }
    cc_obj[jobj] = CUM(i_base[jobj]);
  }
  int iobj1 = 1;
  while (nobj > 1) {
    double alpha_max = 0.0;
    iobj1 = 1;
    int jobj = 1;
    while (jobj < nobj) {
      if (alpha_obj[jobj] > alpha_max) { alpha_max = alpha_obj[jobj]; iobj1 = jobj; }
      jobj = i_next[jobj];
    }
    const int iobj2 = i_next[iobj1];
    for (int l = i_base[iobj1] + 1; l <= i_top[iobj2] - 1; ++l) CUM(l) = CUM(i_base[iobj1]);
    const double cc_pair = alpha_obj[iobj1] * dmax(cc_obj[iobj1], cc_obj[iobj2]) +
                           (1.0 - alpha_obj[iobj1]) * (cc_obj[iobj1] + cc_obj[iobj2] - cc_obj[iobj1] * cc_obj[iobj2]);
    const double scaling = dmin(dmax((cc_pair - cc_obj[iobj1]) / dmax(min_frac, cc_obj[iobj2]), 0.0), 1.0);
    for (int l = i_top[iobj2]; l <= i_base[iobj2]; ++l) CUM(l) = CUM(i_base[iobj1]) + CUM(l) * scaling;
    cc_obj[iobj1] = cc_pair;
    i_base[iobj1] = i_base[iobj2];
    i_next[iobj1] = i_next[iobj2];
    alpha_obj[iobj1] = alpha_obj[iobj2];
    --nobj;
  }
  for (int l = i_base[iobj1] + 1; l <= nlev; ++l) CUM(l) = CUM(i_base[iobj1]);
  for (int l = 1; l <= nlev - 1; ++l) PAIR(l) = dmax(PAIR(l), FR(l) + CUM(l + 1) - CUM(l));
  for (int l = 1; l <= nlev; ++l) CUM(l) = dmin(CUM(l), 1.0);
#undef FR
#undef OP
#undef CUM
#undef PAIR
}

// scheme: 0 = Max-Ran, 1 = Exp-Ran, 2 = Exp-Exp.  Writes cum, pair, opi (strided); returns the total cloud cover
// (0 if below the threshold, radiation_cloud_generator.F90:130-133).
HD double gen_prepare(int scheme, int nlev, int fstride, int stride, const double* frac, const double* overlap_param, bool beta,
                      double decorr_scaling, double frac_threshold, double* cum, double* pair, double* opi,
                      int* ibegin_out, int* iend_out) {
  const double MaxCloudFrac = 1.0 - DBL_EPSILON * 10.0;
  int ibegin = 0, iend = 0;
  if (scheme == 2) {
    cum_pair_exp_exp(nlev, fstride, stride, frac, overlap_param, beta, cum, pair);
    for (int jl = 0; jl < nlev; ++jl)
      if (frac[(size_t)jl * fstride] > 0.0) { if (!ibegin) ibegin = jl + 1; iend = jl + 1; }
  } else {
  double f1 = frac[0];
  double cum_product = sub_rn(1.0, f1);
  cum[0] = f1;
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
                 int32_t* ri1, uint32_t* code, int rowlen, bool exp_exp = false) {
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
    if (exp_exp) {
      // generate_column_exp_exp (radiation_cloud_generator.F90:396-530): one inhomogeneity draw per layer of the whole range
      bool cloudy = true;   // layer itrigger
      out[itrigger - 1] = 1u;
      for (jlev = itrigger + 1; jlev <= c.iend; ++jlev) {
        double r = (double)rcloud[iy] * RM; ++iy;
        double f_prev = c.frac[(size_t)(jlev - 2) * fs], pr = c.pair[(size_t)(jlev - 2) * st];
        if (cloudy) {
          double f_cur = c.frac[(size_t)(jlev - 1) * fs];
          cloudy = mul_rn(r, f_prev) < sub_rn(add_rn(f_cur, f_prev), pr);
        } else {
          double cum_prev = c.cum[(size_t)(jlev - 2) * st];
          double overhang = sub_rn(c.cum[(size_t)(jlev - 1) * st], cum_prev);
          cloudy = mul_rn(r, sub_rn(cum_prev, f_prev)) < sub_rn(sub_rn(pr, overhang), f_prev);
        }
        out[jlev - 1] = cloudy ? 1u : 0u;
      }
      for (int k = 0; k < nrand; ++k) ri1[k] = rs.next_int();
      for (int jc = 1; jc <= nrand; ++jc) {
        double r2 = (double)rs.next_int() * RM;
        if (jc >= 2 && r2 < c.opi[(size_t)(itrigger + jc - 3) * st]) ri1[jc - 1] = ri1[jc - 2];
      }
      for (int k = 0; k < nrand; ++k) out[itrigger - 1 + k] = out[itrigger - 1 + k] ? (0x80000000u | (uint32_t)ri1[k]) : 0u;
      continue;
    }
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
