// cloud_core.h -- cloud optics (liquid: SOCRATES, Slingo / Lindner-Li; ice: Fu, Baran, Baran-2016, Baran-2017, Yi) and the McICA
// stochastic cloud generator.
//
// Reference map:  cloud_optics_layer <- radiation/radiation_cloud_optics.F90:218-523,
//                                       radiation_liquid_optics_socrates.F90:40-80, radiation_liquid_optics_slingo.F90:29-106,
//                                       radiation_ice_optics_fu.F90:42-138, radiation_ice_optics_baran.F90:34-108,
//                                       radiation_ice_optics_baran2017.F90:32-68, radiation_ice_optics_yi.F90:37-145,
//                                       radiation_delta_eddington.h:103-119
//                 RngMix             <- utilities/radiation_random_numbers_mix.F90:142-309 (30-bit lagged Fibonacci)
//                 cum_cloud_cover_*  <- radiation/radiation_cloud_cover.F90:169-300
//                 generate_subcolumn <- radiation/radiation_cloud_generator.F90:262-390
//                 pdf_sample         <- radiation/radiation_pdf_sampler.F90:126-150
//
// Everything that feeds a comparison with a random number uses the *_rn helpers (no FMA contraction) so that the
// sub-column cloud masks are bit-identical to the reference's.
#pragma once
#include <float.h>

#include "hd.h"

namespace ecb {

enum { LIQ_SOCRATES = 1, LIQ_SLINGO = 2 };                                       // config%i_liq_model (radiation_config.F90:108-112)
enum { ICE_FU = 1, ICE_BARAN = 2, ICE_BARAN2016 = 3, ICE_BARAN2017 = 4, ICE_YI = 5 };   // config%i_ice_model (:123-127)
enum { ICE_MAX_COEFF = 69 };                                                     // Yi: 3 x 23 look-up entries per band

struct CloudMeta {
  double liq_lw[16 * 16], liq_sw[14 * 16];  // (nb, ncoeff <= 16) Fortran order: coeff(jb,k) at [(k-1)*nb + jb]
  double ice_lw[16 * ICE_MAX_COEFF], ice_sw[14 * ICE_MAX_COEFF];
  double ice_gen[5];                        // Baran-2017: coeff_gen
  int liq_model, ice_model;
  int pdf_ncdf, pdf_nfsd;
  double pdf_fsd1, pdf_inv_fsd_interval;
  // decoding of the generator's code words into uniform deviates: (code & gen_mask) * gen_scale
  //   default generator: 30-bit numbers of radiation_random_numbers_mix (mask 0x3FFFFFFF, scale 2^-30)
  //   vectorizable generator: MINSTD states 1..2^31-2 of radiation_random_numbers (mask 0x7FFFFFFF, scale 1/(2^31-1))
  unsigned gen_mask; double gen_scale;
};

// Aerosol optics per band (config%aerosol_optics): offsets into one packed device array; phobic (nband, ntype),
// philic (nband, nrh, ntype), Fortran order.
struct AerMeta {
  int ntype, nrh, n_phobic, n_philic;
  int iclass[32], itype[32];      // per aerosol type of the caller: 0 ignored / 1 hydrophobic / 2 hydrophilic; 1-based table type
  double rh_lower[16];
  int me_sw_phobic, ssa_sw_phobic, g_sw_phobic, me_lw_phobic, ssa_lw_phobic, g_lw_phobic;
  int me_sw_philic, ssa_sw_philic, g_sw_philic, me_lw_philic, ssa_lw_philic, g_lw_philic;
};

#define ECB_CO(c, nb, jb, k) ((c)[((k) - 1) * (nb) + (jb)])

// One band of one layer; returns od, scat_od, g of liquid and ice combined as the reference merges them.
// lwp/iwp: in-cloud water paths (kg m-2).  is_sw selects the coefficient set and delta-Eddington on liquid.
struct CloudBandOut { double od, ssa, g; };

HD void liq_socrates(const double* c, int nb, int jb, double lwp, double re_in, double& od, double& scat, double& g) {
  const double MinRe = (double)1.2e-6f, MaxRe = (double)50.0e-6f;  // default-kind literals in the source
  double re = dmax(MinRe, dmin(re_in, MaxRe));
  od = lwp * (ECB_CO(c, nb, jb, 1) + re * (ECB_CO(c, nb, jb, 2) + re * ECB_CO(c, nb, jb, 3))) /
       (1.0 + re * (ECB_CO(c, nb, jb, 4) + re * (ECB_CO(c, nb, jb, 5) + re * ECB_CO(c, nb, jb, 6))));
  scat = od * (1.0 - (ECB_CO(c, nb, jb, 7) + re * (ECB_CO(c, nb, jb, 8) + re * ECB_CO(c, nb, jb, 9))) /
                         (1.0 + re * (ECB_CO(c, nb, jb, 10) + re * ECB_CO(c, nb, jb, 11))));
  g = (ECB_CO(c, nb, jb, 12) + re * (ECB_CO(c, nb, jb, 13) + re * ECB_CO(c, nb, jb, 14))) /
      (1.0 + re * (ECB_CO(c, nb, jb, 15) + re * ECB_CO(c, nb, jb, 16)));
}
HD void ice_fu_sw(const double* c, int nb, int jb, double iwp, double re, double& od, double& scat, double& g) {
  const double MaxG = 1.0 - 10.0 * DBL_EPSILON;
  double de_um = dmin(re, 100.0e-6) * (1.0e6 / 0.64952);
  double inv_de_um = 1.0 / de_um;
  double iwp_gm_2 = iwp * 1000.0;
  od = iwp_gm_2 * (ECB_CO(c, nb, jb, 1) + ECB_CO(c, nb, jb, 2) * inv_de_um);
  scat = od * (1.0 - (ECB_CO(c, nb, jb, 3) + de_um * (ECB_CO(c, nb, jb, 4) + de_um * (ECB_CO(c, nb, jb, 5) + de_um * ECB_CO(c, nb, jb, 6)))));
  g = dmin(ECB_CO(c, nb, jb, 7) + de_um * (ECB_CO(c, nb, jb, 8) + de_um * (ECB_CO(c, nb, jb, 9) + de_um * ECB_CO(c, nb, jb, 10))), MaxG);
}
HD void ice_fu_lw(const double* c, int nb, int jb, double iwp, double re, double& od, double& scat, double& g) {
  const double MaxG = 1.0 - 10.0 * DBL_EPSILON;
  double de_um = dmin(re, 100.0e-6) * (1.0e6 / 0.64952);
  double inv_de_um = 1.0 / de_um;
  double iwp_gm_2 = iwp * 1000.0;
  od = iwp_gm_2 * (ECB_CO(c, nb, jb, 1) + inv_de_um * (ECB_CO(c, nb, jb, 2) + inv_de_um * ECB_CO(c, nb, jb, 3)));
  scat = od - iwp_gm_2 * inv_de_um * (ECB_CO(c, nb, jb, 4) + de_um * (ECB_CO(c, nb, jb, 5) + de_um * (ECB_CO(c, nb, jb, 6) + de_um * ECB_CO(c, nb, jb, 7))));
  g = dmin(ECB_CO(c, nb, jb, 8) + de_um * (ECB_CO(c, nb, jb, 9) + de_um * (ECB_CO(c, nb, jb, 10) + de_um * ECB_CO(c, nb, jb, 11))), MaxG);
}
// radiation_liquid_optics_slingo.F90:29-63 (Slingo 1989, shortwave) and :69-106 (Lindner & Li 2000, longwave)
HD void liq_slingo_sw(const double* c, int nb, int jb, double lwp, double re, double& od, double& scat, double& g) {
  const double lwp_gm_2 = lwp * 1000.0;
  const double re_um = dmin(dmax(4.2, re * 1.0e6), 16.6);
  const double inv_re_um = 1.0 / re_um;
  od = lwp_gm_2 * (ECB_CO(c, nb, jb, 1) + inv_re_um * ECB_CO(c, nb, jb, 2));
  scat = od * (1.0 - ECB_CO(c, nb, jb, 3) - re_um * ECB_CO(c, nb, jb, 4));
  g = ECB_CO(c, nb, jb, 5) + re_um * ECB_CO(c, nb, jb, 6);
}
HD void liq_lindner_li_lw(const double* c, int nb, int jb, double lwp, double re, double& od, double& scat, double& g) {
  const double lwp_gm_2 = lwp * 1000.0;
  const double re_um = dmin(dmax(2.0, re * 1.0e6), 40.0);
  const double inv_re_um = 1.0 / re_um;
  od = lwp_gm_2 * (ECB_CO(c, nb, jb, 1) + re_um * ECB_CO(c, nb, jb, 2) +
                   inv_re_um * (ECB_CO(c, nb, jb, 3) + inv_re_um * (ECB_CO(c, nb, jb, 4) + inv_re_um * ECB_CO(c, nb, jb, 5))));
  scat = od * (1.0 - (ECB_CO(c, nb, jb, 6) + inv_re_um * ECB_CO(c, nb, jb, 7) + re_um * (ECB_CO(c, nb, jb, 8) + re_um * ECB_CO(c, nb, jb, 9))));
  g = ECB_CO(c, nb, jb, 10) + inv_re_um * ECB_CO(c, nb, jb, 11) + re_um * (ECB_CO(c, nb, jb, 12) + re_um * ECB_CO(c, nb, jb, 13));
}
// radiation_ice_optics_baran.F90:34-60: functions of the ice mixing ratio qi (kg/kg)
HD void ice_baran(const double* c, int nb, int jb, double iwp, double qi, double& od, double& scat, double& g) {
  od = iwp * (ECB_CO(c, nb, jb, 1) + ECB_CO(c, nb, jb, 2) / (1.0 + qi * ECB_CO(c, nb, jb, 3)));
  scat = od * (ECB_CO(c, nb, jb, 4) + ECB_CO(c, nb, jb, 5) / (1.0 + qi * ECB_CO(c, nb, jb, 6)));
  g = ECB_CO(c, nb, jb, 7) + ECB_CO(c, nb, jb, 8) / (1.0 + qi * ECB_CO(c, nb, jb, 9));
}
// radiation_ice_optics_baran.F90:66-108 (Baran et al. 2016): functions of qi and the layer temperature
HD void ice_baran2016(const double* c, int nb, int jb, double iwp, double qi, double temperature, double& od, double& scat, double& g) {
  const double T2 = temperature * temperature;
  const double qi_T = (qi < 1.0e-3 ? qi : 1.0e-3) * temperature;
  const double qi_over_T4 = 1.0 / (T2 * T2);
  od = iwp * ECB_CO(c, nb, jb, 1) * qi_over_T4;
  scat = od * (ECB_CO(c, nb, jb, 2) + ECB_CO(c, nb, jb, 3) * qi_T);
  g = ECB_CO(c, nb, jb, 4) + ECB_CO(c, nb, jb, 5) * qi_T;
}
// radiation_ice_optics_baran2017.F90:32-68
HD void ice_baran2017(const double* cg, const double* c, int nb, int jb, double iwp, double qi, double temperature, double& od, double& scat, double& g) {
  const double qi_mod = qi * exp(cg[0] * (temperature - cg[1]));
  const double qi_mod_od = pow(qi_mod, cg[2]), qi_mod_ssa = pow(qi_mod, cg[3]), qi_mod_g = pow(qi_mod, cg[4]);
  od = iwp * (ECB_CO(c, nb, jb, 1) + ECB_CO(c, nb, jb, 2) / (1.0 + qi_mod_od * ECB_CO(c, nb, jb, 3)));
  scat = od * (ECB_CO(c, nb, jb, 4) + ECB_CO(c, nb, jb, 5) / (1.0 + qi_mod_ssa * ECB_CO(c, nb, jb, 6)));
  g = ECB_CO(c, nb, jb, 7) + ECB_CO(c, nb, jb, 8) / (1.0 + qi_mod_g * ECB_CO(c, nb, jb, 9));
}
// radiation_ice_optics_yi.F90:37-89 / :95-145 (Yi et al. 2013): linear interpolation in a 23-entry look-up table over the effective
// diameter (the shortwave and longwave routines are the same expression on their own coefficients)
HD void ice_yi(const double* c, int nb, int jb, double iwp, double re, double& od, double& scat, double& g) {
  const int NSingleCoeffs = 23;
  const double lu_scale = 0.2, lu_offset = 1.0;
  double de_um = re * 2.0e6;
  de_um = dmax(de_um, 10.0);
  de_um = dmin(de_um, 119.99);
  const double iwp_gm_2 = iwp * 1000.0;
  const int lu_idx = (int)floor(de_um * lu_scale - lu_offset);
  const double wts_2 = (de_um * lu_scale - lu_offset) - lu_idx, wts_1 = 1.0 - wts_2;
  od = 0.001 * iwp_gm_2 * (wts_1 * ECB_CO(c, nb, jb, lu_idx) + wts_2 * ECB_CO(c, nb, jb, lu_idx + 1));
  scat = od * (wts_1 * ECB_CO(c, nb, jb, lu_idx + NSingleCoeffs) + wts_2 * ECB_CO(c, nb, jb, lu_idx + NSingleCoeffs + 1));
  g = wts_1 * ECB_CO(c, nb, jb, lu_idx + 2 * NSingleCoeffs) + wts_2 * ECB_CO(c, nb, jb, lu_idx + 2 * NSingleCoeffs + 1);
}
HD void delta_eddington_scat_od(double& od, double& scat, double& g) {
  double f = g * g;
  od = od - scat * f;
  scat = scat * (1.0 - f);
  g = g / (1.0 + g);
}

// Cloud water of one layer as the parameterisations need it: in-cloud water paths (kg m-2), effective radii (m), ice mixing
// ratio (kg/kg) and layer-mean temperature (K; radiation_cloud_optics.F90:397-398)
struct CloudLayerIn { double lwp, iwp, re_liq, re_ice, q_ice, temperature; };

HD void liquid_band(const CloudMeta& C, bool sw, int nb, int jb, const CloudLayerIn& L, double& od, double& scat, double& g) {
  const double* c = sw ? C.liq_sw : C.liq_lw;
  if (C.liq_model == LIQ_SLINGO) { if (sw) liq_slingo_sw(c, nb, jb, L.lwp, L.re_liq, od, scat, g); else liq_lindner_li_lw(c, nb, jb, L.lwp, L.re_liq, od, scat, g); }
  else liq_socrates(c, nb, jb, L.lwp, L.re_liq, od, scat, g);
}
HD void ice_band(const CloudMeta& C, bool sw, int nb, int jb, const CloudLayerIn& L, bool fu_lw_bug, double& od, double& scat, double& g) {
  const double* c = sw ? C.ice_sw : C.ice_lw;
  switch (C.ice_model) {
    case ICE_BARAN: ice_baran(c, nb, jb, L.iwp, L.q_ice, od, scat, g); break;
    case ICE_BARAN2016: ice_baran2016(c, nb, jb, L.iwp, L.q_ice, L.temperature, od, scat, g); break;
    case ICE_BARAN2017: ice_baran2017(C.ice_gen, c, nb, jb, L.iwp, L.q_ice, L.temperature, od, scat, g); break;
    case ICE_YI: ice_yi(c, nb, jb, L.iwp, L.re_ice, od, scat, g); break;
    default:
      if (sw) ice_fu_sw(c, nb, jb, L.iwp, L.re_ice, od, scat, g);
      else { ice_fu_lw(c, nb, jb, L.iwp, L.re_ice, od, scat, g); if (fu_lw_bug) scat = od - scat; }
  }
}

// SW band jb of a cloudy layer (frac > 0): radiation_cloud_optics.F90:325-514
HD CloudBandOut cloud_optics_sw(const CloudMeta& C, int jb, const CloudLayerIn& L, bool delta_scaling_with_gases) {
  double odl = 0, scl = 0, gl = 0, odi = 0, sci = 0, gi = 0;
  if (L.lwp > 0.0) { liquid_band(C, true, 14, jb, L, odl, scl, gl); if (!delta_scaling_with_gases) delta_eddington_scat_od(odl, scl, gl); }
  if (L.iwp > 0.0) { ice_band(C, true, 14, jb, L, false, odi, sci, gi); if (!delta_scaling_with_gases) delta_eddington_scat_od(odi, sci, gi); }
  CloudBandOut o;
  o.od = odl + odi;
  o.g = (gl * scl + gi * sci) / (scl + sci);
  o.ssa = (scl + sci) / (odl + odi);
  return o;
}
// LW band jb of a cloudy layer
HD CloudBandOut cloud_optics_lw(const CloudMeta& C, int jb, const CloudLayerIn& L, bool lw_cloud_scattering, bool fu_lw_bug) {
  double odl = 0, scl = 0, gl = 0, odi = 0, sci = 0, gi = 0;
  if (L.lwp > 0.0) liquid_band(C, false, 16, jb, L, odl, scl, gl);
  if (L.iwp > 0.0) {
    ice_band(C, false, 16, jb, L, fu_lw_bug, odi, sci, gi);
    delta_eddington_scat_od(odi, sci, gi);
  }
  CloudBandOut o;
  if (lw_cloud_scattering) {
    o.od = odl + odi;
    o.g = (scl + sci > 0.0) ? (gl * scl + gi * sci) / (scl + sci) : 0.0;
    o.ssa = (scl + sci) / (odl + odi);
  } else {
    o.od = odl - scl + odi - sci;
    o.g = 0.0; o.ssa = 0.0;
  }
  return o;
}

// ---------------------------------------------------------------------------------------------------------
// utilities/radiation_random_numbers_mix.F90: 30-bit lagged-Fibonacci generator (p=273, q=607)
// ---------------------------------------------------------------------------------------------------------
enum { JPP = 273, JPQ = 607, JPS = 105, JPMM = 30 };
enum { JPNUMSPLIT = (JPQ - 2) / (JPP - 1), JPLENSPLIT = (JPQ - JPP + JPNUMSPLIT - 1) / JPNUMSPLIT };

HD int32_t lfsr_step(int32_t idum, int* top) {
  uint32_t u = (uint32_t)idum;
  *top = (int)((u >> 31) & 1u);
  if (*top) u = ((u ^ 87u) << 1) | 1u;   // IBSET(ISHFT(IEOR(IDUM,87),1),0)
  else u = (u << 1) & ~1u;               // IBCLR(ISHFT(IDUM,1),0)
  return (int32_t)u;
}

// The state array lives wherever the caller puts it (thread-local memory on the GPU); ix is 1-based: ix[1..JPQ].
struct RngMix {
  int32_t* ix;
  int iused;

  HD void refill() {
    const int32_t IVAR = 0x3FFFFFFF;
    for (int jj = 1; jj <= JPP; ++jj) ix[jj] = IVAR & (ix[jj] + ix[jj - JPP + JPQ]);
    for (int jk = 1; jk <= JPNUMSPLIT; ++jk)
      for (int jj = 1 + JPP + (jk - 1) * JPLENSPLIT; jj <= imin((int)JPQ, JPP + jk * JPLENSPLIT); ++jj)
        ix[jj] = IVAR & (ix[jj] + ix[jj - JPP]);
    iused = 0;
  }
  // next number of the stream as the 30-bit integer; uniform deviate = value * 2^-30 (exact in double)
  HD int32_t next_int() {
    if (iused >= JPQ) refill();
    ++iused;
    return ix[iused];
  }
  HD double next() { return (double)next_int() * (1.0 / (double)(1 << JPMM)); }
  HD void skip(int n) {  // discard n numbers
    while (n > 0) {
      if (iused >= JPQ) refill();
      int k = imin(n, JPQ - iused);
      iused += k; n -= k;
    }
  }
  // radiation_random_numbers_mix.F90:142-231 initialize_random_numbers (+ 999-number warm-up)
  HD void init(int32_t kseed) {
    const int32_t JPMASK = 123459876;
    int32_t idum = kseed ^ JPMASK;
    if (idum < 0) idum = (idum == INT32_MIN) ? idum : -idum;
    if (idum == 0) idum = JPMASK;
    int top;
    for (int jj = 1; jj <= 64; ++jj) idum = lfsr_step(idum, &top);
    for (int i = 1; i <= JPQ - 1; ++i) ix[i] = 0;
    ix[2] = (int32_t)(((uint32_t)idum & ((1u << (JPMM - 1)) - 1u)) << 1);
    ix[JPQ] = (int32_t)(((uint32_t)idum >> (JPMM - 1)) & 7u);
    for (int jbit = 1; jbit <= JPMM - 1; ++jbit)
      for (int jj = 3; jj <= JPQ - 1; ++jj) {
        idum = lfsr_step(idum, &top);
        if (top) ix[jj] |= (int32_t)(1u << jbit);
      }
    ix[JPQ - JPS] |= 1;
    iused = JPQ;
    skip(999);
  }
};

// radiation_pdf_sampler.F90:126-150; val is (ncdf, nfsd) Fortran order
HD double pdf_sample(const CloudMeta& C, const double* val, double fsd, double cdf) {
  const int ncdf = C.pdf_ncdf, nfsd = C.pdf_nfsd;
  double wcdf = cdf * (ncdf - 1) + 1.0;
  int icdf = imax(1, imin((int)wcdf, ncdf - 1));
  wcdf = dmax(0.0, dmin(wcdf - icdf, 1.0));
  double wfsd = (fsd - C.pdf_fsd1) * C.pdf_inv_fsd_interval + 1.0;
  int ifsd = imax(1, imin((int)wfsd, nfsd - 1));
  wfsd = dmax(0.0, dmin(wfsd - ifsd, 1.0));
  const double* v = val + (size_t)(ifsd - 1) * ncdf + (icdf - 1);
  return (1.0 - wcdf) * (1.0 - wfsd) * v[0] + (1.0 - wcdf) * wfsd * v[ncdf] + wcdf * (1.0 - wfsd) * v[1] +
         wcdf * wfsd * v[ncdf + 1];
}

}  // namespace ecb
