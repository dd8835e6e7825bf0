/* cloud.c -- oracle restatement of cloud optics and the McICA stochastic cloud generator.  TEST INFRASTRUCTURE.
 * Follows radiation/radiation_cloud_optics.F90:218-523, radiation_liquid_optics_socrates.F90:40-80,
 * radiation_liquid_optics_slingo.F90:29-106, radiation_ice_optics_fu.F90:42-138, radiation_ice_optics_baran.F90:34-108,
 * radiation_ice_optics_baran2017.F90:32-68, radiation_ice_optics_yi.F90:37-145, radiation_delta_eddington.h:103-119, radiation_cloud_generator.F90:37-390,
 * radiation_cloud_cover.F90:231-300, radiation_pdf_sampler.F90:126-150,
 * utilities/radiation_random_numbers_mix.F90:142-309.
 */
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

static inline double dmin(double a, double b) { return a < b ? a : b; }
static inline double dmax(double a, double b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* coeff(jb, k) of a (nb, ncoeff) Fortran array, jb 0-based, k 1-based */
#define CO(c, nb, jb, k) ((c)[((k) - 1) * (nb) + (jb)])

/* radiation_liquid_optics_socrates.F90:40-80 */
static void liq_socrates(int nb, const double* c, double lwp, double re_in, double* od, double* scat_od, double* g) {
  const double MinRe = (double)1.2e-6f, MaxRe = (double)50.0e-6f; /* default-kind literals in the source :31-32 */
  double re = dmax(MinRe, dmin(re_in, MaxRe));
  for (int jb = 0; jb < nb; ++jb) {
    od[jb] = lwp * (CO(c, nb, jb, 1) + re * (CO(c, nb, jb, 2) + re * CO(c, nb, jb, 3))) /
             (1.0 + re * (CO(c, nb, jb, 4) + re * (CO(c, nb, jb, 5) + re * CO(c, nb, jb, 6))));
    scat_od[jb] = od[jb] * (1.0 - (CO(c, nb, jb, 7) + re * (CO(c, nb, jb, 8) + re * CO(c, nb, jb, 9))) /
                                      (1.0 + re * (CO(c, nb, jb, 10) + re * CO(c, nb, jb, 11))));
    g[jb] = (CO(c, nb, jb, 12) + re * (CO(c, nb, jb, 13) + re * CO(c, nb, jb, 14))) /
            (1.0 + re * (CO(c, nb, jb, 15) + re * CO(c, nb, jb, 16)));
  }
}
static const double MaxAsymmetryFactor = 1.0 - 10.0 * DBL_EPSILON; /* radiation_ice_optics_fu.F90:33 */
/* radiation_ice_optics_fu.F90:42-84 */
static void ice_fu_sw(int nb, const double* c, double iwp, double re, double* od, double* scat_od, double* g) {
  double de_um = dmin(re, 100.0e-6) * (1.0e6 / 0.64952);
  double inv_de_um = 1.0 / de_um;
  double iwp_gm_2 = iwp * 1000.0;
  for (int jb = 0; jb < nb; ++jb) {
    od[jb] = iwp_gm_2 * (CO(c, nb, jb, 1) + CO(c, nb, jb, 2) * inv_de_um);
    scat_od[jb] = od[jb] * (1.0 - (CO(c, nb, jb, 3) + de_um * (CO(c, nb, jb, 4) + de_um * (CO(c, nb, jb, 5) + de_um * CO(c, nb, jb, 6)))));
    g[jb] = dmin(CO(c, nb, jb, 7) + de_um * (CO(c, nb, jb, 8) + de_um * (CO(c, nb, jb, 9) + de_um * CO(c, nb, jb, 10))), MaxAsymmetryFactor);
  }
}
/* radiation_ice_optics_fu.F90:90-138 */
static void ice_fu_lw(int nb, const double* c, double iwp, double re, double* od, double* scat_od, double* g) {
  double de_um = dmin(re, 100.0e-6) * (1.0e6 / 0.64952);
  double inv_de_um = 1.0 / de_um;
  double iwp_gm_2 = iwp * 1000.0;
  for (int jb = 0; jb < nb; ++jb) {
    od[jb] = iwp_gm_2 * (CO(c, nb, jb, 1) + inv_de_um * (CO(c, nb, jb, 2) + inv_de_um * CO(c, nb, jb, 3)));
    scat_od[jb] = od[jb] - iwp_gm_2 * inv_de_um * (CO(c, nb, jb, 4) + de_um * (CO(c, nb, jb, 5) + de_um * (CO(c, nb, jb, 6) + de_um * CO(c, nb, jb, 7))));
    g[jb] = dmin(CO(c, nb, jb, 8) + de_um * (CO(c, nb, jb, 9) + de_um * (CO(c, nb, jb, 10) + de_um * CO(c, nb, jb, 11))), MaxAsymmetryFactor);
  }
}
/* radiation_liquid_optics_slingo.F90:29-63: Slingo (1989), shortwave */
static void liq_slingo(int nb, const double* c, double lwp, double re, double* od, double* scat_od, double* g) {
  double lwp_gm_2 = lwp * 1000.0;
  double re_um = dmin(dmax(4.2, re * 1.0e6), 16.6); /* range of validity 4.2-16.6 microns */
  double inv_re_um = 1.0 / re_um;
  for (int jb = 0; jb < nb; ++jb) {
    od[jb] = lwp_gm_2 * (CO(c, nb, jb, 1) + inv_re_um * CO(c, nb, jb, 2));
    scat_od[jb] = od[jb] * (1.0 - CO(c, nb, jb, 3) - re_um * CO(c, nb, jb, 4));
    g[jb] = CO(c, nb, jb, 5) + re_um * CO(c, nb, jb, 6);
  }
}
/* radiation_liquid_optics_slingo.F90:69-106: Lindner & Li (2000), longwave */
static void liq_lindner_li(int nb, const double* c, double lwp, double re, double* od, double* scat_od, double* g) {
  double lwp_gm_2 = lwp * 1000.0;
  double re_um = dmin(dmax(2.0, re * 1.0e6), 40.0); /* range of validity 2-40 microns */
  double inv_re_um = 1.0 / re_um;
  for (int jb = 0; jb < nb; ++jb) {
    od[jb] = lwp_gm_2 * (CO(c, nb, jb, 1) + re_um * CO(c, nb, jb, 2) + inv_re_um * (CO(c, nb, jb, 3) + inv_re_um * (CO(c, nb, jb, 4) + inv_re_um * CO(c, nb, jb, 5))));
    scat_od[jb] = od[jb] * (1.0 - (CO(c, nb, jb, 6) + inv_re_um * CO(c, nb, jb, 7) + re_um * (CO(c, nb, jb, 8) + re_um * CO(c, nb, jb, 9))));
    g[jb] = CO(c, nb, jb, 10) + inv_re_um * CO(c, nb, jb, 11) + re_um * (CO(c, nb, jb, 12) + re_um * CO(c, nb, jb, 13));
  }
}
/* radiation_ice_optics_baran.F90:34-60 */
static void ice_baran(int nb, const double* c, double ice_wp, double qi, double* od, double* scat_od, double* g) {
  for (int jb = 0; jb < nb; ++jb) {
    od[jb] = ice_wp * (CO(c, nb, jb, 1) + CO(c, nb, jb, 2) / (1.0 + qi * CO(c, nb, jb, 3)));
    scat_od[jb] = od[jb] * (CO(c, nb, jb, 4) + CO(c, nb, jb, 5) / (1.0 + qi * CO(c, nb, jb, 6)));
    g[jb] = CO(c, nb, jb, 7) + CO(c, nb, jb, 8) / (1.0 + qi * CO(c, nb, jb, 9));
  }
}
/* radiation_ice_optics_baran.F90:66-108 */
static void ice_baran2016(int nb, const double* c, double ice_wp, double qi, double temperature, double* od, double* scat_od, double* g) {
  double T2 = temperature * temperature, qi_T, qi_over_T4;
  if (qi < 1.0e-3) { qi_T = qi * temperature; qi_over_T4 = 1.0 / (T2 * T2); }
  else { qi_T = 1.0e-3 * temperature; qi_over_T4 = 1.0 / (T2 * T2); }
  for (int jb = 0; jb < nb; ++jb) {
    od[jb] = ice_wp * CO(c, nb, jb, 1) * qi_over_T4;
    scat_od[jb] = od[jb] * (CO(c, nb, jb, 2) + CO(c, nb, jb, 3) * qi_T);
    g[jb] = CO(c, nb, jb, 4) + CO(c, nb, jb, 5) * qi_T;
  }
}
/* radiation_ice_optics_baran2017.F90:32-68 */
static void ice_baran2017(int nb, const double* coeff_gen, const double* c, double ice_wp, double qi, double temperature, double* od, double* scat_od, double* g) {
  double qi_mod = qi * exp(coeff_gen[0] * (temperature - coeff_gen[1]));
  double qi_mod_od = pow(qi_mod, coeff_gen[2]), qi_mod_ssa = pow(qi_mod, coeff_gen[3]), qi_mod_g = pow(qi_mod, coeff_gen[4]);
  for (int jb = 0; jb < nb; ++jb) {
    od[jb] = ice_wp * (CO(c, nb, jb, 1) + CO(c, nb, jb, 2) / (1.0 + qi_mod_od * CO(c, nb, jb, 3)));
    scat_od[jb] = od[jb] * (CO(c, nb, jb, 4) + CO(c, nb, jb, 5) / (1.0 + qi_mod_ssa * CO(c, nb, jb, 6)));
    g[jb] = CO(c, nb, jb, 7) + CO(c, nb, jb, 8) / (1.0 + qi_mod_g * CO(c, nb, jb, 9));
  }
}
/* radiation_ice_optics_yi.F90:37-89 (shortwave) and :95-145 (longwave): the same statements on either coefficient set */
static void ice_yi(int nb, const double* c, double ice_wp, double re, double* od, double* scat_od, double* g) {
  const int NSingleCoeffs = 23;
  const double lu_scale = 0.2, lu_offset = 1.0;
  double de_um = re * 2.0e6;
  de_um = dmax(de_um, 10.0);
  de_um = dmin(de_um, 119.99);
  double iwp_gm_2 = ice_wp * 1000.0;
  int lu_idx = (int)floor(de_um * lu_scale - lu_offset);
  double wts_2 = (de_um * lu_scale - lu_offset) - lu_idx;
  double wts_1 = 1.0 - wts_2;
  for (int jb = 0; jb < nb; ++jb) {
    od[jb] = 0.001 * iwp_gm_2 * (wts_1 * CO(c, nb, jb, lu_idx) + wts_2 * CO(c, nb, jb, lu_idx + 1));
    scat_od[jb] = od[jb] * (wts_1 * CO(c, nb, jb, lu_idx + NSingleCoeffs) + wts_2 * CO(c, nb, jb, lu_idx + NSingleCoeffs + 1));
    g[jb] = wts_1 * CO(c, nb, jb, lu_idx + 2 * NSingleCoeffs) + wts_2 * CO(c, nb, jb, lu_idx + 2 * NSingleCoeffs + 1);
  }
}
/* radiation_delta_eddington.h:103-119 */
static void delta_eddington_scat_od(int n, double* od, double* scat_od, double* g) {
  for (int i = 0; i < n; ++i) {
    double f = g[i] * g[i];
    od[i] = od[i] - scat_od[i] * f;
    scat_od[i] = scat_od[i] * (1.0 - f);
    g[i] = g[i] / (1.0 + g[i]);
  }
}

/* coefficient array of the configured model: "<base>.<tag>" of the stand-alone blob, else <base> (what a host model registered),
 * checked against the coefficient count radiation_cloud_optics.F90:46-216 expects */
static const double* model_coeff(const orc_tables* t, const char* base, const char* tag, int nb, int ncoef) {
  char nm[64];
  const orc_array* a = NULL;
  if (tag[0]) { snprintf(nm, sizeof nm, "%s.%s", base, tag); a = orc_find(t, nm); }
  if (!a) a = orc_find(t, base);
  if (!a || a->dtype != 0) return NULL;
  int64_t n = 1;
  for (int k = 0; k < a->ndim; ++k) n *= a->dims[k];
  return n == (int64_t)nb * ncoef ? (const double*)a->data : NULL;
}

/* radiation_cloud_optics.F90:218-523 for one column.  Outputs [nlev][nb]. */
int orc_cloud_optics(const orc_tables* t, const ecrad_b200_config* cfg, int nlev, const double* p_hl, const double* t_hl,
                     const double* frac, const double* q_liq, const double* q_ice, const double* re_liq,
                     const double* re_ice, double* od_lw, double* ssa_lw, double* g_lw,
                     double* od_sw, double* ssa_sw, double* g_sw) {
  const double AccelDueToGravity = 9.80665;
  static const char* liq_tag[] = {"", "", "slingo"};
  static const int liq_n[][2] = {{0, 0}, {16, 16}, {13, 6}};
  static const char* ice_tag[] = {"", "", "baran", "baran2016", "baran2017", "yi"};
  static const int ice_n[][2] = {{0, 0}, {11, 10}, {9, 9}, {5, 5}, {9, 9}, {69, 69}};
  const int lm = cfg->i_liq_model, im = cfg->i_ice_model;
  if (lm < ECRAD_LIQ_SOCRATES || lm > ECRAD_LIQ_SLINGO || im < ECRAD_ICE_FU || im > ECRAD_ICE_YI) return -1;
  const double* liq_coeff_lw = model_coeff(t, "liq_coeff_lw", liq_tag[lm], NB_LW, liq_n[lm][0]);
  const double* liq_coeff_sw = model_coeff(t, "liq_coeff_sw", liq_tag[lm], NB_SW, liq_n[lm][1]);
  const double* ice_coeff_lw = model_coeff(t, "ice_coeff_lw", ice_tag[im], NB_LW, ice_n[im][0]);
  const double* ice_coeff_sw = model_coeff(t, "ice_coeff_sw", ice_tag[im], NB_SW, ice_n[im][1]);
  const double* ice_coeff_gen = im == ECRAD_ICE_BARAN2017 ? model_coeff(t, "ice_coeff_gen", ice_tag[im], 5, 1) : NULL;
  if (!liq_coeff_lw || !liq_coeff_sw || !ice_coeff_lw || !ice_coeff_sw || (im == ECRAD_ICE_BARAN2017 && !ice_coeff_gen)) return -1;
  memset(od_lw, 0, sizeof(double) * (size_t)nlev * NB_LW);
  memset(ssa_lw, 0, sizeof(double) * (size_t)nlev * NB_LW);
  memset(g_lw, 0, sizeof(double) * (size_t)nlev * NB_LW);
  memset(od_sw, 0, sizeof(double) * (size_t)nlev * NB_SW);
  memset(ssa_sw, 0, sizeof(double) * (size_t)nlev * NB_SW);
  memset(g_sw, 0, sizeof(double) * (size_t)nlev * NB_SW);
  for (int jl = 0; jl < nlev; ++jl) {
    if (!(frac[jl] > 0.0)) continue;
    double od_lw_liq[NB_LW], scat_lw_liq[NB_LW], g_lw_liq[NB_LW], od_lw_ice[NB_LW], scat_lw_ice[NB_LW], g_lw_ice[NB_LW];
    double od_sw_liq[NB_SW], scat_sw_liq[NB_SW], g_sw_liq[NB_SW], od_sw_ice[NB_SW], scat_sw_ice[NB_SW], g_sw_ice[NB_SW];
    /* config%is_homogeneous (radiation_cloud_optics.F90:318-327): the Homogeneous solvers take gridbox-mean water paths */
    const int is_homogeneous = (cfg->do_sw && cfg->i_solver_sw == ECRAD_SOLVER_HOMOGENEOUS) || (cfg->do_lw && cfg->i_solver_lw == ECRAD_SOLVER_HOMOGENEOUS);
    double factor = is_homogeneous ? (p_hl[jl + 1] - p_hl[jl]) / AccelDueToGravity : (p_hl[jl + 1] - p_hl[jl]) / (AccelDueToGravity * frac[jl]);
    double lwp = factor * q_liq[jl], iwp = factor * q_ice[jl];
    if (lwp > 0.0) {
      if (lm == ECRAD_LIQ_SOCRATES) {
        liq_socrates(NB_LW, liq_coeff_lw, lwp, re_liq[jl], od_lw_liq, scat_lw_liq, g_lw_liq);
        liq_socrates(NB_SW, liq_coeff_sw, lwp, re_liq[jl], od_sw_liq, scat_sw_liq, g_sw_liq);
      } else {   /* Slingo */
        liq_lindner_li(NB_LW, liq_coeff_lw, lwp, re_liq[jl], od_lw_liq, scat_lw_liq, g_lw_liq);
        liq_slingo(NB_SW, liq_coeff_sw, lwp, re_liq[jl], od_sw_liq, scat_sw_liq, g_sw_liq);
      }
      if (!cfg->do_sw_delta_scaling_with_gases) delta_eddington_scat_od(NB_SW, od_sw_liq, scat_sw_liq, g_sw_liq);
    } else {
      memset(od_lw_liq, 0, sizeof od_lw_liq); memset(scat_lw_liq, 0, sizeof scat_lw_liq); memset(g_lw_liq, 0, sizeof g_lw_liq);
      memset(od_sw_liq, 0, sizeof od_sw_liq); memset(scat_sw_liq, 0, sizeof scat_sw_liq); memset(g_sw_liq, 0, sizeof g_sw_liq);
    }
    if (iwp > 0.0) {
      const double temperature = 0.5 * (t_hl[jl] + t_hl[jl + 1]);
      if (im == ECRAD_ICE_BARAN) {
        ice_baran(NB_LW, ice_coeff_lw, iwp, q_ice[jl], od_lw_ice, scat_lw_ice, g_lw_ice);
        ice_baran(NB_SW, ice_coeff_sw, iwp, q_ice[jl], od_sw_ice, scat_sw_ice, g_sw_ice);
      } else if (im == ECRAD_ICE_BARAN2016) {
        ice_baran2016(NB_LW, ice_coeff_lw, iwp, q_ice[jl], temperature, od_lw_ice, scat_lw_ice, g_lw_ice);
        ice_baran2016(NB_SW, ice_coeff_sw, iwp, q_ice[jl], temperature, od_sw_ice, scat_sw_ice, g_sw_ice);
      } else if (im == ECRAD_ICE_BARAN2017) {
        ice_baran2017(NB_LW, ice_coeff_gen, ice_coeff_lw, iwp, q_ice[jl], temperature, od_lw_ice, scat_lw_ice, g_lw_ice);
        ice_baran2017(NB_SW, ice_coeff_gen, ice_coeff_sw, iwp, q_ice[jl], temperature, od_sw_ice, scat_sw_ice, g_sw_ice);
      } else if (im == ECRAD_ICE_FU) {
        ice_fu_lw(NB_LW, ice_coeff_lw, iwp, re_ice[jl], od_lw_ice, scat_lw_ice, g_lw_ice);
        if (cfg->do_fu_lw_ice_optics_bug) for (int b = 0; b < NB_LW; ++b) scat_lw_ice[b] = od_lw_ice[b] - scat_lw_ice[b];
        ice_fu_sw(NB_SW, ice_coeff_sw, iwp, re_ice[jl], od_sw_ice, scat_sw_ice, g_sw_ice);
      } else {   /* Yi */
        ice_yi(NB_LW, ice_coeff_lw, iwp, re_ice[jl], od_lw_ice, scat_lw_ice, g_lw_ice);
        ice_yi(NB_SW, ice_coeff_sw, iwp, re_ice[jl], od_sw_ice, scat_sw_ice, g_sw_ice);
      }
      if (!cfg->do_sw_delta_scaling_with_gases) delta_eddington_scat_od(NB_SW, od_sw_ice, scat_sw_ice, g_sw_ice);
      delta_eddington_scat_od(NB_LW, od_lw_ice, scat_lw_ice, g_lw_ice);
    } else {
      memset(od_lw_ice, 0, sizeof od_lw_ice); memset(scat_lw_ice, 0, sizeof scat_lw_ice); memset(g_lw_ice, 0, sizeof g_lw_ice);
      memset(od_sw_ice, 0, sizeof od_sw_ice); memset(scat_sw_ice, 0, sizeof scat_sw_ice); memset(g_sw_ice, 0, sizeof g_sw_ice);
    }
    if (cfg->do_lw_cloud_scattering) {
      for (int b = 0; b < NB_LW; ++b) {
        od_lw[jl * NB_LW + b] = od_lw_liq[b] + od_lw_ice[b];
        if (scat_lw_liq[b] + scat_lw_ice[b] > 0.0)
          g_lw[jl * NB_LW + b] = (g_lw_liq[b] * scat_lw_liq[b] + g_lw_ice[b] * scat_lw_ice[b]) / (scat_lw_liq[b] + scat_lw_ice[b]);
        else g_lw[jl * NB_LW + b] = 0.0;
        ssa_lw[jl * NB_LW + b] = (scat_lw_liq[b] + scat_lw_ice[b]) / (od_lw_liq[b] + od_lw_ice[b]);
      }
    } else {
      for (int b = 0; b < NB_LW; ++b) od_lw[jl * NB_LW + b] = od_lw_liq[b] - scat_lw_liq[b] + od_lw_ice[b] - scat_lw_ice[b];
    }
    for (int b = 0; b < NB_SW; ++b) {
      od_sw[jl * NB_SW + b] = od_sw_liq[b] + od_sw_ice[b];
      g_sw[jl * NB_SW + b] = (g_sw_liq[b] * scat_sw_liq[b] + g_sw_ice[b] * scat_sw_ice[b]) / (scat_sw_liq[b] + scat_sw_ice[b]);
      ssa_sw[jl * NB_SW + b] = (scat_sw_liq[b] + scat_sw_ice[b]) / (od_sw_liq[b] + od_sw_ice[b]);
    }
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------------------
 * utilities/radiation_random_numbers_mix.F90: 30-bit lagged-Fibonacci generator (p=273, q=607)
 * ------------------------------------------------------------------------------------------------- */
#define JPP 273
#define JPQ 607
#define JPS 105
#define JPMM 30
#define JPNUMSPLIT ((JPQ - 2) / (JPP - 1))
#define JPLENSPLIT ((JPQ - JPP + JPNUMSPLIT - 1) / JPNUMSPLIT)
typedef struct { int iused; int32_t ix[JPQ + 1]; double zrm; } rng_stream; /* ix[1..JPQ] */

static inline int32_t lfsr_step(int32_t idum, int* top) {
  uint32_t u = (uint32_t)idum;
  *top = (u >> 31) & 1u;
  if (*top) u = ((u ^ 87u) << 1) | 1u;  /* IBSET(ISHFT(IEOR(IDUM,87),1),0) */
  else u = (u << 1) & ~1u;                /* IBCLR(ISHFT(IDUM,1),0)          */
  return (int32_t)u;
}
static void rng_uniform(rng_stream* s, int n, double* px);

/* radiation_random_numbers_mix.F90:142-231 */
static void rng_init(int32_t kseed, rng_stream* s) {
  const int32_t JPMASK = 123459876;
  int32_t idum = kseed ^ JPMASK;
  if (idum < 0) idum = (idum == INT32_MIN) ? idum : -idum; /* ABS */
  if (idum == 0) idum = JPMASK;
  int top;
  for (int jj = 1; jj <= 64; ++jj) idum = lfsr_step(idum, &top);
  for (int i = 1; i <= JPQ - 1; ++i) s->ix[i] = 0;
  s->ix[2] = (int32_t)(((uint32_t)idum & ((1u << (JPMM - 1)) - 1u)) << 1);   /* ISHFT(IBITS(IDUM,0,JPMM-1),1) */
  s->ix[JPQ] = (int32_t)(((uint32_t)idum >> (JPMM - 1)) & 7u);               /* IBITS(IDUM,JPMM-1,3)          */
  for (int jbit = 1; jbit <= JPMM - 1; ++jbit) {
    for (int jj = 3; jj <= JPQ - 1; ++jj) {
      idum = lfsr_step(idum, &top);
      if (top) s->ix[jj] |= (int32_t)(1u << jbit);
    }
  }
  s->ix[JPQ - JPS] |= 1;
  s->iused = JPQ;
  s->zrm = 1.0 / (double)(1 << JPMM);
  double warm[999];
  rng_uniform(s, 999, warm);
}

/* radiation_random_numbers_mix.F90:261-309 */
static void rng_uniform(rng_stream* s, int n, double* px) {
  const int32_t IVAR = 0x3FFFFFFF;
  int ifilled = 0;
  int hi = imin(JPQ, n + s->iused);
  for (int jj = s->iused + 1; jj <= hi; ++jj) {
    px[jj - s->iused - 1] = s->ix[jj] * s->zrm;
    ifilled++;
  }
  s->iused += ifilled;
  if (ifilled == n) return;
  while (ifilled < n) {
    for (int jj = 1; jj <= JPP; ++jj) s->ix[jj] = IVAR & (s->ix[jj] + s->ix[jj - JPP + JPQ]);
    for (int jk = 1; jk <= JPNUMSPLIT; ++jk)
      for (int jj = 1 + JPP + (jk - 1) * JPLENSPLIT; jj <= imin(JPQ, JPP + jk * JPLENSPLIT); ++jj)
        s->ix[jj] = IVAR & (s->ix[jj] + s->ix[jj - JPP]);
    s->iused = imin(JPQ, n - ifilled);
    for (int i = 1; i <= s->iused; ++i) px[ifilled + i - 1] = s->ix[i] * s->zrm;
    ifilled += s->iused;
  }
}

/* radiation_pdf_sampler.F90:126-150 sample_from_pdf */
static double pdf_sample(const orc_tables* t, double fsd, double cdf) {
  const int ncdf = t->pdf_ncdf, nfsd = t->pdf_nfsd;
  double wcdf = cdf * (ncdf - 1) + 1.0;
  int icdf = imax(1, imin((int)wcdf, ncdf - 1));
  wcdf = dmax(0.0, dmin(wcdf - icdf, 1.0));
  double wfsd = (fsd - t->pdf_fsd1) * t->pdf_inv_fsd_interval + 1.0;
  int ifsd = imax(1, imin((int)wfsd, nfsd - 1));
  wfsd = dmax(0.0, dmin(wfsd - ifsd, 1.0));
#define VAL(i, j) (t->pdf_val[((j) - 1) * (size_t)ncdf + ((i) - 1)])
  return (1.0 - wcdf) * (1.0 - wfsd) * VAL(icdf, ifsd) + (1.0 - wcdf) * wfsd * VAL(icdf, ifsd + 1) +
         wcdf * (1.0 - wfsd) * VAL(icdf + 1, ifsd) + wcdf * wfsd * VAL(icdf + 1, ifsd + 1);
#undef VAL
}

static const double MaxCloudFrac = 1.0 - DBL_EPSILON * 10.0; /* radiation_cloud_cover.F90:40 */

/* radiation_cloud_cover.F90:53-69 */
static double beta2alpha(double beta, double f1, double f2) {
  if (beta < 1.0) {
    double d = fabs(f1 - f2);
    return beta + (1.0 - beta) * d / (d + 1.0 / beta - 1.0);
  }
  return 1.0;
}
/* radiation_cloud_cover.F90:231-300 cum_cloud_cover_exp_ran (1 column); arrays 0-based */
static void cum_cloud_cover_exp_ran(int nlev, const double* frac, const double* overlap_param, int beta,
                                    double* cum, double* pair) {
  double cum_product = 1.0 - frac[0];
  cum[0] = frac[0];
  for (int jl = 0; jl < nlev - 1; ++jl) {
    double alpha = beta ? beta2alpha(overlap_param[jl], frac[jl], frac[jl + 1]) : overlap_param[jl];
    pair[jl] = alpha * dmax(frac[jl], frac[jl + 1]) + (1.0 - alpha) * (frac[jl] + frac[jl + 1] - frac[jl] * frac[jl + 1]);
    if (frac[jl] >= MaxCloudFrac) cum_product = 0.0;
    else cum_product = cum_product * (1.0 - pair[jl]) / (1.0 - frac[jl]);
    cum[jl + 1] = 1.0 - cum_product;
  }
}
/* radiation_cloud_cover.F90:169-225 cum_cloud_cover_max_ran */
static void cum_cloud_cover_max_ran(int nlev, const double* frac, double* cum, double* pair) {
  double cum_product = 1.0 - frac[0];
  cum[0] = frac[0];
  for (int jl = 0; jl < nlev - 1; ++jl) {
    if (frac[jl] >= MaxCloudFrac) cum_product = 0.0;
    else cum_product = cum_product * (1.0 - dmax(frac[jl], frac[jl + 1])) / (1.0 - frac[jl]);
    cum[jl + 1] = 1.0 - cum_product;
    pair[jl] = dmax(frac[jl], frac[jl + 1]);
  }
}

/* radiation_cloud_cover.F90:339-623 cum_cloud_cover_exp_exp (1 column); arrays 0-based, object indices 1-based as in
 * the source */
static void cum_cloud_cover_exp_exp(int nlev, const double* frac, const double* overlap_param, int beta,
                                    double* cum, double* pair) {
  const double min_frac = 1.0e-6;
  int* i_top = (int*)malloc(sizeof(int) * (size_t)(nlev + 2) * 4);
  int *i_max = i_top + nlev + 2, *i_base = i_max + nlev + 2, *i_next = i_base + nlev + 2;
  double* cc_obj = (double*)malloc(sizeof(double) * (size_t)(nlev + 2) * 3);
  double *alpha_obj = cc_obj + nlev + 2, *alpha = alpha_obj + nlev + 2;
#define FR(l) frac[(l) - 1]
#define CUM(l) cum[(l) - 1]
#define PAIR(l) pair[(l) - 1]
  int jlev = 1, nobj = 0;
  while (jlev <= nlev) {
    if (FR(jlev) > min_frac) {
      nobj++;
      i_top[nobj] = jlev;
      jlev++;
      while (jlev <= nlev) { if (FR(jlev) < FR(jlev - 1)) break; jlev++; }
      i_max[nobj] = jlev - 1;
      while (jlev <= nlev) { if (FR(jlev) > FR(jlev - 1) || FR(jlev) <= min_frac) break; jlev++; }
      i_base[nobj] = jlev - 1;
      i_next[nobj] = nobj + 1;
    } else jlev++;
  }
  for (int l = 0; l < nlev; ++l) cum[l] = 0.0;
  for (int l = 0; l < nlev - 1; ++l) pair[l] = 0.0;
  if (nobj > 0) {
    for (int l = 1; l <= nlev - 1; ++l) {
      alpha[l] = beta ? beta2alpha(overlap_param[l - 1], FR(l), FR(l + 1)) : overlap_param[l - 1];
      PAIR(l) = alpha[l] * dmax(FR(l), FR(l + 1)) + (1.0 - alpha[l]) * (FR(l) + FR(l + 1) - FR(l) * FR(l + 1));
    }
    for (int jobj = 1; jobj <= nobj - 1; ++jobj) {
      double prod = 1.0;
      for (int l = i_max[jobj]; l <= i_max[jobj + 1] - 1; ++l) prod = prod * alpha[l];
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
      int jobj = 1, count = 1;
      /* "do while (jobj < nobj)" walks the linked list; the source compares the list index with the remaining count */
      while (jobj < nobj) {
        if (alpha_obj[jobj] > alpha_max) { alpha_max = alpha_obj[jobj]; iobj1 = jobj; }
        jobj = i_next[jobj];
        (void)count;
      }
      int iobj2 = i_next[iobj1];
      for (int l = i_base[iobj1] + 1; l <= i_top[iobj2] - 1; ++l) CUM(l) = CUM(i_base[iobj1]);
      double cc_pair = alpha_obj[iobj1] * dmax(cc_obj[iobj1], cc_obj[iobj2]) +
                       (1.0 - alpha_obj[iobj1]) * (cc_obj[iobj1] + cc_obj[iobj2] - cc_obj[iobj1] * cc_obj[iobj2]);
      double scaling = dmin(dmax((cc_pair - cc_obj[iobj1]) / dmax(min_frac, cc_obj[iobj2]), 0.0), 1.0);
      for (int l = i_top[iobj2]; l <= i_base[iobj2]; ++l) CUM(l) = CUM(i_base[iobj1]) + CUM(l) * scaling;
      cc_obj[iobj1] = cc_pair;
      i_base[iobj1] = i_base[iobj2];
      i_next[iobj1] = i_next[iobj2];
      alpha_obj[iobj1] = alpha_obj[iobj2];
      nobj--;
    }
    for (int l = i_base[iobj1] + 1; l <= nlev; ++l) CUM(l) = CUM(i_base[iobj1]);
    for (int l = 1; l <= nlev - 1; ++l) PAIR(l) = dmax(PAIR(l), FR(l) + CUM(l + 1) - CUM(l));
    for (int l = 1; l <= nlev; ++l) CUM(l) = dmin(CUM(l), 1.0);
  }
#undef FR
#undef CUM
#undef PAIR
  free(i_top); free(cc_obj);
}

/* radiation_cloud_generator.F90:396-530 generate_column_exp_exp */
static void generate_column_exp_exp(const orc_tables* t, int ng, int nlev, int ig, rng_stream* rs, const double* frac,
                                    const double* pair, const double* cum, const double* overhang, const double* fsd,
                                    const double* overlap_param_inhom, int itrigger, int iend, double* od_scaling,
                                    double* rand_cloud, double* rand_inhom1, double* rand_inhom2) {
#define F(a, l) ((a)[(l) - 1])
  int* is_cloudy = (int*)calloc((size_t)nlev + 2, sizeof(int));
  int iy = 0;
  is_cloudy[itrigger] = 1;
  rng_uniform(rs, iend + 1 - itrigger, rand_cloud);
  for (int jlev = itrigger + 1; jlev <= iend; ++jlev) {
    iy++;
    if (is_cloudy[jlev - 1]) {
      if (rand_cloud[iy - 1] * F(frac, jlev - 1) < F(frac, jlev) + F(frac, jlev - 1) - F(pair, jlev - 1)) is_cloudy[jlev] = 1;
    } else {
      if (rand_cloud[iy - 1] * (F(cum, jlev - 1) - F(frac, jlev - 1)) < F(pair, jlev - 1) - F(overhang, jlev - 1) - F(frac, jlev - 1))
        is_cloudy[jlev] = 1;
    }
  }
  const int n = iend + 1 - itrigger;
  rng_uniform(rs, n, rand_inhom1);
  rng_uniform(rs, n, rand_inhom2);
  for (int jc = 2; jc <= n; ++jc)
    if (rand_inhom2[jc - 1] < F(overlap_param_inhom, iend - n + jc - 1)) rand_inhom1[jc - 1] = rand_inhom1[jc - 2];
  for (int k = 0; k < n; ++k) {
    int lev = itrigger + k;
    od_scaling[(size_t)(lev - 1) * ng + ig] = is_cloudy[lev] ? pdf_sample(t, F(fsd, lev), rand_inhom1[k]) : 0.0;
  }
  free(is_cloudy);
#undef F
}

/* radiation_cloud_generator.F90:262-390 generate_column_exp_ran; levels 1-based like the source via macros */
static void generate_column_exp_ran(const orc_tables* t, int ng, int nlev, int ig, rng_stream* rs, const double* frac,
                                    const double* pair, const double* cum, const double* overhang,
                                    const double* fsd, const double* overlap_param_inhom, int itrigger, int iend,
                                    double* od_scaling, double* rand_cloud, double* rand_inhom1, double* rand_inhom2) {
#define F(a, l) ((a)[(l) - 1])
  int n_layers_to_scale = 1;
  int iy = 0;
  (void)nlev;
  rng_uniform(rs, iend + 1 - itrigger, rand_cloud);
  for (int jlev = itrigger + 1; jlev <= iend + 1; ++jlev) {
    int do_fill = 0;
    if (jlev <= iend) {
      iy++;
      if (n_layers_to_scale > 0) {
        if (rand_cloud[iy - 1] * F(frac, jlev - 1) < F(frac, jlev) + F(frac, jlev - 1) - F(pair, jlev - 1)) n_layers_to_scale++;
        else do_fill = 1;
      } else {
        if (rand_cloud[iy - 1] * (F(cum, jlev - 1) - F(frac, jlev - 1)) < F(pair, jlev - 1) - F(overhang, jlev - 1) - F(frac, jlev - 1))
          n_layers_to_scale = 1;
      }
    } else do_fill = 1;
    if (do_fill) {
      rng_uniform(rs, n_layers_to_scale, rand_inhom1);
      rng_uniform(rs, n_layers_to_scale, rand_inhom2);
      for (int jc = 2; jc <= n_layers_to_scale; ++jc)
        if (rand_inhom2[jc - 1] < F(overlap_param_inhom, jlev - n_layers_to_scale + jc - 2)) rand_inhom1[jc - 1] = rand_inhom1[jc - 2];
      for (int k = 0; k < n_layers_to_scale; ++k) {
        int lev = jlev - n_layers_to_scale + k; /* 1-based level */
        od_scaling[(size_t)(lev - 1) * ng + ig] = pdf_sample(t, F(fsd, lev), rand_inhom1[k]);
      }
      n_layers_to_scale = 0;
    }
  }
#undef F
}

/* ---- vectorizable generator: radiation_random_numbers.F90 (rng_type, IRngMinstdVector) + radiation_cloud_generator.F90:587-734 ---- */
#define MINSTD_A 48271.0
#define MINSTD_M 2147483647.0
#define MINSTD_A0 16807.0

/* rng_type%initialize, radiation_random_numbers.F90:96-150 (state kept in double precision, USE_REAL_RNG_STATE) */
static void minstd_init(int32_t iseed, int nstreams, double* istate) {
  const double rseed = fabs((double)iseed);
  for (int jstr = 1; jstr <= nstreams; ++jstr)
    istate[jstr - 1] = (double)llround(fmod(rseed * jstr * (1.0 - 0.05 * jstr + 0.005 * (double)(jstr * jstr)) * MINSTD_A0, MINSTD_M));
  for (int jstr = 0; jstr < nstreams; ++jstr) istate[jstr] = fmod(MINSTD_A * istate[jstr], MINSTD_M);
}
/* one block of nstreams numbers (uniform_distribution_1d / one jblock of _2d), :157-176 */
static void minstd_block(int nstreams, double* istate, double* randnum) {
  const double scale = 1.0 / MINSTD_M;
  for (int i = 0; i < nstreams; ++i) {
    istate[i] = fmod(MINSTD_A * istate[i], MINSTD_M);
    randnum[i] = scale * istate[i];
  }
}

/* generate_columns_exp_ran: all g-points at once; ibegin/iend 1-based; od_scaling [nlev][ng] (already zero) */
static void generate_columns_exp_ran(const orc_tables* t, int ng, int nlev, int32_t iseed, double total_cloud_cover, double frac_threshold,
                                     const double* frac, const double* pair, const double* cum, const double* overhang,
                                     const double* fsd, const double* opi, int ibegin, int iend, double* od_scaling) {
  (void)nlev;
  const int n = iend - ibegin + 1;
  double* istate = (double*)malloc(sizeof(double) * (size_t)ng * (3 * (size_t)n + 3));
  double* trigger = istate + ng;
  double* rand_cloud = trigger + ng;                  /* [n][ng]   levels ibegin..iend */
  double* rand_inhom = rand_cloud + (size_t)n * ng;   /* [n+1][ng] levels ibegin-1..iend */
  double* rand_inhom2 = rand_inhom + (size_t)(n + 1) * ng;
  char *is_cloud = (char*)calloc((size_t)ng * 2, 1), *found_cloud = is_cloud + ng;
  minstd_init(iseed, ng, istate);
  minstd_block(ng, istate, trigger);
  for (int jl = ibegin; jl <= iend; ++jl) if (frac[jl - 1] >= frac_threshold) minstd_block(ng, istate, rand_cloud + (size_t)(jl - ibegin) * ng);
  for (int k = 0; k <= n; ++k) minstd_block(ng, istate, rand_inhom + (size_t)k * ng);
  for (int jl = ibegin; jl <= iend; ++jl) if (frac[jl - 1] >= frac_threshold) minstd_block(ng, istate, rand_inhom2 + (size_t)(jl - ibegin) * ng);
  for (int g = 0; g < ng; ++g) trigger[g] = trigger[g] * total_cloud_cover;
  for (int jl = ibegin; jl <= iend; ++jl) {
    if (frac[jl - 1] >= frac_threshold) {
      const double* rc = rand_cloud + (size_t)(jl - ibegin) * ng;
      const double* r2 = rand_inhom2 + (size_t)(jl - ibegin) * ng;
      double* ri = rand_inhom + (size_t)(jl - ibegin + 1) * ng;   /* this level */
      const double* ri_above = ri - ng;
      for (int g = 0; g < ng; ++g) {
        const int prev_cloud = is_cloud[g];
        const int first_cloud = (trigger[g] <= cum[jl - 1]) && !found_cloud[g];
        found_cloud[g] = found_cloud[g] || first_cloud;
        int test = 0;
        if (jl >= 2) {   /* (the reference evaluates this at jl = ibegin = 1 with out-of-range indices; it cannot matter there: found implies first) */
          test = prev_cloud ? (rc[g] * frac[jl - 2] < frac[jl - 1] + frac[jl - 2] - pair[jl - 2])
                            : (rc[g] * (cum[jl - 2] - frac[jl - 2]) < pair[jl - 2] - overhang[jl - 2] - frac[jl - 2]);
        }
        is_cloud[g] = first_cloud || (found_cloud[g] && test);
        if (is_cloud[g]) { if (jl >= 2 && r2[g] < opi[jl - 2] && prev_cloud) ri[g] = ri_above[g]; }
        else ri[g] = 0.0;
      }
      /* sample_from_pdf_masked_block, radiation_pdf_sampler.F90:267-310 */
      for (int g = 0; g < ng; ++g) od_scaling[(size_t)(jl - 1) * ng + g] = ri[g] > 0.0 ? pdf_sample(t, fsd[jl - 1], ri[g]) : 0.0;
    } else {
      for (int g = 0; g < ng; ++g) is_cloud[g] = 0;
    }
  }
  free(istate); free(is_cloud);
}

/* radiation_cloud_generator.F90:37-255 cloud_generator.  od_scaling is [nlev][ng]. */
void orc_cloud_generator(const orc_tables* t, int ng, int nlev, int i_overlap_scheme, int32_t iseed,
                         double frac_threshold, const double* frac, const double* overlap_param,
                         double decorrelation_scaling, const double* fractional_std, int use_beta_overlap,
                         int use_vectorizable_generator, double* od_scaling, double* total_cloud_cover) {
  double* cum = (double*)malloc(sizeof(double) * (size_t)nlev * 7);
  double* pair = cum + nlev; double* overhang = pair + nlev; double* opi = overhang + nlev;
  double* rand_cloud = opi + nlev; double* ri1 = rand_cloud + nlev; double* ri2 = ri1 + nlev;
  if (i_overlap_scheme == ECRAD_OVERLAP_EXP_RAN) cum_cloud_cover_exp_ran(nlev, frac, overlap_param, use_beta_overlap, cum, pair);
  else if (i_overlap_scheme == ECRAD_OVERLAP_EXP_EXP) cum_cloud_cover_exp_exp(nlev, frac, overlap_param, use_beta_overlap, cum, pair);
  else cum_cloud_cover_max_ran(nlev, frac, cum, pair);
  double tcc = cum[nlev - 1];
  for (int jl = 0; jl < nlev - 1; ++jl) overhang[jl] = cum[jl + 1] - cum[jl];
  if (tcc < frac_threshold) {
    tcc = 0.0;
  } else {
    int jlev = 1;
    while (frac[jlev - 1] <= 0.0) jlev++;
    int ibegin = jlev, iend = jlev;
    for (jlev = jlev + 1; jlev <= nlev; ++jlev) if (frac[jlev - 1] > 0.0) iend = jlev;
    for (int jl = 0; jl < nlev - 1; ++jl) opi[jl] = overlap_param[jl];
    for (jlev = ibegin; jlev <= iend - 1; ++jlev)
      if (overlap_param[jlev - 1] > 0.0) opi[jlev - 1] = pow(overlap_param[jlev - 1], 1.0 / decorrelation_scaling);
    memset(od_scaling, 0, sizeof(double) * (size_t)ng * nlev);
    if (use_vectorizable_generator) {   /* :235-249 (not available with Exp-Exp: refused by orc_radiation) */
      generate_columns_exp_ran(t, ng, nlev, iseed, tcc, frac_threshold, frac, pair, cum, overhang, fractional_std, opi, ibegin, iend, od_scaling);
      *total_cloud_cover = tcc;
      free(cum);
      return;
    }
    rng_stream rs;
    rng_init(iseed, &rs);
    double* rand_top = (double*)malloc(sizeof(double) * (size_t)ng);
    rng_uniform(&rs, ng, rand_top);
    for (int jg = 0; jg < ng; ++jg) {
      double trigger = rand_top[jg] * tcc;
      jlev = ibegin;
      while (trigger > cum[jlev - 1] && jlev < iend) jlev++;
      if (i_overlap_scheme == ECRAD_OVERLAP_EXP_EXP)
        generate_column_exp_exp(t, ng, nlev, jg, &rs, frac, pair, cum, overhang, fractional_std, opi, jlev, iend, od_scaling,
                                rand_cloud, ri1, ri2);
      else
        generate_column_exp_ran(t, ng, nlev, jg, &rs, frac, pair, cum, overhang, fractional_std, opi, jlev, iend, od_scaling,
                                rand_cloud, ri1, ri2);
    }
    free(rand_top);
  }
  *total_cloud_cover = tcc;
  free(cum);
}
