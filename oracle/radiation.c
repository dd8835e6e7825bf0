/* radiation.c -- oracle restatement of radiation() and the McICA / Cloudless solvers.  TEST INFRASTRUCTURE.
 * Follows radiation/radiation_interface.F90:200-510, radiation_single_level.F90:216-365 (get_albedos),
 * radiation_ifs_rrtm.F90:216-613 (gas_optics) and :618-852 (planck_function_atmos/_surf),
 * radiation_cloud.F90:700-740 (crop_cloud_fraction), radiation_mcica_lw.F90:39-419, radiation_mcica_sw.F90:41-408,
 * radiation_cloudless_lw.F90, radiation_cloudless_sw.F90, radiation_lw_derivatives.F90:43-130,
 * radiation_flux.F90:397-577 (calc_surface_spectral).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "oracle.h"

#define A2(p, jcol, j) ((p)[(size_t)(j) * ncol + (jcol)]) /* (ncol, n) column-fastest, 0-based */

/* spectral sizes come from the configuration here (RRTMG 140/112 g-points in 16/14 bands; ecCKD: bands == g-points) */
#undef NG_LW
#undef NG_SW
#undef NB_LW
#undef NB_SW
#define NG_LW (cfg->n_g_lw)
#define NG_SW (cfg->n_g_sw)
#define NB_LW (cfg->n_bands_lw)
#define NB_SW (cfg->n_bands_sw)

typedef struct {
  double *od_lw, *planck_hl, *lw_emission, *lw_albedo;         /* [nlev][140], [nlev+1][140], [140], [140] */
  double *ssa_lw, *g_lw;                                       /* [nlev][140] gas + aerosol (do_lw_aerosol_scattering; zero otherwise) */
  double *od_sw, *ssa_sw, *g_sw, *incoming_sw, *alb_dir, *alb_diff;   /* [nlev][112] x3, [112] x3 */
  double *od_lw_cloud, *ssa_lw_cloud, *g_lw_cloud, *od_sw_cloud, *ssa_sw_cloud, *g_sw_cloud; /* [nlev][nb] */
  double *w;  /* scratch pool */
} col_work;

/* get_albedos, radiation_single_level.F90:216-365 (paths used by test/ifs/configCY49R1.nam) */
static int get_albedos(const orc_tables* t, const ecrad_b200_config* cfg, int ncol, int jcol,
                       const ecrad_b200_inputs* in, double* alb_dir, double* alb_diff, double* lw_albedo) {
  if (cfg->do_nearest_spectral_sw_albedo ? !t->i_albedo_from_band_sw : !t->sw_albedo_weights) return 2;
  const int nalb = cfg->n_albedo_sw;
  double band[NB_SW], band_dir[NB_SW];
  if (cfg->do_nearest_spectral_sw_albedo)   /* radiation_single_level.F90:266-285: albedo of the nearest interval */
    for (int jb = 0; jb < NB_SW; ++jb) {
      band[jb] = A2(in->sw_albedo, jcol, t->i_albedo_from_band_sw[jb] - 1);
      band_dir[jb] = in->sw_albedo_direct ? A2(in->sw_albedo_direct, jcol, t->i_albedo_from_band_sw[jb] - 1) : band[jb];
    }
  else
  for (int jb = 0; jb < NB_SW; ++jb) {
    band[jb] = 0.0; band_dir[jb] = 0.0;
    for (int ja = 0; ja < nalb; ++ja) {
      double w = t->sw_albedo_weights[(size_t)jb * nalb + ja];
      if (w != 0.0) {
        band[jb] = band[jb] + w * A2(in->sw_albedo, jcol, ja);
        if (in->sw_albedo_direct) band_dir[jb] = band_dir[jb] + w * A2(in->sw_albedo_direct, jcol, ja);
      }
    }
  }
  for (int g = 0; g < NG_SW; ++g) {
    int jb = t->band_sw[g];
    alb_diff[g] = band[jb];
    alb_dir[g] = in->sw_albedo_direct ? band_dir[jb] : band[jb];
  }
  if (cfg->do_nearest_spectral_lw_emiss) {
    if (!t->i_emiss_from_band_lw) return 2;
    for (int g = 0; g < NG_LW; ++g)
      lw_albedo[g] = 1.0 - A2(in->lw_emissivity, jcol, t->i_emiss_from_band_lw[t->band_lw[g]] - 1);
  } else {
    /* weighted emissivity intervals, radiation_single_level.F90:322-345 */
    if (!t->lw_emiss_weights) return 2;
    const int nem = cfg->n_emiss_lw;
    double lband[NB_LW];
    for (int jb = 0; jb < NB_LW; ++jb) {
      lband[jb] = 0.0;
      for (int ja = 0; ja < nem; ++ja) {
        double w = t->lw_emiss_weights[(size_t)jb * nem + ja];
        if (w != 0.0) lband[jb] = lband[jb] + w * (1.0 - A2(in->lw_emissivity, jcol, ja));
      }
    }
    for (int g = 0; g < NG_LW; ++g) lw_albedo[g] = lband[t->band_lw[g]];
  }
  return 0;
}

/* Planck function at a temperature for all 16 bands: radiation_ifs_rrtm.F90:676-699 */
static void planck_bands(const orc_tables* t, double temperature, double* store) {
  const double zfluxfac = 2.0 * asin(1.0) * 1.0e4;
  int ind; double frac;
  if (temperature < 339.0 && temperature >= 160.0) {
    ind = (int)(temperature - 159.0);
    frac = temperature - (int)temperature;
  } else if (temperature >= 339.0) {
    ind = 180; frac = temperature - 339.0;
  } else {
    ind = 1; frac = 0.0;
  }
  for (int jb = 0; jb < 16; ++jb) {
    double factor = zfluxfac * t->delwave[jb];
    const double* tp = t->totplnk + (size_t)jb * 181;
    store[jb] = factor * (tp[ind - 1] + frac * (tp[ind] - tp[ind - 1]));
  }
}

/* gas_optics for one column, radiation_ifs_rrtm.F90:216-613.  Outputs in ecRad level order (0 = top). */
static void gas_optics_column(const orc_tables* t, const ecrad_b200_config* cfg, int ncol, int nlev, int jcol,
                              const ecrad_b200_inputs* in, const double* lw_albedo, double* od_lw, double* planck_hl,
                              double* lw_emission, double* od_sw, double* ssa_sw, double* incoming_sw) {
  /* radiation_interface.F90:333-355: each gas model fills the spectra it is configured for */
  const int ckd_lw = cfg->i_gas_model_lw == ECRAD_GAS_ECCKD, ckd_sw = cfg->i_gas_model_sw == ECRAD_GAS_ECCKD;
  if (ckd_lw || ckd_sw)
    orc_ecckd_gas_optics_column(t, cfg, ncol, nlev, jcol, in, lw_albedo, od_lw, planck_hl, lw_emission, od_sw, ssa_sw, incoming_sw);
  if (ckd_lw && ckd_sw) return;
  ecrad_b200_config cfg_rrtmg = *cfg;
  cfg_rrtmg.do_lw = cfg->do_lw && !ckd_lw; cfg_rrtmg.do_sw = cfg->do_sw && !ckd_sw;
  cfg = &cfg_rrtmg;
  double* buf = (double*)malloc(sizeof(double) * (size_t)(nlev + 1) * 16);
  double *p_hl = buf, *t_hl = p_hl + (nlev + 1), *p_fl = t_hl + (nlev + 1), *t_fl = p_fl + nlev;
  double* gas[9];
  const double* src[9] = {in->h2o_mmr, in->co2_mmr, in->ch4_mmr, in->n2o_mmr, in->cfc11_mmr, in->cfc12_mmr,
                          in->hcfc22_mmr, in->ccl4_mmr, in->o3_mmr};
  for (int i = 0; i < 9; ++i) gas[i] = t_fl + nlev + (size_t)i * nlev;
  for (int jl = 0; jl <= nlev; ++jl) { p_hl[jl] = A2(in->pressure_hl, jcol, jl); t_hl[jl] = A2(in->temperature_hl, jcol, jl); }
  for (int jl = 0; jl < nlev; ++jl) {
    p_fl[jl] = 0.5 * (p_hl[jl] + p_hl[jl + 1]);
    t_fl[jl] = 0.5 * (t_hl[jl] + t_hl[jl + 1]);
    for (int i = 0; i < 9; ++i) gas[i][jl] = A2(src[i], jcol, jl);
  }
  orc_lay_lw* lay = (orc_lay_lw*)malloc(sizeof(orc_lay_lw) * (size_t)nlev);
  orc_prepare_gases(nlev, p_hl, t_hl, p_fl, t_fl, gas[0], gas[1], gas[2], gas[3], gas[4], gas[5], gas[6], gas[7], gas[8], lay);
  if (cfg->do_lw) {
    int laytrop;
    double* tau = (double*)malloc(sizeof(double) * (size_t)nlev * NG_LW * 2);
    double* pfrac = tau + (size_t)nlev * NG_LW;
    orc_setcoef_lw(t, nlev, lay, &laytrop);
    orc_taumol_lw(t, nlev, lay, laytrop, tau, pfrac);
    /* planck_function_atmos :618-752 */
    double store[16];
    for (int jlev = 1; jlev <= nlev + 1; ++jlev) {
      planck_bands(t, t_hl[jlev - 1], store);
      int ilay = (jlev == 1) ? nlev : nlev + 2 - jlev; /* RRTMG layer (1-based) whose PFRAC is used */
      for (int g = 0; g < NG_LW; ++g)
        planck_hl[(size_t)(jlev - 1) * NG_LW + g] = store[t->band_lw[g]] * pfrac[(size_t)(ilay - 1) * NG_LW + g];
    }
    /* planck_function_surf :757-852 and lw_emission :466 */
    planck_bands(t, in->skin_temperature[jcol], store);
    for (int g = 0; g < NG_LW; ++g) {
      lw_emission[g] = store[t->band_lw[g]] * pfrac[g];
      lw_emission[g] = lw_emission[g] * (1.0 - lw_albedo[g]);
    }
    /* :506-511 un-reverse + clamp */
    for (int jl = 1; jl <= nlev; ++jl)
      for (int g = 0; g < NG_LW; ++g) {
        double v = tau[(size_t)(nlev - jl) * NG_LW + g];
        od_lw[(size_t)(jl - 1) * NG_LW + g] = v > cfg->min_gas_od_lw ? v : cfg->min_gas_od_lw;
      }
    free(tau);
  }
  if (cfg->do_sw) {
    memset(od_sw, 0, sizeof(double) * (size_t)nlev * NG_SW);
    memset(ssa_sw, 0, sizeof(double) * (size_t)nlev * NG_SW);
    for (int g = 0; g < NG_SW; ++g) incoming_sw[g] = 0.0;
    if (in->cos_sza[jcol] > 0.0) {
      int laytrop;
      orc_lay_sw* ls = (orc_lay_sw*)malloc(sizeof(orc_lay_sw) * (size_t)nlev);
      double* zod = (double*)malloc(sizeof(double) * (size_t)nlev * NG_SW * 2);
      double* zssa = zod + (size_t)nlev * NG_SW;
      double incsol[NG_SW];
      orc_setcoef_sw(t, nlev, lay, ls, &laytrop);
      orc_taumol_sw(t, nlev, ls, laytrop, zod, zssa, incsol);
      double sum = 0.0;
      for (int g = 0; g < NG_SW; ++g) sum = sum + incsol[g];
      double scale = in->solar_irradiance / sum;  /* :557-565 */
      for (int jl = 1; jl <= nlev; ++jl)
        for (int g = 0; g < NG_SW; ++g) {
          double v = zod[(size_t)(jl - 1) * NG_SW + g];
          od_sw[(size_t)(nlev - jl) * NG_SW + g] = v > cfg->min_gas_od_sw ? v : cfg->min_gas_od_sw;
          ssa_sw[(size_t)(nlev - jl) * NG_SW + g] = zssa[(size_t)(jl - 1) * NG_SW + g];
        }
      for (int g = 0; g < NG_SW; ++g) incoming_sw[g] = scale * incsol[g];
      free(zod); free(ls);
    }
  }
  free(lay); free(buf);
}

int orc_gas_optics_column(const orc_tables* t, const ecrad_b200_config* cfg, int ncol, int nlev, int jcol,
                          const ecrad_b200_inputs* in, double* od_lw, double* planck_hl, double* lw_emission,
                          double* od_sw, double* ssa_sw, double* incoming_sw) {
  double alb_dir[NG_SW], alb_diff[NG_SW], lw_albedo[NG_LW];
  int rc = get_albedos(t, cfg, ncol, jcol - 1, in, alb_dir, alb_diff, lw_albedo);
  if (rc) return rc;
  gas_optics_column(t, cfg, ncol, nlev, jcol - 1, in, lw_albedo, od_lw, planck_hl, lw_emission, od_sw, ssa_sw, incoming_sw);
  return 0;
}

/* add_aerosol_optics, radiation_aerosol_optics.F90:487-826 (band-wise aerosol properties, no LW aerosol scattering,
 * do_cloud_aerosol_per_*_g_point = false).  Modifies od_sw, ssa_sw, g_sw, od_lw of one column in place. */
static void add_aerosol_optics(const orc_tables* t, const ecrad_b200_config* cfg, int ncol, int nlev, int jcol,
                               const ecrad_b200_inputs* in, double* od_lw, double* ssa_lw, double* g_lw, double* od_sw, double* ssa_sw, double* g_sw) {
  const int ntype = cfg->n_aerosol_types, nrh = t->aer_nrh;
  const double OneOverAccelDueToGravity = 1.0 / 9.80665;
  double* od_sw_aer = (double*)calloc((size_t)nlev * (3 * NB_SW + 3 * NB_LW), sizeof(double));
  double *scat_sw_aer = od_sw_aer + (size_t)nlev * NB_SW, *scat_g_sw_aer = scat_sw_aer + (size_t)nlev * NB_SW;
  double* od_lw_aer = scat_g_sw_aer + (size_t)nlev * NB_SW;
  double *scat_lw_aer = od_lw_aer + (size_t)nlev * NB_LW, *scat_g_lw_aer = scat_lw_aer + (size_t)nlev * NB_LW;
  int* irhs = (int*)malloc(sizeof(int) * (size_t)nlev);
  double* factor = (double*)malloc(sizeof(double) * (size_t)nlev);
  for (size_t i = 0; i < (size_t)nlev * NG_SW; ++i) g_sw[i] = 0.0;
  for (int jl = 0; jl < nlev; ++jl) {
    /* gas%mixing_ratio(:,:,IH2O) is a volume mixing ratio under ecCKD: radiation_aerosol_optics.F90:590-600 converts */
    double h2o = A2(in->h2o_mmr, jcol, jl);
    if (cfg->i_gas_model_lw == ECRAD_GAS_ECCKD && cfg->i_gas_model_sw == ECRAD_GAS_ECCKD) h2o = h2o * (18.0152833 / 28.970);
    double rh = h2o / A2(in->h2o_sat_liq, jcol, jl);
    int irh;   /* calc_rh_index, radiation_aerosol_optics_data.F90:640-664 (1-based) */
    if (rh > t->aer_rh_lower[nrh - 1]) irh = nrh;
    else { irh = 1; while (rh > t->aer_rh_lower[irh]) irh++; }
    irhs[jl] = irh;
    factor[jl] = (A2(in->pressure_hl, jcol, jl + 1) - A2(in->pressure_hl, jcol, jl)) * OneOverAccelDueToGravity;
  }
  for (int jt = 0; jt < ntype; ++jt) {
    const int itype = t->aer_itype[jt] - 1, iclass = t->aer_iclass[jt];
    if (iclass == 0) continue;
    for (int jl = 0; jl < nlev; ++jl) {
      const double mixing_ratio = in->aerosol_mmr[((size_t)jt * nlev + jl) * ncol + jcol];
      const size_t isw = iclass == 1 ? (size_t)itype * NB_SW : ((size_t)itype * nrh + (irhs[jl] - 1)) * NB_SW;
      const size_t ilw = iclass == 1 ? (size_t)itype * NB_LW : ((size_t)itype * nrh + (irhs[jl] - 1)) * NB_LW;
      const double *me_sw = (iclass == 1 ? t->aer_me_sw_phobic : t->aer_me_sw_philic) + isw;
      const double *ss_sw = (iclass == 1 ? t->aer_ssa_sw_phobic : t->aer_ssa_sw_philic) + isw;
      const double *gg_sw = (iclass == 1 ? t->aer_g_sw_phobic : t->aer_g_sw_philic) + isw;
      const double *me_lw = (iclass == 1 ? t->aer_me_lw_phobic : t->aer_me_lw_philic) + ilw;
      const double *ss_lw = (iclass == 1 ? t->aer_ssa_lw_phobic : t->aer_ssa_lw_philic) + ilw;
      const double *gg_lw = (iclass == 1 ? t->aer_g_lw_phobic : t->aer_g_lw_philic) + ilw;
      for (int jb = 0; jb < NB_SW; ++jb) {
        double local_od_sw = factor[jl] * mixing_ratio * me_sw[jb];
        od_sw_aer[jl * NB_SW + jb] = od_sw_aer[jl * NB_SW + jb] + local_od_sw;
        scat_sw_aer[jl * NB_SW + jb] = scat_sw_aer[jl * NB_SW + jb] + local_od_sw * ss_sw[jb];
        scat_g_sw_aer[jl * NB_SW + jb] = scat_g_sw_aer[jl * NB_SW + jb] + local_od_sw * ss_sw[jb] * gg_sw[jb];
      }
      if (cfg->do_lw_aerosol_scattering) {   /* radiation_aerosol_optics.F90:657-670, :697-710 */
        for (int jb = 0; jb < NB_LW; ++jb) {
          double local_od_lw = factor[jl] * mixing_ratio * me_lw[jb];
          od_lw_aer[jl * NB_LW + jb] = od_lw_aer[jl * NB_LW + jb] + local_od_lw;
          scat_lw_aer[jl * NB_LW + jb] = scat_lw_aer[jl * NB_LW + jb] + local_od_lw * ss_lw[jb];
          scat_g_lw_aer[jl * NB_LW + jb] = scat_g_lw_aer[jl * NB_LW + jb] + local_od_lw * ss_lw[jb] * gg_lw[jb];
        }
      } else
      for (int jb = 0; jb < NB_LW; ++jb)
        od_lw_aer[jl * NB_LW + jb] = od_lw_aer[jl * NB_LW + jb] + factor[jl] * mixing_ratio * me_lw[jb] * (1.0 - ss_lw[jb]);
    }
  }
  if (!cfg->do_sw_delta_scaling_with_gases) {
    /* delta_eddington_extensive_vec, radiation_delta_eddington.h:74-96 (1.0e-24 is a default-kind literal) */
    for (int i = 0; i < nlev * NB_SW; ++i) {
      double den = scat_sw_aer[i] > (double)1.0e-24f ? scat_sw_aer[i] : (double)1.0e-24f;
      double g = scat_g_sw_aer[i] / den;
      double f = g * g;
      od_sw_aer[i] = od_sw_aer[i] - scat_sw_aer[i] * f;
      scat_sw_aer[i] = scat_sw_aer[i] * (1.0 - f);
      scat_g_sw_aer[i] = scat_sw_aer[i] * g / (1.0 + g);
    }
  }
  if (cfg->do_sw)
    for (int jl = 0; jl < nlev; ++jl)
      for (int g = 0; g < NG_SW; ++g) {
        const int ib = t->band_sw[g];
        const size_t i = (size_t)jl * NG_SW + g;
        double local_od = od_sw[i] + od_sw_aer[jl * NB_SW + ib];
        if (local_od > 0.0 && od_sw_aer[jl * NB_SW + ib] > 0.0) {
          double local_scat = ssa_sw[i] * od_sw[i] + scat_sw_aer[jl * NB_SW + ib];
          if (local_scat > 0.0) g_sw[i] = scat_g_sw_aer[jl * NB_SW + ib] / local_scat;
          ssa_sw[i] = local_scat / local_od;
          od_sw[i] = local_od;
        }
      }
  if (cfg->do_lw && cfg->do_lw_aerosol_scattering) {
    /* radiation_aerosol_optics.F90:778-801: delta-Eddington scaling of the aerosol, then the merge (gases do not scatter) */
    for (int i = 0; i < nlev * NB_LW; ++i) {
      double den = scat_lw_aer[i] > (double)1.0e-24f ? scat_lw_aer[i] : (double)1.0e-24f;
      double g = scat_g_lw_aer[i] / den;
      double f = g * g;
      od_lw_aer[i] = od_lw_aer[i] - scat_lw_aer[i] * f;
      scat_lw_aer[i] = scat_lw_aer[i] * (1.0 - f);
      scat_g_lw_aer[i] = scat_lw_aer[i] * g / (1.0 + g);
    }
    for (int jl = 0; jl < nlev; ++jl)
      for (int g = 0; g < NG_LW; ++g) {
        const int ib = t->band_lw[g];
        const size_t i = (size_t)jl * NG_LW + g;
        double local_od = od_lw[i] + od_lw_aer[jl * NB_LW + ib];
        if (local_od > 0.0 && od_lw_aer[jl * NB_LW + ib] > 0.0) {
          if (scat_lw_aer[jl * NB_LW + ib] > 0.0) g_lw[i] = scat_g_lw_aer[jl * NB_LW + ib] / scat_lw_aer[jl * NB_LW + ib];
          ssa_lw[i] = scat_lw_aer[jl * NB_LW + ib] / local_od;
          od_lw[i] = local_od;
        }
      }
  } else if (cfg->do_lw)
    for (int jl = 0; jl < nlev; ++jl)
      for (int g = 0; g < NG_LW; ++g)
        od_lw[(size_t)jl * NG_LW + g] = od_lw[(size_t)jl * NG_LW + g] + od_lw_aer[jl * NB_LW + (t->band_lw[g])];
  free(od_sw_aer); free(irhs); free(factor);
}

/* radiation_lw_derivatives.F90:43-84 / :93-130 */
static void lw_derivatives(int ng, int nlev, int ncol, int jcol, const double* trans, const double* flux_up_surf,
                           double weight, int modify, double* out) {
  double* dg = (double*)malloc(sizeof(double) * (size_t)ng);
  double s = 0.0;
  for (int g = 0; g < ng; ++g) s = s + flux_up_surf[g];
  for (int g = 0; g < ng; ++g) dg[g] = flux_up_surf[g] / s;
  A2(out, jcol, nlev) = 1.0;
  for (int jl = nlev - 1; jl >= 0; --jl) {
    double sum = 0.0;
    for (int g = 0; g < ng; ++g) { dg[g] = dg[g] * trans[(size_t)jl * ng + g]; sum = sum + dg[g]; }
    if (modify) A2(out, jcol, jl) = (1.0 - weight) * A2(out, jcol, jl) + weight * sum;
    else A2(out, jcol, jl) = sum;
  }
  free(dg);
}

static void sum_g(int ng, int nlev1, const double* f, int ncol, int jcol, double* out) {
  if (!out) return;
  for (int jl = 0; jl < nlev1; ++jl) {
    double s = 0.0;
    for (int g = 0; g < ng; ++g) s = s + f[(size_t)jl * ng + g];
    A2(out, jcol, jl) = s;
  }
}
/* indexed_sum_profile, radiation_flux.F90:820-855: band profile (nband, ncol, nlev+1) */
static void band_profile(int ng, int nb, int nlev1, const int32_t* band, const double* f, int ncol, int jcol,
                         double* out, int add) {
  if (!out) return;
  for (int jl = 0; jl < nlev1; ++jl) {
    double* o = out + ((size_t)jl * ncol + jcol) * nb;
    if (!add) for (int b = 0; b < nb; ++b) o[b] = 0.0;
    for (int g = 0; g < ng; ++g) o[band[g]] = o[band[g]] + f[(size_t)jl * ng + g];
  }
}

#define OUTG(p, ng, g) ((p)[(size_t)jcol * (ng) + (g)])

/* ----- LW: radiation_mcica_lw.F90:39-419 and radiation_cloudless_lw.F90 ----- */
static void solver_lw(const orc_tables* t, const ecrad_b200_config* cfg, int ncol, int nlev, int jcol,
                      const ecrad_b200_inputs* in, ecrad_b200_outputs* out, const col_work* w, const double* frac) {
  const int ng = NG_LW;
  const size_t nl = (size_t)nlev * ng, nl1 = (size_t)(nlev + 1) * ng;
  double* pool = (double*)malloc(sizeof(double) * (8 * nl + 4 * nl1 + nl + 4 * ng));
  double *ref_clear = pool, *trans_clear = ref_clear + nl, *ref = trans_clear + nl, *trans = ref + nl;
  double *su_clear = trans + nl, *sd_clear = su_clear + nl, *su = sd_clear + nl, *sd = su + nl;
  double *fu = sd + nl, *fd = fu + nl1, *fu_clear = fd + nl1, *fd_clear = fu_clear + nl1;
  double *od_scaling = fd_clear + nl1, *od_total = od_scaling + nl, *ssa_total = od_total + ng, *g_total = ssa_total + ng;
  const double* planck = w->planck_hl;
  if (cfg->do_lw_aerosol_scattering) {
    /* radiation_mcica_lw.F90:160-173, radiation_cloudless_lw.F90:102-116: two-stream with scattering in every layer, adding method */
    orc_calc_ref_trans_lw(ng * nlev, w->od_lw, w->ssa_lw, w->g_lw, planck, planck + ng, ref_clear, trans_clear, su_clear, sd_clear);
    orc_adding_ica_lw(ng, nlev, ref_clear, trans_clear, su_clear, sd_clear, w->lw_emission, w->lw_albedo, fu_clear, fd_clear);
  } else {
  /* clear sky: no-scattering (do_lw_aerosol_scattering = false) */
  orc_calc_no_scattering_transmittance_lw(ng * nlev, w->od_lw, planck, planck + ng, trans_clear, su_clear, sd_clear);
  memset(ref_clear, 0, sizeof(double) * nl);
  orc_calc_fluxes_no_scattering_lw(ng, nlev, trans_clear, su_clear, sd_clear, w->lw_emission, w->lw_albedo, fu_clear, fd_clear);
  }
  sum_g(ng, nlev + 1, fu_clear, ncol, jcol, out->lw_up_clear);
  sum_g(ng, nlev + 1, fd_clear, ncol, jcol, out->lw_dn_clear);
  for (int g = 0; g < ng; ++g) {
    if (out->lw_dn_surf_clear_g) OUTG(out->lw_dn_surf_clear_g, ng, g) = fd_clear[nl1 - ng + g];
    if (out->lw_up_toa_clear_g) OUTG(out->lw_up_toa_clear_g, ng, g) = fu_clear[g];
  }
  if (cfg->i_solver_lw == ECRAD_SOLVER_CLOUDLESS) {
    /* radiation_cloudless_lw.F90:118-160: all-sky = clear-sky */
    sum_g(ng, nlev + 1, fu_clear, ncol, jcol, out->lw_up);
    sum_g(ng, nlev + 1, fd_clear, ncol, jcol, out->lw_dn);
    for (int g = 0; g < ng; ++g) {
      if (out->lw_dn_surf_g) OUTG(out->lw_dn_surf_g, ng, g) = fd_clear[nl1 - ng + g];
      if (out->lw_up_toa_g) OUTG(out->lw_up_toa_g, ng, g) = fu_clear[g];
    }
    band_profile(ng, NB_LW, nlev + 1, t->band_lw, fu_clear, ncol, jcol, out->lw_up_band, 0);
    band_profile(ng, NB_LW, nlev + 1, t->band_lw, fd_clear, ncol, jcol, out->lw_dn_band, 0);
    if (cfg->do_lw_derivatives && out->lw_derivatives)
      lw_derivatives(ng, nlev, ncol, jcol, trans_clear, fu_clear + nl1 - ng, 0.0, 0, out->lw_derivatives);
    free(pool);
    return;
  }
  double tcc;
  double *fsd = (double*)malloc(sizeof(double) * (size_t)nlev * 2), *op = fsd + nlev;
  for (int jl = 0; jl < nlev; ++jl) fsd[jl] = A2(in->fractional_std, jcol, jl);
  for (int jl = 0; jl < nlev - 1; ++jl) op[jl] = A2(in->overlap_param, jcol, jl);
  orc_cloud_generator(t, ng, nlev, cfg->i_overlap_scheme, in->iseed[jcol] + 997, cfg->cloud_fraction_threshold, frac, op,
                      cfg->cloud_inhom_decorr_scaling, fsd, cfg->use_beta_overlap, cfg->use_vectorizable_generator, od_scaling, &tcc);
  free(fsd);
  if (out->cloud_cover_lw) out->cloud_cover_lw[jcol] = tcc;
  if (tcc >= cfg->cloud_fraction_threshold) {
    int* is_clear = (int*)malloc(sizeof(int) * (size_t)nlev);
    int i_cloud_top = nlev + 1;
    for (int jl = 0; jl < nlev; ++jl) {
      is_clear[jl] = 1;
      if (frac[jl] >= cfg->cloud_fraction_threshold) {
        is_clear[jl] = 0;
        if (i_cloud_top > jl + 1) i_cloud_top = jl + 1;
        for (int g = 0; g < ng; ++g) {
          int jb = t->band_lw[g];
          double od_cloud_new = od_scaling[(size_t)jl * ng + g] * w->od_lw_cloud[jl * NB_LW + jb];
          od_total[g] = w->od_lw[(size_t)jl * ng + g] + od_cloud_new;
          ssa_total[g] = 0.0; g_total[g] = 0.0;
          if (cfg->do_lw_cloud_scattering && cfg->do_lw_aerosol_scattering) {   /* radiation_mcica_lw.F90:260-280 */
            if (od_total[g] > 0.0) {
              const size_t i = (size_t)jl * ng + g;
              double scat_od_total = w->ssa_lw[i] * w->od_lw[i] + w->ssa_lw_cloud[jl * NB_LW + jb] * od_cloud_new;
              ssa_total[g] = scat_od_total / od_total[g];
              if (scat_od_total > 0.0)
                g_total[g] = (w->g_lw[i] * w->ssa_lw[i] * w->od_lw[i] + w->g_lw_cloud[jl * NB_LW + jb] * w->ssa_lw_cloud[jl * NB_LW + jb] * od_cloud_new) / scat_od_total;
            }
          } else if (cfg->do_lw_cloud_scattering) {
            if (od_total[g] > 0.0) {
              double scat_od = w->ssa_lw_cloud[jl * NB_LW + jb] * od_cloud_new;
              ssa_total[g] = scat_od / od_total[g];
              if (scat_od > 0.0) g_total[g] = w->g_lw_cloud[jl * NB_LW + jb] * w->ssa_lw_cloud[jl * NB_LW + jb] * od_cloud_new / scat_od;
            }
          }
        }
        if (cfg->do_lw_cloud_scattering)
          orc_calc_ref_trans_lw(ng, od_total, ssa_total, g_total, planck + (size_t)jl * ng, planck + (size_t)(jl + 1) * ng,
                                ref + (size_t)jl * ng, trans + (size_t)jl * ng, su + (size_t)jl * ng, sd + (size_t)jl * ng);
        else
          orc_calc_no_scattering_transmittance_lw(ng, od_total, planck + (size_t)jl * ng, planck + (size_t)(jl + 1) * ng,
                                                  trans + (size_t)jl * ng, su + (size_t)jl * ng, sd + (size_t)jl * ng);
      } else {
        for (int g = 0; g < ng; ++g) {
          size_t i = (size_t)jl * ng + g;
          ref[i] = ref_clear[i]; trans[i] = trans_clear[i]; su[i] = su_clear[i]; sd[i] = sd_clear[i];
        }
      }
    }
    if (cfg->do_lw_aerosol_scattering)   /* radiation_mcica_lw.F90:324-329: scattering in all layers */
      orc_adding_ica_lw(ng, nlev, ref, trans, su, sd, w->lw_emission, w->lw_albedo, fu, fd);
    else if (cfg->do_lw_cloud_scattering)
      orc_fast_adding_ica_lw(ng, nlev, ref, trans, su, sd, w->lw_emission, w->lw_albedo, is_clear, i_cloud_top, fd_clear, fu, fd);
    else
      orc_calc_fluxes_no_scattering_lw(ng, nlev, trans, su, sd, w->lw_emission, w->lw_albedo, fu, fd);
    free(is_clear);
    for (int jl = 0; jl <= nlev; ++jl) {
      double s_up = 0.0, s_dn = 0.0;
      for (int g = 0; g < ng; ++g) { s_up = s_up + fu[(size_t)jl * ng + g]; s_dn = s_dn + fd[(size_t)jl * ng + g]; }
      if (out->lw_up) A2(out->lw_up, jcol, jl) = tcc * s_up + (1.0 - tcc) * A2(out->lw_up_clear, jcol, jl);
      if (out->lw_dn) A2(out->lw_dn, jcol, jl) = tcc * s_dn + (1.0 - tcc) * A2(out->lw_dn_clear, jcol, jl);
    }
    for (int g = 0; g < ng; ++g) {
      if (out->lw_dn_surf_g) OUTG(out->lw_dn_surf_g, ng, g) = tcc * fd[nl1 - ng + g] + (1.0 - tcc) * fd_clear[nl1 - ng + g];
      if (out->lw_up_toa_g) OUTG(out->lw_up_toa_g, ng, g) = tcc * fu[g] + (1.0 - tcc) * fu_clear[g];
    }
    if (cfg->do_lw_derivatives && out->lw_derivatives) {
      lw_derivatives(ng, nlev, ncol, jcol, trans, fu + nl1 - ng, 0.0, 0, out->lw_derivatives);
      if (tcc < 1.0 - cfg->cloud_fraction_threshold)
        lw_derivatives(ng, nlev, ncol, jcol, trans_clear, fu_clear + nl1 - ng, 1.0 - tcc, 1, out->lw_derivatives);
    }
  } else {
    for (int jl = 0; jl <= nlev; ++jl) {
      if (out->lw_up) A2(out->lw_up, jcol, jl) = A2(out->lw_up_clear, jcol, jl);
      if (out->lw_dn) A2(out->lw_dn, jcol, jl) = A2(out->lw_dn_clear, jcol, jl);
    }
    for (int g = 0; g < ng; ++g) {
      if (out->lw_dn_surf_g) OUTG(out->lw_dn_surf_g, ng, g) = fd_clear[nl1 - ng + g];
      if (out->lw_up_toa_g) OUTG(out->lw_up_toa_g, ng, g) = fu_clear[g];
    }
    if (cfg->do_lw_derivatives && out->lw_derivatives)
      lw_derivatives(ng, nlev, ncol, jcol, trans_clear, fu_clear + nl1 - ng, 0.0, 0, out->lw_derivatives);
  }
  free(pool);
}

/* delta_eddington, radiation_delta_eddington.h:20-37 (do_sw_delta_scaling_with_gases: applied to the gas-aerosol(-cloud) mixture
 * inside the solver instead of to the aerosol and cloud parts on their own) */
static void delta_eddington_vec(int n, double* od, double* ssa, double* g) {
  for (int i = 0; i < n; ++i) {
    const double f = g[i] * g[i];
    od[i] = od[i] * (1.0 - ssa[i] * f);
    ssa[i] = ssa[i] * (1.0 - f) / (1.0 - ssa[i] * f);
    g[i] = g[i] / (1.0 + g[i]);
  }
}

/* ----- SW: radiation_mcica_sw.F90:41-408 and radiation_cloudless_sw.F90 ----- */
static void solver_sw(const orc_tables* t, const ecrad_b200_config* cfg, int ncol, int nlev, int jcol,
                      const ecrad_b200_inputs* in, ecrad_b200_outputs* out, const col_work* w, const double* frac) {
  const int ng = NG_SW;
  const size_t nl = (size_t)nlev * ng, nl1 = (size_t)(nlev + 1) * ng;
  const double cos_sza = in->cos_sza[jcol];
  if (!(cos_sza > 0.0)) {
    for (int jl = 0; jl <= nlev; ++jl) {
      if (out->sw_up) A2(out->sw_up, jcol, jl) = 0.0;
      if (out->sw_dn) A2(out->sw_dn, jcol, jl) = 0.0;
      if (out->sw_dn_direct) A2(out->sw_dn_direct, jcol, jl) = 0.0;
      if (out->sw_up_clear) A2(out->sw_up_clear, jcol, jl) = 0.0;
      if (out->sw_dn_clear) A2(out->sw_dn_clear, jcol, jl) = 0.0;
      if (out->sw_dn_direct_clear) A2(out->sw_dn_direct_clear, jcol, jl) = 0.0;
    }
    double* gs[6] = {out->sw_dn_diffuse_surf_g, out->sw_dn_direct_surf_g, out->sw_up_toa_g,
                     out->sw_dn_diffuse_surf_clear_g, out->sw_dn_direct_surf_clear_g, out->sw_up_toa_clear_g};
    for (int k = 0; k < 6; ++k) if (gs[k]) for (int g = 0; g < ng; ++g) OUTG(gs[k], ng, g) = 0.0;
    double* bs[3] = {out->sw_up_band, out->sw_dn_band, out->sw_dn_direct_band};
    for (int k = 0; k < 3; ++k)
      if (bs[k]) for (int jl = 0; jl <= nlev; ++jl) for (int b = 0; b < NB_SW; ++b) bs[k][((size_t)jl * ncol + jcol) * NB_SW + b] = 0.0;
    return;
  }
  double* pool = (double*)malloc(sizeof(double) * (11 * nl + 3 * nl1 + 3 * ng));
  double *ref_clear = pool, *trans_clear = ref_clear + nl, *ref = trans_clear + nl, *trans = ref + nl;
  double *rdir_clear = trans + nl, *tdd_clear = rdir_clear + nl, *rdir = tdd_clear + nl, *tdd = rdir + nl;
  double *tdir_clear = tdd + nl, *tdir = tdir_clear + nl, *od_scaling = tdir + nl;
  double *fu = od_scaling + nl, *fdd = fu + nl1, *fdir = fdd + nl1;
  double *od_total = fdir + nl1, *ssa_total = od_total + ng, *g_total = ssa_total + ng;
  const double* gzero = w->g_sw;   /* g_sw: zero without aerosols (radiation_interface.F90:395), else from add_aerosol_optics */
  const int cloudless = (cfg->i_solver_sw == ECRAD_SOLVER_CLOUDLESS);
  if (cfg->do_sw_delta_scaling_with_gases) {
    /* radiation_mcica_sw.F90:165-180 / radiation_cloudless_sw.F90:126-146: scale the gas-aerosol mixture layer by layer */
    for (int jl = 0; jl < nlev; ++jl) {
      for (int g = 0; g < ng; ++g) { od_total[g] = w->od_sw[(size_t)jl * ng + g]; ssa_total[g] = w->ssa_sw[(size_t)jl * ng + g]; g_total[g] = gzero[(size_t)jl * ng + g]; }
      delta_eddington_vec(ng, od_total, ssa_total, g_total);
      if (cloudless)
        orc_calc_reflectance_transmittance_sw(ng, cos_sza, od_total, ssa_total, g_total, ref_clear + (size_t)jl * ng, trans_clear + (size_t)jl * ng,
                                              rdir_clear + (size_t)jl * ng, tdd_clear + (size_t)jl * ng, tdir_clear + (size_t)jl * ng);
      else
        orc_calc_ref_trans_sw(ng, cos_sza, od_total, ssa_total, g_total, ref_clear + (size_t)jl * ng, trans_clear + (size_t)jl * ng,
                              rdir_clear + (size_t)jl * ng, tdd_clear + (size_t)jl * ng, tdir_clear + (size_t)jl * ng);
    }
  } else if (cloudless) {
    for (int jl = 0; jl < nlev; ++jl)
      orc_calc_reflectance_transmittance_sw(ng, cos_sza, w->od_sw + (size_t)jl * ng, w->ssa_sw + (size_t)jl * ng, gzero + (size_t)jl * ng,
                                            ref_clear + (size_t)jl * ng, trans_clear + (size_t)jl * ng, rdir_clear + (size_t)jl * ng,
                                            tdd_clear + (size_t)jl * ng, tdir_clear + (size_t)jl * ng);
  } else {
    orc_calc_ref_trans_sw(ng * nlev, cos_sza, w->od_sw, w->ssa_sw, gzero, ref_clear, trans_clear, rdir_clear, tdd_clear, tdir_clear);
  }
  orc_adding_ica_sw(ng, nlev, w->incoming_sw, w->alb_diff, w->alb_dir, cos_sza, ref_clear, trans_clear, rdir_clear, tdd_clear,
                    tdir_clear, fu, fdd, fdir);
  for (int jl = 0; jl <= nlev; ++jl) {
    double s_up = 0.0, s_dd = 0.0, s_dir = 0.0;
    for (int g = 0; g < ng; ++g) { s_up = s_up + fu[(size_t)jl * ng + g]; s_dd = s_dd + fdd[(size_t)jl * ng + g]; s_dir = s_dir + fdir[(size_t)jl * ng + g]; }
    if (out->sw_up_clear) A2(out->sw_up_clear, jcol, jl) = s_up;
    if (out->sw_dn_clear) A2(out->sw_dn_clear, jcol, jl) = s_dd + s_dir;
    if (out->sw_dn_direct_clear) A2(out->sw_dn_direct_clear, jcol, jl) = s_dir;
  }
  for (int g = 0; g < ng; ++g) {
    if (out->sw_dn_diffuse_surf_clear_g) OUTG(out->sw_dn_diffuse_surf_clear_g, ng, g) = fdd[nl1 - ng + g];
    if (out->sw_dn_direct_surf_clear_g) OUTG(out->sw_dn_direct_surf_clear_g, ng, g) = fdir[nl1 - ng + g];
    if (out->sw_up_toa_clear_g) OUTG(out->sw_up_toa_clear_g, ng, g) = fu[g];
  }
  if (cloudless) {
    for (int jl = 0; jl <= nlev; ++jl) {
      if (out->sw_up) A2(out->sw_up, jcol, jl) = A2(out->sw_up_clear, jcol, jl);
      if (out->sw_dn) A2(out->sw_dn, jcol, jl) = A2(out->sw_dn_clear, jcol, jl);
      if (out->sw_dn_direct) A2(out->sw_dn_direct, jcol, jl) = A2(out->sw_dn_direct_clear, jcol, jl);
    }
    for (int g = 0; g < ng; ++g) {
      if (out->sw_dn_diffuse_surf_g) OUTG(out->sw_dn_diffuse_surf_g, ng, g) = fdd[nl1 - ng + g];
      if (out->sw_dn_direct_surf_g) OUTG(out->sw_dn_direct_surf_g, ng, g) = fdir[nl1 - ng + g];
      if (out->sw_up_toa_g) OUTG(out->sw_up_toa_g, ng, g) = fu[g];
    }
    band_profile(ng, NB_SW, nlev + 1, t->band_sw, fu, ncol, jcol, out->sw_up_band, 0);
    band_profile(ng, NB_SW, nlev + 1, t->band_sw, fdir, ncol, jcol, out->sw_dn_direct_band, 0);
    band_profile(ng, NB_SW, nlev + 1, t->band_sw, fdir, ncol, jcol, out->sw_dn_band, 0);
    band_profile(ng, NB_SW, nlev + 1, t->band_sw, fdd, ncol, jcol, out->sw_dn_band, 1);
    free(pool);
    return;
  }
  double tcc;
  double *fsd = (double*)malloc(sizeof(double) * (size_t)nlev * 2), *op = fsd + nlev;
  for (int jl = 0; jl < nlev; ++jl) fsd[jl] = A2(in->fractional_std, jcol, jl);
  for (int jl = 0; jl < nlev - 1; ++jl) op[jl] = A2(in->overlap_param, jcol, jl);
  orc_cloud_generator(t, ng, nlev, cfg->i_overlap_scheme, in->iseed[jcol], cfg->cloud_fraction_threshold, frac, op,
                      cfg->cloud_inhom_decorr_scaling, fsd, cfg->use_beta_overlap, cfg->use_vectorizable_generator, od_scaling, &tcc);
  free(fsd);
  if (out->cloud_cover_sw) out->cloud_cover_sw[jcol] = tcc;
  if (tcc >= cfg->cloud_fraction_threshold) {
    /* keep the clear-sky g-point surface/TOA fluxes before fu/fdd/fdir are overwritten */
    double* keep = (double*)malloc(sizeof(double) * 3 * (size_t)ng);
    for (int g = 0; g < ng; ++g) { keep[g] = fdd[nl1 - ng + g]; keep[ng + g] = fdir[nl1 - ng + g]; keep[2 * ng + g] = fu[g]; }
    for (int jl = 0; jl < nlev; ++jl) {
      if (frac[jl] >= cfg->cloud_fraction_threshold) {
        for (int g = 0; g < ng; ++g) {
          int jb = t->band_sw[g];
          size_t i = (size_t)jl * ng + g;
          double od_cloud_new = od_scaling[i] * w->od_sw_cloud[jl * NB_SW + jb];
          od_total[g] = w->od_sw[i] + od_cloud_new;
          ssa_total[g] = 0.0; g_total[g] = 0.0;
          if (od_total[g] > 0.0) {
            double scat_od = w->ssa_sw[i] * w->od_sw[i] + w->ssa_sw_cloud[jl * NB_SW + jb] * od_cloud_new;
            ssa_total[g] = scat_od / od_total[g];
            if (scat_od > 0.0)
              g_total[g] = (gzero[i] * w->ssa_sw[i] * w->od_sw[i] + w->g_sw_cloud[jl * NB_SW + jb] * w->ssa_sw_cloud[jl * NB_SW + jb] * od_cloud_new) / scat_od;
          }
        }
        if (cfg->do_sw_delta_scaling_with_gases) delta_eddington_vec(ng, od_total, ssa_total, g_total);   /* radiation_mcica_sw.F90:274-278 */
        orc_calc_ref_trans_sw(ng, cos_sza, od_total, ssa_total, g_total, ref + (size_t)jl * ng, trans + (size_t)jl * ng,
                              rdir + (size_t)jl * ng, tdd + (size_t)jl * ng, tdir + (size_t)jl * ng);
      } else {
        for (int g = 0; g < ng; ++g) {
          size_t i = (size_t)jl * ng + g;
          ref[i] = ref_clear[i]; trans[i] = trans_clear[i]; rdir[i] = rdir_clear[i]; tdd[i] = tdd_clear[i]; tdir[i] = tdir_clear[i];
        }
      }
    }
    orc_adding_ica_sw(ng, nlev, w->incoming_sw, w->alb_diff, w->alb_dir, cos_sza, ref, trans, rdir, tdd, tdir, fu, fdd, fdir);
    for (int jl = 0; jl <= nlev; ++jl) {
      double s_up = 0.0, s_dd = 0.0, s_dir = 0.0;
      for (int g = 0; g < ng; ++g) { s_up = s_up + fu[(size_t)jl * ng + g]; s_dd = s_dd + fdd[(size_t)jl * ng + g]; s_dir = s_dir + fdir[(size_t)jl * ng + g]; }
      if (out->sw_up) A2(out->sw_up, jcol, jl) = tcc * s_up + (1.0 - tcc) * A2(out->sw_up_clear, jcol, jl);
      if (out->sw_dn) A2(out->sw_dn, jcol, jl) = tcc * (s_dd + s_dir) + (1.0 - tcc) * A2(out->sw_dn_clear, jcol, jl);
      if (out->sw_dn_direct) A2(out->sw_dn_direct, jcol, jl) = tcc * s_dir + (1.0 - tcc) * A2(out->sw_dn_direct_clear, jcol, jl);
    }
    for (int g = 0; g < ng; ++g) {
      if (out->sw_dn_diffuse_surf_g) OUTG(out->sw_dn_diffuse_surf_g, ng, g) = tcc * fdd[nl1 - ng + g] + (1.0 - tcc) * keep[g];
      if (out->sw_dn_direct_surf_g) OUTG(out->sw_dn_direct_surf_g, ng, g) = tcc * fdir[nl1 - ng + g] + (1.0 - tcc) * keep[ng + g];
      if (out->sw_up_toa_g) OUTG(out->sw_up_toa_g, ng, g) = tcc * fu[g] + (1.0 - tcc) * keep[2 * ng + g];
    }
    free(keep);
  } else {
    for (int jl = 0; jl <= nlev; ++jl) {
      if (out->sw_up) A2(out->sw_up, jcol, jl) = A2(out->sw_up_clear, jcol, jl);
      if (out->sw_dn) A2(out->sw_dn, jcol, jl) = A2(out->sw_dn_clear, jcol, jl);
      if (out->sw_dn_direct) A2(out->sw_dn_direct, jcol, jl) = A2(out->sw_dn_direct_clear, jcol, jl);
    }
    for (int g = 0; g < ng; ++g) {
      if (out->sw_dn_diffuse_surf_g) OUTG(out->sw_dn_diffuse_surf_g, ng, g) = fdd[nl1 - ng + g];
      if (out->sw_dn_direct_surf_g) OUTG(out->sw_dn_direct_surf_g, ng, g) = fdir[nl1 - ng + g];
      if (out->sw_up_toa_g) OUTG(out->sw_up_toa_g, ng, g) = fu[g];
    }
  }
  free(pool);
}

/* ----- Homogeneous solvers: radiation_homogeneous_sw.F90:36-379, radiation_homogeneous_lw.F90:36-319.  Clouds fill the gridbox in
 * every layer whose (cropped) fraction reaches the threshold; cloud_optics has then computed gridbox-mean water paths
 * (config%is_homogeneous, radiation_cloud_optics.F90:318-327).  No LW aerosol scattering, no SW delta scaling with gases. ----- */
static void solver_homog(const orc_tables* t, const ecrad_b200_config* cfg, int ncol, int nlev, int jcol, const ecrad_b200_inputs* in,
                         ecrad_b200_outputs* out, const col_work* w, const double* frac, int sw) {
  const int ng = sw ? NG_SW : NG_LW, nb = sw ? NB_SW : NB_LW, nl1n = nlev + 1;
  const size_t nl = (size_t)nlev * ng, nl1 = (size_t)nl1n * ng;
  const int32_t* band = sw ? t->band_sw : t->band_lw;
  int is_cloudy_profile = 0;
  for (int jl = 0; jl < nlev; ++jl) if (frac[jl] >= cfg->cloud_fraction_threshold) is_cloudy_profile = 1;
  if (sw) {
    const double cos_sza = in->cos_sza[jcol];
    double* prof[6] = {out->sw_up, out->sw_dn, out->sw_dn_direct, out->sw_up_clear, out->sw_dn_clear, out->sw_dn_direct_clear};
    double* gs[6] = {out->sw_dn_diffuse_surf_g, out->sw_dn_direct_surf_g, out->sw_up_toa_g,
                     out->sw_dn_diffuse_surf_clear_g, out->sw_dn_direct_surf_clear_g, out->sw_up_toa_clear_g};
    double* bs[3] = {out->sw_up_band, out->sw_dn_band, out->sw_dn_direct_band};
    if (!(cos_sza > 0.0)) {
      for (int k = 0; k < 6; ++k) if (prof[k]) for (int jl = 0; jl <= nlev; ++jl) A2(prof[k], jcol, jl) = 0.0;
      for (int k = 0; k < 6; ++k) if (gs[k]) for (int g = 0; g < ng; ++g) OUTG(gs[k], ng, g) = 0.0;
      for (int k = 0; k < 3; ++k)
        if (bs[k]) for (int jl = 0; jl <= nlev; ++jl) for (int b = 0; b < nb; ++b) bs[k][((size_t)jl * ncol + jcol) * nb + b] = 0.0;
      return;
    }
    double* pool = (double*)malloc(sizeof(double) * (5 * nl + 3 * nl1 + 3 * (size_t)ng));
    double *ref = pool, *trans = ref + nl, *rdir = trans + nl, *tdd = rdir + nl, *tdir = tdd + nl;
    double *fu = tdir + nl, *fdd = fu + nl1, *fdir = fdd + nl1, *od_total = fdir + nl1, *ssa_total = od_total + ng, *g_total = ssa_total + ng;
    for (int jl = 0; jl < nlev; ++jl) {
      for (int g = 0; g < ng; ++g) { od_total[g] = w->od_sw[(size_t)jl * ng + g]; ssa_total[g] = w->ssa_sw[(size_t)jl * ng + g]; g_total[g] = w->g_sw[(size_t)jl * ng + g]; }
      if (cfg->do_sw_delta_scaling_with_gases) delta_eddington_vec(ng, od_total, ssa_total, g_total);   /* radiation_homogeneous_sw.F90:145-175 */
      orc_calc_reflectance_transmittance_sw(ng, cos_sza, od_total, ssa_total, g_total, ref + (size_t)jl * ng, trans + (size_t)jl * ng,
                                            rdir + (size_t)jl * ng, tdd + (size_t)jl * ng, tdir + (size_t)jl * ng);
    }
    for (int pass = 0; pass < 2; ++pass) {   /* 0: clear sky, 1: all sky */
      if (pass == 1) {
        if (!is_cloudy_profile) {   /* all-sky = clear-sky */
          for (int k = 0; k < 3; ++k) if (prof[k] && prof[k + 3]) for (int jl = 0; jl <= nlev; ++jl) A2(prof[k], jcol, jl) = A2(prof[k + 3], jcol, jl);
          for (int k = 0; k < 3; ++k) if (gs[k] && gs[k + 3]) for (int g = 0; g < ng; ++g) OUTG(gs[k], ng, g) = OUTG(gs[k + 3], ng, g);
          /* (band profiles: same sums as below on the unchanged clear-sky arrays) */
        } else {
          for (int jl = 0; jl < nlev; ++jl) {
            if (!(frac[jl] >= cfg->cloud_fraction_threshold)) continue;
            for (int g = 0; g < ng; ++g) {
              const size_t i = (size_t)jl * ng + g;
              const int ib = band[g];
              const double od_cloud_g = w->od_sw_cloud[jl * nb + ib];
              od_total[g] = w->od_sw[i] + od_cloud_g;
              ssa_total[g] = 0.0; g_total[g] = 0.0;
              if (od_total[g] > 0.0) ssa_total[g] = (w->ssa_sw[i] * w->od_sw[i] + w->ssa_sw_cloud[jl * nb + ib] * od_cloud_g) / od_total[g];
              if (ssa_total[g] > 0.0 && od_total[g] > 0.0)
                g_total[g] = (w->g_sw[i] * w->ssa_sw[i] * w->od_sw[i] + w->g_sw_cloud[jl * nb + ib] * w->ssa_sw_cloud[jl * nb + ib] * od_cloud_g) /
                             (ssa_total[g] * od_total[g]);
            }
            if (cfg->do_sw_delta_scaling_with_gases) delta_eddington_vec(ng, od_total, ssa_total, g_total);   /* :257-259 */
            orc_calc_reflectance_transmittance_sw(ng, cos_sza, od_total, ssa_total, g_total, ref + (size_t)jl * ng, trans + (size_t)jl * ng,
                                                  rdir + (size_t)jl * ng, tdd + (size_t)jl * ng, tdir + (size_t)jl * ng);
          }
        }
      }
      if (pass == 1 && !is_cloudy_profile) { /* fluxes of the clear-sky pass are still in fu/fdd/fdir */ }
      else orc_adding_ica_sw(ng, nlev, w->incoming_sw, w->alb_diff, w->alb_dir, cos_sza, ref, trans, rdir, tdd, tdir, fu, fdd, fdir);
      if (!(pass == 1 && !is_cloudy_profile)) {
        double *o_up = prof[pass ? 0 : 3], *o_dn = prof[pass ? 1 : 4], *o_dir = prof[pass ? 2 : 5];
        for (int jl = 0; jl <= nlev; ++jl) {
          double s_up = 0.0, s_dd = 0.0, s_dir = 0.0;
          for (int g = 0; g < ng; ++g) { s_up = s_up + fu[(size_t)jl * ng + g]; s_dd = s_dd + fdd[(size_t)jl * ng + g]; s_dir = s_dir + fdir[(size_t)jl * ng + g]; }
          if (o_up) A2(o_up, jcol, jl) = s_up;
          if (o_dn) A2(o_dn, jcol, jl) = s_dd + s_dir;
          if (o_dir) A2(o_dir, jcol, jl) = s_dir;
        }
        for (int g = 0; g < ng; ++g) {
          if (gs[pass ? 0 : 3]) OUTG(gs[pass ? 0 : 3], ng, g) = fdd[nl1 - ng + g];
          if (gs[pass ? 1 : 4]) OUTG(gs[pass ? 1 : 4], ng, g) = fdir[nl1 - ng + g];
          if (gs[pass ? 2 : 5]) OUTG(gs[pass ? 2 : 5], ng, g) = fu[g];
        }
      }
      if (pass == 1) {   /* all-sky band profiles (the clear-sky ones are not part of the C-ABI) */
        band_profile(ng, nb, nl1n, band, fu, ncol, jcol, out->sw_up_band, 0);
        band_profile(ng, nb, nl1n, band, fdir, ncol, jcol, out->sw_dn_direct_band, 0);
        band_profile(ng, nb, nl1n, band, fdir, ncol, jcol, out->sw_dn_band, 0);
        band_profile(ng, nb, nl1n, band, fdd, ncol, jcol, out->sw_dn_band, 1);
      }
    }
    free(pool);
  } else {
    double* pool = (double*)malloc(sizeof(double) * (4 * nl + 2 * nl1 + 3 * (size_t)ng));
    double *ref = pool, *trans = ref + nl, *su = trans + nl, *sd = su + nl, *fu = sd + nl, *fd = fu + nl1;
    double *od_total = fd + nl1, *ssa_total = od_total + ng, *g_total = ssa_total + ng;
    const double* planck = w->planck_hl;
    orc_calc_no_scattering_transmittance_lw(ng * nlev, w->od_lw, planck, planck + ng, trans, su, sd);
    memset(ref, 0, sizeof(double) * nl);
    orc_calc_fluxes_no_scattering_lw(ng, nlev, trans, su, sd, w->lw_emission, w->lw_albedo, fu, fd);
    sum_g(ng, nl1n, fu, ncol, jcol, out->lw_up_clear);
    sum_g(ng, nl1n, fd, ncol, jcol, out->lw_dn_clear);
    for (int g = 0; g < ng; ++g) {
      if (out->lw_dn_surf_clear_g) OUTG(out->lw_dn_surf_clear_g, ng, g) = fd[nl1 - ng + g];
      if (out->lw_up_toa_clear_g) OUTG(out->lw_up_toa_clear_g, ng, g) = fu[g];
    }
    if (is_cloudy_profile) {
      for (int jl = 0; jl < nlev; ++jl) {
        if (!(frac[jl] >= cfg->cloud_fraction_threshold)) continue;
        for (int g = 0; g < ng; ++g) {
          const size_t i = (size_t)jl * ng + g;
          const int ib = band[g];
          const double od_cloud_g = w->od_lw_cloud[jl * nb + ib];
          od_total[g] = w->od_lw[i] + od_cloud_g;
          ssa_total[g] = 0.0; g_total[g] = 0.0;
          if (cfg->do_lw_cloud_scattering) {
            if (od_total[g] > 0.0) ssa_total[g] = w->ssa_lw_cloud[jl * nb + ib] * od_cloud_g / od_total[g];
            if (ssa_total[g] > 0.0 && od_total[g] > 0.0)
              g_total[g] = w->g_lw_cloud[jl * nb + ib] * w->ssa_lw_cloud[jl * nb + ib] * od_cloud_g / (ssa_total[g] * od_total[g]);
          }
        }
        if (cfg->do_lw_cloud_scattering)
          orc_calc_ref_trans_lw(ng, od_total, ssa_total, g_total, planck + (size_t)jl * ng, planck + (size_t)(jl + 1) * ng, ref + (size_t)jl * ng,
                                trans + (size_t)jl * ng, su + (size_t)jl * ng, sd + (size_t)jl * ng);
        else
          orc_calc_no_scattering_transmittance_lw(ng, od_total, planck + (size_t)jl * ng, planck + (size_t)(jl + 1) * ng, trans + (size_t)jl * ng,
                                                  su + (size_t)jl * ng, sd + (size_t)jl * ng);
      }
      if (cfg->do_lw_cloud_scattering) orc_adding_ica_lw(ng, nlev, ref, trans, su, sd, w->lw_emission, w->lw_albedo, fu, fd);
      else orc_calc_fluxes_no_scattering_lw(ng, nlev, trans, su, sd, w->lw_emission, w->lw_albedo, fu, fd);
    }
    sum_g(ng, nl1n, fu, ncol, jcol, out->lw_up);
    sum_g(ng, nl1n, fd, ncol, jcol, out->lw_dn);
    for (int g = 0; g < ng; ++g) {
      if (out->lw_dn_surf_g) OUTG(out->lw_dn_surf_g, ng, g) = fd[nl1 - ng + g];
      if (out->lw_up_toa_g) OUTG(out->lw_up_toa_g, ng, g) = fu[g];
    }
    band_profile(ng, nb, nl1n, band, fu, ncol, jcol, out->lw_up_band, 0);
    band_profile(ng, nb, nl1n, band, fd, ncol, jcol, out->lw_dn_band, 0);
    if (cfg->do_lw_derivatives && out->lw_derivatives) lw_derivatives(ng, nlev, ncol, jcol, trans, fu + nl1 - ng, 0.0, 0, out->lw_derivatives);
    free(pool);
  }
  (void)in;
}

/* ----- Tripleclouds wrappers: scatter the per-column results of tripleclouds.c into flux_type ----- */
static void solver_tc(const orc_tables* t, const ecrad_b200_config* cfg, int ncol, int nlev, int jcol, const ecrad_b200_inputs* in,
                      ecrad_b200_outputs* out, const col_work* w, const double* frac, int sw) {
  const int ng = sw ? NG_SW : NG_LW, nb = sw ? NB_SW : NB_LW, nl1 = nlev + 1;
  double* buf = (double*)calloc((size_t)7 * nl1 + 6 * ng + 3 * (size_t)nl1 * ng + 2 * nlev, sizeof(double));
  orc_tc_out o; memset(&o, 0, sizeof o);
  o.up = buf; o.dn = o.up + nl1; o.dn_direct = o.dn + nl1; o.up_clear = o.dn_direct + nl1; o.dn_clear = o.up_clear + nl1;
  o.dn_direct_clear = o.dn_clear + nl1; o.lw_deriv = o.dn_direct_clear + nl1;
  o.up_toa_g = o.lw_deriv + nl1; o.up_toa_clear_g = o.up_toa_g + ng; o.dn_diffuse_surf_g = o.up_toa_clear_g + ng;
  o.dn_direct_surf_g = o.dn_diffuse_surf_g + ng; o.dn_diffuse_surf_clear_g = o.dn_direct_surf_g + ng; o.dn_direct_surf_clear_g = o.dn_diffuse_surf_clear_g + ng;
  o.up_g_prof = o.dn_direct_surf_clear_g + ng; o.dn_dif_g_prof = o.up_g_prof + (size_t)nl1 * ng; o.dn_dir_g_prof = o.dn_dif_g_prof + (size_t)nl1 * ng;
  double *fsd = o.dn_dir_g_prof + (size_t)nl1 * ng, *op = fsd + nlev;
  for (int jl = 0; jl < nlev; ++jl) fsd[jl] = A2(in->fractional_std, jcol, jl);
  for (int jl = 0; jl < nlev - 1; ++jl) op[jl] = A2(in->overlap_param, jcol, jl);
  /* SPARTACUS also reads the half-level pressure/temperature (layer depth) and the cloud effective sizes */
  const int spartacus = (sw ? cfg->i_solver_sw : cfg->i_solver_lw) == ECRAD_SOLVER_SPARTACUS;
  double *sp = NULL, *p_hl = NULL, *t_hl = NULL, *ics = NULL, *iis = NULL;
  if (spartacus) {
    sp = (double*)malloc(sizeof(double) * (2 * (size_t)nl1 + 2 * (size_t)nlev));
    p_hl = sp; t_hl = p_hl + nl1;
    for (int jl = 0; jl < nl1; ++jl) { p_hl[jl] = A2(in->pressure_hl, jcol, jl); t_hl[jl] = A2(in->temperature_hl, jcol, jl); }
    if (in->inv_cloud_effective_size) { ics = t_hl + nl1; for (int jl = 0; jl < nlev; ++jl) ics[jl] = A2(in->inv_cloud_effective_size, jcol, jl); }
    if (in->inv_inhom_effective_size) { iis = t_hl + nl1 + nlev; for (int jl = 0; jl < nlev; ++jl) iis[jl] = A2(in->inv_inhom_effective_size, jcol, jl); }
  }
  if (!sw) {
    if (spartacus)
      orc_spartacus_lw(t, cfg, nlev, p_hl, t_hl, frac, fsd, op, ics, iis, w->od_lw, w->ssa_lw, w->g_lw, w->planck_hl, w->od_lw_cloud,
                       w->ssa_lw_cloud, w->g_lw_cloud, w->lw_emission, w->lw_albedo, &o);
    else
    orc_tripleclouds_lw(t, cfg, nlev, frac, fsd, op, w->od_lw, w->planck_hl, w->od_lw_cloud, w->ssa_lw_cloud, w->g_lw_cloud,
                        w->lw_emission, w->lw_albedo, &o);
    if (out->cloud_cover_lw) out->cloud_cover_lw[jcol] = o.cloud_cover;
    for (int jl = 0; jl < nl1; ++jl) {
      if (out->lw_up) A2(out->lw_up, jcol, jl) = o.up[jl];
      if (out->lw_dn) A2(out->lw_dn, jcol, jl) = o.dn[jl];
      if (out->lw_up_clear) A2(out->lw_up_clear, jcol, jl) = o.up_clear[jl];
      if (out->lw_dn_clear) A2(out->lw_dn_clear, jcol, jl) = o.dn_clear[jl];
      if (cfg->do_lw_derivatives && out->lw_derivatives) A2(out->lw_derivatives, jcol, jl) = o.lw_deriv[jl];
    }
    for (int g = 0; g < ng; ++g) {
      if (out->lw_dn_surf_g) OUTG(out->lw_dn_surf_g, ng, g) = o.dn_diffuse_surf_g[g];
      if (out->lw_dn_surf_clear_g) OUTG(out->lw_dn_surf_clear_g, ng, g) = o.dn_diffuse_surf_clear_g[g];
      if (out->lw_up_toa_g) OUTG(out->lw_up_toa_g, ng, g) = o.up_toa_g[g];
      if (out->lw_up_toa_clear_g) OUTG(out->lw_up_toa_clear_g, ng, g) = o.up_toa_clear_g[g];
    }
    band_profile(ng, nb, nl1, t->band_lw, o.up_g_prof, ncol, jcol, out->lw_up_band, 0);
    band_profile(ng, nb, nl1, t->band_lw, o.dn_dif_g_prof, ncol, jcol, out->lw_dn_band, 0);
  } else {
    const double mu0 = in->cos_sza[jcol];
    if (mu0 < 1.0e-10) {
      /* night: fluxes zeroed; cloud_cover_sw is still the overlap-matrix value (computed before the column loop) */
      double (*reg)[3] = malloc(sizeof(double[3]) * nlev), (*ods)[3] = malloc(sizeof(double[3]) * nlev);
      double (*U)[3][3] = malloc(sizeof(double[3][3]) * nl1), (*V)[3][3] = malloc(sizeof(double[3][3]) * nl1);
      extern void orc_region_properties(int, const double*, const double*, double, int, double (*)[3], double (*)[3]);
      extern void orc_overlap_matrices(int, double (*)[3], const double*, double, double, int, double (*)[3][3], double (*)[3][3], double*);
      orc_region_properties(nlev, frac, fsd, cfg->cloud_fraction_threshold, cfg->i_cloud_pdf_shape == ECRAD_PDF_LOGNORMAL, reg, ods);
      if (spartacus && cfg->n_regions == 2)   /* radiation_regions.F90:105-110 (see two_region_properties in spartacus.c) */
        for (int jl = 0; jl < nlev; ++jl) { reg[jl][1] = frac[jl]; reg[jl][0] = 1.0 - reg[jl][1]; reg[jl][2] = 0.0; }
      orc_overlap_matrices(nlev, reg, op, cfg->cloud_inhom_decorr_scaling, cfg->cloud_fraction_threshold, cfg->use_beta_overlap, U, V, &o.cloud_cover);
      free(reg); free(ods); free(U); free(V);
    } else if (spartacus) {
      orc_spartacus_sw(t, cfg, nlev, mu0, p_hl, t_hl, frac, fsd, op, ics, iis, w->od_sw, w->ssa_sw, w->g_sw, w->od_sw_cloud, w->ssa_sw_cloud,
                       w->g_sw_cloud, w->incoming_sw, w->alb_diff, w->alb_dir, &o);
    } else {
      orc_tripleclouds_sw(t, cfg, nlev, mu0, frac, fsd, op, w->od_sw, w->ssa_sw, w->g_sw, w->od_sw_cloud, w->ssa_sw_cloud, w->g_sw_cloud,
                          w->incoming_sw, w->alb_diff, w->alb_dir, &o);
    }
    if (out->cloud_cover_sw) out->cloud_cover_sw[jcol] = o.cloud_cover;
    for (int jl = 0; jl < nl1; ++jl) {
      if (out->sw_up) A2(out->sw_up, jcol, jl) = o.up[jl];
      if (out->sw_dn) A2(out->sw_dn, jcol, jl) = o.dn[jl];
      if (out->sw_dn_direct) A2(out->sw_dn_direct, jcol, jl) = o.dn_direct[jl];
      if (out->sw_up_clear) A2(out->sw_up_clear, jcol, jl) = o.up_clear[jl];
      if (out->sw_dn_clear) A2(out->sw_dn_clear, jcol, jl) = o.dn_clear[jl];
      if (out->sw_dn_direct_clear) A2(out->sw_dn_direct_clear, jcol, jl) = o.dn_direct_clear[jl];
    }
    for (int g = 0; g < ng; ++g) {
      if (out->sw_dn_diffuse_surf_g) OUTG(out->sw_dn_diffuse_surf_g, ng, g) = o.dn_diffuse_surf_g[g];
      if (out->sw_dn_direct_surf_g) OUTG(out->sw_dn_direct_surf_g, ng, g) = o.dn_direct_surf_g[g];
      if (out->sw_dn_diffuse_surf_clear_g) OUTG(out->sw_dn_diffuse_surf_clear_g, ng, g) = o.dn_diffuse_surf_clear_g[g];
      if (out->sw_dn_direct_surf_clear_g) OUTG(out->sw_dn_direct_surf_clear_g, ng, g) = o.dn_direct_surf_clear_g[g];
      if (out->sw_up_toa_g) OUTG(out->sw_up_toa_g, ng, g) = o.up_toa_g[g];
      if (out->sw_up_toa_clear_g) OUTG(out->sw_up_toa_clear_g, ng, g) = o.up_toa_clear_g[g];
      /* radiation_tripleclouds_sw.F90:444 (sunlit columns only; the other solvers never set it) */
      if (out->sw_dn_toa_g && !spartacus && !(mu0 < 1.0e-10)) OUTG(out->sw_dn_toa_g, ng, g) = w->incoming_sw[g] * mu0;
    }
    /* band profiles: up; dn = mu0*direct + diffuse; dn_direct = mu0*direct (radiation_tripleclouds_sw.F90:604-624) */
    if (out->sw_up_band || out->sw_dn_band || out->sw_dn_direct_band) {
      band_profile(ng, nb, nl1, t->band_sw, o.up_g_prof, ncol, jcol, out->sw_up_band, 0);
      band_profile(ng, nb, nl1, t->band_sw, o.dn_dir_g_prof, ncol, jcol, out->sw_dn_direct_band, 0);
      if (out->sw_dn_direct_band)
        for (int jl = 0; jl < nl1; ++jl) for (int b = 0; b < nb; ++b) out->sw_dn_direct_band[((size_t)jl * ncol + jcol) * nb + b] *= mu0;
      band_profile(ng, nb, nl1, t->band_sw, o.dn_dir_g_prof, ncol, jcol, out->sw_dn_band, 0);
      if (out->sw_dn_band) {
        for (int jl = 0; jl < nl1; ++jl) for (int b = 0; b < nb; ++b) out->sw_dn_band[((size_t)jl * ncol + jcol) * nb + b] *= mu0;
        band_profile(ng, nb, nl1, t->band_sw, o.dn_dif_g_prof, ncol, jcol, out->sw_dn_band, 1);
      }
    }
  }
  free(buf); free(sp);
}

/* radiation_flux.F90:579-660 calc_toa_spectral: band sums (indexed_sum) of the per-g-point top-of-atmosphere fluxes.  sw_dn_toa_g is only
 * set by the Tripleclouds solver (radiation_tripleclouds_sw.F90:444); for the other solvers the reference sums an array it never wrote. */
static void toa_spectral(const orc_tables* t, const ecrad_b200_config* cfg, int jcol, double mu0, ecrad_b200_outputs* out) {
  if (!cfg->do_toa_spectral_flux) return;
  for (int k = 0; k < 5; ++k) {
    const int sw = k < 3;
    if (sw ? !cfg->do_sw : !cfg->do_lw) continue;
    const double* src = k == 0 ? out->sw_dn_toa_g : k == 1 ? out->sw_up_toa_g : k == 2 ? out->sw_up_toa_clear_g : k == 3 ? out->lw_up_toa_g : out->lw_up_toa_clear_g;
    double* dst = k == 0 ? out->sw_dn_toa_band : k == 1 ? out->sw_up_toa_band : k == 2 ? out->sw_up_toa_clear_band : k == 3 ? out->lw_up_toa_band : out->lw_up_toa_clear_band;
    if (!src || !dst || ((k == 2 || k == 4) && !cfg->do_clear) || (k == 0 && cfg->i_solver_sw != ECRAD_SOLVER_TRIPLECLOUDS)) continue;
    if (k == 0 && mu0 < 1.0e-10) continue;   /* night column: sw_dn_toa_g was never set */
    const int ng = sw ? NG_SW : NG_LW, nb = sw ? NB_SW : NB_LW;
    const int32_t* band = sw ? t->band_sw : t->band_lw;
    double* o = dst + (size_t)jcol * nb;
    for (int b = 0; b < nb; ++b) o[b] = 0.0;
    for (int g = 0; g < ng; ++g) o[band[g]] = o[band[g]] + src[(size_t)jcol * ng + g];
  }
}

/* radiation_flux.F90:397-577 calc_surface_spectral (paths used by the test namelists) */
static void surface_spectral(const orc_tables* t, const ecrad_b200_config* cfg, int jcol, ecrad_b200_outputs* out) {
  if (cfg->do_sw && cfg->do_surface_sw_spectral_flux && out->sw_dn_surf_band && out->sw_dn_direct_surf_band &&
      out->sw_dn_diffuse_surf_g && out->sw_dn_direct_surf_g) {
    for (int pass = 0; pass < 2; ++pass) {
      double *dirb = pass ? out->sw_dn_direct_surf_clear_band : out->sw_dn_direct_surf_band;
      double *totb = pass ? out->sw_dn_surf_clear_band : out->sw_dn_surf_band;
      const double *dirg = pass ? out->sw_dn_direct_surf_clear_g : out->sw_dn_direct_surf_g;
      const double *difg = pass ? out->sw_dn_diffuse_surf_clear_g : out->sw_dn_diffuse_surf_g;
      if (!dirb || !totb || !dirg || !difg || (pass && !cfg->do_clear)) continue;
      double* db = dirb + (size_t)jcol * NB_SW; double* tb = totb + (size_t)jcol * NB_SW;
      for (int b = 0; b < NB_SW; ++b) { db[b] = 0.0; tb[b] = 0.0; }
      for (int g = 0; g < NG_SW; ++g) db[t->band_sw[g]] = db[t->band_sw[g]] + dirg[(size_t)jcol * NG_SW + g];
      for (int g = 0; g < NG_SW; ++g) tb[t->band_sw[g]] = tb[t->band_sw[g]] + difg[(size_t)jcol * NG_SW + g];
      for (int b = 0; b < NB_SW; ++b) tb[b] = tb[b] + db[b];
    }
  }
  if (cfg->do_sw && cfg->do_canopy_fluxes_sw && out->sw_dn_diffuse_surf_canopy && out->sw_dn_direct_surf_canopy &&
      out->sw_dn_surf_band && !cfg->do_nearest_spectral_sw_albedo && t->sw_albedo_weights) {
    const int nalb = cfg->n_albedo_sw;
    double* dif = out->sw_dn_diffuse_surf_canopy + (size_t)jcol * nalb;
    double* dir = out->sw_dn_direct_surf_canopy + (size_t)jcol * nalb;
    for (int a = 0; a < nalb; ++a) { dif[a] = 0.0; dir[a] = 0.0; }
    for (int jb = 0; jb < NB_SW; ++jb)
      for (int a = 0; a < nalb; ++a) {
        double wgt = t->sw_albedo_weights[(size_t)jb * nalb + a];
        if (wgt != 0.0) {
          dif[a] = dif[a] + wgt * out->sw_dn_surf_band[(size_t)jcol * NB_SW + jb];
          dir[a] = dir[a] + wgt * out->sw_dn_direct_surf_band[(size_t)jcol * NB_SW + jb];
        }
      }
    for (int a = 0; a < nalb; ++a) dif[a] = dif[a] - dir[a];
  }
  if (cfg->do_sw && cfg->do_canopy_fluxes_sw && out->sw_dn_diffuse_surf_canopy && out->sw_dn_direct_surf_canopy &&
      cfg->do_nearest_spectral_sw_albedo && t->i_albedo_from_band_sw && out->sw_dn_diffuse_surf_g && out->sw_dn_direct_surf_g) {
    /* radiation_flux.F90:479-497: indexed_sum of the per-g-point surface fluxes into the albedo intervals */
    const int nalb = cfg->n_canopy_bands_sw;
    double* dif = out->sw_dn_diffuse_surf_canopy + (size_t)jcol * nalb;
    double* dir = out->sw_dn_direct_surf_canopy + (size_t)jcol * nalb;
    for (int a = 0; a < nalb; ++a) { dif[a] = 0.0; dir[a] = 0.0; }
    for (int g = 0; g < NG_SW; ++g) {
      const int a = t->i_albedo_from_band_sw[t->band_sw[g]] - 1;
      dir[a] = dir[a] + out->sw_dn_direct_surf_g[(size_t)jcol * NG_SW + g];
      dif[a] = dif[a] + out->sw_dn_diffuse_surf_g[(size_t)jcol * NG_SW + g];
    }
  }
  if (cfg->do_lw && cfg->do_canopy_fluxes_lw && out->lw_dn_surf_canopy && out->lw_dn_surf_g &&
      cfg->do_nearest_spectral_lw_emiss && t->i_emiss_from_band_lw) {
    const int ne = cfg->n_canopy_bands_lw;
    double* c = out->lw_dn_surf_canopy + (size_t)jcol * ne;
    for (int a = 0; a < ne; ++a) c[a] = 0.0;
    for (int g = 0; g < NG_LW; ++g) {
      int a = t->i_emiss_from_band_lw[t->band_lw[g]] - 1;
      c[a] = c[a] + out->lw_dn_surf_g[(size_t)jcol * NG_LW + g];
    }
  }
  if (cfg->do_lw && cfg->do_canopy_fluxes_lw && out->lw_dn_surf_canopy && out->lw_dn_surf_g &&
      !cfg->do_nearest_spectral_lw_emiss && t->lw_emiss_weights) {
    /* radiation_flux.F90:545-571: band sums of lw_dn_surf_g, then the emissivity-interval weights */
    const int ne = cfg->n_emiss_lw;
    double lband[NB_LW];
    for (int b = 0; b < NB_LW; ++b) lband[b] = 0.0;
    for (int g = 0; g < NG_LW; ++g) lband[t->band_lw[g]] = lband[t->band_lw[g]] + out->lw_dn_surf_g[(size_t)jcol * NG_LW + g];
    double* c = out->lw_dn_surf_canopy + (size_t)jcol * ne;
    for (int a = 0; a < ne; ++a) c[a] = 0.0;
    for (int jb = 0; jb < NB_LW; ++jb)
      for (int a = 0; a < ne; ++a) {
        double wgt = t->lw_emiss_weights[(size_t)jb * ne + a];
        if (wgt != 0.0) c[a] = c[a] + wgt * lband[jb];
      }
  }
}

/* props != NULL: stop where radiation() calls save_radiative_properties (radiation_interface.F90:405-425) and copy the optical
 * properties out instead of solving */
static int radiation_column(const orc_tables* t, const ecrad_b200_config* cfg, int ncol, int nlev, int jcol,
                            const ecrad_b200_inputs* in, ecrad_b200_outputs* out, const ecrad_b200_radiative_properties* props) {
  col_work w;
  size_t n = 3 * (size_t)nlev * NG_LW + (size_t)(nlev + 1) * NG_LW + 2 * NG_LW + 3 * (size_t)nlev * NG_SW + 3 * NG_SW +
             3 * (size_t)nlev * NB_LW + 3 * (size_t)nlev * NB_SW + 6 * (size_t)nlev;
  w.w = (double*)malloc(sizeof(double) * n);
  double* p = w.w;
  w.od_lw = p; p += (size_t)nlev * NG_LW;  w.planck_hl = p; p += (size_t)(nlev + 1) * NG_LW;
  w.ssa_lw = p; p += (size_t)nlev * NG_LW;  w.g_lw = p; p += (size_t)nlev * NG_LW;
  for (size_t i = 0; i < 2 * (size_t)nlev * NG_LW; ++i) w.ssa_lw[i] = 0.0;
  w.lw_emission = p; p += NG_LW;           w.lw_albedo = p; p += NG_LW;
  w.od_sw = p; p += (size_t)nlev * NG_SW;  w.ssa_sw = p; p += (size_t)nlev * NG_SW;  w.g_sw = p; p += (size_t)nlev * NG_SW;
  w.incoming_sw = p; p += NG_SW; w.alb_dir = p; p += NG_SW; w.alb_diff = p; p += NG_SW;
  w.od_lw_cloud = p; p += (size_t)nlev * NB_LW; w.ssa_lw_cloud = p; p += (size_t)nlev * NB_LW; w.g_lw_cloud = p; p += (size_t)nlev * NB_LW;
  w.od_sw_cloud = p; p += (size_t)nlev * NB_SW; w.ssa_sw_cloud = p; p += (size_t)nlev * NB_SW; w.g_sw_cloud = p; p += (size_t)nlev * NB_SW;
  double *frac = p, *qliq = frac + nlev, *qice = qliq + nlev, *rel = qice + nlev, *rei = rel + nlev, *phl = rei + nlev;
  double* phl_full = (double*)malloc(sizeof(double) * (size_t)(nlev + 1));
  (void)phl;
  int rc = get_albedos(t, cfg, ncol, jcol, in, w.alb_dir, w.alb_diff, w.lw_albedo);
  if (rc) { free(w.w); free(phl_full); return rc; }
  gas_optics_column(t, cfg, ncol, nlev, jcol, in, w.lw_albedo, w.od_lw, w.planck_hl, w.lw_emission, w.od_sw, w.ssa_sw, w.incoming_sw);
  for (size_t i = 0; i < (size_t)nlev * NG_SW; ++i) w.g_sw[i] = 0.0;
  if (cfg->use_aerosols) add_aerosol_optics(t, cfg, ncol, nlev, jcol, in, w.od_lw, w.ssa_lw, w.g_lw, w.od_sw, w.ssa_sw, w.g_sw);
  for (int jl = 0; jl <= nlev; ++jl) phl_full[jl] = A2(in->pressure_hl, jcol, jl);
  if (cfg->do_clouds) {
    /* crop_cloud_fraction, radiation_cloud.F90:700-740 (mutates the caller's array) */
    for (int jl = 0; jl < nlev; ++jl) {
      double sum_mr = 0.0;
      sum_mr = sum_mr + A2(in->q_liq, jcol, jl);
      sum_mr = sum_mr + A2(in->q_ice, jcol, jl);
      if (A2(in->cloud_fraction, jcol, jl) < cfg->cloud_fraction_threshold || sum_mr < cfg->cloud_mixing_ratio_threshold)
        A2(in->cloud_fraction, jcol, jl) = 0.0;
      frac[jl] = A2(in->cloud_fraction, jcol, jl);
      qliq[jl] = A2(in->q_liq, jcol, jl); qice[jl] = A2(in->q_ice, jcol, jl);
      rel[jl] = A2(in->re_liq, jcol, jl); rei[jl] = A2(in->re_ice, jcol, jl);
    }
    if (cfg->use_general_cloud_optics)   /* radiation_interface.F90:378-392 */
      orc_general_cloud_optics(t, cfg, nlev, phl_full, frac, qliq, qice, rel, rei, w.od_lw_cloud, w.ssa_lw_cloud, w.g_lw_cloud,
                               w.od_sw_cloud, w.ssa_sw_cloud, w.g_sw_cloud);
    else
    {
      double* thl_full = (double*)malloc(sizeof(double) * (size_t)(nlev + 1));
      for (int jl = 0; jl <= nlev; ++jl) thl_full[jl] = A2(in->temperature_hl, jcol, jl);
      rc = orc_cloud_optics(t, cfg, nlev, phl_full, thl_full, frac, qliq, qice, rel, rei, w.od_lw_cloud, w.ssa_lw_cloud, w.g_lw_cloud,
                            w.od_sw_cloud, w.ssa_sw_cloud, w.g_sw_cloud);
      free(thl_full);
      if (rc) { free(w.w); free(phl_full); return rc; }
    }
  } else {
    for (int jl = 0; jl < nlev; ++jl) frac[jl] = 0.0;
  }
  /* SPARTACUS on RRTMG works on g-points reordered by approximately increasing gas optical depth (radiation_ifs_rrtm.F90:50-68,
   * :122-133, :167-177): the gas-optics arrays leave radiation_ifs_rrtm.F90:481-505, :571-590, :717-737, :838-845 in that order, every
   * later step (aerosols and albedos are element-wise, so permuting after them is the same thing) indexes bands through
   * i_band_from_reordered_g, and the per-g-point outputs of flux_type are in the reordered order too. */
  orc_tables* tp = NULL;
  const int reorder_lw = !t->is_ecckd_lw && cfg->do_lw && cfg->i_solver_lw == ECRAD_SOLVER_SPARTACUS;
  const int reorder_sw = !t->is_ecckd_sw && cfg->do_sw && cfg->i_solver_sw == ECRAD_SOLVER_SPARTACUS;
  if (reorder_lw || reorder_sw) {
    const orc_array* pl = orc_find(t, "i_g_from_reordered_g_lw");
    const orc_array* ps = orc_find(t, "i_g_from_reordered_g_sw");
    if ((reorder_lw && !pl) || (reorder_sw && !ps)) { fprintf(stderr, "oracle: i_g_from_reordered_g_lw/sw missing from the table directory\n"); free(w.w); free(phl_full); return 13; }
    tp = (orc_tables*)malloc(sizeof(orc_tables));
    memcpy(tp, t, sizeof(orc_tables));
    double* tmp = (double*)malloc(sizeof(double) * (size_t)(NG_LW > NG_SW ? NG_LW : NG_SW));
    if (reorder_lw) {
      const int32_t* perm = (const int32_t*)pl->data;
      const int ng = NG_LW;
      double* rows[4] = {w.od_lw, w.ssa_lw, w.g_lw, w.planck_hl};
      const int nrows[4] = {nlev, nlev, nlev, nlev + 1};
      for (int a = 0; a < 4; ++a)
        for (int r = 0; r < nrows[a]; ++r) {
          double* x = rows[a] + (size_t)r * ng;
          for (int j = 0; j < ng; ++j) tmp[j] = x[perm[j] - 1];
          memcpy(x, tmp, sizeof(double) * ng);
        }
      double* one[2] = {w.lw_emission, w.lw_albedo};
      for (int a = 0; a < 2; ++a) { for (int j = 0; j < ng; ++j) tmp[j] = one[a][perm[j] - 1]; memcpy(one[a], tmp, sizeof(double) * ng); }
      for (int j = 0; j < ng; ++j) tp->band_lw[j] = t->band_lw[perm[j] - 1];
    }
    if (reorder_sw) {
      const int32_t* perm = (const int32_t*)ps->data;
      const int ng = NG_SW;
      double* rows[3] = {w.od_sw, w.ssa_sw, w.g_sw};
      for (int a = 0; a < 3; ++a)
        for (int r = 0; r < nlev; ++r) {
          double* x = rows[a] + (size_t)r * ng;
          for (int j = 0; j < ng; ++j) tmp[j] = x[perm[j] - 1];
          memcpy(x, tmp, sizeof(double) * ng);
        }
      double* one[3] = {w.incoming_sw, w.alb_dir, w.alb_diff};
      for (int a = 0; a < 3; ++a) { for (int j = 0; j < ng; ++j) tmp[j] = one[a][perm[j] - 1]; memcpy(one[a], tmp, sizeof(double) * ng); }
      for (int j = 0; j < ng; ++j) tp->band_sw[j] = t->band_sw[perm[j] - 1];
    }
    free(tmp);
    t = tp;
  }
  if (props) {
    const size_t c = (size_t)jcol;
#define PUT(dst, src, n) do { if (props->dst) memcpy(props->dst + c * (size_t)(n), (src), sizeof(double) * (size_t)(n)); } while (0)
    if (cfg->do_lw) {
      PUT(planck_hl, w.planck_hl, (size_t)(nlev + 1) * NG_LW); PUT(lw_emission, w.lw_emission, NG_LW); PUT(lw_albedo, w.lw_albedo, NG_LW);
      PUT(od_lw, w.od_lw, (size_t)nlev * NG_LW); PUT(ssa_lw, w.ssa_lw, (size_t)nlev * NG_LW); PUT(g_lw, w.g_lw, (size_t)nlev * NG_LW);
      PUT(od_lw_cloud, w.od_lw_cloud, (size_t)nlev * NB_LW); PUT(ssa_lw_cloud, w.ssa_lw_cloud, (size_t)nlev * NB_LW); PUT(g_lw_cloud, w.g_lw_cloud, (size_t)nlev * NB_LW);
    }
    if (cfg->do_sw) {
      PUT(sw_albedo_direct, w.alb_dir, NG_SW); PUT(sw_albedo_diffuse, w.alb_diff, NG_SW); PUT(incoming_sw, w.incoming_sw, NG_SW);
      PUT(od_sw, w.od_sw, (size_t)nlev * NG_SW); PUT(ssa_sw, w.ssa_sw, (size_t)nlev * NG_SW); PUT(g_sw, w.g_sw, (size_t)nlev * NG_SW);
      PUT(od_sw_cloud, w.od_sw_cloud, (size_t)nlev * NB_SW); PUT(ssa_sw_cloud, w.ssa_sw_cloud, (size_t)nlev * NB_SW); PUT(g_sw_cloud, w.g_sw_cloud, (size_t)nlev * NB_SW);
    }
#undef PUT
    free(w.w); free(phl_full); free(tp);
    return 0;
  }
  if (cfg->do_lw && cfg->i_solver_lw == ECRAD_SOLVER_HOMOGENEOUS) solver_homog(t, cfg, ncol, nlev, jcol, in, out, &w, frac, 0);
  else if (cfg->do_lw) { if (cfg->i_solver_lw == ECRAD_SOLVER_TRIPLECLOUDS || cfg->i_solver_lw == ECRAD_SOLVER_SPARTACUS) solver_tc(t, cfg, ncol, nlev, jcol, in, out, &w, frac, 0); else solver_lw(t, cfg, ncol, nlev, jcol, in, out, &w, frac); }
  if (cfg->do_sw && cfg->i_solver_sw == ECRAD_SOLVER_HOMOGENEOUS) solver_homog(t, cfg, ncol, nlev, jcol, in, out, &w, frac, 1);
  else if (cfg->do_sw) { if (cfg->i_solver_sw == ECRAD_SOLVER_TRIPLECLOUDS || cfg->i_solver_sw == ECRAD_SOLVER_SPARTACUS) solver_tc(t, cfg, ncol, nlev, jcol, in, out, &w, frac, 1); else solver_sw(t, cfg, ncol, nlev, jcol, in, out, &w, frac); }
  surface_spectral(t, cfg, jcol, out);
  toa_spectral(t, cfg, jcol, (cfg->do_sw && in->cos_sza) ? in->cos_sza[jcol] : 1.0, out);
  free(w.w); free(phl_full); free(tp);
  return 0;
}

int orc_radiation(const orc_tables* t, const ecrad_b200_config* cfg, int ncol, int nlev, int istartcol, int iendcol,
                  const ecrad_b200_inputs* in, ecrad_b200_outputs* out, int nthreads) {
  const int lw_plain = cfg->i_solver_lw == ECRAD_SOLVER_MCICA || cfg->i_solver_lw == ECRAD_SOLVER_CLOUDLESS || cfg->i_solver_lw == ECRAD_SOLVER_SPARTACUS;
  if ((cfg->do_lw_aerosol_scattering && cfg->do_lw && (!lw_plain || t->is_ecckd_lw || !cfg->do_lw_cloud_scattering)) || (cfg->do_sw_delta_scaling_with_gases && cfg->do_sw && cfg->i_solver_sw == ECRAD_SOLVER_SPARTACUS) ||
      (cfg->use_vectorizable_generator && cfg->i_overlap_scheme == ECRAD_OVERLAP_EXP_EXP)) {
    fprintf(stderr, "oracle: configuration outside the restated path\n");
    return 10;
  }
  if ((cfg->i_gas_model_lw == ECRAD_GAS_ECCKD) != (t->is_ecckd_lw != 0) || (cfg->i_gas_model_sw == ECRAD_GAS_ECCKD) != (t->is_ecckd_sw != 0)) {
    fprintf(stderr, "oracle: the table directory does not hold the gas models the configuration names\n");
    return 14;
  }
  if ((t->is_ecckd_lw || t->is_ecckd_sw) && !cfg->use_general_cloud_optics) { fprintf(stderr, "oracle: ecCKD needs use_general_cloud_optics\n"); return 14; }
  if (!out->lw_up_clear || !out->lw_dn_clear || !out->sw_up_clear || !out->sw_dn_clear || !out->sw_dn_direct_clear) {
    fprintf(stderr, "oracle: clear-sky flux outputs are required (do_clear)\n");
    return 11;
  }
  if (cfg->use_aerosols && (!in->aerosol_mmr || !in->h2o_sat_liq || !t->aer_me_sw_phobic || !t->aer_iclass || !t->aer_itype)) {
    fprintf(stderr, "oracle: aerosol inputs/tables missing\n");
    return 12;
  }
  int err = 0;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) reduction(| : err)
#endif
  for (int jcol = istartcol - 1; jcol < iendcol; ++jcol) err |= radiation_column(t, cfg, ncol, nlev, jcol, in, out, NULL);
  (void)nthreads;
  return err;
}

/* the optical properties radiation() would hand to save_radiative_properties (radiation_interface.F90:405-425, radiation_save.F90:716-726),
 * columns istartcol..iendcol; in->cloud_fraction is cropped in place like in orc_radiation */
int orc_radiative_properties(const orc_tables* t, const ecrad_b200_config* cfg, int ncol, int nlev, int istartcol, int iendcol,
                             const ecrad_b200_inputs* in, const ecrad_b200_radiative_properties* props) {
  if (!props) return 12;
  if ((cfg->i_gas_model_lw == ECRAD_GAS_ECCKD) != (t->is_ecckd_lw != 0) || (cfg->i_gas_model_sw == ECRAD_GAS_ECCKD) != (t->is_ecckd_sw != 0)) return 14;
  int err = 0;
  for (int jcol = istartcol - 1; jcol < iendcol; ++jcol) err |= radiation_column(t, cfg, ncol, nlev, jcol, in, NULL, props);
  return err;
}
