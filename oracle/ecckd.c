/* ecckd.c -- oracle restatement of the ecCKD gas optics and the generalised cloud optics.  TEST INFRASTRUCTURE.
 * Follows radiation/radiation_ecckd_interface.F90:174-324 (gas_optics), radiation_ecckd.F90:457-654
 * (calc_optical_depth_ckd_model), :900-928 (calc_planck_function), :935-964 (calc_incoming_sw),
 * radiation_general_cloud_optics.F90:134-287 (general_cloud_optics) and
 * radiation_general_cloud_optics_data.F90:249-330 (add_optical_properties), same operation order.
 * Gas arrays hold VOLUME mixing ratios here (set_gas_units of the ecCKD interface, radiation_ecckd_interface.F90:148-163).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#define A2(p, jcol, j) ((p)[(size_t)(j) * ncol + (jcol)])

static const double AccelDueToGravity = 9.80665;   /* radiation_constants.F90 */
static const double AirMolarMass = 28.970;         /* radiation_gas_constants.F90:42 */

static inline double dmin(double a, double b) { return a < b ? a : b; }
static inline double dmax(double a, double b) { return a > b ? a : b; }

/* mole fraction array of gas `code` (radiation_gas_constants.F90:26-39) in the C-ABI input struct, NULL if absent */
static const double* gas_array(const ecrad_b200_inputs* in, int code) {
  switch (code) {
    case 1: return in->h2o_mmr; case 2: return in->co2_mmr; case 3: return in->o3_mmr; case 4: return in->n2o_mmr;
    case 6: return in->ch4_mmr; case 8: return in->cfc11_mmr; case 9: return in->cfc12_mmr; case 10: return in->hcfc22_mmr;
    case 11: return in->ccl4_mmr; default: return NULL;
  }
}

/* calc_optical_depth_ckd_model for one column; od (and rayleigh for SW) are [nlev][ng] */
/* GasMolarMass(0:12), radiation_gas_constants.F90:43-56 */
static const double GasMolarMass[13] = {0.0, 18.0152833, 44.011, 47.9982, 44.013, 28.0101, 16.043, 31.9988, 137.3686, 120.914, 86.469, 153.823, 46.0055};

/* concentration_scaling: NULL when the gas arrays hold volume mixing ratios; else gas%get_scaling(IVolumeMixingRatio)
 * (radiation_gas.F90:471-486: AirMolarMass / GasMolarMass for arrays in mass mixing ratio), indexed by gas code */
static void ckd_optical_depth(const orc_ckd_model* m, int ncol, int nlev, int jcol, const ecrad_b200_inputs* in,
                              const double* temperature_fl, const double* concentration_scaling, double* od, double* rayleigh) {
  const int ng = m->ng;
  const size_t sp = (size_t)ng, st = (size_t)ng * m->npress, sc = st * m->ntemp;   /* strides of pressure, temperature, concentration */
  const double global_multiplier = 1.0 / (AccelDueToGravity * 0.001 * AirMolarMass);
  for (int jl = 0; jl < nlev; ++jl) {
    const double p1 = A2(in->pressure_hl, jcol, jl), p2 = A2(in->pressure_hl, jcol, jl + 1);
    const double log_pressure_fl = log(0.5 * (p1 + p2));
    double pindex1 = (log_pressure_fl - m->log_pressure1) / m->d_log_pressure;
    pindex1 = 1.0 + dmax(0.0, dmin(pindex1, m->npress - 1.0001));
    const int ip1 = (int)pindex1;
    const double pw2 = pindex1 - ip1, pw1 = 1.0 - pw2;
    const double temperature1 = pw1 * m->temperature1[ip1 - 1] + pw2 * m->temperature1[ip1];
    double tindex1 = (temperature_fl[jl] - temperature1) / m->d_temperature;
    tindex1 = 1.0 + dmax(0.0, dmin(tindex1, m->ntemp - 1.0001));
    const int it1 = (int)tindex1;
    const double tw2 = tindex1 - it1, tw1 = 1.0 - tw2;
    const double simple_multiplier = global_multiplier * (p2 - p1);
    double* o = od + (size_t)jl * ng;
    for (int g = 0; g < ng; ++g) o[g] = 0.0;
    for (int jg = 0; jg < m->ngas; ++jg) {
      const orc_ckd_gas* gas = &m->gas[jg];
      const double* src = gas_array(in, gas->code);
      const double mf = src ? A2(src, jcol, jl) : 0.0;
      const double scaling = concentration_scaling ? concentration_scaling[gas->code] : 1.0;   /* local_concentration_scaling(igascode) */
      const double* k00 = gas->molar_abs + (size_t)(ip1 - 1) * sp + (size_t)(it1 - 1) * st;   /* (:, ip1, it1) */
      const double *k10 = k00 + sp, *k01 = k00 + st, *k11 = k00 + sp + st;
      if (gas->dep == ORC_CONC_LUT) {
        const double mole_frac1 = exp(gas->log_mole_frac1);
        const double log_conc = log(dmax(mf * scaling, mole_frac1));
        double cindex1 = (log_conc - gas->log_mole_frac1) / gas->d_log_mole_frac;
        cindex1 = 1.0 + dmax(0.0, dmin(cindex1, gas->n_mole_frac - 1.0001));
        const int ic1 = (int)cindex1;
        const double cw2 = cindex1 - ic1, cw1 = 1.0 - cw2;
        const size_t c0 = (size_t)(ic1 - 1) * sc, c1 = c0 + sc;
        const double mult = simple_multiplier * mf * scaling;
        const double w000 = cw1 * tw1 * pw1, w100 = cw1 * tw1 * pw2, w010 = cw1 * tw2 * pw1, w110 = cw1 * tw2 * pw2;
        const double w001 = cw2 * tw1 * pw1, w101 = cw2 * tw1 * pw2, w011 = cw2 * tw2 * pw1, w111 = cw2 * tw2 * pw2;
        for (int g = 0; g < ng; ++g)
          o[g] = o[g] + mult * (w000 * k00[c0 + g] + w100 * k10[c0 + g] + w010 * k01[c0 + g] + w110 * k11[c0 + g] +
                                w001 * k00[c1 + g] + w101 * k10[c1 + g] + w011 * k01[c1 + g] + w111 * k11[c1 + g]);
      } else {
        double multiplier;
        if (gas->dep == ORC_CONC_LINEAR) multiplier = simple_multiplier * mf * scaling;   /* :560-562 */
        else if (gas->dep == ORC_CONC_RELATIVE_LINEAR) multiplier = simple_multiplier * (mf * scaling - gas->reference_mole_frac);
        else multiplier = simple_multiplier;
        for (int g = 0; g < ng; ++g)
          o[g] = o[g] + multiplier * (tw1 * (pw1 * k00[g] + pw2 * k10[g]) + tw2 * (pw1 * k01[g] + pw2 * k11[g]));
      }
    }
    for (int g = 0; g < ng; ++g) o[g] = dmax(0.0, o[g]);
    if (rayleigh)
      for (int g = 0; g < ng; ++g) rayleigh[(size_t)jl * ng + g] = global_multiplier * (p2 - p1) * m->rayleigh_molar_scat[g];
  }
}

/* calc_planck_function for one temperature */
static void ckd_planck(const orc_ckd_model* m, double temperature, double* planck) {
  const int ng = m->ng;
  double tindex1 = (temperature - m->temperature1_planck) * (1.0 / m->d_temperature_planck);
  if (tindex1 >= 0) {
    tindex1 = 1.0 + tindex1;
    int it1 = (int)tindex1; if (it1 > m->nplanck - 1) it1 = m->nplanck - 1;
    const double tw2 = tindex1 - it1, tw1 = 1.0 - tw2;
    const double *a = m->planck_function + (size_t)(it1 - 1) * ng, *b = a + ng;
    for (int g = 0; g < ng; ++g) planck[g] = tw1 * a[g] + tw2 * b[g];
  } else {
    for (int g = 0; g < ng; ++g) planck[g] = m->planck_function[g] * (temperature / m->temperature1_planck);
  }
}

void orc_ecckd_gas_optics_column(const orc_tables* t, const ecrad_b200_config* cfg, int ncol, int nlev, int jcol,
                                 const ecrad_b200_inputs* in, const double* lw_albedo, double* od_lw, double* planck_hl,
                                 double* lw_emission, double* od_sw, double* ssa_sw, double* incoming_sw) {
  double* temperature_fl = (double*)malloc(sizeof(double) * (size_t)nlev);
  for (int jl = 0; jl < nlev; ++jl) {   /* pressure-weighted, radiation_ecckd_interface.F90:239-245 */
    const double p1 = A2(in->pressure_hl, jcol, jl), p2 = A2(in->pressure_hl, jcol, jl + 1);
    temperature_fl[jl] = (A2(in->temperature_hl, jcol, jl) * p1 + A2(in->temperature_hl, jcol, jl + 1) * p2) / (p1 + p2);
  }
  /* set_gas_units (radiation_interface.F90:164-186): mass mixing ratios as soon as one spectrum uses RRTMG; ecCKD then scales
   * (radiation_ecckd_interface.F90:249-255) */
  double scaling_buf[13];
  const double* scaling = NULL;
  if (cfg->i_gas_model_sw != ECRAD_GAS_ECCKD || cfg->i_gas_model_lw != ECRAD_GAS_ECCKD) {
    scaling_buf[0] = 1.0;
    for (int j = 1; j <= 12; ++j) scaling_buf[j] = 1.0 * AirMolarMass / GasMolarMass[j];
    scaling = scaling_buf;
  }
  if (cfg->do_sw && cfg->i_gas_model_sw == ECRAD_GAS_ECCKD) {
    const orc_ckd_model* m = &t->ckd_sw;
    const int ng = m->ng;
    ckd_optical_depth(m, ncol, nlev, jcol, in, temperature_fl, scaling, od_sw, ssa_sw);
    for (size_t i = 0; i < (size_t)nlev * ng; ++i) {   /* :272-279 */
      od_sw[i] = od_sw[i] + ssa_sw[i];
      ssa_sw[i] = ssa_sw[i] / od_sw[i];
    }
    if (t->solar_cycle_multiplier != 0.0 && m->norm_amplitude_solar_irradiance)   /* calc_incoming_sw, radiation_ecckd.F90:950-954 */
      for (int g = 0; g < ng; ++g)
        incoming_sw[g] = in->solar_irradiance * (m->norm_solar_irradiance[g] + t->solar_cycle_multiplier * m->norm_amplitude_solar_irradiance[g]);
    else
      for (int g = 0; g < ng; ++g) incoming_sw[g] = in->solar_irradiance * m->norm_solar_irradiance[g];
  }
  if (cfg->do_lw && cfg->i_gas_model_lw == ECRAD_GAS_ECCKD) {
    const orc_ckd_model* m = &t->ckd_lw;
    const int ng = m->ng;
    ckd_optical_depth(m, ncol, nlev, jcol, in, temperature_fl, scaling, od_lw, NULL);
    for (int jl = 0; jl <= nlev; ++jl) ckd_planck(m, A2(in->temperature_hl, jcol, jl), planck_hl + (size_t)jl * ng);
    ckd_planck(m, in->skin_temperature[jcol], lw_emission);
    for (int g = 0; g < ng; ++g) lw_emission[g] = lw_emission[g] * (1.0 - lw_albedo[g]);
  }
  free(temperature_fl);
}

/* add_optical_properties for one cloud type of one column (scattering form); od/scat_od/scat_g are [nlev][ng] */
static void gco_add(const orc_gco* c, int ng, int nlev, const double* frac, const double* water_path, const double* re,
                    double* od, double* scat_od, double* scat_g) {
  for (int jl = 0; jl < nlev; ++jl) {
    if (!(frac[jl] > 0.0)) continue;
    const double re_index = dmax(1.0, dmin(1.0 + (re[jl] - c->re0) / c->dre, c->nre - 0.0001));
    const int ire = (int)re_index;
    const double weight2 = re_index - ire, weight1 = 1.0 - weight2;
    const double *me1 = c->mass_ext + (size_t)(ire - 1) * ng, *me2 = me1 + ng;
    const double *ss1 = c->ssa + (size_t)(ire - 1) * ng, *ss2 = ss1 + ng;
    const double *as1 = c->asymmetry + (size_t)(ire - 1) * ng, *as2 = as1 + ng;
    for (int g = 0; g < ng; ++g) {
      const size_t i = (size_t)jl * ng + g;
      double od_local = water_path[jl] * (weight1 * me1[g] + weight2 * me2[g]);
      od[i] = od[i] + od_local;
      if (scat_od) {
        od_local = od_local * (weight1 * ss1[g] + weight2 * ss2[g]);
        scat_od[i] = scat_od[i] + od_local;
        scat_g[i] = scat_g[i] + od_local * (weight1 * as1[g] + weight2 * as2[g]);
      }
    }
  }
}
/* no-scattering form (:316-326): adds the ABSORPTION optical depth where water_path > 0 */
static void gco_add_absorption(const orc_gco* c, int ng, int nlev, const double* water_path, const double* re, double* od) {
  for (int jl = 0; jl < nlev; ++jl) {
    if (!(water_path[jl] > 0.0)) continue;
    const double re_index = dmax(1.0, dmin(1.0 + (re[jl] - c->re0) / c->dre, c->nre - 0.0001));
    const int ire = (int)re_index;
    const double weight2 = re_index - ire, weight1 = 1.0 - weight2;
    const double *me1 = c->mass_ext + (size_t)(ire - 1) * ng, *me2 = me1 + ng;
    const double *ss1 = c->ssa + (size_t)(ire - 1) * ng, *ss2 = ss1 + ng;
    for (int g = 0; g < ng; ++g) {
      const size_t i = (size_t)jl * ng + g;
      od[i] = od[i] + water_path[jl] * (weight1 * me1[g] + weight2 * me2[g]) * (1.0 - (weight1 * ss1[g] + weight2 * ss2[g]));
    }
  }
}

/* delta_eddington_extensive (elemental form, radiation_delta_eddington.h:46-69) */
static void delta_eddington_extensive(int n, double* od, double* scat_od, double* scat_od_g) {
  for (int i = 0; i < n; ++i) {
    double g = scat_od[i] > 0.0 ? scat_od_g[i] / scat_od[i] : 0.0;
    double f = g * g;
    od[i] = od[i] - scat_od[i] * f;
    scat_od[i] = scat_od[i] * (1.0 - f);
    scat_od_g[i] = scat_od[i] * g / (1.0 + g);
  }
}

void orc_general_cloud_optics(const orc_tables* t, const ecrad_b200_config* cfg, int nlev, const double* p_hl,
                              const double* frac, const double* q_liq, const double* q_ice, const double* re_liq,
                              const double* re_ice, double* od_lw, double* ssa_lw, double* g_lw,
                              double* od_sw, double* ssa_sw, double* g_sw) {
  const int nlw = cfg->n_bands_lw, nsw = cfg->n_bands_sw;
  memset(od_lw, 0, sizeof(double) * (size_t)nlev * nlw);
  memset(od_sw, 0, sizeof(double) * (size_t)nlev * nsw);
  memset(ssa_sw, 0, sizeof(double) * (size_t)nlev * nsw);
  memset(g_sw, 0, sizeof(double) * (size_t)nlev * nsw);
  memset(ssa_lw, 0, sizeof(double) * (size_t)nlev * nlw);
  memset(g_lw, 0, sizeof(double) * (size_t)nlev * nlw);
  double* water_path = (double*)malloc(sizeof(double) * (size_t)nlev);
  for (int jt = 0; jt < 2; ++jt) {
    const double* q = jt ? q_ice : q_liq;
    const double* re = jt ? re_ice : re_liq;
    for (int jl = 0; jl < nlev; ++jl)   /* in-cloud water path, :191-197 */
      water_path[jl] = ((cfg->do_sw && cfg->i_solver_sw == ECRAD_SOLVER_HOMOGENEOUS) || (cfg->do_lw && cfg->i_solver_lw == ECRAD_SOLVER_HOMOGENEOUS))
                           ? q[jl] * (p_hl[jl + 1] - p_hl[jl]) * (1.0 / AccelDueToGravity)   /* config%is_homogeneous */
                           : q[jl] * (p_hl[jl + 1] - p_hl[jl]) * (1.0 / (AccelDueToGravity * dmax(cfg->cloud_fraction_threshold, frac[jl])));
    if (cfg->do_lw) {
      if (cfg->do_lw_cloud_scattering) gco_add(&t->gco_lw[jt], nlw, nlev, frac, water_path, re, od_lw, ssa_lw, g_lw);
      else gco_add_absorption(&t->gco_lw[jt], nlw, nlev, water_path, re, od_lw);
    }
    if (cfg->do_sw) gco_add(&t->gco_sw[jt], nsw, nlev, frac, water_path, re, od_sw, ssa_sw, g_sw);
  }
  free(water_path);
  for (int jl = 0; jl < nlev; ++jl) {
    if (!(frac[jl] > 0.0)) continue;
    if (cfg->do_lw && cfg->do_lw_cloud_scattering) {
      double *o = od_lw + (size_t)jl * nlw, *s = ssa_lw + (size_t)jl * nlw, *g = g_lw + (size_t)jl * nlw;
      delta_eddington_extensive(nlw, o, s, g);
      for (int i = 0; i < nlw; ++i) { g[i] = g[i] / dmax(s[i], 1.0e-15); s[i] = s[i] / dmax(o[i], 1.0e-15); }
    }
    if (cfg->do_sw) {
      double *o = od_sw + (size_t)jl * nsw, *s = ssa_sw + (size_t)jl * nsw, *g = g_sw + (size_t)jl * nsw;
      if (!cfg->do_sw_delta_scaling_with_gases) delta_eddington_extensive(nsw, o, s, g);
      for (int i = 0; i < nsw; ++i) { g[i] = g[i] / dmax(s[i], 1.0e-15); s[i] = s[i] / dmax(o[i], 1.0e-15); }
    }
  }
}
