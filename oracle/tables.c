/* tables.c -- loader for the "ETB1" table blob (format: ecrad_b200/tables.py).  Oracle = test infrastructure.
 * The arrays are what setup_radiation leaves in module storage (radiation_ifs_rrtm.F90:34-213 setup_gas_optics,
 * ifsrrtm/yoerrta*.F90, yoesrta*.F90); indices below use the Fortran shapes. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

typedef struct { char name[48]; int32_t dtype; int32_t ndim; int64_t dims[4]; int64_t offset; } etb_entry;

const orc_array* orc_find(const orc_tables* t, const char* name) {
  for (int i = 0; i < t->n; ++i) if (!strcmp(t->arr[i].name, name)) return &t->arr[i];
  return NULL;
}

int orc_tables_add(orc_tables* t, const char* name, int dtype, int ndim, const int64_t* dims, const void* data) {
  t->arr = (orc_array*)realloc(t->arr, sizeof(orc_array) * (size_t)(t->n + 1));
  orc_array* a = &t->arr[t->n++];
  memset(a, 0, sizeof(*a));
  strncpy(a->name, name, 47);
  a->dtype = dtype; a->ndim = ndim;
  for (int i = 0; i < 4; ++i) a->dims[i] = i < ndim ? dims[i] : 1;
  a->data = data;
  return 0;
}

static const double* D(const orc_tables* t, const char* fmt, int b) {
  char nm[64]; snprintf(nm, sizeof nm, fmt, b);
  const orc_array* a = orc_find(t, nm);
  return a ? (const double*)a->data : NULL;
}
static const double* Dreq(const orc_tables* t, const char* nm, int* err) {
  const orc_array* a = orc_find(t, nm);
  if (!a) { fprintf(stderr, "oracle: missing table %s\n", nm); *err = 1; return NULL; }
  return (const double*)a->data;
}

static void resolve_common(orc_tables* t, int* perr);

/* ecCKD blob (tools/extract_ecckd_tables.py): "ckd_{lw,sw}_*", "gco_{lw,sw}_{0,1}_*" */
static void resolve_ckd_model(orc_tables* t, const char* pre, orc_ckd_model* m, int* err) {
  char nm[64];
  snprintf(nm, sizeof nm, "%smeta", pre);
  const double* meta = Dreq(t, nm, err);
  if (!meta) return;
  m->ng = (int)meta[0]; m->npress = (int)meta[1]; m->ntemp = (int)meta[2]; m->nplanck = (int)meta[3]; m->ngas = (int)meta[4];
  m->log_pressure1 = meta[5]; m->d_log_pressure = meta[6]; m->d_temperature = meta[7];
  m->temperature1_planck = meta[8]; m->d_temperature_planck = meta[9]; m->is_sw = meta[10] != 0.0;
  snprintf(nm, sizeof nm, "%stemperature1", pre); m->temperature1 = Dreq(t, nm, err);
  if (m->is_sw) {
    snprintf(nm, sizeof nm, "%snorm_solar_irradiance", pre); m->norm_solar_irradiance = Dreq(t, nm, err);
    snprintf(nm, sizeof nm, "%srayleigh_molar_scat", pre); m->rayleigh_molar_scat = Dreq(t, nm, err);
    snprintf(nm, sizeof nm, "%snorm_amplitude_solar_irradiance", pre);
    { const orc_array* a = orc_find(t, nm); m->norm_amplitude_solar_irradiance = a ? (const double*)a->data : NULL; }
  } else {
    snprintf(nm, sizeof nm, "%splanck_function", pre); m->planck_function = Dreq(t, nm, err);
  }
  snprintf(nm, sizeof nm, "%sgas_meta", pre);
  const double* gm = Dreq(t, nm, err);
  for (int j = 0; gm && j < m->ngas && j < 16; ++j) {
    orc_ckd_gas* g = &m->gas[j];
    g->code = (int)gm[6 * j]; g->dep = (int)gm[6 * j + 1]; g->reference_mole_frac = gm[6 * j + 2];
    g->n_mole_frac = (int)gm[6 * j + 3]; g->log_mole_frac1 = gm[6 * j + 4]; g->d_log_mole_frac = gm[6 * j + 5];
    snprintf(nm, sizeof nm, "%sgas%d_molar_abs", pre, j); g->molar_abs = Dreq(t, nm, err);
  }
}

/* general cloud optics tables, "gco_{lw,sw}_{0,1}_*": on the g-points' bands for ecCKD, on the 16/14 RRTMG bands in the
 * RRTMG blob (tools/extract_rrtmg_tables.py general_cloud_tables; radiation_general_cloud_optics.F90:36-141) */
static void resolve_gco(orc_tables* t, int required, int* err) {
  for (int jt = 0; jt < 2; ++jt)
    for (int sw = 0; sw < 2; ++sw) {
      orc_gco* c = sw ? &t->gco_sw[jt] : &t->gco_lw[jt];
      char nm[64];
      snprintf(nm, sizeof nm, "gco_%s_%d_meta", sw ? "sw" : "lw", jt);
      if (!required && !orc_find(t, nm)) { c->nre = 0; c->mass_ext = c->ssa = c->asymmetry = NULL; continue; }
      const double* meta = Dreq(t, nm, err);
      if (meta) { c->nre = (int)meta[0]; c->re0 = meta[1]; c->dre = meta[2]; }
      snprintf(nm, sizeof nm, "gco_%s_%d_mass_ext", sw ? "sw" : "lw", jt); c->mass_ext = Dreq(t, nm, err);
      snprintf(nm, sizeof nm, "gco_%s_%d_ssa", sw ? "sw" : "lw", jt); c->ssa = Dreq(t, nm, err);
      snprintf(nm, sizeof nm, "gco_%s_%d_asymmetry", sw ? "sw" : "lw", jt); c->asymmetry = Dreq(t, nm, err);
    }
}

/* calc_incoming_sw, radiation_ecckd.F90:946-962: a non-zero multiplier needs the solar-cycle amplitude */
int orc_set_solar_cycle_multiplier(orc_tables* t, double multiplier) {
  if (multiplier != 0.0 && !(t->is_ecckd_sw && t->ckd_sw.norm_amplitude_solar_irradiance)) return 1;
  t->solar_cycle_multiplier = multiplier;
  return 0;
}

int orc_tables_resolve(orc_tables* t) {
  int err = 0;
  /* a spectrum uses ecCKD when its model is in the directory; a directory with one ecCKD spectrum next to the RRTMG tables is the
   * mixed configuration of radiation_interface.F90:333-355 (test/ifs/configCY49R1_mixed.nam) */
  t->is_ecckd_lw = orc_find(t, "ckd_lw_meta") != NULL;
  t->is_ecckd_sw = orc_find(t, "ckd_sw_meta") != NULL;
  t->is_ecckd = t->is_ecckd_lw && t->is_ecckd_sw;
  if (t->is_ecckd_lw) resolve_ckd_model(t, "ckd_lw_", &t->ckd_lw, &err);
  if (t->is_ecckd_sw) resolve_ckd_model(t, "ckd_sw_", &t->ckd_sw, &err);
  if (t->is_ecckd) {
    resolve_gco(t, 1, &err);
    for (int g = 0; g < 256; ++g) { t->band_lw[g] = g; t->band_sw[g] = g; }   /* radiation_ecckd_interface.F90:60-63 */
    resolve_common(t, &err);
    return err;
  }
  resolve_gco(t, 0, &err);
  for (int b = 1; b <= 16; ++b) {
    t->absa_lw[b] = D(t, "lw%d_ABSA", b);  t->absb_lw[b] = D(t, "lw%d_ABSB", b);
    t->selfref_lw[b] = D(t, "lw%d_SELFREF", b); t->forref_lw[b] = D(t, "lw%d_FORREF", b);
    t->fracrefa_lw[b] = D(t, "lw%d_FRACREFA", b); t->fracrefb_lw[b] = D(t, "lw%d_FRACREFB", b);
  }
  t->ka_mn2_1 = Dreq(t, "lw1_KA_MN2", &err);   t->kb_mn2_1 = Dreq(t, "lw1_KB_MN2", &err);
  t->ka_mn2o_3 = Dreq(t, "lw3_KA_MN2O", &err); t->kb_mn2o_3 = Dreq(t, "lw3_KB_MN2O", &err);
  t->ka_mo3_5 = Dreq(t, "lw5_KA_MO3", &err);   t->ccl4_5 = Dreq(t, "lw5_CCL4", &err);
  t->cfc11adj_6 = Dreq(t, "lw6_CFC11ADJ", &err); t->cfc12_6 = Dreq(t, "lw6_CFC12", &err);
  t->ka_mco2_6 = Dreq(t, "lw6_KA_MCO2", &err);
  t->ka_mco2_7 = Dreq(t, "lw7_KA_MCO2", &err); t->kb_mco2_7 = Dreq(t, "lw7_KB_MCO2", &err);
  t->ka_mco2_8 = Dreq(t, "lw8_KA_MCO2", &err); t->kb_mco2_8 = Dreq(t, "lw8_KB_MCO2", &err);
  t->ka_mn2o_8 = Dreq(t, "lw8_KA_MN2O", &err); t->kb_mn2o_8 = Dreq(t, "lw8_KB_MN2O", &err);
  t->ka_mo3_8 = Dreq(t, "lw8_KA_MO3", &err);   t->cfc12_8 = Dreq(t, "lw8_CFC12", &err);
  t->cfc22adj_8 = Dreq(t, "lw8_CFC22ADJ", &err);
  t->ka_mn2o_9 = Dreq(t, "lw9_KA_MN2O", &err); t->kb_mn2o_9 = Dreq(t, "lw9_KB_MN2O", &err);
  t->ka_mo2_11 = Dreq(t, "lw11_KA_MO2", &err); t->kb_mo2_11 = Dreq(t, "lw11_KB_MO2", &err);
  t->ka_mco2_13 = Dreq(t, "lw13_KA_MCO2", &err); t->ka_mco_13 = Dreq(t, "lw13_KA_MCO", &err);
  t->kb_mo3_13 = Dreq(t, "lw13_KB_MO3", &err); t->ka_mn2_15 = Dreq(t, "lw15_KA_MN2", &err);
  t->totplnk = Dreq(t, "lw_TOTPLNK", &err); t->delwave = Dreq(t, "lw_DELWAVE", &err);
  t->preflog_lw = Dreq(t, "lw_PREFLOG", &err); t->tref_lw = Dreq(t, "lw_TREF", &err);
  t->chi_mls = Dreq(t, "lw_CHI_MLS", &err);
  t->ngb_lw = (const int32_t*)Dreq(t, "lw_NGB", &err); t->ngc_lw = (const int32_t*)Dreq(t, "lw_NGC", &err);
  for (int b = 16; b <= 29; ++b) {
    t->absa_sw[b] = D(t, "sw%d_ABSA", b); t->absb_sw[b] = D(t, "sw%d_ABSB", b);
    t->selfref_sw[b] = D(t, "sw%d_SELFREFC", b); t->forref_sw[b] = D(t, "sw%d_FORREFC", b);
    t->sfluxref_sw[b] = D(t, "sw%d_SFLUXREFC", b);
    t->rayl_sw[b] = D(t, "sw%d_RAYL", b); t->raylc_sw[b] = D(t, "sw%d_RAYLC", b);
    char nm[64]; snprintf(nm, sizeof nm, "sw%d_FORREFC", b);
    const orc_array* a = orc_find(t, nm); t->nfor_sw[b] = a ? (int)a->dims[0] : 0;
    const double* s = D(t, b == 16 ? "sw%d_STRRAT1" : "sw%d_STRRAT", b); t->strrat_sw[b] = s ? s[0] : 0.0;
    snprintf(nm, sizeof nm, "sw%d_LAYREFFR", b);
    a = orc_find(t, nm); t->layreffr_sw[b] = a ? ((const int32_t*)a->data)[0] : 0;
  }
  t->absch4_20 = Dreq(t, "sw20_ABSCH4C", &err);
  t->abso3a_24 = Dreq(t, "sw24_ABSO3AC", &err); t->abso3b_24 = Dreq(t, "sw24_ABSO3BC", &err);
  t->raylac_24 = Dreq(t, "sw24_RAYLAC", &err);  t->raylbc_24 = Dreq(t, "sw24_RAYLBC", &err);
  t->abso3a_25 = Dreq(t, "sw25_ABSO3AC", &err); t->abso3b_25 = Dreq(t, "sw25_ABSO3BC", &err);
  t->absco2_29 = Dreq(t, "sw29_ABSCO2C", &err); t->absh2o_29 = Dreq(t, "sw29_ABSH2OC", &err);
  { const double* s = Dreq(t, "sw23_GIVFAC", &err); t->givfac_23 = s ? s[0] : 0; }
  { const double* s = Dreq(t, "sw27_SCALEKUR", &err); t->scalekur_27 = s ? s[0] : 0; }
  t->preflog_sw = Dreq(t, "sw_PREFLOG", &err); t->tref_sw = Dreq(t, "sw_TREF", &err);
  t->ngb_sw = (const int32_t*)Dreq(t, "sw_NGBSW", &err); t->ngc_sw = (const int32_t*)Dreq(t, "sw_NGC", &err);
  t->liq_coeff_lw = Dreq(t, "liq_coeff_lw", &err); t->liq_coeff_sw = Dreq(t, "liq_coeff_sw", &err);
  t->ice_coeff_lw = Dreq(t, "ice_coeff_lw", &err); t->ice_coeff_sw = Dreq(t, "ice_coeff_sw", &err);
  if (t->ngb_lw) for (int g = 0; g < NG_LW; ++g) t->band_lw[g] = t->ngb_lw[g] - 1;
  if (t->ngb_sw) for (int g = 0; g < NG_SW; ++g) t->band_sw[g] = t->ngb_sw[g] - 16;
  if (t->is_ecckd_lw) for (int g = 0; g < 256; ++g) t->band_lw[g] = g;
  if (t->is_ecckd_sw) for (int g = 0; g < 256; ++g) t->band_sw[g] = g;
  resolve_common(t, &err);
  return err;
}

/* tables shared by both gas models: McICA PDF look-up table, aerosol optics, config-derived surface weights */
static void resolve_common(orc_tables* t, int* perr) {
  int err = 0;
  t->pdf_val = Dreq(t, "pdf_val", &err);
  const orc_array* pv = orc_find(t, "pdf_val");
  const double* fsd = Dreq(t, "pdf_fsd", &err);
  if (pv && fsd) {
    /* radiation_pdf_sampler.F90:83-93 */
    t->pdf_ncdf = (int)pv->dims[0]; t->pdf_nfsd = (int)pv->dims[1];
    t->pdf_fsd1 = fsd[0]; t->pdf_inv_fsd_interval = 1.0 / (fsd[1] - fsd[0]);
  }
  {
    const char* nm[12] = {"aer_mass_ext_sw_phobic", "aer_ssa_sw_phobic", "aer_g_sw_phobic", "aer_mass_ext_lw_phobic", "aer_ssa_lw_phobic",
                          "aer_g_lw_phobic", "aer_mass_ext_sw_philic", "aer_ssa_sw_philic", "aer_g_sw_philic", "aer_mass_ext_lw_philic",
                          "aer_ssa_lw_philic", "aer_g_lw_philic"};
    const double** dst[12] = {&t->aer_me_sw_phobic, &t->aer_ssa_sw_phobic, &t->aer_g_sw_phobic, &t->aer_me_lw_phobic, &t->aer_ssa_lw_phobic,
                              &t->aer_g_lw_phobic, &t->aer_me_sw_philic, &t->aer_ssa_sw_philic, &t->aer_g_sw_philic, &t->aer_me_lw_philic,
                              &t->aer_ssa_lw_philic, &t->aer_g_lw_philic};
    for (int i = 0; i < 12; ++i) { const orc_array* a = orc_find(t, nm[i]); *dst[i] = a ? (const double*)a->data : NULL; }
    const orc_array* rh = orc_find(t, "aer_rh_lower");
    t->aer_rh_lower = rh ? (const double*)rh->data : NULL; t->aer_nrh = rh ? (int)rh->dims[0] : 0;
    const orc_array* ic = orc_find(t, "aerosol_iclass"); t->aer_iclass = ic ? (const int32_t*)ic->data : NULL;
    const orc_array* it = orc_find(t, "aerosol_itype");  t->aer_itype = it ? (const int32_t*)it->data : NULL;
  }
  const orc_array* w = orc_find(t, "sw_albedo_weights");
  t->sw_albedo_weights = w ? (const double*)w->data : NULL;
  { const orc_array* ia = orc_find(t, "i_albedo_from_band_sw"); t->i_albedo_from_band_sw = ia ? (const int32_t*)ia->data : NULL; }
  const orc_array* e = orc_find(t, "i_emiss_from_band_lw");
  t->i_emiss_from_band_lw = e ? (const int32_t*)e->data : NULL;
  const orc_array* ew = orc_find(t, "lw_emiss_weights");
  t->lw_emiss_weights = ew ? (const double*)ew->data : NULL;
  if (err) *perr = 1;
}

orc_tables* orc_tables_load(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "oracle: cannot open %s\n", path); return NULL; }
  fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
  char* buf = (char*)malloc((size_t)sz);
  if (fread(buf, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(buf); return NULL; }
  fclose(f);
  if (memcmp(buf, "ETB1", 4)) { free(buf); fprintf(stderr, "oracle: %s is not ETB1\n", path); return NULL; }
  uint32_t n; memcpy(&n, buf + 4, 4);
  orc_tables* t = (orc_tables*)calloc(1, sizeof(orc_tables));
  t->blob = buf;
  const etb_entry* e = (const etb_entry*)(buf + 8);
  for (uint32_t i = 0; i < n; ++i) orc_tables_add(t, e[i].name, e[i].dtype, e[i].ndim, e[i].dims, buf + e[i].offset);
  return t;
}

void orc_tables_free(orc_tables* t) {
  if (!t) return;
  free(t->arr); free(t->blob); free(t);
}
