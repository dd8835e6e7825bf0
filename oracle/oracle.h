/* oracle.h -- CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This directory is the parity oracle: a plain-C, double-precision, one-column-at-a-time restatement of
 * ecRad's gas optics -> cloud optics -> McICA / Cloudless solvers, following the reference routine by routine
 * in the same operation order (each function cites the file:line it follows).  It is compiled with
 * -ffp-contract=off (no FMA contraction) like the reference's "Bit" build (cmake/ecrad_compile_flags.cmake:19).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (ecrad_b200/csrc) never includes, links or calls anything in here.
 *
 * Parity pin: tests/test_oracle_golden.py checks it against the reference's own golden outputs
 * test/ifs/ecrad_meridian_{cloudless,noaer}_out_REFERENCE.nc (float32, copied to tests/golden/).
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>
#include "../include/ecrad_b200.h"

#define NG_LW 140
#define NG_SW 112
#define NB_LW 16
#define NB_SW 14

/* ---- ecCKD model (radiation_ecckd.F90:34-118 ckd_model_type, radiation_ecckd_gas.F90:38-72 ckd_gas_type) ---- */
enum { ORC_CONC_NONE = 0, ORC_CONC_LINEAR = 1, ORC_CONC_LUT = 2, ORC_CONC_RELATIVE_LINEAR = 3 };
typedef struct {
  int code, dep, n_mole_frac;       /* code: radiation_gas_constants.F90:26-39, 0 = composite of well-mixed gases */
  double reference_mole_frac, log_mole_frac1, d_log_mole_frac;
  const double* molar_abs;          /* (ng, npress, ntemp [, n_mole_frac]) */
} orc_ckd_gas;
typedef struct {
  int ng, npress, ntemp, nplanck, ngas, is_sw;
  double log_pressure1, d_log_pressure, d_temperature, temperature1_planck, d_temperature_planck;
  const double *temperature1, *planck_function /*(ng, nplanck)*/, *norm_solar_irradiance, *rayleigh_molar_scat;
  const double* norm_amplitude_solar_irradiance;   /* read_spectral_solar_cycle, radiation_ecckd.F90:421-431; NULL if not in the directory */
  orc_ckd_gas gas[16];
} orc_ckd_model;
/* generalised cloud optics of one cloud type (radiation_general_cloud_optics_data.F90:33-62) */
typedef struct { int nre; double re0, dre; const double *mass_ext, *ssa, *asymmetry; /* (ng, nre) */ } orc_gco;

/* ---- named-array tables (ETB1 blob) ---- */
typedef struct { char name[48]; int dtype; int ndim; int64_t dims[4]; const void* data; } orc_array;
typedef struct orc_tables {
  int n; orc_array* arr; void* blob;
  /* resolved pointers, filled by orc_tables_resolve */
  const double *absa_lw[17], *absb_lw[17], *selfref_lw[17], *forref_lw[17], *fracrefa_lw[17], *fracrefb_lw[17];
  const double *ka_mn2_1, *kb_mn2_1, *ka_mn2o_3, *kb_mn2o_3, *ka_mo3_5, *ccl4_5, *cfc11adj_6, *cfc12_6, *ka_mco2_6;
  const double *ka_mco2_7, *kb_mco2_7, *ka_mco2_8, *kb_mco2_8, *ka_mn2o_8, *kb_mn2o_8, *ka_mo3_8, *cfc12_8, *cfc22adj_8;
  const double *ka_mn2o_9, *kb_mn2o_9, *ka_mo2_11, *kb_mo2_11, *ka_mco2_13, *ka_mco_13, *kb_mo3_13, *ka_mn2_15;
  const double *totplnk, *delwave, *preflog_lw, *tref_lw, *chi_mls;
  const int32_t *ngb_lw, *ngc_lw;
  const double *absa_sw[30], *absb_sw[30], *selfref_sw[30], *forref_sw[30], *sfluxref_sw[30], *rayl_sw[30], *raylc_sw[30];
  double strrat_sw[30]; int layreffr_sw[30]; int nfor_sw[30];
  const double *absch4_20, *abso3a_24, *abso3b_24, *raylac_24, *raylbc_24, *abso3a_25, *abso3b_25, *absco2_29, *absh2o_29;
  double givfac_23, scalekur_27;
  const double *preflog_sw, *tref_sw;
  const int32_t *ngb_sw, *ngc_sw;
  const double *liq_coeff_lw, *liq_coeff_sw, *ice_coeff_lw, *ice_coeff_sw;
  const double *pdf_val; int pdf_ncdf, pdf_nfsd; double pdf_fsd1, pdf_inv_fsd_interval;
  /* aerosol optics per band (config%aerosol_optics): phobic (nband, ntype), philic (nband, nrh, ntype) */
  const double *aer_me_sw_phobic, *aer_ssa_sw_phobic, *aer_g_sw_phobic, *aer_me_lw_phobic, *aer_ssa_lw_phobic, *aer_g_lw_phobic;
  const double *aer_me_sw_philic, *aer_ssa_sw_philic, *aer_g_sw_philic, *aer_me_lw_philic, *aer_ssa_lw_philic, *aer_g_lw_philic;
  const double *aer_rh_lower; int aer_nrh;
  const int32_t *aer_iclass, *aer_itype;  /* (n_aerosol_types): 0 ignored / 1 hydrophobic / 2 hydrophilic; 1-based type */
  const double *sw_albedo_weights;      /* (n_albedo_sw, n_bands_sw) */
  const int32_t *i_emiss_from_band_lw;  /* (n_bands_lw), 1-based */
  const int32_t *i_albedo_from_band_sw; /* (n_bands_sw), 1-based; do_nearest_spectral_sw_albedo */
  const double *lw_emiss_weights;       /* (n_emiss_lw, n_bands_lw), used when !do_nearest_spectral_lw_emiss */
  /* 0-based band of each g-point: RRTMG ngb-1 / ngb-16; ecCKD with per-g-point cloud/aerosol optics: identity */
  int32_t band_lw[256], band_sw[256];
  /* ecCKD gas optics + generalised cloud optics (blob of tools/extract_ecckd_tables.py) */
  double solar_cycle_multiplier;        /* single_level%spectral_solar_cycle_multiplier (orc_set_solar_cycle_multiplier), default 0 */
  int is_ecckd, is_ecckd_lw, is_ecckd_sw;   /* both spectra / this spectrum (mixed gas models: one of the two) */
  orc_ckd_model ckd_lw, ckd_sw;
  orc_gco gco_lw[2], gco_sw[2];         /* cloud types: 0 liquid (mie_droplet), 1 ice (baum-general-habit-mixture) */
} orc_tables;

orc_tables* orc_tables_load(const char* path);
int  orc_tables_add(orc_tables* t, const char* name, int dtype, int ndim, const int64_t* dims, const void* data);
int  orc_tables_resolve(orc_tables* t);
void orc_tables_free(orc_tables* t);
int  orc_set_solar_cycle_multiplier(orc_tables* t, double multiplier);   /* non-zero needs an ecCKD shortwave model with the amplitude table */
const orc_array* orc_find(const orc_tables* t, const char* name);

/* ---- per-column gas-optics state (RRTMG layer order: 1 = bottom) ---- */
typedef struct {
  /* rrtm_prepare_gases */
  double pavel, tavel, coldry, wbroad, wkl[8] /*1..7*/, wx[5] /*1..4*/;
  /* rrtm_setcoef_140gp */
  int jp, jt, jt1, indself, indfor, indminor;
  double fac00, fac01, fac10, fac11, forfac, forfrac, selffac, selffrac, scaleminor, scaleminorn2, minorfrac;
  double colh2o, colco2, colo3, coln2o, colch4, colo2, co2mult, colbrd;
  double rat_h2oco2, rat_h2oco2_1, rat_h2oo3, rat_h2oo3_1, rat_h2on2o, rat_h2on2o_1, rat_h2och4, rat_h2och4_1,
         rat_n2oco2, rat_n2oco2_1, rat_o3co2, rat_o3co2_1;
} orc_lay_lw;

typedef struct {
  int jp, jt, jt1, indself, indfor;
  double fac00, fac01, fac10, fac11, forfac, forfrac, selffac, selffrac;
  double colh2o, colco2, colo3, colch4, colo2, colmol;
} orc_lay_sw;

/* rrtmg_lw.c / rrtmg_sw.c */
void orc_prepare_gases(int nlev, const double* p_hl, const double* t_hl, const double* p_fl, const double* t_fl,
                       const double* q, const double* co2, const double* ch4, const double* n2o,
                       const double* cfc11, const double* cfc12, const double* hcfc22, const double* ccl4,
                       const double* o3, orc_lay_lw* lay);
void orc_setcoef_lw(const orc_tables* t, int nlev, orc_lay_lw* lay, int* laytrop);
void orc_taumol_lw(const orc_tables* t, int nlev, const orc_lay_lw* lay, int laytrop,
                   double* tau /*[nlev][140] rrtmg layer order*/, double* pfrac /*[nlev][140]*/);
void orc_setcoef_sw(const orc_tables* t, int nlev, const orc_lay_lw* gas, orc_lay_sw* lay, int* laytrop);
void orc_taumol_sw(const orc_tables* t, int nlev, const orc_lay_sw* lay, int laytrop,
                   double* od /*[nlev][112] rrtmg order*/, double* ssa, double* incsol /*[112]*/);

/* solvers.c */
void orc_calc_ref_trans_lw(int ng, const double* od, const double* ssa, const double* g, const double* planck_top,
                           const double* planck_bot, double* ref, double* trans, double* source_up, double* source_dn);
void orc_calc_no_scattering_transmittance_lw(int ng, const double* od, const double* planck_top,
                           const double* planck_bot, double* trans, double* source_up, double* source_dn);
void orc_calc_ref_trans_sw(int ng, double mu0, const double* od, const double* ssa, const double* g,
                           double* ref_diff, double* trans_diff, double* ref_dir, double* trans_dir_diff,
                           double* trans_dir_dir);
void orc_calc_reflectance_transmittance_sw(int ng, double mu0, const double* od, const double* ssa, const double* g,
                           double* ref_diff, double* trans_diff, double* ref_dir, double* trans_dir_diff,
                           double* trans_dir_dir);
void orc_adding_ica_sw(int ng, int nlev, const double* incoming, const double* alb_diff, const double* alb_dir,
                       double cos_sza, const double* ref, const double* trans, const double* ref_dir,
                       const double* trans_dir_diff, const double* trans_dir_dir,
                       double* flux_up, double* flux_dn_diffuse, double* flux_dn_direct);
void orc_calc_fluxes_no_scattering_lw(int ng, int nlev, const double* trans, const double* source_up,
                       const double* source_dn, const double* emission, const double* albedo,
                       double* flux_up, double* flux_dn);
void orc_adding_ica_lw(int ng, int nlev, const double* ref, const double* trans, const double* source_up, const double* source_dn,
                       const double* emission, const double* albedo_surf, double* flux_up, double* flux_dn);
void orc_fast_adding_ica_lw(int ng, int nlev, const double* ref, const double* trans, const double* source_up,
                       const double* source_dn, const double* emission, const double* albedo,
                       const int* is_clear_sky_layer, int i_cloud_top, const double* flux_dn_clear,
                       double* flux_up, double* flux_dn);

/* cloud.c */
/* returns 0, or -1 when the coefficient arrays of the configured liquid / ice model are not in the table directory */
int orc_cloud_optics(const orc_tables* t, const ecrad_b200_config* cfg, int nlev, const double* p_hl, const double* t_hl,
                     const double* frac, const double* q_liq, const double* q_ice, const double* re_liq,
                     const double* re_ice, double* od_lw, double* ssa_lw, double* g_lw,
                     double* od_sw, double* ssa_sw, double* g_sw);
void orc_cloud_generator(const orc_tables* t, int ng, int nlev, int i_overlap_scheme, int32_t iseed,
                         double frac_threshold, const double* frac, const double* overlap_param,
                         double decorrelation_scaling, const double* fractional_std, int use_beta_overlap,
                         int use_vectorizable_generator, double* od_scaling /*[nlev][ng]*/, double* total_cloud_cover);

/* ecckd.c */
void orc_ecckd_gas_optics_column(const orc_tables* t, const ecrad_b200_config* cfg, int ncol, int nlev, int jcol,
                                 const ecrad_b200_inputs* in, const double* lw_albedo, double* od_lw, double* planck_hl,
                                 double* lw_emission, double* od_sw, double* ssa_sw, double* incoming_sw);
void orc_general_cloud_optics(const orc_tables* t, const ecrad_b200_config* cfg, int nlev, const double* p_hl,
                              const double* frac, const double* q_liq, const double* q_ice, const double* re_liq,
                              const double* re_ice, double* od_lw, double* ssa_lw, double* g_lw,
                              double* od_sw, double* ssa_sw, double* g_sw);

/* tripleclouds.c */
typedef struct {
  double cloud_cover;
  double *up, *dn, *dn_direct, *up_clear, *dn_clear, *dn_direct_clear;                   /* [nlev+1] sums over g */
  double *up_toa_g, *up_toa_clear_g, *dn_diffuse_surf_g, *dn_direct_surf_g, *dn_diffuse_surf_clear_g, *dn_direct_surf_clear_g; /* [ng] */
  double *up_g_prof, *dn_dif_g_prof, *dn_dir_g_prof;   /* optional [nlev+1][ng]: per-g totals over regions (band profiles) */
  double *lw_deriv;                                    /* optional [nlev+1] */
} orc_tc_out;
void orc_tripleclouds_sw(const orc_tables* t, const ecrad_b200_config* cfg, int nlev, double mu0, const double* frac,
                         const double* fsd, const double* overlap_param, const double* od, const double* ssa, const double* g,
                         const double* od_cloud, const double* ssa_cloud, const double* g_cloud, const double* incoming,
                         const double* alb_diff, const double* alb_dir, orc_tc_out* o);
void orc_tripleclouds_lw(const orc_tables* t, const ecrad_b200_config* cfg, int nlev, const double* frac, const double* fsd,
                         const double* overlap_param, const double* od, const double* planck_hl, const double* od_cloud,
                         const double* ssa_cloud, const double* g_cloud, const double* emission, const double* albedo,
                         orc_tc_out* o);

/* spartacus.c (outputs as tripleclouds.c) */
void orc_spartacus_sw(const orc_tables* t, const ecrad_b200_config* cfg, int nlev, double mu0, const double* p_hl, const double* t_hl,
                      const double* frac, const double* fsd, const double* overlap_param, const double* inv_cloud_size,
                      const double* inv_inhom_size, const double* od, const double* ssa, const double* g, const double* od_cloud,
                      const double* ssa_cloud, const double* g_cloud, const double* incoming, const double* alb_diff,
                      const double* alb_dir, orc_tc_out* o);
void orc_spartacus_lw(const orc_tables* t, const ecrad_b200_config* cfg, int nlev, const double* p_hl, const double* t_hl,
                      const double* frac, const double* fsd, const double* overlap_param, const double* inv_cloud_size,
                      const double* inv_inhom_size, const double* od, const double* ssa, const double* g, const double* planck_hl,
                      const double* od_cloud, const double* ssa_cloud, const double* g_cloud, const double* emission, const double* albedo,
                      orc_tc_out* o);
void orc_expm(int m, double* a, int sw_pattern);                                     /* test hooks for the matrix routines */
void orc_fast_expm_exchange_3(double a, double b, double c, double d, double* r);

/* radiation.c */
int orc_radiation(const orc_tables* t, const ecrad_b200_config* cfg, int ncol, int nlev, int istartcol, int iendcol,
                  const ecrad_b200_inputs* in, ecrad_b200_outputs* out, int nthreads);
int orc_radiative_properties(const orc_tables* t, const ecrad_b200_config* cfg, int ncol, int nlev, int istartcol, int iendcol,
                             const ecrad_b200_inputs* in, const ecrad_b200_radiative_properties* props);
/* stage dump for localising differences: gas optics of ONE column (1-based jcol), ecRad level order */
int orc_gas_optics_column(const orc_tables* t, const ecrad_b200_config* cfg, int ncol, int nlev, int jcol,
                  const ecrad_b200_inputs* in, double* od_lw, double* planck_hl, double* lw_emission,
                  double* od_sw, double* ssa_sw, double* incoming_sw);
#endif
