// kernels.cuh -- launch interface between api.cu (host orchestration) and kernels.cu (sm_100a kernels).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cloud_core.h"
#include "ecckd_core.h"
#include "gas_core.h"
#include "sp_core.h"

namespace ecb {

// Device view of the inputs of radiation(): column-fastest arrays with leading dimension ld, already offset to the
// first column of the tile (element (c, j) at p[j*ld + c]).
struct DevIn {
  const double *cos_sza, *skin_t, *sw_albedo, *sw_albedo_direct, *lw_emissivity;
  const int32_t* iseed;
  const double *p_hl, *t_hl;
  const double* gas[9];  // h2o co2 ch4 n2o cfc11 cfc12 hcfc22 ccl4 o3 (mass mixing ratios with RRTMG, mole fractions with ecCKD)
  double* frac;          // in/out (cropped)
  const double *q_liq, *q_ice, *re_liq, *re_ice, *overlap, *fsd;
  const double *aerosol_mmr, *h2o_sat_liq;   // (ld, nlev, ntype), (ld, nlev); only with aerosols
  const double *inv_cloud_size, *inv_inhom_size;   // (ld, nlev) cloud%inv_cloud_effective_size / inv_inhom_effective_size (SPARTACUS); may be NULL
  double solar_irradiance;
  int ld;
};

// Device view of flux_type; NULL = not allocated.  Profiles (ld, nlev+1) column-fastest; *_g (ng, ld) g fastest.
struct DevOut {
  double *lw_up, *lw_dn, *lw_up_clear, *lw_dn_clear;
  double *sw_up, *sw_dn, *sw_dn_direct, *sw_up_clear, *sw_dn_clear, *sw_dn_direct_clear;
  double *lw_derivatives, *cloud_cover_lw, *cloud_cover_sw;
  double *lw_dn_surf_g, *lw_dn_surf_clear_g, *lw_up_toa_g, *lw_up_toa_clear_g;
  double *sw_dn_diffuse_surf_g, *sw_dn_direct_surf_g, *sw_dn_diffuse_surf_clear_g, *sw_dn_direct_surf_clear_g;
  double *sw_up_toa_g, *sw_up_toa_clear_g;
  double *sw_dn_surf_band, *sw_dn_direct_surf_band, *sw_dn_surf_clear_band, *sw_dn_direct_surf_clear_band;
  double *sw_dn_diffuse_surf_canopy, *sw_dn_direct_surf_canopy, *lw_dn_surf_canopy;
  double *lw_up_band, *lw_dn_band, *sw_up_band, *sw_dn_band, *sw_dn_direct_band;  // (nband, ld, nlev+1)
  double *sw_dn_toa_g;                                                            // (ng_sw, ld), Tripleclouds only
  double *sw_dn_toa_band, *sw_up_toa_band, *sw_up_toa_clear_band, *lw_up_toa_band, *lw_up_toa_clear_band;   // (nband, ld): calc_toa_spectral
  int ld;
};

// Read-only tables on the device.
struct DevTables {
  const GasMeta* meta;
  const double *lwtab, *swtab;
  const CloudMeta* cloud;
  const double* pdf_val;
  const double* sw_albedo_weights;     // (n_albedo_sw, 14)
  const int32_t* i_emiss_from_band_lw; // (n_bands_lw), 1-based; do_nearest_spectral_lw_emiss
  const double* lw_emiss_weights;      // (n_emiss_lw, n_bands_lw); !do_nearest_spectral_lw_emiss
  const AerMeta* aer;                  // aerosol optics (NULL tables if no aerosols)
  const double* aertab;
  const CkdMeta* ckd;                  // ecCKD gas optics + generalised cloud optics (NULL with RRTMG)
  const double* ckdtab;
};

// Scalars of config_type the kernels read.
struct DevCfg {
  int solver_sw, solver_lw, overlap_scheme;
  int do_sw, do_lw, do_clouds, do_lw_cloud_scattering, do_lw_derivatives;
  int do_sw_delta_scaling_with_gases, do_fu_lw_ice_optics_bug, use_beta_overlap;
  int do_lw_aerosol_scattering;   // RRTMG + McICA / Cloudless (scan solvers): gas + aerosol ssa_lw, g_lw per g-point
  int do_surface_sw_spectral_flux, do_canopy_fluxes_sw, do_canopy_fluxes_lw, do_clear;
  int n_albedo_sw, n_emiss_lw, n_canopy_bands_sw, n_canopy_bands_lw;
  int use_aerosols, n_aerosol_types;
  int do_save_spectral_flux;
  int use_vectorizable_generator;
  int do_nearest_spectral_lw_emiss;
  int ckd_lw, ckd_sw;            // this spectrum's gas optics is ecCKD (else RRTMG-IFS); one of each = mixed gas models
  int gas_mmr;                   // the gas arrays hold mass mixing ratios (any RRTMG spectrum, set_gas_units radiation_interface.F90:164-186): ecCKD scales them itself
  int use_general_cloud_optics;  // cloud optics from the generalised look-up tables (always with ecCKD; an option with RRTMG, per band)
  int do_toa_spectral_flux;
  int pdf_gamma;                 // config%i_cloud_pdf_shape == IPdfShapeGamma (regions of Tripleclouds / SPARTACUS)
  int is_homogeneous;            // config%is_homogeneous: Homogeneous solvers (gridbox-mean cloud water paths, clouds fill the box)
  int ckd_ngas_lw, ckd_nlut_lw, ckd_ngas_sw, ckd_nlut_sw;   // ecCKD: gases / look-up-table gases per model (shared-memory sizing)
  int ng_lw, ng_sw, nb_lw, nb_sw;   // spectral sizes: RRTMG 140/112/16/14; ecCKD ng = nb = 32/64/96
  double cloud_fraction_threshold, cloud_mixing_ratio_threshold, min_gas_od_lw, min_gas_od_sw, cloud_inhom_decorr_scaling;
  double solar_cycle_multiplier;   // single_level%spectral_solar_cycle_multiplier (ecCKD shortwave, use_spectral_solar_cycle); 0 = mean spectrum
  SpCfg sp;                      // SPARTACUS scalars
};

enum { LW_SCR_ARRAYS = 5, SW_SCR_ARRAYS = 10 };

// Per-column bookkeeping of the RRTMG gas optics (gas_col_kernel): LAYTROP of both spectra, ecRad layer index that supplies the solar
// source function of each shortwave band (-1: none).
struct GasCol { int laytrop_lw, laytrop_sw; int lsol[NB_SW]; };

// Per-tile scratch (nc = columns in the tile).
struct Work {
  double *od_lw, *planck, *emission, *lw_albedo;  // [nc][nlev][140], [nc][nlev+1][140], [nc][140], [nc][140]
  double *od_sw, *ssa_sw, *incoming;              // [nc][nlev][112] x2, [nc][112]
  double *g_sw;                                   // [nc][nlev][112] asymmetry factor of gas+aerosol (NULL without aerosols: g = 0)
  double *ssa_lw, *g_lw;                          // [nc][140][ls] gas + aerosol, do_lw_aerosol_scattering only (layout of od_lw; NULL otherwise)
  double *aer_sw, *aer_lw;                        // aerosol band optics [nc][nlev][3][14] (od, scat, scat*g), [nc][nlev][16] (absorption od)
  double *cl_lw, *cl_sw;                          // cloud optics per band [nc][nlev][3][16], [nc][nlev][3][14]
  double *cum, *pair, *opi;                       // [nlev][nc] column-fastest
  double* tcc;                                    // [nc]
  int *ibegin, *iend, *ict;                       // [nc]
  uint32_t *code_lw, *code_sw;                    // [nc][ng][nlev]
  double *scr_lw, *scr_sw;                        // [nc][LW_SCR_ARRAYS*nlev*140], [nc][SW_SCR_ARRAYS*nlev*112] (separate: LW and SW chains run concurrently)
  double *sw_band_dir;                            // [nc][nlev+1][14] mu0 * per-band direct-beam sums (spectral flux profiles)
  double *tc_reg, *tc_ods, *tc_u, *tc_v, *tc_cc;  // Tripleclouds: [nc][nlev][3] x2, [nc][nlev+1][9] x2, [nc]
  double *lev_lw, *lev_sw;                        // [nc][LWLEV_NF / SWLEV_NF][nlev] per-layer gas-optics state (gas_prep_kernel; gas_core.h lwlev_load)
  double *lw_sums, *lw_carry;                     // [nc][6][nlev+1], [nc][4][140] (LW kernels)
  double *sw_sums, *sw_carry;                     // [nc][6][nlev+1] g-point sums per half-level, [nc][4][112] per-g carries between SW kernels
  // Layout of the gas optical properties (od_*, ssa_sw, g_sw, planck), per spectrum: 0 = [column][layer][g] (g fastest: the
  // lanes-are-g-points kernels), 1 = [column][g][ls] (layer fastest, row stride ls >= nlev+1, a multiple of 4: the scan solvers,
  // whose warps read one contiguous row per g-point)
  int layout_b_lw, layout_b_sw, ls;
  uint8_t* gas_jp;                                // [nc][nlev] reference-pressure index jp (bits 0-6), longwave "below LAYTROP" flag (bit 7)
  GasCol* gas_col;                                // [nc]
  int* sunlit;                                    // [1 + nc] number of sunlit columns of the tile, then their indices (any order)
};

void init_generator_constants();   // once per process/device, before the first generator launch

// Launchers.  All enqueue on `st` and return the number of kernels launched.
int launch_gas_prep(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st);
int launch_aerosol(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st);
int launch_gas_lw(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st);
int launch_gas_sw(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st);
// band-wise RRTMG gas optics from shared-memory table images: gas_band.cu
int launch_gas_col(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st);
int launch_gas_lw_band(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st);
int launch_gas_sw_band(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st);
// ecCKD gas optics (+ per-g-point aerosol merge), generalised cloud optics: ecckd.cu
int launch_ckd_lw(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st);
int launch_ckd_sw(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st);
int launch_general_cloud_optics(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st);
int launch_cloud(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st);
size_t tc_scratch_doubles_lw(int nlev, int ng);
size_t tc_scratch_doubles_sw(int nlev, int ng);
int launch_tc_prep(const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st);
int launch_tc_lw(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, const Work& w, int nc, int nlev, cudaStream_t st);
int launch_tc_sw(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, const Work& w, int nc, int nlev, cudaStream_t st);
// SPARTACUS: solver_sp.cu
size_t sp_scratch_doubles_lw(int nlev, int ng);
size_t sp_scratch_doubles_sw(int nlev, int ng);
int launch_sp_lw(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, const Work& w, int nc, int nlev, cudaStream_t st);
int launch_sp_sw(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, const Work& w, int nc, int nlev, cudaStream_t st);
// Device view of ecrad_b200_radiative_properties for one tile: the reference's element order, column slowest; NULL = skipped.
struct DevProps {
  double *planck_hl, *lw_emission, *lw_albedo, *sw_albedo_direct, *sw_albedo_diffuse, *incoming_sw;
  double *od_lw, *ssa_lw, *g_lw, *od_sw, *ssa_sw, *g_sw;
  double *od_lw_cloud, *ssa_lw_cloud, *g_lw_cloud, *od_sw_cloud, *ssa_sw_cloud, *g_sw_cloud;
};
int launch_radprops_gather(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, const DevProps& p, int nc, int nlev, cudaStream_t st);   // save_radiative_properties
int launch_toa_spectral(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, int nc, bool sw, cudaStream_t st);   // flux%calc_toa_spectral
int scan_max_levels();   // most layers the scan solvers take (32 lanes x layers per lane)
int launch_solver_lw_scan(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, const Work& w, int nc, int nlev, cudaStream_t st);
int launch_solver_sw_scan(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, const Work& w, int nc, int nlev, cudaStream_t st);
int launch_solver_lw(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, const Work& w, int nc, int nlev, cudaStream_t st);
int launch_solver_sw(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, const Work& w, int nc, int nlev, cudaStream_t st);

}  // namespace ecb
