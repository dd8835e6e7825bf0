/* ecrad_b200.h -- C-ABI of the B200-native ecRad hot path (libecrad_b200.so).
 *
 * Drop-in boundary for the three public procedures of the reference's `module radiation_interface`
 * (radiation/radiation_interface.F90:29):
 *     setup_radiation(config)                     :37    -> ecrad_b200_setup      (called at its tail, :153)
 *     radiation(ncol,nlev,istartcol,iendcol,...)  :200   -> ecrad_b200_radiation  (replaces the body :318-505)
 *     (no finaliser in the reference)                    -> ecrad_b200_finalize
 * The Fortran host keeps its derived types; a ~150-line ISO_C_BINDING shim passes c_loc()
 * of their contiguous components.  No Fortran descriptors, no torch types: plain pointers, ints, doubles.
 *
 * Array conventions (identical to the reference's memory layout, so no host-side repacking):
 *   - every (ncol, n) array is column-fastest ("Fortran order"): element (jcol, j) at [ (j-1)*ncol + (jcol-1) ]
 *   - the leading dimension is always the FULL ncol; only columns istartcol..iendcol (1-based, inclusive,
 *     exactly the reference's arguments) are read or written
 *   - per-g-point outputs are (ng, ncol), g-point fastest, as in radiation_flux.F90:38-118
 *   - levels are ordered top-of-atmosphere first (the shim keeps the reference's radiation_reverse in Fortran,
 *     radiation_interface.F90:310-317)
 *   - NULL output pointer == "component not allocated" (allocation rules radiation_flux.F90:147-300)
 */
#ifndef ECRAD_B200_H
#define ECRAD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Enumerations: numeric values are those of the reference's `enum, bind(c)` blocks. */
enum { ECRAD_SOLVER_CLOUDLESS = 0, ECRAD_SOLVER_HOMOGENEOUS = 1, ECRAD_SOLVER_MCICA = 2,
       ECRAD_SOLVER_SPARTACUS = 3, ECRAD_SOLVER_TRIPLECLOUDS = 4 };        /* radiation_config.F90:51-54   */
enum { ECRAD_GAS_MONOCHROMATIC = 0, ECRAD_GAS_IFSRRTMG = 1, ECRAD_GAS_ECCKD = 2 }; /* radiation_config.F90:86-88 */
enum { ECRAD_OVERLAP_MAX_RAN = 0, ECRAD_OVERLAP_EXP_RAN = 1, ECRAD_OVERLAP_EXP_EXP = 2 }; /* radiation_cloud_cover.F90:30-33 */
enum { ECRAD_LIQ_SOCRATES = 1, ECRAD_LIQ_SLINGO = 2 };                    /* radiation_config.F90:108-112 (Jahangir = 3 and Nielsen = 4
                                                                             are not dispatched by radiation_cloud_optics.F90 either) */
enum { ECRAD_ICE_FU = 1, ECRAD_ICE_BARAN = 2, ECRAD_ICE_BARAN2016 = 3,
       ECRAD_ICE_BARAN2017 = 4, ECRAD_ICE_YI = 5 };                       /* radiation_config.F90:123-127 */
enum { ECRAD_PDF_LOGNORMAL = 0, ECRAD_PDF_GAMMA = 1 };                    /* radiation_config.F90:134-138 */
/* SPARTACUS shortwave entrapment, config%i_3d_sw_entrapment (radiation_config.F90:69-77) */
enum { ECRAD_ENTRAPMENT_ZERO = 0, ECRAD_ENTRAPMENT_EDGE_ONLY = 1, ECRAD_ENTRAPMENT_EXPLICIT = 2,
       ECRAD_ENTRAPMENT_EXPLICIT_NON_FRACTAL = 3, ECRAD_ENTRAPMENT_MAXIMUM = 4 };

/* POD copy of the scalars of `config_type` (radiation_config.F90:163-649) that the hot path reads. */
typedef struct ecrad_b200_config {
  int32_t struct_bytes;              /* = sizeof(ecrad_b200_config), ABI check                           */
  int32_t i_solver_sw, i_solver_lw;  /* config%i_solver_sw/lw                                            */
  int32_t i_gas_model_sw, i_gas_model_lw;
  int32_t i_overlap_scheme;
  int32_t i_liq_model, i_ice_model;
  int32_t do_sw, do_lw, do_sw_direct, do_clear, do_clouds, use_aerosols;
  int32_t do_lw_cloud_scattering, do_lw_aerosol_scattering, do_lw_derivatives;
  int32_t do_sw_delta_scaling_with_gases, do_fu_lw_ice_optics_bug;
  int32_t use_beta_overlap, use_vectorizable_generator;
  int32_t do_surface_sw_spectral_flux, do_canopy_fluxes_sw, do_canopy_fluxes_lw, do_save_spectral_flux;
  int32_t do_nearest_spectral_sw_albedo, do_nearest_spectral_lw_emiss;
  int32_t n_g_sw, n_g_lw, n_bands_sw, n_bands_lw;
  int32_t n_albedo_sw;               /* size(config%sw_albedo_weights,1) == size(single_level%sw_albedo,2) */
  int32_t n_emiss_lw;                /* size(single_level%lw_emissivity,2)                               */
  int32_t n_canopy_bands_sw, n_canopy_bands_lw;
  int32_t n_aerosol_types;           /* config%n_aerosol_types == size(aerosol%mixing_ratio,3); 0 without aerosols  */
  /* SPARTACUS (radiation_config.F90:225-411); the number of regions is n_regions at the end of the struct */
  int32_t do_3d_effects, i_3d_sw_entrapment, do_3d_lw_multilayer_effects;
  double cloud_fraction_threshold;   /* config%cloud_fraction_threshold      (default 1e-6)              */
  double cloud_mixing_ratio_threshold; /*                                     (default 1e-9)              */
  double min_gas_od_lw, min_gas_od_sw; /* radiation_config.F90:244-245                                   */
  double cloud_inhom_decorr_scaling;
  double max_gas_od_3d, max_cloud_od, max_3d_transfer_rate, min_cloud_effective_size;  /* defaults 8, 16, 10, 100 m */
  double overhead_sun_factor, overhang_factor, clear_to_thick_fraction;                /* defaults 0, 0, 0          */
  int32_t do_lw_side_emissivity, use_expm_everywhere;                                  /* defaults 1, 0             */
  int32_t do_toa_spectral_flux;      /* config%do_toa_spectral_flux (default 0) */
  int32_t i_cloud_pdf_shape;         /* config%i_cloud_pdf_shape (radiation_config.F90:134-138): 0 lognormal, 1 gamma (default); shapes the
                                      * two cloudy regions of Tripleclouds / SPARTACUS; McICA takes it through the 'pdf_val' table */
  int32_t n_regions;                 /* config%nregions (radiation_config.F90:268): 3 (default) or 2; read by the SPARTACUS solvers only
                                      * (Tripleclouds has three regions by construction, radiation_tripleclouds_sw.F90) */
  int32_t use_general_cloud_optics;  /* config%use_general_cloud_optics (radiation_config.F90:185): cloud optics from the look-up tables of
                                      * radiation_general_cloud_optics.F90 ("gco_*" tables) instead of the band parameterisations selected
                                      * by i_liq_model / i_ice_model.  Must be 1 with ecCKD; 0 or 1 with RRTMG-IFS (tables per band) */
} ecrad_b200_config;

/* Read-only tables: a directory of named arrays, Fortran element order.  Names are listed in DESIGN.md
 * ("table directory"); for the RRTMG path they are the module variables of ifsrrtm/yoerrta1..16.F90,
 * yoesrta16..29.F90, yoerrtwn.F90, yoerrtrf.F90, yoesrtwn.F90 plus config%cloud_optics%*, config%pdf_sampler%*,
 * config%sw_albedo_weights (or, with do_nearest_spectral_sw_albedo, config%i_albedo_from_band_sw as "i_albedo_from_band_sw"),
 * config%i_emiss_from_band_lw or config%lw_emiss_weights ("lw_emiss_weights", when
 * do_nearest_spectral_lw_emiss is false), and (with aerosols) config%aerosol_optics%{mass_ext,ssa,g}_{sw,lw}_
 * {phobic,philic}, %rh_lower, %iclass, %itype ("aer_*", "aerosol_iclass", "aerosol_itype").  Data are COPIED by ecrad_b200_setup. */
typedef struct ecrad_b200_tables ecrad_b200_tables;
ecrad_b200_tables* ecrad_b200_tables_create(void);
/* dtype: 0 = float64, 1 = int32.  dims[ndim] in Fortran order (dims[0] fastest).  Returns 0 on success. */
int  ecrad_b200_tables_add(ecrad_b200_tables* t, const char* name, int dtype, int ndim,
                           const int64_t* dims, const void* data);
/* For the ECCKD gas model the directory holds config%gas_optics_{lw,sw} ("ckd_lw_*", "ckd_sw_*": look-up-table grids,
 * molar_abs per gas, planck_function, norm_solar_irradiance, rayleigh_molar_scat), config%cloud_optics_{lw,sw}(1:2)
 * ("gco_*": mass_ext, ssa, asymmetry per g-point and effective radius) and the aerosol arrays per g-point; names and
 * shapes in tools/extract_ecckd_tables.py.
 * Load an "ETB1" blob written by tools/extract_rrtmg_tables.py / extract_ecckd_tables.py (stand-alone use without a Fortran host). */
int  ecrad_b200_tables_load_file(ecrad_b200_tables* t, const char* path);
/* Same from memory (e.g. the blob received by an MPI/NCCL broadcast from the rank that read it; mirrors MPL_BROADCAST in
 * ifsrrtm/rrtm_kgb1.F90:38-51). */
int  ecrad_b200_tables_load_memory(ecrad_b200_tables* t, const void* blob, int64_t nbytes);
void ecrad_b200_tables_free(ecrad_b200_tables* t);

/* Inputs of radiation(): components of single_level_type (radiation_single_level.F90:29-102),
 * thermodynamics_type (radiation_thermodynamics.F90:29-49), gas_type (radiation_gas.F90:36-79; mass mixing
 * ratios, i.e. after set_gas_units), cloud_type (radiation_cloud.F90:33-96). */
typedef struct ecrad_b200_inputs {
  int32_t struct_bytes;
  int32_t reserved;
  double solar_irradiance;            /* single_level%solar_irradiance                                   */
  const double* cos_sza;              /* (ncol)                                                          */
  const double* skin_temperature;     /* (ncol)                                                          */
  const double* sw_albedo;            /* (ncol, n_albedo_sw)   diffuse                                   */
  const double* sw_albedo_direct;     /* (ncol, n_albedo_sw)   or NULL -> same as diffuse                */
  const double* lw_emissivity;        /* (ncol, n_emiss_lw)                                              */
  const int32_t* iseed;               /* (ncol)                                                          */
  const double* pressure_hl;          /* (ncol, nlev+1)  Pa                                              */
  const double* temperature_hl;       /* (ncol, nlev+1)  K                                               */
  /* gas%mixing_ratio(:,:,IGAS) slices AFTER set_gas_units (radiation_interface.F90:164-193), each (ncol, nlev); gas codes
   * radiation_gas_constants.F90:26-39.  RRTMG-IFS: mass mixing ratios, kg/kg (radiation_ifs_rrtm.F90 set_gas_units);
   * ECCKD: volume mixing ratios, mol/mol (radiation_ecckd_interface.F90:148-163) -- the field names keep the RRTMG spelling. */
  const double* h2o_mmr; const double* co2_mmr; const double* o3_mmr; const double* n2o_mmr;
  const double* ch4_mmr; const double* cfc11_mmr; const double* cfc12_mmr; const double* hcfc22_mmr;
  const double* ccl4_mmr;
  double*       cloud_fraction;       /* (ncol, nlev)  IN/OUT: cropped like cloud%crop_cloud_fraction    */
  const double* q_liq;                /* (ncol, nlev)  cloud%mixing_ratio(:,:,1)                         */
  const double* q_ice;                /* (ncol, nlev)  cloud%mixing_ratio(:,:,2)                         */
  const double* re_liq;               /* (ncol, nlev)  cloud%effective_radius(:,:,1)  m                  */
  const double* re_ice;               /* (ncol, nlev)  cloud%effective_radius(:,:,2)  m                  */
  const double* overlap_param;        /* (ncol, nlev-1)                                                  */
  const double* fractional_std;       /* (ncol, nlev)                                                    */
  /* aerosol_type (radiation_aerosol.F90) and thermodynamics%h2o_sat_liq; only read when cfg.use_aerosols */
  const double* aerosol_mmr;          /* (ncol, nlev, n_aerosol_types)  aerosol%mixing_ratio, levels 1..nlev */
  const double* h2o_sat_liq;          /* (ncol, nlev)  saturation mass mixing ratio w.r.t. liquid (calc_saturation_wrt_liquid) */
  /* cloud%inv_cloud_effective_size, cloud%inv_inhom_effective_size (radiation_cloud.F90:74-88), m-1; only read by the
   * SPARTACUS solvers; NULL = not allocated (no 3D effects / inhomogeneity size = cloud size) */
  const double* inv_cloud_effective_size;   /* (ncol, nlev) */
  const double* inv_inhom_effective_size;   /* (ncol, nlev) */
} ecrad_b200_inputs;

/* Outputs: components of flux_type (radiation_flux.F90:38-118).  Any pointer may be NULL. */
typedef struct ecrad_b200_outputs {
  int32_t struct_bytes;
  int32_t reserved;
  double *lw_up, *lw_dn, *lw_up_clear, *lw_dn_clear;                 /* (ncol, nlev+1) */
  double *sw_up, *sw_dn, *sw_dn_direct;                              /* (ncol, nlev+1) */
  double *sw_up_clear, *sw_dn_clear, *sw_dn_direct_clear;            /* (ncol, nlev+1) */
  double *lw_derivatives;                                            /* (ncol, nlev+1) */
  double *cloud_cover_lw, *cloud_cover_sw;                           /* (ncol)         */
  double *lw_dn_surf_g, *lw_dn_surf_clear_g, *lw_up_toa_g, *lw_up_toa_clear_g;        /* (n_g_lw, ncol) */
  double *sw_dn_diffuse_surf_g, *sw_dn_direct_surf_g;                                   /* (n_g_sw, ncol) */
  double *sw_dn_diffuse_surf_clear_g, *sw_dn_direct_surf_clear_g;                       /* (n_g_sw, ncol) */
  double *sw_up_toa_g, *sw_up_toa_clear_g;                                              /* (n_g_sw, ncol) */
  double *sw_dn_surf_band, *sw_dn_direct_surf_band;                  /* (n_bands_sw, ncol) */
  double *sw_dn_surf_clear_band, *sw_dn_direct_surf_clear_band;      /* (n_bands_sw, ncol) */
  double *sw_dn_diffuse_surf_canopy, *sw_dn_direct_surf_canopy;      /* (n_canopy_bands_sw, ncol) */
  double *lw_dn_surf_canopy;                                         /* (n_canopy_bands_lw, ncol) */
  /* per-band profiles (n_bands, ncol, nlev+1), filled by the Cloudless and Tripleclouds solvers when do_save_spectral_flux
   * (the McICA solver has none in the reference either); with ECCKD n_bands == n_g, i.e. one profile per g-point */
  double *lw_up_band, *lw_dn_band, *sw_up_band, *sw_dn_band, *sw_dn_direct_band;
  /* top-of-atmosphere spectral fluxes, flux%calc_toa_spectral (radiation_flux.F90:579-660) when cfg.do_toa_spectral_flux: band sums of
   * the *_toa_g arrays above (which must then be allocated).  sw_dn_toa_g (n_g_sw, ncol) is only set by the Tripleclouds solver
   * in the reference (radiation_tripleclouds_sw.F90:444), and so is sw_dn_toa_band here. */
  double *sw_dn_toa_g;                                               /* (n_g_sw, ncol) */
  double *sw_dn_toa_band, *sw_up_toa_band, *sw_up_toa_clear_band;    /* (n_bands_sw, ncol) */
  double *lw_up_toa_band, *lw_up_toa_clear_band;                     /* (n_bands_lw, ncol) */
} ecrad_b200_outputs;

/* Create the device-side state: copies `cfg` and the tables to the GPU selected by cudaGetDevice().
 * Returns 0 on success; on failure *handle is NULL and ecrad_b200_last_error(NULL) describes why.
 * There is NO CPU fallback: without a CUDA device this fails. */
int ecrad_b200_setup(const ecrad_b200_config* cfg, const ecrad_b200_tables* tab, void** handle);

/* Host-buffer entry: H2D of columns istartcol..iendcol, kernels, D2H into `out` (and the cropped
 * cloud_fraction back into in->cloud_fraction).  Thread-safe per handle (internally serialised).
 * Error convention: non-zero return + message (reference: radiation_abort, utilities/radiation_io.F90:45-73). */
int ecrad_b200_radiation(void* handle, int ncol, int nlev, int istartcol, int iendcol,
                         const ecrad_b200_inputs* in, ecrad_b200_outputs* out);

/* Single-precision boundary for hosts built with JPRB = JPRM (ifsaux/parkind1.F90:40-49): the same structs, but every `double*`
 * member points to FLOAT data (iseed stays int32, solar_irradiance a double scalar).  The arrays cross PCIe as float (half the
 * bytes), are widened on the device, and the kernels compute in double precision exactly as for ecrad_b200_radiation; the outputs
 * are rounded to float on the device.  Results therefore equal the double-precision path rounded to float (within the 1e-3 W m-2
 * of the single-precision target by construction); a single-precision COMPUTE path does not exist. */
int ecrad_b200_radiation_sp(void* handle, int ncol, int nlev, int istartcol, int iendcol,
                            const ecrad_b200_inputs* in, ecrad_b200_outputs* out);

/* Device-resident entry: every pointer in `in`/`out` is a DEVICE pointer with leading dimension `ncol`,
 * all ncol columns are processed, work is enqueued on `cuda_stream` (a cudaStream_t) and not synchronised. */
int ecrad_b200_radiation_device(void* handle, int ncol, int nlev,
                                const ecrad_b200_inputs* in, ecrad_b200_outputs* out, void* cuda_stream);

/* Same with separate leading dimensions: input arrays are (ld_in, n), output arrays (ld_out, n) / (n, ld_out) / (nband, ld_out,
 * nlev+1), every pointer already offset to the caller's first column; ncol columns are processed.  This is the multi-GPU
 * entry of a host model that wants all fluxes in one place: rank r passes pointers into the column slice it owns of arrays
 * that live on ANOTHER GPU of the NVLink domain (peer-mapped: cudaIpcOpenMemHandle / cudaDeviceEnablePeerAccess), and the flux
 * kernels store their results straight into that slice over NVLink -- the gather of the reference's block decomposition
 * (driver/ecrad_driver.F90:345-354) fused into the kernels that produce the fluxes, no separate collective. */
int ecrad_b200_radiation_device_ld(void* handle, int ncol, int nlev, int ld_in, int ld_out,
                                   const ecrad_b200_inputs* in, ecrad_b200_outputs* out, void* cuda_stream);

/* Blocked (NPROMA) entry: the layout of the IFS-style drivers (driver/ifs_blocking.F90, driver/ecrad_ifs_driver_blocked.F90:367-481,
 * ifs/radiation_scheme.F90), in which every variable of a block of `nproma` columns is a group of consecutive "fields" of one array
 * zrgp(nproma, nfields, nblocks) -- column index fastest, then field, then block.  Here the fields are the arrays of
 * ecrad_b200_inputs / ecrad_b200_outputs: in_field[k] (k = position of the pointer in ecrad_b200_inputs, 0 = cos_sza ... 27 =
 * inv_inhom_effective_size) is the 0-based first field of that input in zrgp_in, or -1 if absent; it occupies as many fields as the
 * array has rows ((ncol, rows) arrays: rows; aerosol_mmr: nlev * n_aerosol_types, type slowest); iseed travels as a real like
 * every IFS field.  out_field[k] likewise for the 41 outputs in zrgp_out: (ncol, nlev+1) profiles take nlev+1 fields, (n, ncol)
 * arrays n fields, (nband, ncol, nlev+1) profiles (nlev+1) * nband fields (band fastest).  Columns beyond ncol_total in the last
 * block are ignored / left untouched.  One host-to-device copy of the whole input array, a device kernel that unpacks it into the
 * column-fastest arrays the radiation kernels read, the same kernels as ecrad_b200_radiation, a pack kernel, one copy back:
 * results are bit-identical to ecrad_b200_radiation on the same columns.  (cloud_fraction is cropped on the device only.) */
typedef struct ecrad_b200_block_layout {
  int32_t struct_bytes;
  int32_t nproma, nblocks, nfields_in, nfields_out;
  int32_t in_field[28];
  int32_t out_field[41];
  double  solar_irradiance;
} ecrad_b200_block_layout;
int ecrad_b200_radiation_blocked(void* handle, int ncol_total, int nlev, const ecrad_b200_block_layout* layout,
                                 const double* zrgp_in, double* zrgp_out);

/* The intermediate optical properties radiation() hands to save_radiative_properties when config%do_save_radiative_properties is
 * set (radiation_interface.F90:405-425, radiation_save.F90:716-1100): what the gas, aerosol and cloud optics produced for the
 * solvers.  All arrays are host arrays in the reference's element order, spectral index fastest, column slowest, and are written for
 * columns istartcol..iendcol only; a NULL member is skipped.  The g-point order is the solver's (reordered for SPARTACUS on RRTMG-IFS,
 * like every per-g-point array of flux_type).  ssa_lw / g_lw are defined with do_lw_aerosol_scattering only (zero otherwise, as the
 * reference leaves them); the cloud properties of layers without cloud are zero (except where the reference's no-scattering
 * generalised cloud optics leaves absorption in cropped layers that hold condensate: reproduced).  Night columns hold what the
 * reference holds: aerosols on zero gas optical depth with RRTMG-IFS (radiation_ifs_rrtm.F90:531-594), full values with ecCKD. */
typedef struct ecrad_b200_radiative_properties {
  double* planck_hl;          /* (n_g_lw, nlev+1, ncol) */
  double* lw_emission;        /* (n_g_lw, ncol) */
  double* lw_albedo;          /* (n_g_lw, ncol) */
  double* sw_albedo_direct;   /* (n_g_sw, ncol) */
  double* sw_albedo_diffuse;  /* (n_g_sw, ncol) */
  double* incoming_sw;        /* (n_g_sw, ncol) */
  double* od_lw;  double* ssa_lw;  double* g_lw;                      /* (n_g_lw, nlev, ncol) gases + aerosols */
  double* od_sw;  double* ssa_sw;  double* g_sw;                      /* (n_g_sw, nlev, ncol) */
  double* od_lw_cloud;  double* ssa_lw_cloud;  double* g_lw_cloud;    /* (n_bands_lw, nlev, ncol) in-cloud, delta-Eddington scaled */
  double* od_sw_cloud;  double* ssa_sw_cloud;  double* g_sw_cloud;    /* (n_bands_sw, nlev, ncol) */
} ecrad_b200_radiative_properties;
/* Runs the optics stages of radiation() (no solver) on the same inputs and copies the properties out.  A diagnostic: synchronous,
 * not pipelined.  cloud_fraction is NOT cropped in the caller's array by this call.  Returns 0 or an error (ecrad_b200_last_error). */
int ecrad_b200_save_radiative_properties(void* handle, int ncol, int nlev, int istartcol, int iendcol,
                                         const ecrad_b200_inputs* inputs, const ecrad_b200_radiative_properties* props);

/* single_level%spectral_solar_cycle_multiplier (radiation_single_level.F90:71; config%use_spectral_solar_cycle, radiation_config.F90:174)
 * for the radiation calls that follow: -1 = solar minimum, +1 = solar maximum, 0 (the default) = the mean spectrum.  ecCKD shortwave only:
 * incoming_sw = solar_irradiance * (norm_solar_irradiance + multiplier * norm_amplitude_solar_irradiance), calc_incoming_sw
 * radiation_ecckd.F90:935-964; the amplitude is the table 'ckd_sw_norm_amplitude_solar_irradiance' (read_spectral_solar_cycle, :295-451).
 * A non-zero multiplier without that table, or with RRTMG-IFS in the shortwave, is an error like in the reference. */
int ecrad_b200_set_solar_cycle_multiplier(void* handle, double multiplier);

/* Tuning/diagnostic options.  Returns 0 on success.
 *   "serial"            0/1: run a tile's kernels on one stream instead of the three overlapping chains (per-kernel timing)
 *   "tile_cols", "tile_cols_device"   columns per internal tile of the host / device entry
 *   "edge_cols", "tail_tiles", "tile_ramp"   host entry: columns in the first / last tile (default tile_cols / 4), number of such short
 *                       tiles at the end (1), tile sizes doubling from the edge size (0); measured defaults, DESIGN.md section 4c.
 *                       (Environment ECRAD_B200_TIMELINE=1 prints the H2D / kernel / D2H intervals of every tile of a call.)
 *   "register_host"     0/1: page-lock the caller's input and output arrays (cudaHostRegister, cached per pointer until
 *                       ecrad_b200_finalize or register_host = 0), so that ordinary (pageable) Fortran allocatables are copied
 *                       asynchronously like pinned memory; the arrays must stay allocated while registered
 *   "gas_variant"       RRTMG gas optics kernel per spectrum (bit 0 longwave, bit 1 shortwave): band-wise on TMA-staged
 *                       shared-memory table images / one CTA per column
 *   "scan_solvers"      0/1: McICA / Cloudless solvers with the adding method as warp scans (no scratch) / lanes = g-points */
int ecrad_b200_set_option(void* handle, const char* key, int value);

/* Measured fp64 multiply-add throughput of the current device in TFLOP/s (2 flops per FMA): the ALU roofline bench.py reports
 * next to the HBM one.  A measurement helper, not part of the radiation path. */
int ecrad_b200_measure_fp64(double* tflops);

/* Number of kernels launched by this handle so far (bench.py's gpu_launches). */
int64_t ecrad_b200_kernel_launches(void* handle);
/* Event-timed duration (ms) of the last call's kernels, by stage; returns number of stages written. */
int ecrad_b200_last_stage_ms(void* handle, float* ms, int max_stages);
const char* ecrad_b200_stage_name(int stage);

void ecrad_b200_finalize(void* handle);
const char* ecrad_b200_last_error(void* handle);
const char* ecrad_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ECRAD_B200_H */
