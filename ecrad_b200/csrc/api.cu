// api.cu -- the C-ABI of libecrad_b200.so (include/ecrad_b200.h): table directory, setup, radiation (host and
// device-resident entries), finalize.  Host orchestration only: tiling of the column range, H2D/D2H staging
// pipelined against the kernels on three streams, kernel launches.  There is NO CPU compute path in here.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/ecrad_b200.h"
#include "kernels.cuh"
#include "tables.h"

using namespace ecb;

namespace {

thread_local std::string g_last_error;  // errors without a handle (setup failures)

struct Buf {   // grow-only device buffer
  void* p = nullptr; size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

enum { N_IN = 28, N_OUT = 41, N_STAGE = 5, N_WORK_MAX = 40 };
const char* kStageNames[N_STAGE] = {"gas_optics_lw", "gas_optics_sw", "cloud_optics_generator", "solver_lw", "solver_sw"};

struct Slot {            // device staging of one tile's inputs and outputs
  Buf in[N_IN], out[N_OUT];
  Buf in32[N_IN], out32[N_OUT];   // single-precision host arrays (ecrad_b200_radiation_sp): staged as float, converted on the device
  cudaEvent_t h2d_done = nullptr, compute_done = nullptr, d2h_done = nullptr;
  bool used = false;
};

struct Handle {
  ecrad_b200_config cfg;
  DevCfg dcfg;
  DevTables T;
  std::vector<void*> table_allocs;
  int device = 0;
  bool has_solar_cycle = false;  // the ecCKD shortwave model came with norm_amplitude_solar_irradiance
  int tile_cols = 4096;          // host entry: columns per tile (H2D / kernels / D2H of consecutive tiles overlap)
  int edge_cols = 1024;          // host entry: at most this many columns in the first and the last tile
  bool edge_explicit = false;    // edge_cols was set through set_option: not capped at tile_cols / 4
  int tail_tiles = 1;            // host entry: this many edge-sized tiles at the end of the call
  int tile_ramp = 0;             // host entry: 1 = tiles double from the edge size up to tile_cols (and halve again at the end)
  int tile_cols_device = 16384;  // device entry: only bounds the scratch (about 3 MB per column); bigger tiles = fewer partial waves
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  // Two compute sets (scratch + three streams + fork/join events): consecutive tiles of the host entry alternate between them,
  // so the kernels of tile t+1 fill the SMs that the tail of tile t leaves idle.
  cudaStream_t s_comp[2] = {nullptr, nullptr}, s_aux1[2] = {nullptr, nullptr}, s_aux2[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork[2] = {nullptr, nullptr}, ev_cloud[2] = {nullptr, nullptr}, ev_sw_done[2] = {nullptr, nullptr};
  // RRTMG gas optics, per spectrum (bit 0 longwave, bit 1 shortwave): band-wise kernel on TMA-staged shared-memory table images
  // (gas_band.cu) or one CTA per column gathering table rows through L1 (kernels.cu).  Measured on the B200 (10 000 columns):
  // longwave 3.05 vs 4.05 ms, shortwave 2.25 vs 1.95 ms -> default 1.
  int gas_variant = 1;
  cudaEvent_t ev_dev_done = nullptr;   // end of the last device-entry call: the next call (on any stream) waits for it before reusing the scratch
  bool dev_pending = false;
  int register_host = 0;  // host entry: page-lock the caller's arrays (cudaHostRegister, cached per pointer) so that pageable Fortran allocatables
                          // are copied asynchronously at full PCIe speed like pinned memory
  std::vector<std::pair<void*, size_t>> registered;
  int scan_solvers = 0;   // McICA / Cloudless solvers as warp scans (solver_scan.cu); 0: the lanes-are-g-points kernels (solver_sw.cu, solver_lw.cu)
  int serial = 0;   // 1: all kernels of a tile on one stream (per-kernel timing); 0: LW chain, SW chain and cloud chain overlap
  Slot slot[2];
  Buf work[2][N_WORK_MAX];
  Buf blocked;            // blocked entry: device copies of the two zrgp arrays + the column-layout arrays unpacked from / packed into them
  Work w[2];
  int w_cols[2] = {0, 0}, w_nlev[2] = {0, 0}, w_scan[2] = {-1, -1};
  std::mutex mu;
  std::string err;
  int64_t launches = 0;
  std::vector<cudaEvent_t> ev;       // stage boundary events of the last call: (N_STAGE+1) per tile
  int ev_tiles = 0;
};

int fail(Handle* h, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  if (h) h->err = buf;
  g_last_error = buf;
  return 1;
}
#define CK(h, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(h, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)
// Inside the pipelined host entry an error must not leave asynchronous copies into the caller's arrays in flight.
#define CKD(h, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { drain(h); return fail(h, "%s: %s", #call, cudaGetErrorString(e_)); } } while (0)

void unregister_all(Handle* h) {
  for (auto& r : h->registered) cudaHostUnregister(r.first);
  h->registered.clear();
  cudaGetLastError();
}
// Page-lock [p, p + bytes) once; a later call with the same pointer and a size that fits is free.  Failure is not an error: the
// copy then takes the (slower, staged) pageable route.
void register_range(Handle* h, const void* p, size_t bytes) {
  if (!p || !bytes) return;
  for (auto& r : h->registered)
    if (r.first == p) {
      if (r.second >= bytes) return;
      cudaHostUnregister(r.first); r.first = nullptr; r.second = 0;
    }
  if (cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterDefault) == cudaSuccess) {
    for (auto& r : h->registered) if (!r.first) { r = {const_cast<void*>(p), bytes}; return; }
    h->registered.push_back({const_cast<void*>(p), bytes});
  } else {
    cudaGetLastError();   // already registered by the caller (overlap), or not registrable: leave it
  }
}

// Wait for everything this handle has enqueued (error paths of the host entry: no copy may still be writing the caller's arrays
// or reading them once the call has returned).
void drain(Handle* h) {
  for (cudaStream_t q : {h->s_h2d, h->s_d2h, h->s_comp[0], h->s_comp[1], h->s_aux1[0], h->s_aux1[1], h->s_aux2[0], h->s_aux2[1]})
    if (q) cudaStreamSynchronize(q);
  cudaGetLastError();
}

template <class Tp>
int upload(Handle* h, const Tp* src, size_t n, const Tp** dst) {
  void* p = nullptr;
  CK(h, cudaMalloc(&p, n * sizeof(Tp)));
  h->table_allocs.push_back(p);
  CK(h, cudaMemcpy(p, src, n * sizeof(Tp), cudaMemcpyHostToDevice));
  *dst = (const Tp*)p;
  return 0;
}

// Which solver / model combinations have kernels.
int check_config(Handle* h, const ecrad_b200_config& c) {
  if (c.struct_bytes != (int32_t)sizeof(ecrad_b200_config)) return fail(h, "ecrad_b200_config: struct_bytes mismatch (ABI)");
  auto solver_ok = [](int s) { return s >= ECRAD_SOLVER_CLOUDLESS && s <= ECRAD_SOLVER_TRIPLECLOUDS; };
  if ((c.do_sw && !solver_ok(c.i_solver_sw)) || (c.do_lw && !solver_ok(c.i_solver_lw))) return fail(h, "unknown solver");
  if (c.do_lw && c.do_sw && ((c.i_solver_sw == ECRAD_SOLVER_HOMOGENEOUS) != (c.i_solver_lw == ECRAD_SOLVER_HOMOGENEOUS)))   // radiation_config.F90:1341-1349
    return fail(h, "if one solver is \"Homogeneous\" then the other must be");
  {   // radiation_config.F90:1134-1141
    auto regions = [](int s) { return s == ECRAD_SOLVER_SPARTACUS || s == ECRAD_SOLVER_TRIPLECLOUDS; };
    if (((c.do_sw && regions(c.i_solver_sw)) || (c.do_lw && regions(c.i_solver_lw))) && c.i_overlap_scheme != ECRAD_OVERLAP_EXP_RAN)
      return fail(h, "SPARTACUS/Tripleclouds solvers can only do Exponential-Random overlap");
  }
  if ((c.do_sw && c.i_solver_sw == ECRAD_SOLVER_SPARTACUS) || (c.do_lw && c.i_solver_lw == ECRAD_SOLVER_SPARTACUS)) {
    if (c.do_sw && c.i_solver_sw == ECRAD_SOLVER_SPARTACUS && c.do_sw_delta_scaling_with_gases)   // radiation_config.F90:1336-1339
      return fail(h, "SW delta-Eddington scaling with gases not possible with SPARTACUS solver");
    if (c.i_3d_sw_entrapment < ECRAD_ENTRAPMENT_ZERO || c.i_3d_sw_entrapment > ECRAD_ENTRAPMENT_MAXIMUM) return fail(h, "unknown sw_entrapment");
    if (!(c.min_cloud_effective_size > 0.0) || !(c.max_cloud_od > 0.0)) return fail(h, "SPARTACUS: min_cloud_effective_size and max_cloud_od must be positive");
  }
  // one gas model per spectrum (radiation_interface.F90:333-355); RRTMG-IFS in one and ECCKD in the other = test/ifs/configCY49R1_mixed.nam
  auto known = [](int m) { return m == ECRAD_GAS_IFSRRTMG || m == ECRAD_GAS_ECCKD; };
  if (!known(c.i_gas_model_lw) || !known(c.i_gas_model_sw)) return fail(h, "gas model not available in this build (RRTMG-IFS or ECCKD)");
  const bool ckd_lw = c.i_gas_model_lw == ECRAD_GAS_ECCKD, ckd_sw = c.i_gas_model_sw == ECRAD_GAS_ECCKD;
  if (c.i_cloud_pdf_shape != ECRAD_PDF_GAMMA && c.i_cloud_pdf_shape != ECRAD_PDF_LOGNORMAL) return fail(h, "unknown cloud PDF shape");
  if (c.n_regions != 2 && c.n_regions != 3) return fail(h, "n_regions must be 2 or 3 (radiation_config.F90:268)");
  if (c.n_regions == 2 && c.do_sw && c.do_lw && (c.i_solver_sw == ECRAD_SOLVER_SPARTACUS) != (c.i_solver_lw == ECRAD_SOLVER_SPARTACUS))
    return fail(h, "n_regions = 2 needs SPARTACUS in both spectra or in neither (Tripleclouds has three regions by construction and the two solvers share the region properties here)");
  if (c.i_overlap_scheme != ECRAD_OVERLAP_EXP_RAN && c.i_overlap_scheme != ECRAD_OVERLAP_MAX_RAN && c.i_overlap_scheme != ECRAD_OVERLAP_EXP_EXP)
    return fail(h, "unknown overlap scheme");
  if (c.do_lw_aerosol_scattering && c.do_lw) {
    // radiation_interface.F90:84-88
    if (!c.do_lw_cloud_scattering) return fail(h, "longwave aerosol scattering requires longwave cloud scattering");
    const bool plain = c.i_solver_lw == ECRAD_SOLVER_MCICA || c.i_solver_lw == ECRAD_SOLVER_CLOUDLESS || c.i_solver_lw == ECRAD_SOLVER_SPARTACUS;
    if (!plain || ckd_lw || (c.i_solver_lw == ECRAD_SOLVER_CLOUDLESS && c.do_save_spectral_flux))
      return fail(h, "do_lw_aerosol_scattering is available with the McICA, Cloudless and SPARTACUS longwave solvers on RRTMG-IFS gas optics");
  }
  if (c.use_aerosols && (c.n_aerosol_types < 1 || c.n_aerosol_types > 32)) return fail(h, "use_aerosols needs 1..32 aerosol types");
  if (c.use_vectorizable_generator && c.i_overlap_scheme == ECRAD_OVERLAP_EXP_EXP)   // radiation_cloud_generator.F90:239-242
    return fail(h, "the vectorizable cloud generator is not available with Exp-Exp overlap");
  if (!ckd_lw && !ckd_sw && !c.use_general_cloud_optics) {
    // radiation_cloud_optics.F90:345-372 dispatches SOCRATES and Slingo only; the five ice models of :376-447
    // (with the generalised look-up tables the two model codes are not read)
    if (c.i_liq_model != ECRAD_LIQ_SOCRATES && c.i_liq_model != ECRAD_LIQ_SLINGO) return fail(h, "liquid optics model not available (SOCRATES and Slingo are)");
    if (c.i_ice_model < ECRAD_ICE_FU || c.i_ice_model > ECRAD_ICE_YI) return fail(h, "ice optics model not available (Fu-IFS, Baran, Baran2016, Baran2017 and Yi are)");
  }
  // RRTMG: 140 / 112 g-points in 16 / 14 bands.  ecCKD: generalised cloud + aerosol optics per g-point
  // (do_cloud_aerosol_per_{sw,lw}_g_point): bands == g-points
  auto ok = [](int n) { return n == 32 || n == 64 || n == 96; };
  if (!ckd_lw && (c.n_g_lw != NG_LW || c.n_bands_lw != NB_LW)) return fail(h, "unexpected RRTMG spectral dimensions");
  if (!ckd_sw && (c.n_g_sw != NG_SW || c.n_bands_sw != NB_SW)) return fail(h, "unexpected RRTMG spectral dimensions");
  if ((ckd_lw && (!ok(c.n_g_lw) || c.n_bands_lw != c.n_g_lw)) || (ckd_sw && (!ok(c.n_g_sw) || c.n_bands_sw != c.n_g_sw)))
    return fail(h, "ECCKD: models with 32, 64 or 96 g-points and cloud/aerosol optics per g-point (n_bands == n_g) are built in");
  // radiation_cloud_optics.F90:66-79 would abort: the band parameterisations have 16 + 14 RRTMG bands
  if ((ckd_lw || ckd_sw) && !c.use_general_cloud_optics) return fail(h, "ECCKD needs use_general_cloud_optics (the band parameterisations are defined on the RRTMG bands)");
  return 0;
}

enum { N_WORK = 40 };
// Which spectra run the scan solvers (and therefore want their gas optical properties laid out [column][g][layer]).
bool use_scan(const Handle* h, bool sw, int nlev) {
  const ecrad_b200_config& c = h->cfg;
  if (!sw && c.do_lw && c.do_lw_aerosol_scattering && c.i_solver_lw != ECRAD_SOLVER_SPARTACUS)
    return nlev <= scan_max_levels();   // scattering in every layer: the general (scan) adding method
  if (!h->scan_solvers || !(sw ? c.do_sw : c.do_lw)) return false;
  const int sol = sw ? c.i_solver_sw : c.i_solver_lw;
  if (sol != ECRAD_SOLVER_MCICA && sol != ECRAD_SOLVER_CLOUDLESS) return false;
  if (sol == ECRAD_SOLVER_CLOUDLESS && c.do_save_spectral_flux) return false;   // per-band profiles: the band-summing kernels
  if (sw ? h->dcfg.ckd_sw : h->dcfg.ckd_lw) return false;
  return nlev <= scan_max_levels();
}
// bytes of every per-tile scratch array for `cols` columns (all linear in cols)
void work_sizes(const Handle* h, int cols, int nlev, size_t* out) {
  const size_t nc = (size_t)cols, nlp = (size_t)((nlev + 3) & ~3);
  const size_t nl = (size_t)((nlev + 1 + 3) & ~3);   // row stride of the [g][layer] layout; also covers [layer][g] with nlev + 1 levels
  // the Homogeneous solvers run on the Tripleclouds kernels
  const bool tc_lw = h->cfg.do_lw && (h->cfg.i_solver_lw == ECRAD_SOLVER_TRIPLECLOUDS || h->cfg.i_solver_lw == ECRAD_SOLVER_HOMOGENEOUS),
             tc_sw = h->cfg.do_sw && (h->cfg.i_solver_sw == ECRAD_SOLVER_TRIPLECLOUDS || h->cfg.i_solver_sw == ECRAD_SOLVER_HOMOGENEOUS);
  const bool sp_lw = h->cfg.do_lw && h->cfg.i_solver_lw == ECRAD_SOLVER_SPARTACUS, sp_sw = h->cfg.do_sw && h->cfg.i_solver_sw == ECRAD_SOLVER_SPARTACUS;
  const bool tc = tc_lw || tc_sw || sp_lw || sp_sw;   // region fractions and overlap matrices (tc_prep_kernel)
  const size_t NG_LW = (size_t)h->cfg.n_g_lw, NG_SW = (size_t)h->cfg.n_g_sw, NB_LW = (size_t)h->cfg.n_bands_lw, NB_SW = (size_t)h->cfg.n_bands_sw;
  const bool ckd_lw = h->dcfg.ckd_lw, ckd_sw = h->dcfg.ckd_sw, ckd = ckd_lw && ckd_sw;   // ckd: no RRTMG spectrum at all
  const bool lwscat = h->cfg.do_lw && h->cfg.do_lw_aerosol_scattering;
  const size_t sz[N_WORK] = {
      8 * nc * nl * NG_LW, 8 * nc * nl * NG_LW, 8 * nc * NG_LW, 8 * nc * NG_LW,              // od_lw planck emission lw_albedo
      8 * nc * nl * NG_SW, 8 * nc * nl * NG_SW, 8 * nc * NG_SW,                              // od_sw ssa_sw incoming
      8 * nc * nl * 3 * NB_LW, 8 * nc * nl * 3 * NB_SW,                                      // cl_lw cl_sw
      8 * nc * nl, 8 * nc * nl, 8 * nc * nl,                                                 // cum pair opi
      8 * nc, 4 * nc, 4 * nc, 4 * nc,                                                        // tcc ibegin iend ict
      4 * nc * NG_LW * nlp, 4 * nc * NG_SW * nlp,                                            // code_lw code_sw
      8 * nc * (sp_lw ? sp_scratch_doubles_lw(nlev, (int)NG_LW) : tc_lw ? tc_scratch_doubles_lw(nlev, (int)NG_LW) : use_scan(h, false, nlev) ? 0 : LW_SCR_ARRAYS * nl * NG_LW),            // scr_lw (the scan solvers keep the adding-method state in registers)
      8 * nc * 6 * (nl + 1), 8 * nc * 4 * NG_SW,                                             // sw_sums sw_carry
      8 * nc * 6 * (nl + 1), 8 * nc * 4 * NG_LW,                                             // lw_sums lw_carry
      8 * nc * (sp_sw ? sp_scratch_doubles_sw(nlev, (int)NG_SW) : tc_sw ? tc_scratch_doubles_sw(nlev, (int)NG_SW) : use_scan(h, true, nlev) ? 0 : SW_SCR_ARRAYS * nl * NG_SW),            // scr_sw
      ckd ? 0 : 8 * (size_t)LWLEV_NF * nc * nl, ckd ? 0 : 8 * (size_t)SWLEV_NF * nc * nl,    // lev_lw lev_sw (RRTMG)
      h->cfg.use_aerosols ? 8 * nc * nl * NG_SW : 0, (h->cfg.use_aerosols && !ckd_sw) ? 8 * nc * nl * 3 * NB_SW : 0,   // g_sw aer_sw
      (h->cfg.use_aerosols && !ckd_lw) ? 8 * nc * nl * NB_LW * (lwscat ? 3 : 1) : 0,            // aer_lw (ecCKD merges aerosols per g-point inside its gas kernels)
      0,                                                                                     // (sw_band_dir: no longer used)
      tc ? 8 * nc * nl * 3 : 0, tc ? 8 * nc * nl * 3 : 0, tc ? 8 * nc * (nl + 1) * 9 : 0, tc ? 8 * nc * (nl + 1) * 9 : 0, tc ? 8 * nc : 0,  // tc_reg tc_ods tc_u tc_v tc_cc
      ckd ? 0 : nc * nl, ckd ? 0 : sizeof(GasCol) * nc, ckd ? 0 : 4 * (nc + 1),              // gas_jp gas_col sunlit (RRTMG)
      lwscat ? 8 * nc * nl * NG_LW : 0, lwscat ? 8 * nc * nl * NG_LW : 0};                   // ssa_lw g_lw (do_lw_aerosol_scattering)
  for (int i = 0; i < N_WORK; ++i) out[i] = sz[i];
}
size_t work_bytes_per_column(const Handle* h, int nlev) {
  size_t sz[N_WORK], tot = 0;
  work_sizes(h, 1, nlev, sz);
  for (int i = 0; i < N_WORK; ++i) tot += sz[i];
  return tot;
}

int ensure_work(Handle* h, int set, int cols, int nlev) {
  if (cols <= h->w_cols[set] && nlev == h->w_nlev[set] && h->scan_solvers == h->w_scan[set]) return 0;
  if (nlev != h->w_nlev[set] || h->scan_solvers != h->w_scan[set]) {   // different array sizes per column: start over
    cudaDeviceSynchronize();
    for (auto& b : h->work[set]) b.release();
    h->w_cols[set] = 0;
  }
  h->w_scan[set] = h->scan_solvers;
  size_t sz[N_WORK];
  work_sizes(h, cols, nlev, sz);
  for (int i = 0; i < N_WORK; ++i) CK(h, h->work[set][i].reserve(sz[i]));
  Work& w = h->w[set];
  w.od_lw = (double*)h->work[set][0].p; w.planck = (double*)h->work[set][1].p; w.emission = (double*)h->work[set][2].p; w.lw_albedo = (double*)h->work[set][3].p;
  w.od_sw = (double*)h->work[set][4].p; w.ssa_sw = (double*)h->work[set][5].p; w.incoming = (double*)h->work[set][6].p;
  w.cl_lw = (double*)h->work[set][7].p; w.cl_sw = (double*)h->work[set][8].p;
  w.cum = (double*)h->work[set][9].p; w.pair = (double*)h->work[set][10].p; w.opi = (double*)h->work[set][11].p;
  w.tcc = (double*)h->work[set][12].p; w.ibegin = (int*)h->work[set][13].p; w.iend = (int*)h->work[set][14].p; w.ict = (int*)h->work[set][15].p;
  w.code_lw = (uint32_t*)h->work[set][16].p; w.code_sw = (uint32_t*)h->work[set][17].p;
  w.scr_lw = (double*)h->work[set][18].p; w.scr_sw = (double*)h->work[set][23].p;
  w.lev_lw = (double*)h->work[set][24].p; w.lev_sw = (double*)h->work[set][25].p;
  w.g_sw = (double*)h->work[set][26].p; w.aer_sw = (double*)h->work[set][27].p; w.aer_lw = (double*)h->work[set][28].p;
  w.sw_band_dir = (double*)h->work[set][29].p;
  w.tc_reg = (double*)h->work[set][30].p; w.tc_ods = (double*)h->work[set][31].p; w.tc_u = (double*)h->work[set][32].p;
  w.tc_v = (double*)h->work[set][33].p; w.tc_cc = (double*)h->work[set][34].p;
  w.sw_sums = (double*)h->work[set][19].p; w.sw_carry = (double*)h->work[set][20].p;
  w.lw_sums = (double*)h->work[set][21].p; w.lw_carry = (double*)h->work[set][22].p;
  w.gas_jp = (uint8_t*)h->work[set][35].p; w.gas_col = (GasCol*)h->work[set][36].p; w.sunlit = (int*)h->work[set][37].p;
  w.ssa_lw = (double*)h->work[set][38].p; w.g_lw = (double*)h->work[set][39].p;
  w.ls = (nlev + 1 + 3) & ~3;
  h->w_cols[set] = cols; h->w_nlev[set] = nlev;
  return 0;
}

// Kernels of one tile; `in`/`out` are device views whose column 0 is the first column of the tile.
// Three independent chains: cloud (prep -> optics -> generator), LW (gas -> down -> up -> flux), SW (gas -> direct ->
// adding -> flux); the solvers wait for the cloud chain.  With serial == 0 they run on three streams forked from and
// joined into `st`, so that latency-bound and fp64-bound kernels share the SMs.
// ev: 2 events per stage (start, end), stages = gas_lw, gas_sw, cloud, solver_lw, solver_sw.
int run_tile(Handle* h, int set, const DevIn& in, const DevOut& out, int nc, int nlev, cudaStream_t st, cudaEvent_t* ev, bool optics_only = false) {
  const DevCfg& c = h->dcfg;
  int n = 0;
  const bool par = !h->serial;
  h->w[set].layout_b_lw = use_scan(h, false, nlev);
  h->w[set].layout_b_sw = use_scan(h, true, nlev);
  cudaStream_t s_lw = st, s_sw = par ? h->s_aux1[set] : st, s_cl = par ? h->s_aux2[set] : st;
  const bool ckd_lw = c.ckd_lw != 0, ckd_sw = c.ckd_sw != 0;
  // what the RRTMG kernels see: with mixed gas models only the RRTMG spectrum is theirs (the other one has other spectral sizes)
  DevCfg crr = c;
  crr.do_lw = c.do_lw && !ckd_lw; crr.do_sw = c.do_sw && !ckd_sw;
  if (!(ckd_lw && ckd_sw)) {
    n += launch_gas_prep(h->T, crr, in, h->w[set], nc, nlev, st);   // shared by the LW and SW gas-optics kernels
    if ((h->gas_variant & 3) || c.do_lw_aerosol_scattering) n += launch_gas_col(h->T, crr, in, h->w[set], nc, nlev, st);
    if (c.use_aerosols) n += launch_aerosol(h->T, crr, in, h->w[set], nc, nlev, st);
  }
  if (par) {
    CK(h, cudaEventRecord(h->ev_fork[set], st));
    CK(h, cudaStreamWaitEvent(s_sw, h->ev_fork[set], 0));
    CK(h, cudaStreamWaitEvent(s_cl, h->ev_fork[set], 0));
  }
  // cloud chain
  CK(h, cudaEventRecord(ev[4], s_cl));
  if (c.do_clouds) n += launch_cloud(h->T, c, in, h->w[set], nc, nlev, s_cl);
  auto regions = [](int s) { return s == ECRAD_SOLVER_TRIPLECLOUDS || s == ECRAD_SOLVER_SPARTACUS || s == ECRAD_SOLVER_HOMOGENEOUS; };
  if ((c.do_lw && regions(c.solver_lw)) || (c.do_sw && regions(c.solver_sw)))
    n += launch_tc_prep(c, in, h->w[set], nc, nlev, s_cl);
  CK(h, cudaEventRecord(ev[5], s_cl));
  if (par) CK(h, cudaEventRecord(h->ev_cloud[set], s_cl));
  // LW chain
  CK(h, cudaEventRecord(ev[0], s_lw));
  if (c.do_lw) n += ckd_lw ? launch_ckd_lw(h->T, c, in, h->w[set], nc, nlev, s_lw) : ((h->gas_variant & 1) || c.do_lw_aerosol_scattering) ? launch_gas_lw_band(h->T, crr, in, h->w[set], nc, nlev, s_lw)
                                                                                                  : launch_gas_lw(h->T, crr, in, h->w[set], nc, nlev, s_lw);
  CK(h, cudaEventRecord(ev[1], s_lw));
  // SW chain
  CK(h, cudaEventRecord(ev[2], s_sw));
  if (c.do_sw) n += ckd_sw ? launch_ckd_sw(h->T, c, in, h->w[set], nc, nlev, s_sw) : (h->gas_variant & 2) ? launch_gas_sw_band(h->T, crr, in, h->w[set], nc, nlev, s_sw)
                                                                                                  : launch_gas_sw(h->T, crr, in, h->w[set], nc, nlev, s_sw);
  CK(h, cudaEventRecord(ev[3], s_sw));
  if (par) { CK(h, cudaStreamWaitEvent(s_lw, h->ev_cloud[set], 0)); CK(h, cudaStreamWaitEvent(s_sw, h->ev_cloud[set], 0)); }
  CK(h, cudaEventRecord(ev[6], s_lw));
  if (c.do_lw && !optics_only) n += launch_solver_lw(h->T, c, in, out, h->w[set], nc, nlev, s_lw);
  if (c.do_lw && c.do_toa_spectral_flux && !optics_only) n += launch_toa_spectral(h->T, c, in, out, nc, false, s_lw);
  CK(h, cudaEventRecord(ev[7], s_lw));
  CK(h, cudaEventRecord(ev[8], s_sw));
  if (c.do_sw && !optics_only) n += launch_solver_sw(h->T, c, in, out, h->w[set], nc, nlev, s_sw);
  if (c.do_sw && c.do_toa_spectral_flux && !optics_only) n += launch_toa_spectral(h->T, c, in, out, nc, true, s_sw);
  CK(h, cudaEventRecord(ev[9], s_sw));
  if (par) {
    CK(h, cudaEventRecord(h->ev_sw_done[set], s_sw));
    CK(h, cudaStreamWaitEvent(st, h->ev_sw_done[set], 0));
  }
  CK(h, cudaGetLastError());
  h->launches += n;
  return 0;
}

int ensure_events(Handle* h, int tiles) {
  while ((int)h->ev.size() < tiles * 2 * N_STAGE) {
    cudaEvent_t e; CK(h, cudaEventCreate(&e)); h->ev.push_back(e);
  }
  h->ev_tiles = tiles;
  return 0;
}

struct InDesc { const void* host; int rows; int elem; };   // column-fastest (ncol, rows)
struct OutDesc { double* host; int kind; int rows; };      // kind 0: (ncol, rows) profile; 1: (rows, ncol) per-column block; 2: (nb, ncol, nlev+1)

int check_args(Handle* h, int ncol, int nlev, int istartcol, int iendcol, const ecrad_b200_inputs* in, const ecrad_b200_outputs* out) {
  if (!in || !out) return fail(h, "null inputs/outputs");
  if (in->struct_bytes != (int32_t)sizeof(ecrad_b200_inputs) || out->struct_bytes != (int32_t)sizeof(ecrad_b200_outputs))
    return fail(h, "ecrad_b200_inputs/outputs: struct_bytes mismatch (ABI)");
  if (ncol < 1 || nlev < 2 || nlev > 256 || istartcol < 1 || iendcol > ncol || iendcol < istartcol)
    return fail(h, "bad dimensions ncol=%d nlev=%d istartcol=%d iendcol=%d (nlev <= 256)", ncol, nlev, istartcol, iendcol);
  const ecrad_b200_config& c = h->cfg;
  if (!in->pressure_hl || !in->temperature_hl || !in->h2o_mmr || !in->co2_mmr || !in->o3_mmr || !in->n2o_mmr || !in->ch4_mmr ||
      !in->cfc11_mmr || !in->cfc12_mmr || !in->hcfc22_mmr || !in->ccl4_mmr)
    return fail(h, "missing thermodynamics/gas input");
  if (c.do_lw && (!in->skin_temperature || !in->lw_emissivity)) return fail(h, "missing LW surface input");
  if (c.do_sw && (!in->cos_sza || !in->sw_albedo)) return fail(h, "missing SW surface input");
  if (c.do_clouds && (!in->cloud_fraction || !in->q_liq || !in->q_ice || !in->re_liq || !in->re_ice || !in->overlap_param ||
                      !in->fractional_std || !in->iseed))
    return fail(h, "missing cloud input");
  if (c.use_aerosols && (!in->aerosol_mmr || !in->h2o_sat_liq)) return fail(h, "missing aerosol input (aerosol_mmr, h2o_sat_liq)");
  if (c.do_3d_effects && ((c.do_sw && c.i_solver_sw == ECRAD_SOLVER_SPARTACUS) || (c.do_lw && c.i_solver_lw == ECRAD_SOLVER_SPARTACUS)) &&
      !in->inv_cloud_effective_size)
    return fail(h, "SPARTACUS with do_3d_effects needs cloud%%inv_cloud_effective_size");
  if (c.do_lw && (!out->lw_up || !out->lw_dn)) return fail(h, "flux%%lw_up/lw_dn must be allocated");
  if (c.do_sw && (!out->sw_up || !out->sw_dn)) return fail(h, "flux%%sw_up/sw_dn must be allocated");
  if (c.do_toa_spectral_flux && ((c.do_sw && out->sw_up_toa_band && !out->sw_up_toa_g) || (c.do_lw && out->lw_up_toa_band && !out->lw_up_toa_g)))
    return fail(h, "do_toa_spectral_flux: flux%%sw_up_toa_g / lw_up_toa_g must be allocated");
  return 0;
}

void fill_descs(const ecrad_b200_config& c, int nlev, const ecrad_b200_inputs* in, const ecrad_b200_outputs* out, InDesc* id, OutDesc* od) {
  const int nl = nlev, nl1 = nlev + 1;
  const bool sp = (c.do_sw && c.i_solver_sw == ECRAD_SOLVER_SPARTACUS) || (c.do_lw && c.i_solver_lw == ECRAD_SOLVER_SPARTACUS);
  const InDesc ins[N_IN] = {
      {in->cos_sza, 1, 8}, {in->skin_temperature, 1, 8}, {in->sw_albedo, c.n_albedo_sw, 8}, {in->sw_albedo_direct, c.n_albedo_sw, 8},
      {in->lw_emissivity, c.n_emiss_lw, 8}, {in->iseed, 1, 4}, {in->pressure_hl, nl1, 8}, {in->temperature_hl, nl1, 8},
      {in->h2o_mmr, nl, 8}, {in->co2_mmr, nl, 8}, {in->ch4_mmr, nl, 8}, {in->n2o_mmr, nl, 8}, {in->cfc11_mmr, nl, 8},
      {in->cfc12_mmr, nl, 8}, {in->hcfc22_mmr, nl, 8}, {in->ccl4_mmr, nl, 8}, {in->o3_mmr, nl, 8},
      {in->cloud_fraction, nl, 8}, {in->q_liq, nl, 8}, {in->q_ice, nl, 8}, {in->re_liq, nl, 8}, {in->re_ice, nl, 8},
      {in->overlap_param, nl - 1, 8}, {in->fractional_std, nl, 8},
      {c.use_aerosols ? in->aerosol_mmr : nullptr, nl * c.n_aerosol_types, 8}, {c.use_aerosols ? in->h2o_sat_liq : nullptr, nl, 8},
      {sp ? in->inv_cloud_effective_size : nullptr, nl, 8}, {sp ? in->inv_inhom_effective_size : nullptr, nl, 8}};
  for (int i = 0; i < N_IN; ++i) id[i] = ins[i];
  const OutDesc outs[N_OUT] = {
      {out->lw_up, 0, nl1}, {out->lw_dn, 0, nl1}, {out->lw_up_clear, 0, nl1}, {out->lw_dn_clear, 0, nl1},
      {out->sw_up, 0, nl1}, {out->sw_dn, 0, nl1}, {out->sw_dn_direct, 0, nl1},
      {out->sw_up_clear, 0, nl1}, {out->sw_dn_clear, 0, nl1}, {out->sw_dn_direct_clear, 0, nl1},
      {out->lw_derivatives, 0, nl1}, {out->cloud_cover_lw, 1, 1}, {out->cloud_cover_sw, 1, 1},
      {out->lw_dn_surf_g, 1, c.n_g_lw}, {out->lw_dn_surf_clear_g, 1, c.n_g_lw}, {out->lw_up_toa_g, 1, c.n_g_lw}, {out->lw_up_toa_clear_g, 1, c.n_g_lw},
      {out->sw_dn_diffuse_surf_g, 1, c.n_g_sw}, {out->sw_dn_direct_surf_g, 1, c.n_g_sw}, {out->sw_dn_diffuse_surf_clear_g, 1, c.n_g_sw},
      {out->sw_dn_direct_surf_clear_g, 1, c.n_g_sw}, {out->sw_up_toa_g, 1, c.n_g_sw}, {out->sw_up_toa_clear_g, 1, c.n_g_sw},
      {out->sw_dn_surf_band, 1, c.n_bands_sw}, {out->sw_dn_direct_surf_band, 1, c.n_bands_sw}, {out->sw_dn_surf_clear_band, 1, c.n_bands_sw},
      {out->sw_dn_direct_surf_clear_band, 1, c.n_bands_sw},
      {out->sw_dn_diffuse_surf_canopy, 1, c.n_canopy_bands_sw}, {out->sw_dn_direct_surf_canopy, 1, c.n_canopy_bands_sw},
      {out->lw_dn_surf_canopy, 1, c.n_canopy_bands_lw},
      {out->lw_up_band, 2, c.n_bands_lw}, {out->lw_dn_band, 2, c.n_bands_lw}, {out->sw_up_band, 2, c.n_bands_sw}, {out->sw_dn_band, 2, c.n_bands_sw},
      {out->sw_dn_direct_band, 2, c.n_bands_sw},
      {out->sw_dn_toa_g, 1, c.n_g_sw}, {out->sw_dn_toa_band, 1, c.n_bands_sw}, {out->sw_up_toa_band, 1, c.n_bands_sw},
      {out->sw_up_toa_clear_band, 1, c.n_bands_sw}, {out->lw_up_toa_band, 1, c.n_bands_lw}, {out->lw_up_toa_clear_band, 1, c.n_bands_lw}};
  for (int i = 0; i < N_OUT; ++i) od[i] = outs[i];
  // config%do_sw_direct = false: flux%sw_dn_direct, sw_dn_direct_clear and sw_dn_direct_band are not allocated (radiation_flux.F90:208-240)
  // and every solver tests allocated() before storing them; the direct beam itself is always part of the solution
  if (!c.do_sw_direct)
    for (int i = 0; i < N_OUT; ++i)
      if (od[i].host && (od[i].host == out->sw_dn_direct || od[i].host == out->sw_dn_direct_clear || od[i].host == out->sw_dn_direct_band)) od[i].host = nullptr;
}

// Build the kernel-facing views from 28 input / 41 output base pointers (device) with leading dimension ld.
void make_views(void* const* ip, void* const* op, int ld, int ld_out, double solar_irradiance, DevIn& di, DevOut& dout) {
  di.cos_sza = (const double*)ip[0]; di.skin_t = (const double*)ip[1]; di.sw_albedo = (const double*)ip[2];
  di.sw_albedo_direct = (const double*)ip[3]; di.lw_emissivity = (const double*)ip[4]; di.iseed = (const int32_t*)ip[5];
  di.p_hl = (const double*)ip[6]; di.t_hl = (const double*)ip[7];
  for (int k = 0; k < 9; ++k) di.gas[k] = (const double*)ip[8 + k];
  di.frac = (double*)ip[17]; di.q_liq = (const double*)ip[18]; di.q_ice = (const double*)ip[19];
  di.re_liq = (const double*)ip[20]; di.re_ice = (const double*)ip[21]; di.overlap = (const double*)ip[22]; di.fsd = (const double*)ip[23];
  di.aerosol_mmr = (const double*)ip[24]; di.h2o_sat_liq = (const double*)ip[25];
  di.inv_cloud_size = (const double*)ip[26]; di.inv_inhom_size = (const double*)ip[27];
  di.solar_irradiance = solar_irradiance; di.ld = ld;
  double** o = (double**)&dout;
  for (int k = 0; k < N_OUT; ++k) o[k] = (double*)op[k];
  dout.ld = ld_out;
}
static_assert(sizeof(DevOut) >= N_OUT * sizeof(double*) + sizeof(int), "DevOut layout");
static_assert(offsetof(DevOut, lw_up_toa_clear_band) == (N_OUT - 1) * sizeof(double*), "DevOut must list the 41 outputs in ABI order");

// ---- blocked (NPROMA) layout: zrgp(nproma, nfields, nblocks) <-> column-fastest arrays -------------------------------------
struct BlockJob { void* dev; int field0, rows, kind, is_int; };   // kind 0: (ncol, rows); 1: (rows, ncol); 2: (nb = rows, ncol, nlev+1)
struct BlockJobs { BlockJob j[48]; int n; };
// one thread per (column, field row): column fastest, so both sides are read / written in runs of nproma
__global__ void block_unpack_kernel(BlockJobs J, const double* __restrict__ z, int nproma, int nfields, int ncol, int nlev1) {
  const BlockJob& b = J.j[blockIdx.y];
  const int nrows = b.kind == 2 ? b.rows * nlev1 : b.rows;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)ncol * nrows) return;
  const int c = (int)(i % ncol), r = (int)(i / ncol);
  const double v = z[((size_t)(c / nproma) * nfields + b.field0 + r) * nproma + (c % nproma)];
  if (b.is_int) ((int32_t*)b.dev)[(size_t)r * ncol + c] = (int32_t)v;
  else if (b.kind == 0) ((double*)b.dev)[(size_t)r * ncol + c] = v;
  else if (b.kind == 1) ((double*)b.dev)[(size_t)c * b.rows + r] = v;
  else ((double*)b.dev)[((size_t)(r / b.rows) * ncol + c) * b.rows + (r % b.rows)] = v;
}
__global__ void block_pack_kernel(BlockJobs J, double* __restrict__ z, int nproma, int nfields, int ncol, int nlev1) {
  const BlockJob& b = J.j[blockIdx.y];
  const int nrows = b.kind == 2 ? b.rows * nlev1 : b.rows;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)ncol * nrows) return;
  const int c = (int)(i % ncol), r = (int)(i / ncol);
  const double* src = (const double*)b.dev;
  double v;
  if (b.kind == 0) v = src[(size_t)r * ncol + c];
  else if (b.kind == 1) v = src[(size_t)c * b.rows + r];
  else v = src[((size_t)(r / b.rows) * ncol + c) * b.rows + (r % b.rows)];   // field r = level * nband + band
  z[((size_t)(c / nproma) * nfields + b.field0 + r) * nproma + (c % nproma)] = v;
}

// ---- single-precision boundary: float <-> double conversion of whole staging buffers, one launch for all arrays of a tile ----
struct CvtJob { const void* src; void* dst; long long n; };
struct CvtJobs { CvtJob j[48]; int n; };
__global__ void convert_kernel(CvtJobs J, int to_double) {
  const CvtJob& b = J.j[blockIdx.y];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < b.n; i += (long long)gridDim.x * blockDim.x) {
    if (to_double) ((double*)b.dst)[i] = (double)((const float*)b.src)[i];
    else ((float*)b.dst)[i] = (float)((const double*)b.src)[i];
  }
}

// fp64 multiply-add throughput of the device, measured: 8 independent chains per thread, enough CTAs to fill every SM
__global__ void __launch_bounds__(256) fp64_fma_probe_kernel(double* out, int iters, double b, double c) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

}  // namespace

// =========================================================================================================
extern "C" {

int ecrad_b200_measure_fp64(double* tflops) {
  if (!tflops) return 1;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return fail(nullptr, "ecrad_b200_measure_fp64: no CUDA device");
  const int blocks = sms * 8, threads = 256, iters = 2048;
  double* buf = nullptr;
  if (cudaMalloc(&buf, sizeof(double) * blocks * threads) != cudaSuccess) return fail(nullptr, "ecrad_b200_measure_fp64: out of memory");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    fp64_fma_probe_kernel<<<blocks, threads>>>(buf, iters, 0.999999, 1.0e-6);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(buf);
  if (cudaGetLastError() != cudaSuccess) return fail(nullptr, "ecrad_b200_measure_fp64: kernel failed");
  *tflops = 2.0 * 8.0 * iters * (double)blocks * threads / (best * 1e-3) / 1e12;
  return 0;
}

ecrad_b200_tables* ecrad_b200_tables_create(void) { return new (std::nothrow) ecrad_b200_tables(); }
int ecrad_b200_tables_add(ecrad_b200_tables* t, const char* name, int dtype, int ndim, const int64_t* dims, const void* data) {
  if (!t) return 1;
  try { return t->add(name, dtype, ndim, dims, data); } catch (const std::exception& ex) { return fail(nullptr, "ecrad_b200_tables_add('%s'): %s", name ? name : "", ex.what()); }
}
int ecrad_b200_tables_load_file(ecrad_b200_tables* t, const char* path) {
  if (!t || !path) return 1;
  int rc;
  try { rc = t->load_file(path); } catch (const std::exception& ex) { return fail(nullptr, "cannot load table blob '%s': %s", path, ex.what()); }
  if (rc) fail(nullptr, "cannot load table blob '%s' (rc=%d)", path, rc);
  return rc;
}
int ecrad_b200_tables_load_memory(ecrad_b200_tables* t, const void* blob, int64_t nbytes) {
  if (!t || !blob || nbytes < 8) return 1;
  int rc;
  try { rc = t->load_memory((const char*)blob, (size_t)nbytes); } catch (const std::exception& ex) { return fail(nullptr, "table blob in memory: %s", ex.what()); }
  if (rc) fail(nullptr, "table blob in memory is not a valid ETB1 image (rc=%d)", rc);
  return rc;
}
void ecrad_b200_tables_free(ecrad_b200_tables* t) { delete t; }

const char* ecrad_b200_version(void) { return "ecrad_b200 0.3 (sm_100a; RRTMG-IFS / ecCKD gas optics; McICA, Tripleclouds, SPARTACUS, Homogeneous, Cloudless; fp64)"; }
const char* ecrad_b200_last_error(void* handle) {
  if (handle) return ((Handle*)handle)->err.c_str();
  return g_last_error.c_str();
}
const char* ecrad_b200_stage_name(int stage) { return (stage >= 0 && stage < N_STAGE) ? kStageNames[stage] : ""; }

int ecrad_b200_setup(const ecrad_b200_config* cfg, const ecrad_b200_tables* tab, void** handle) {
  if (handle) *handle = nullptr;
  if (!cfg || !tab || !handle) return fail(nullptr, "ecrad_b200_setup: null argument");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev < 1)
    return fail(nullptr, "ecrad_b200_setup: no CUDA device (%s); this library has no CPU path", e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
  Handle* h = new (std::nothrow) Handle();
  if (!h) return fail(nullptr, "out of memory");
  h->cfg = *cfg;
  if (check_config(h, *cfg)) { delete h; return 1; }
  PackedTables P;
  try { pack_tables(*tab, P, cfg->i_liq_model, cfg->i_ice_model, cfg->use_general_cloud_optics != 0); } catch (const std::exception& ex) { fail(nullptr, "table directory incomplete: %s", ex.what()); delete h; return 1; }
  for (int g = 0; g < NG_LW; ++g) P.meta.rank_lw[g] = (short)g;
  for (int g = 0; g < NG_SW; ++g) P.meta.rank_sw[g] = (short)g;
  {
    // SPARTACUS treats the g-points up to the first one whose gas optical depth exceeds max_gas_od_3d with the matrix exponential,
    // "assuming that the g-points have been reordered in approximate order of gas optical depth" (radiation_spartacus_sw.F90:462-478):
    // radiation_ifs_rrtm.F90:122-130 / :167-174 reorder them for this solver only.  The kernels keep the arrays in RRTMG order
    // and carry each g-point's position in that sequence; the per-g-point outputs are written at that position.
    for (int sw = 0; sw < 2; ++sw) {
      if (!(sw ? (cfg->do_sw && cfg->i_solver_sw == ECRAD_SOLVER_SPARTACUS) : (cfg->do_lw && cfg->i_solver_lw == ECRAD_SOLVER_SPARTACUS))) continue;
      if (sw ? P.ckd_sw : P.ckd_lw) continue;   // (an ecCKD spectrum stays in its own order)
      const char* nm = sw ? "i_g_from_reordered_g_sw" : "i_g_from_reordered_g_lw";
      const int ng = sw ? NG_SW : NG_LW;
      const auto* a = tab->find(nm);
      if (!a || a->dtype != 1 || a->data.size() != (size_t)ng * 4) { fail(nullptr, "SPARTACUS on RRTMG-IFS needs the table '%s' (%d int32)", nm, ng); delete h; return 1; }
      const int32_t* perm = (const int32_t*)a->data.data();
      std::vector<int> seen(ng, 0);
      for (int j = 0; j < ng; ++j) {
        if (perm[j] < 1 || perm[j] > ng || seen[perm[j] - 1]++) { fail(nullptr, "'%s' is not a permutation of 1..%d", nm, ng); delete h; return 1; }
        (sw ? P.meta.rank_sw : P.meta.rank_lw)[perm[j] - 1] = (short)j;
      }
    }
  }
  if (P.ckd_lw != (cfg->i_gas_model_lw == ECRAD_GAS_ECCKD) || P.ckd_sw != (cfg->i_gas_model_sw == ECRAD_GAS_ECCKD)) {
    fail(nullptr, "the table directory holds %s / %s tables (longwave / shortwave) but the configuration names other gas models",
         P.ckd_lw ? "ecCKD" : "RRTMG", P.ckd_sw ? "ecCKD" : "RRTMG"); delete h; return 1;
  }
  if ((P.ckd_lw && cfg->do_lw && P.ng_lw != cfg->n_g_lw) || (P.ckd_sw && cfg->do_sw && P.ng_sw != cfg->n_g_sw)) {
    fail(nullptr, "ecCKD tables have %d/%d g-points (LW/SW), the configuration says %d/%d", P.ng_lw, P.ng_sw, cfg->n_g_lw, cfg->n_g_sw); delete h; return 1;
  }
  if (cfg->do_nearest_spectral_sw_albedo) {
    // Nearest-interval mapping (get_albedos, radiation_single_level.F90:266-285; calc_surface_spectral, radiation_flux.F90:479-497) is the
    // weighted mapping with a single weight of 1 per band: 0 + 1*albedo is exact, and the canopy fluxes collect the same bands.
    if ((int)P.i_albedo_from_band_sw.size() != cfg->n_bands_sw || cfg->n_albedo_sw < 1) {
      fail(nullptr, "do_nearest_spectral_sw_albedo: table 'i_albedo_from_band_sw' (n_bands_sw) is required"); delete h; return 1;
    }
    P.n_albedo_sw = cfg->n_albedo_sw;
    P.sw_albedo_weights.assign((size_t)P.n_albedo_sw * cfg->n_bands_sw, 0.0);
    for (int jb = 0; jb < cfg->n_bands_sw; ++jb) {
      const int ia = P.i_albedo_from_band_sw[jb];
      if (ia < 1 || ia > P.n_albedo_sw) { fail(nullptr, "i_albedo_from_band_sw out of range"); delete h; return 1; }
      P.sw_albedo_weights[(size_t)jb * P.n_albedo_sw + ia - 1] = 1.0;
    }
  }
  if (P.sw_albedo_weights.empty() || P.n_albedo_sw != cfg->n_albedo_sw) {
    fail(nullptr, "table 'sw_albedo_weights' (n_albedo_sw x n_bands_sw) is required"); delete h; return 1;
  }
  if (cfg->do_nearest_spectral_lw_emiss ? P.i_emiss_from_band_lw.empty() : (P.lw_emiss_weights.empty() || P.n_emiss_lw != cfg->n_emiss_lw)) {
    fail(nullptr, "table 'i_emiss_from_band_lw' (n_bands_lw) or 'lw_emiss_weights' (n_emiss_lw x n_bands_lw) is required"); delete h; return 1;
  }
  if (cfg->do_nearest_spectral_lw_emiss)   // the kernels index lw_emissivity(ncol, n_emiss_lw) with this map
    for (int v : P.i_emiss_from_band_lw)
      if (v < 1 || v > cfg->n_emiss_lw) { fail(nullptr, "i_emiss_from_band_lw out of range 1..n_emiss_lw = %d", cfg->n_emiss_lw); delete h; return 1; }
  cudaGetDevice(&h->device);
  int rc = 0;
  rc |= upload(h, &P.meta, 1, &h->T.meta);
  h->T.lwtab = h->T.swtab = nullptr; h->T.ckd = nullptr; h->T.ckdtab = nullptr;
  h->T.i_emiss_from_band_lw = nullptr; h->T.lw_emiss_weights = nullptr;
  if (P.ckd_lw || P.ckd_sw || cfg->use_general_cloud_optics) {   // ecCKD models and/or the generalised cloud optics look-up tables (per g-point or per RRTMG band)
    rc |= upload(h, &P.ckd, 1, &h->T.ckd);
    rc |= upload(h, P.ckdtab.data(), P.ckdtab.size(), &h->T.ckdtab);
  }
  if (!P.is_ecckd) {
    rc |= upload(h, P.lwtab.data(), P.lwtab.size(), &h->T.lwtab);
    rc |= upload(h, P.swtab.data(), P.swtab.size(), &h->T.swtab);
  }
  if (cfg->use_vectorizable_generator) { P.cloud.gen_mask = 0x7FFFFFFFu; P.cloud.gen_scale = 1.0 / 2147483647.0; }
  else { P.cloud.gen_mask = 0x3FFFFFFFu; P.cloud.gen_scale = 1.0 / 1073741824.0; }
  rc |= upload(h, &P.cloud, 1, &h->T.cloud);
  rc |= upload(h, P.pdf_val.data(), P.pdf_val.size(), &h->T.pdf_val);
  rc |= upload(h, P.sw_albedo_weights.data(), P.sw_albedo_weights.size(), &h->T.sw_albedo_weights);
  if (!P.i_emiss_from_band_lw.empty()) rc |= upload(h, P.i_emiss_from_band_lw.data(), P.i_emiss_from_band_lw.size(), &h->T.i_emiss_from_band_lw);
  if (!P.lw_emiss_weights.empty()) rc |= upload(h, P.lw_emiss_weights.data(), P.lw_emiss_weights.size(), &h->T.lw_emiss_weights);
  h->T.aer = nullptr; h->T.aertab = nullptr;
  if (cfg->use_aerosols) {
    if (P.aer.ntype != cfg->n_aerosol_types || P.aertab.empty()) {
      fail(nullptr, "use_aerosols: tables 'aerosol_iclass'/'aerosol_itype' (n_aerosol_types) and 'aer_*' are required"); ecrad_b200_finalize(h); return 1;
    }
    rc |= upload(h, &P.aer, 1, &h->T.aer);
    rc |= upload(h, P.aertab.data(), P.aertab.size(), &h->T.aertab);
  }
  if (rc) { g_last_error = h->err; ecrad_b200_finalize(h); return 1; }
  DevCfg& d = h->dcfg;
  d.solver_sw = cfg->i_solver_sw; d.solver_lw = cfg->i_solver_lw; d.overlap_scheme = cfg->i_overlap_scheme;
  d.do_sw = cfg->do_sw; d.do_lw = cfg->do_lw; d.do_clouds = cfg->do_clouds;
  d.do_lw_cloud_scattering = cfg->do_lw_cloud_scattering; d.do_lw_derivatives = cfg->do_lw_derivatives;
  d.do_sw_delta_scaling_with_gases = cfg->do_sw_delta_scaling_with_gases; d.do_fu_lw_ice_optics_bug = cfg->do_fu_lw_ice_optics_bug;
  d.do_lw_aerosol_scattering = cfg->do_lw && cfg->do_lw_aerosol_scattering;
  d.use_beta_overlap = cfg->use_beta_overlap; d.do_surface_sw_spectral_flux = cfg->do_surface_sw_spectral_flux;
  d.do_canopy_fluxes_sw = cfg->do_canopy_fluxes_sw; d.do_canopy_fluxes_lw = cfg->do_canopy_fluxes_lw; d.do_clear = cfg->do_clear;
  d.n_albedo_sw = cfg->n_albedo_sw; d.n_emiss_lw = cfg->n_emiss_lw;
  d.n_canopy_bands_sw = cfg->n_canopy_bands_sw; d.n_canopy_bands_lw = cfg->n_canopy_bands_lw;
  d.use_aerosols = cfg->use_aerosols; d.n_aerosol_types = cfg->n_aerosol_types;
  d.do_save_spectral_flux = cfg->do_save_spectral_flux;
  d.use_vectorizable_generator = cfg->use_vectorizable_generator;
  d.do_nearest_spectral_lw_emiss = cfg->do_nearest_spectral_lw_emiss;
  d.ckd_lw = P.ckd_lw; d.ckd_sw = P.ckd_sw;
  d.solar_cycle_multiplier = 0.0;
  h->has_solar_cycle = P.ckd_sw && P.ckd.sw.off_solar_amp >= 0;
  d.gas_mmr = !(P.ckd_lw && P.ckd_sw);
  d.use_general_cloud_optics = cfg->use_general_cloud_optics != 0;
  d.do_toa_spectral_flux = cfg->do_toa_spectral_flux;
  d.pdf_gamma = cfg->i_cloud_pdf_shape == ECRAD_PDF_GAMMA;
  d.is_homogeneous = (cfg->do_sw && cfg->i_solver_sw == ECRAD_SOLVER_HOMOGENEOUS) || (cfg->do_lw && cfg->i_solver_lw == ECRAD_SOLVER_HOMOGENEOUS);   // radiation_config.F90:1351-1356
  d.ng_lw = cfg->n_g_lw; d.ng_sw = cfg->n_g_sw; d.nb_lw = cfg->n_bands_lw; d.nb_sw = cfg->n_bands_sw;
  d.ckd_ngas_lw = P.ckd.lw.ngas; d.ckd_nlut_lw = P.ckd.lw.nlut; d.ckd_ngas_sw = P.ckd.sw.ngas; d.ckd_nlut_sw = P.ckd.sw.nlut;
  d.cloud_fraction_threshold = cfg->cloud_fraction_threshold; d.cloud_mixing_ratio_threshold = cfg->cloud_mixing_ratio_threshold;
  d.min_gas_od_lw = cfg->min_gas_od_lw; d.min_gas_od_sw = cfg->min_gas_od_sw;
  d.cloud_inhom_decorr_scaling = cfg->cloud_inhom_decorr_scaling;
  d.sp.do_3d_effects = cfg->do_3d_effects; d.sp.entrapment = cfg->i_3d_sw_entrapment;
  d.sp.do_3d_lw_multilayer_effects = cfg->do_3d_lw_multilayer_effects; d.sp.do_lw_side_emissivity = cfg->do_lw_side_emissivity; d.sp.use_expm_everywhere = cfg->use_expm_everywhere;
  d.sp.two_regions = cfg->n_regions == 2 && ((cfg->do_sw && cfg->i_solver_sw == ECRAD_SOLVER_SPARTACUS) || (cfg->do_lw && cfg->i_solver_lw == ECRAD_SOLVER_SPARTACUS));
  d.sp.max_gas_od_3d = cfg->max_gas_od_3d; d.sp.max_cloud_od = cfg->max_cloud_od; d.sp.max_3d_transfer_rate = cfg->max_3d_transfer_rate;
  d.sp.min_cloud_effective_size = cfg->min_cloud_effective_size; d.sp.overhead_sun_factor = cfg->overhead_sun_factor;
  d.sp.overhang_factor = cfg->overhang_factor; d.sp.clear_to_thick_fraction = cfg->clear_to_thick_fraction;
  {
    // host-entry tile (measured, two overlapping compute sets; round 2 sweep in DESIGN.md section 4c): 4096 columns, first and
    // last tile 1024 (10 000 columns: 1024 + 3 x 2651 + 1024).  Smaller tiles lose to partial waves of the 128-column gas blocks,
    // bigger ones expose the first tile's H2D and the last tile's D2H.
    int t = 4096;
    // SPARTACUS keeps 3x3 matrices per (layer, g-point) between its kernels: about 5x the scratch per column
    if ((cfg->do_sw && cfg->i_solver_sw == ECRAD_SOLVER_SPARTACUS) || (cfg->do_lw && cfg->i_solver_lw == ECRAD_SOLVER_SPARTACUS)) {
      t /= 4;
      h->tile_cols_device = 4096;   // measured (columns/s at 20 000 / 50 000 columns): 2048 -> 100 k / 100 k, 4096 -> 105 k / 104 k, a third of the free memory -> 105 k / 93 k
    }
    h->tile_cols = t; h->edge_cols = t / 4;
  }
  if (const char* s = getenv("ECRAD_B200_TILE")) { int v = atoi(s); if (v > 0) h->tile_cols = h->tile_cols_device = v; }
  if (const char* s = getenv("ECRAD_B200_EDGE")) { int v = atoi(s); if (v > 0) h->edge_cols = v; }
  if (cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->s_comp[0], cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->s_comp[1], cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->s_aux1[1], cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->s_aux2[1], cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->s_aux1[0], cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->s_aux2[0], cudaStreamNonBlocking) != cudaSuccess) {
    fail(nullptr, "cannot create CUDA streams"); ecrad_b200_finalize(h); return 1;
  }
  for (auto& s : h->slot) {
    cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&s.compute_done, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&s.d2h_done, cudaEventDisableTiming);
  }
  for (int k = 0; k < 2; ++k) {
    cudaEventCreateWithFlags(&h->ev_fork[k], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_cloud[k], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_sw_done[k], cudaEventDisableTiming);
  }
  cudaEventCreateWithFlags(&h->ev_dev_done, cudaEventDisableTiming);
  if (const char* s2 = getenv("ECRAD_B200_SERIAL")) h->serial = atoi(s2) != 0;
  if (const char* s2 = getenv("ECRAD_B200_SCAN")) h->scan_solvers = atoi(s2) != 0;
  if (const char* s2 = getenv("ECRAD_B200_GAS")) h->gas_variant = atoi(s2) & 3;
  init_generator_constants();
  *handle = h;
  return 0;
}

void ecrad_b200_finalize(void* handle) {
  Handle* h = (Handle*)handle;
  if (!h) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  unregister_all(h);
  for (void* p : h->table_allocs) cudaFree(p);
  for (auto& ws : h->work) for (auto& b : ws) b.release();
  h->blocked.release();
  for (auto& s : h->slot) {
    for (auto& b : s.in) b.release();
    for (auto& b : s.out) b.release();
    for (auto& b : s.in32) b.release();
    for (auto& b : s.out32) b.release();
    if (s.h2d_done) cudaEventDestroy(s.h2d_done);
    if (s.compute_done) cudaEventDestroy(s.compute_done);
    if (s.d2h_done) cudaEventDestroy(s.d2h_done);
  }
  for (auto e : h->ev) cudaEventDestroy(e);
  if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
  for (cudaStream_t q : {h->s_comp[0], h->s_comp[1], h->s_aux1[0], h->s_aux1[1], h->s_aux2[0], h->s_aux2[1]}) if (q) cudaStreamDestroy(q);
  if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
  if (h->ev_dev_done) cudaEventDestroy(h->ev_dev_done);
  for (cudaEvent_t e : {h->ev_fork[0], h->ev_cloud[0], h->ev_sw_done[0], h->ev_fork[1], h->ev_cloud[1], h->ev_sw_done[1]}) if (e) cudaEventDestroy(e);
  delete h;
}

int64_t ecrad_b200_kernel_launches(void* handle) { return handle ? ((Handle*)handle)->launches : 0; }

int ecrad_b200_last_stage_ms(void* handle, float* ms, int max_stages) {
  Handle* h = (Handle*)handle;
  if (!h || !ms) return 0;
  std::lock_guard<std::mutex> lk(h->mu);
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  int n = max_stages < N_STAGE ? max_stages : N_STAGE;
  for (int s = 0; s < n; ++s) ms[s] = 0.f;
  for (int t = 0; t < h->ev_tiles; ++t)
    for (int s = 0; s < n; ++s) {
      float v = 0.f;
      if (cudaEventElapsedTime(&v, h->ev[(t * N_STAGE + s) * 2], h->ev[(t * N_STAGE + s) * 2 + 1]) == cudaSuccess) ms[s] += v;
    }
  return n;
}

int ecrad_b200_set_option(void* handle, const char* key, int value) {
  Handle* h = (Handle*)handle;
  if (!h || !key) return 1;
  std::lock_guard<std::mutex> lk(h->mu);
  if (!strcmp(key, "serial")) { h->serial = value != 0; return 0; }
  if (!strcmp(key, "register_host")) { h->register_host = value != 0; if (!value) unregister_all(h); return 0; }
  if (!strcmp(key, "scan_solvers")) { h->scan_solvers = value != 0; return 0; }
  if (!strcmp(key, "gas_variant")) { h->gas_variant = value & 3; return 0; }
  if (!strcmp(key, "tile_cols")) { if (value < 1) return fail(h, "tile_cols must be positive"); h->tile_cols = value; return 0; }
  if (!strcmp(key, "edge_cols")) { if (value < 1) return fail(h, "edge_cols must be positive"); h->edge_cols = value; h->edge_explicit = true; return 0; }
  if (!strcmp(key, "tile_ramp")) { h->tile_ramp = value != 0; return 0; }
  if (!strcmp(key, "tail_tiles")) { if (value < 1 || value > 8) return fail(h, "tail_tiles must be 1..8"); h->tail_tiles = value; return 0; }
  if (!strcmp(key, "tile_cols_device")) { if (value < 1) return fail(h, "tile_cols_device must be positive"); h->tile_cols_device = value; return 0; }
  return fail(h, "unknown option '%s'", key);
}

// single_level%spectral_solar_cycle_multiplier (radiation_single_level.F90:71) for the calls that follow: -1 solar minimum .. +1 maximum.
// calc_incoming_sw (radiation_ecckd.F90:946-962) stops when it is non-zero and the model carries no solar-cycle amplitude.
int ecrad_b200_set_solar_cycle_multiplier(void* handle, double multiplier) {
  Handle* h = (Handle*)handle;
  if (!h) return fail(nullptr, "ecrad_b200_set_solar_cycle_multiplier: null handle");
  std::lock_guard<std::mutex> lk(h->mu);
  if (multiplier != 0.0) {
    if (!h->dcfg.ckd_sw) return fail(h, "solar cycle only available with ecCKD gas optics model");   // radiation_config.F90:1200-1203
    if (!h->has_solar_cycle) return fail(h, "calc_incoming_sw: no information present on solar cycle (table 'ckd_sw_norm_amplitude_solar_irradiance')");
  }
  drain(h);
  h->dcfg.solar_cycle_multiplier = multiplier;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// host-buffer entry
// ---------------------------------------------------------------------------------------------------------
static int host_entry(Handle* h, int ncol, int nlev, int istartcol, int iendcol, const ecrad_b200_inputs* in, ecrad_b200_outputs* out, bool sp);

int ecrad_b200_radiation(void* handle, int ncol, int nlev, int istartcol, int iendcol, const ecrad_b200_inputs* in,
                         ecrad_b200_outputs* out) {
  Handle* h = (Handle*)handle;
  if (!h) return fail(nullptr, "ecrad_b200_radiation: null handle");
  std::lock_guard<std::mutex> lk(h->mu);
  return host_entry(h, ncol, nlev, istartcol, iendcol, in, out, false);
}
int ecrad_b200_radiation_sp(void* handle, int ncol, int nlev, int istartcol, int iendcol, const ecrad_b200_inputs* in,
                            ecrad_b200_outputs* out) {
  Handle* h = (Handle*)handle;
  if (!h) return fail(nullptr, "ecrad_b200_radiation_sp: null handle");
  std::lock_guard<std::mutex> lk(h->mu);
  return host_entry(h, ncol, nlev, istartcol, iendcol, in, out, true);
}

// rb = bytes per real of the caller's arrays: 8, or 4 for the single-precision boundary (float arrays behind the struct's pointers)
static int host_entry(Handle* h, int ncol, int nlev, int istartcol, int iendcol, const ecrad_b200_inputs* in, ecrad_b200_outputs* out, bool sp) {
  const size_t rb = sp ? 4 : 8;
  if (check_args(h, ncol, nlev, istartcol, iendcol, in, out)) return 1;
  CK(h, cudaSetDevice(h->device));
  const ecrad_b200_config& c = h->cfg;
  const int c_first = istartcol - 1, ntot = iendcol - istartcol + 1;
  // Tile schedule.  The first tile's H2D and the last tile's D2H are the only copies that cannot hide behind kernels, so both
  // edge tiles are short (a quarter of tile_cols); the columns in between are split into equal tiles of at most tile_cols.
  std::vector<int> tile_first, tile_n;
  {
    const int edge_max = h->edge_explicit ? h->tile_cols : h->tile_cols / 4;
    const int edge = edge_max < h->edge_cols ? edge_max : h->edge_cols;
    int pos = 0;
    auto push = [&](int n) { tile_first.push_back(pos); tile_n.push_back(n); pos += n; };
    std::vector<int> ramp;   // sizes of the growing tiles at the start (mirrored at the end)
    if (edge >= 64) {
      ramp.push_back(edge);
      if (h->tile_ramp) for (int e = 2 * edge; e < h->tile_cols; e *= 2) ramp.push_back(e);
    }
    int nramp = 0;
    for (int e : ramp) nramp += e;
    const int ntail = nramp + (h->tail_tiles - 1) * edge;
    if (!ramp.empty() && ntot >= nramp + ntail + h->tile_cols / 2) {
      for (size_t i = 0; i < ramp.size(); ++i) push(ramp[i]);
      const int rest = ntot - nramp - ntail, k = (rest + h->tile_cols - 1) / h->tile_cols;
      for (int i = 0; i < k; ++i) push(rest / k + (i < rest % k ? 1 : 0));
      for (size_t i = ramp.size(); i-- > 0;) push(ramp[i]);
      for (int i = 1; i < h->tail_tiles; ++i) push(edge);
    } else {
      const int k = (ntot + h->tile_cols - 1) / h->tile_cols;
      for (int i = 0; i < k; ++i) push(ntot / k + (i < ntot % k ? 1 : 0));
    }
  }
  const int ntiles = (int)tile_n.size();
  int cap = 0;
  for (int n : tile_n) cap = n > cap ? n : cap;
  if (ensure_work(h, 0, cap, nlev)) return 1;
  if (ntiles > 1 && ensure_work(h, 1, cap, nlev)) return 1;
  if (ensure_events(h, ntiles)) return 1;
  InDesc id[N_IN]; OutDesc od[N_OUT];
  fill_descs(c, nlev, in, out, id, od);
  const bool mcica_sw = c.do_sw && c.i_solver_sw == ECRAD_SOLVER_MCICA, mcica_lw = c.do_lw && c.i_solver_lw == ECRAD_SOLVER_MCICA;
  // outputs the kernels do not produce in this configuration are left untouched on the host
  auto out_active = [&](int k) {
    if (!od[k].host) return false;
    const bool lw = (k <= 3) || k == 10 || k == 11 || (k >= 13 && k <= 16) || k == 29 || k == 30 || k == 31 || k == 39 || k == 40;
    if (lw && !c.do_lw) return false;
    if (!lw && !c.do_sw) return false;
    if (k == 11) return mcica_lw || (c.do_lw && (c.i_solver_lw == ECRAD_SOLVER_TRIPLECLOUDS || c.i_solver_lw == ECRAD_SOLVER_SPARTACUS));
    if (k == 12) return mcica_sw || (c.do_sw && (c.i_solver_sw == ECRAD_SOLVER_TRIPLECLOUDS || c.i_solver_sw == ECRAD_SOLVER_SPARTACUS));
    if (k == 10) return c.do_lw_derivatives != 0;
    if (k >= 23 && k <= 26) return c.do_surface_sw_spectral_flux != 0 && (k < 25 || c.do_clear);
    if (k == 27 || k == 28) return c.do_canopy_fluxes_sw != 0;
    if (k == 29) return c.do_canopy_fluxes_lw != 0;
    if (k == 35) return c.i_solver_sw == ECRAD_SOLVER_TRIPLECLOUDS;                                  // sw_dn_toa_g: only Tripleclouds sets it
    if (k == 36) return c.do_toa_spectral_flux != 0 && c.i_solver_sw == ECRAD_SOLVER_TRIPLECLOUDS;
    if (k >= 37) return c.do_toa_spectral_flux != 0 && (c.do_clear || (k != 38 && k != 40));
    if (k >= 30) {   // per-band profiles: Cloudless and Tripleclouds solvers with do_save_spectral_flux
      const int sol = k <= 31 ? c.i_solver_lw : c.i_solver_sw;
      return c.do_save_spectral_flux != 0 && sol != ECRAD_SOLVER_MCICA;   // every solver but McICA stores per-band profiles
    }
    return true;
  };
  if (h->register_host) {
    for (int k = 0; k < N_IN; ++k) if (id[k].host && id[k].rows > 0) register_range(h, id[k].host, (id[k].elem == 8 ? rb : 4) * ncol * id[k].rows);
    for (int k = 0; k < N_OUT; ++k)
      if (out_active(k)) register_range(h, od[k].host, rb * (size_t)ncol * od[k].rows * (od[k].kind == 2 ? (size_t)(nlev + 1) : 1));
  }
  for (auto& s : h->slot) s.used = false;
  // ECRAD_B200_TIMELINE=1: per-tile begin/end of H2D, kernels and D2H of this call, printed to stderr (a tuning aid)
  static const bool timeline = getenv("ECRAD_B200_TIMELINE") != nullptr;
  std::vector<cudaEvent_t> tl;
  if (timeline) { tl.resize((size_t)ntiles * 6 + 1); for (auto& e : tl) cudaEventCreate(&e); cudaEventRecord(tl[(size_t)ntiles * 6], h->s_h2d); }
  for (int t = 0; t < ntiles; ++t) {
    Slot& s = h->slot[t & 1];
    const int c0 = c_first + tile_first[t], nt = tile_n[t];
    void* ip[N_IN]; void* op[N_OUT];
    // ---- H2D (this slot's buffers are free once the kernels and copies of tile t-2 are done) ----
    if (s.used) { CKD(h, cudaStreamWaitEvent(h->s_h2d, s.compute_done, 0)); CKD(h, cudaStreamWaitEvent(h->s_h2d, s.d2h_done, 0)); }
    CvtJobs jin, jout; jin.n = jout.n = 0;
    long long cvt_in_max = 0, cvt_out_max = 0;
    if (timeline) cudaEventRecord(tl[t * 6 + 0], h->s_h2d);
    for (int k = 0; k < N_IN; ++k) {
      ip[k] = nullptr;
      if (!id[k].host || id[k].rows <= 0) continue;
      const size_t el = (size_t)id[k].elem;
      CKD(h, s.in[k].reserve(el * cap * id[k].rows));
      ip[k] = s.in[k].p;
      const bool cvt = sp && el == 8;
      void* dst = s.in[k].p;
      const size_t hel = cvt ? 4 : el;   // bytes per element on the host
      if (cvt) {
        CKD(h, s.in32[k].reserve(4 * (size_t)cap * id[k].rows));
        dst = s.in32[k].p;
        jin.j[jin.n++] = {dst, s.in[k].p, (long long)cap * id[k].rows};
        if ((long long)cap * id[k].rows > cvt_in_max) cvt_in_max = (long long)cap * id[k].rows;
      }
      CKD(h, cudaMemcpy2DAsync(dst, hel * cap, (const char*)id[k].host + hel * c0, hel * ncol, hel * nt, id[k].rows, cudaMemcpyHostToDevice, h->s_h2d));
    }
    for (int k = 0; k < N_OUT; ++k) {
      op[k] = nullptr;
      if (!out_active(k)) continue;
      const size_t per_col = od[k].kind == 0 ? (size_t)od[k].rows : od[k].kind == 1 ? (size_t)od[k].rows : (size_t)od[k].rows * (nlev + 1);
      CKD(h, s.out[k].reserve(8 * per_col * cap));
      op[k] = s.out[k].p;
      if (sp) {
        CKD(h, s.out32[k].reserve(4 * per_col * cap));
        jout.j[jout.n++] = {s.out[k].p, s.out32[k].p, (long long)per_col * cap};
        if ((long long)per_col * cap > cvt_out_max) cvt_out_max = (long long)per_col * cap;
      }
    }
    // night columns keep the caller's cloud_cover_sw (the reference does not touch it): stage the current values; likewise
    // sw_dn_toa_g / sw_dn_toa_band of night columns (Tripleclouds sets them for sunlit columns only)
    for (int k : {12, 35, 36}) {
      if (!op[k]) continue;
      const size_t n = (size_t)nt * od[k].rows;
      void* dst = sp ? s.out32[k].p : op[k];
      CKD(h, cudaMemcpyAsync(dst, (const char*)od[k].host + rb * (size_t)c0 * od[k].rows, rb * n, cudaMemcpyHostToDevice, h->s_h2d));
      if (sp) { jin.j[jin.n++] = {dst, op[k], (long long)n}; if ((long long)n > cvt_in_max) cvt_in_max = (long long)n; }
    }
    CKD(h, cudaEventRecord(s.h2d_done, h->s_h2d));
    if (timeline) cudaEventRecord(tl[t * 6 + 1], h->s_h2d);
    // ---- kernels ----
    DevIn di; DevOut dout;
    make_views(ip, op, cap, cap, in->solar_irradiance, di, dout);
    const int set = t & 1;   // staging slot and compute set alternate together
    if (h->dev_pending) CKD(h, cudaStreamWaitEvent(h->s_comp[set], h->ev_dev_done, 0));
    CKD(h, cudaStreamWaitEvent(h->s_comp[set], s.h2d_done, 0));
    if (s.used) CKD(h, cudaStreamWaitEvent(h->s_comp[set], s.d2h_done, 0));
    if (timeline) cudaEventRecord(tl[t * 6 + 2], h->s_comp[set]);
    if (sp && jin.n) {
      convert_kernel<<<dim3((unsigned)((cvt_in_max + 1023) / 1024 > 592 ? 592 : (cvt_in_max + 1023) / 1024), jin.n), 256, 0, h->s_comp[set]>>>(jin, 1);
      h->launches += 1;
    }
    if (run_tile(h, set, di, dout, nt, nlev, h->s_comp[set], &h->ev[t * 2 * N_STAGE])) { drain(h); return 1; }
    if (sp) {
      if (c.do_clouds && ip[17]) {   // the cropped cloud fraction goes back as float through the input's float staging buffer
        jout.j[jout.n++] = {ip[17], s.in32[17].p, (long long)cap * nlev};
        if ((long long)cap * nlev > cvt_out_max) cvt_out_max = (long long)cap * nlev;
      }
      if (jout.n) {
        convert_kernel<<<dim3((unsigned)((cvt_out_max + 1023) / 1024 > 592 ? 592 : (cvt_out_max + 1023) / 1024), jout.n), 256, 0, h->s_comp[set]>>>(jout, 0);
        h->launches += 1;
      }
    }
    CKD(h, cudaEventRecord(s.compute_done, h->s_comp[set]));
    if (timeline) cudaEventRecord(tl[t * 6 + 3], h->s_comp[set]);
    // ---- D2H ----
    CKD(h, cudaStreamWaitEvent(h->s_d2h, s.compute_done, 0));
    if (timeline) cudaEventRecord(tl[t * 6 + 4], h->s_d2h);
    for (int k = 0; k < N_OUT; ++k) {
      if (!op[k]) continue;
      const char* src = (const char*)(sp ? s.out32[k].p : op[k]);
      char* hst = (char*)od[k].host;
      if (od[k].kind == 0)
        CKD(h, cudaMemcpy2DAsync(hst + rb * c0, rb * (size_t)ncol, src, rb * (size_t)cap, rb * (size_t)nt, od[k].rows, cudaMemcpyDeviceToHost, h->s_d2h));
      else if (od[k].kind == 1)
        CKD(h, cudaMemcpyAsync(hst + rb * (size_t)c0 * od[k].rows, src, rb * (size_t)nt * od[k].rows, cudaMemcpyDeviceToHost, h->s_d2h));
      else   // (nband, ncol, nlev+1): row = half-level, nt*nband contiguous values per row
        CKD(h, cudaMemcpy2DAsync(hst + rb * (size_t)c0 * od[k].rows, rb * (size_t)ncol * od[k].rows, src, rb * (size_t)cap * od[k].rows,
                                rb * (size_t)nt * od[k].rows, nlev + 1, cudaMemcpyDeviceToHost, h->s_d2h));
    }
    if (c.do_clouds && ip[17])   // cropped cloud fraction back into the caller's array (cloud%crop_cloud_fraction)
      CKD(h, cudaMemcpy2DAsync((char*)in->cloud_fraction + rb * c0, rb * (size_t)ncol, sp ? s.in32[17].p : ip[17], rb * (size_t)cap, rb * (size_t)nt, nlev,
                              cudaMemcpyDeviceToHost, h->s_d2h));
    CKD(h, cudaEventRecord(s.d2h_done, h->s_d2h));
    if (timeline) cudaEventRecord(tl[t * 6 + 5], h->s_d2h);
    s.used = true;
  }
  h->dev_pending = false;   // (the compute streams waited for it, and they are drained below)
  CKD(h, cudaStreamSynchronize(h->s_d2h));
  CKD(h, cudaStreamSynchronize(h->s_comp[0]));
  CKD(h, cudaStreamSynchronize(h->s_comp[1]));
  if (timeline) {
    const cudaEvent_t t0 = tl[(size_t)ntiles * 6];
    fprintf(stderr, "ecrad_b200 timeline (ms after the call's first H2D was queued): tile columns | H2D | kernels | D2H\n");
    for (int t = 0; t < ntiles; ++t) {
      float v[6];
      for (int k = 0; k < 6; ++k) cudaEventElapsedTime(&v[k], t0, tl[t * 6 + k]);
      fprintf(stderr, "  %2d %5d | %6.2f-%6.2f | %6.2f-%6.2f | %6.2f-%6.2f\n", t, tile_n[t], v[0], v[1], v[2], v[3], v[4], v[5]);
    }
    for (auto& e : tl) cudaEventDestroy(e);
  }
  CKD(h, cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// device-resident entry: all pointers are device pointers with leading dimension ncol; not synchronised
// ---------------------------------------------------------------------------------------------------------
int ecrad_b200_radiation_device(void* handle, int ncol, int nlev, const ecrad_b200_inputs* in, ecrad_b200_outputs* out, void* cuda_stream) {
  return ecrad_b200_radiation_device_ld(handle, ncol, nlev, ncol, ncol, in, out, cuda_stream);
}

static int device_entry_locked(Handle* h, int ncol, int nlev, int ld_in, int ld_out, const ecrad_b200_inputs* in, ecrad_b200_outputs* out,
                               void* cuda_stream) {
  if (check_args(h, ncol, nlev, 1, ncol, in, out)) return 1;
  if (ld_in < ncol || ld_out < ncol) return fail(h, "leading dimensions (%d, %d) smaller than ncol = %d", ld_in, ld_out, ncol);
  CK(h, cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const ecrad_b200_config& c = h->cfg;
  int tile = h->tile_cols_device;
  if (ncol > h->w_cols[0] || nlev != h->w_nlev[0]) {   // growing the scratch: stay within a third of the memory that is free now
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) {
      const size_t per_col = work_bytes_per_column(h, nlev);
      const size_t fit = (fr / 3 + (size_t)h->w_cols[0] * per_col) / (per_col ? per_col : 1);
      if ((size_t)tile > fit) tile = fit < 256 ? 256 : (int)fit;
    }
  }
  const int ntiles = (ncol + tile - 1) / tile;
  const int cap = (ncol + ntiles - 1) / ntiles;
  if (ensure_work(h, 0, cap, nlev)) return 1;
  if (ensure_events(h, ntiles)) return 1;
  InDesc id[N_IN]; OutDesc od[N_OUT];
  fill_descs(c, nlev, in, out, id, od);
  // the scratch (set 0) may still be in use by an earlier device-entry call on another stream
  if (h->dev_pending) CK(h, cudaStreamWaitEvent(st, h->ev_dev_done, 0));
  for (int t = 0; t < ntiles; ++t) {
    const int c0 = t * cap, nt = (ncol - c0) < cap ? (ncol - c0) : cap;
    void* ip[N_IN]; void* op[N_OUT];
    for (int k = 0; k < N_IN; ++k) ip[k] = id[k].host ? (void*)((const char*)id[k].host + (size_t)id[k].elem * c0) : nullptr;
    for (int k = 0; k < N_OUT; ++k) {
      op[k] = nullptr;
      if (!od[k].host) continue;
      op[k] = od[k].kind == 0 ? (void*)(od[k].host + c0) : (void*)(od[k].host + (size_t)c0 * od[k].rows);
    }
    DevIn di; DevOut dout;
    make_views(ip, op, ld_in, ld_out, in->solar_irradiance, di, dout);
    if (run_tile(h, 0, di, dout, nt, nlev, st, &h->ev[t * 2 * N_STAGE])) return 1;
  }
  CK(h, cudaEventRecord(h->ev_dev_done, st));
  h->dev_pending = true;
  return 0;
}

int ecrad_b200_radiation_device_ld(void* handle, int ncol, int nlev, int ld_in, int ld_out, const ecrad_b200_inputs* in,
                                   ecrad_b200_outputs* out, void* cuda_stream) {
  Handle* h = (Handle*)handle;
  if (!h) return fail(nullptr, "ecrad_b200_radiation_device: null handle");
  std::lock_guard<std::mutex> lk(h->mu);
  return device_entry_locked(h, ncol, nlev, ld_in, ld_out, in, out, cuda_stream);
}

// ---------------------------------------------------------------------------------------------------------
// save_radiative_properties (radiation_interface.F90:405-425): the optics stages only, then a gather into the reference's layout.
// A diagnostic: plain synchronous copies, tiles of at most 2048 columns on compute set 0.
// ---------------------------------------------------------------------------------------------------------
int ecrad_b200_save_radiative_properties(void* handle, int ncol, int nlev, int istartcol, int iendcol, const ecrad_b200_inputs* in,
                                         const ecrad_b200_radiative_properties* props) {
  Handle* h = (Handle*)handle;
  if (!h) return fail(nullptr, "ecrad_b200_save_radiative_properties: null handle");
  std::lock_guard<std::mutex> lk(h->mu);
  if (!in || !props) return fail(h, "ecrad_b200_save_radiative_properties: null argument");
  ecrad_b200_outputs none; memset(&none, 0, sizeof(none));
  InDesc id[N_IN]; OutDesc od[N_OUT];
  if (ncol < 1 || nlev < 2 || nlev > 256 || istartcol < 1 || iendcol > ncol || istartcol > iendcol) return fail(h, "ecrad_b200_save_radiative_properties: bad dimensions or column range");
  fill_descs(h->cfg, nlev, in, &none, id, od);
  {
    // the arrays radiation() requires (check_args), minus the outputs
    const ecrad_b200_config& c = h->cfg;
    const bool need[N_IN] = {c.do_sw != 0, c.do_lw != 0, c.do_sw != 0, false, c.do_lw != 0, c.do_clouds != 0, true, true,
                             true, true, true, true, true, true, true, true, true,
                             c.do_clouds != 0, c.do_clouds != 0, c.do_clouds != 0, c.do_clouds != 0, c.do_clouds != 0, c.do_clouds != 0, c.do_clouds != 0,
                             c.use_aerosols != 0, c.use_aerosols != 0, false, false};
    for (int k = 0; k < N_IN; ++k) if (need[k] && !id[k].host) return fail(h, "ecrad_b200_save_radiative_properties: a required input array (index %d of ecrad_b200_inputs) is null", k);
  }
  CK(h, cudaSetDevice(h->device));
  drain(h);
  const ecrad_b200_config& c = h->cfg;
  const int n = iendcol - istartcol + 1, first = istartcol - 1;
  const int cap = n < 2048 ? n : 2048;
  if (ensure_work(h, 0, cap, nlev)) return 1;
  if (ensure_events(h, 1)) return 1;
  cudaStream_t st = h->s_comp[0];
  if (h->dev_pending) CK(h, cudaStreamWaitEvent(st, h->ev_dev_done, 0));
  std::vector<void*> tmp;
  auto dalloc = [&](size_t bytes) -> void* { void* p = nullptr; if (cudaMalloc(&p, bytes ? bytes : 8) != cudaSuccess) return nullptr; tmp.push_back(p); return p; };
  auto release = [&]() { for (void* p : tmp) cudaFree(p); tmp.clear(); };
  void* ip[N_IN]; void* op[N_OUT];
  for (int k = 0; k < N_OUT; ++k) op[k] = nullptr;
  for (int k = 0; k < N_IN; ++k) {
    ip[k] = nullptr;
    if (!id[k].host) continue;
    ip[k] = dalloc((size_t)id[k].rows * cap * id[k].elem);
    if (!ip[k]) { release(); return fail(h, "ecrad_b200_save_radiative_properties: out of device memory"); }
  }
  struct PD { double* host; size_t per_col; double* dev; };
  const size_t gl = (size_t)c.n_g_lw, gs = (size_t)c.n_g_sw, bl = (size_t)c.n_bands_lw, bs = (size_t)c.n_bands_sw, nl = (size_t)nlev;
  PD pd[18] = {{props->planck_hl, gl * (nl + 1)}, {props->lw_emission, gl}, {props->lw_albedo, gl}, {props->sw_albedo_direct, gs},
               {props->sw_albedo_diffuse, gs}, {props->incoming_sw, gs}, {props->od_lw, gl * nl}, {props->ssa_lw, gl * nl}, {props->g_lw, gl * nl},
               {props->od_sw, gs * nl}, {props->ssa_sw, gs * nl}, {props->g_sw, gs * nl}, {props->od_lw_cloud, bl * nl}, {props->ssa_lw_cloud, bl * nl},
               {props->g_lw_cloud, bl * nl}, {props->od_sw_cloud, bs * nl}, {props->ssa_sw_cloud, bs * nl}, {props->g_sw_cloud, bs * nl}};
  const bool is_lw[18] = {true, true, true, false, false, false, true, true, true, false, false, false, true, true, true, false, false, false};
  for (int k = 0; k < 18; ++k) {
    pd[k].dev = nullptr;
    if (!pd[k].host || !(is_lw[k] ? c.do_lw : c.do_sw)) continue;
    pd[k].dev = (double*)dalloc(pd[k].per_col * cap * 8);
    if (!pd[k].dev) { release(); return fail(h, "ecrad_b200_save_radiative_properties: out of device memory"); }
  }
  int rc = 0;
  double* ones = nullptr;
  if (c.do_sw && h->dcfg.ckd_sw) {
    ones = (double*)dalloc((size_t)cap * 8);
    const std::vector<double> one((size_t)cap, 1.0);
    if (!ones || cudaMemcpy(ones, one.data(), (size_t)cap * 8, cudaMemcpyHostToDevice) != cudaSuccess) { release(); return fail(h, "ecrad_b200_save_radiative_properties: out of device memory"); }
  }
  for (int c0 = 0; c0 < n && !rc; c0 += cap) {
    const int nt = (n - c0) < cap ? (n - c0) : cap;
    for (int k = 0; k < N_IN && !rc; ++k) {
      if (!ip[k]) continue;
      const size_t rb = (size_t)id[k].elem;
      if (cudaMemcpy2D(ip[k], rb * cap, (const char*)id[k].host + rb * (first + c0), rb * (size_t)ncol, rb * nt, id[k].rows, cudaMemcpyHostToDevice) != cudaSuccess) rc = 1;
    }
    if (rc || cudaDeviceSynchronize() != cudaSuccess) { rc = 1; break; }   // (pageable sources: the DMA may trail the call's return)
    DevIn di; DevOut dout;
    make_views(ip, op, cap, cap, in->solar_irradiance, di, dout);
    // ecCKD computes the shortwave properties of every column, sunlit or not (radiation_ecckd_interface.F90:257-292); the kernels skip
    // night columns because no solver reads them, so this pass marks every column sunlit.  (RRTMG-IFS skips them in the reference too.)
    if (ones) di.cos_sza = ones;
    if (run_tile(h, 0, di, dout, nt, nlev, st, &h->ev[0], true)) { rc = 2; break; }
    DevProps dp = {pd[0].dev, pd[1].dev, pd[2].dev, pd[3].dev, pd[4].dev, pd[5].dev, pd[6].dev, pd[7].dev, pd[8].dev,
                   pd[9].dev, pd[10].dev, pd[11].dev, pd[12].dev, pd[13].dev, pd[14].dev, pd[15].dev, pd[16].dev, pd[17].dev};
    launch_radprops_gather(h->T, h->dcfg, di, h->w[0], dp, nt, nlev, st);
    h->launches += 1;
    if (cudaStreamSynchronize(st) != cudaSuccess) { rc = 1; break; }
    for (int k = 0; k < 18 && !rc; ++k)
      if (pd[k].dev && cudaMemcpy(pd[k].host + pd[k].per_col * (size_t)(first + c0), pd[k].dev, pd[k].per_col * nt * 8, cudaMemcpyDeviceToHost) != cudaSuccess) rc = 1;
  }
  const cudaError_t e = cudaGetLastError();
  release();
  if (rc == 2) return 1;   // (run_tile recorded the message)
  if (rc || e != cudaSuccess) return fail(h, "ecrad_b200_save_radiative_properties: %s", cudaGetErrorString(e != cudaSuccess ? e : cudaErrorUnknown));
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// blocked (NPROMA) entry
// ---------------------------------------------------------------------------------------------------------
int ecrad_b200_radiation_blocked(void* handle, int ncol_total, int nlev, const ecrad_b200_block_layout* lay, const double* zrgp_in, double* zrgp_out) {
  Handle* h = (Handle*)handle;
  if (!h) return fail(nullptr, "ecrad_b200_radiation_blocked: null handle");
  if (!lay || !zrgp_in || !zrgp_out) return fail(h, "ecrad_b200_radiation_blocked: null argument");
  if (lay->struct_bytes != (int32_t)sizeof(ecrad_b200_block_layout)) return fail(h, "ecrad_b200_block_layout: struct_bytes mismatch (ABI)");
  const int nproma = lay->nproma, nblocks = lay->nblocks;
  if (nproma < 1 || nblocks < 1 || ncol_total < 1 || ncol_total > (long long)nproma * nblocks || ncol_total <= (long long)nproma * (nblocks - 1))
    return fail(h, "ecrad_b200_radiation_blocked: ncol_total = %d does not fit %d blocks of %d columns", ncol_total, nblocks, nproma);
  const ecrad_b200_config& c = h->cfg;
  // the column-layout view of what the slabs hold: reuse the host-entry descriptors with dummy non-null "host" pointers
  ecrad_b200_inputs in; ecrad_b200_outputs out;
  memset(&in, 0, sizeof in); memset(&out, 0, sizeof out);
  in.struct_bytes = (int32_t)sizeof in; out.struct_bytes = (int32_t)sizeof out;
  in.solar_irradiance = lay->solar_irradiance;
  {
    const void** ip = (const void**)&in.cos_sza;
    for (int k = 0; k < N_IN; ++k) ip[k] = lay->in_field[k] >= 0 ? (const void*)zrgp_in : nullptr;
    double** op = (double**)&out.lw_up;
    for (int k = 0; k < N_OUT; ++k) op[k] = lay->out_field[k] >= 0 ? zrgp_out : nullptr;
  }
  std::lock_guard<std::mutex> lk(h->mu);
  if (check_args(h, ncol_total, nlev, 1, ncol_total, &in, &out)) return 1;
  CK(h, cudaSetDevice(h->device));
  InDesc id[N_IN]; OutDesc od[N_OUT];
  fill_descs(c, nlev, &in, &out, id, od);
  const size_t nin = (size_t)nproma * lay->nfields_in * nblocks, nout = (size_t)nproma * lay->nfields_out * nblocks;
  // device memory: the two slabs + column-layout arrays of every present input / output
  size_t need = 8 * (nin + nout);
  for (int k = 0; k < N_IN; ++k) if (id[k].host && id[k].rows > 0) need += ((size_t)id[k].elem * ncol_total * id[k].rows + 255) & ~(size_t)255;
  for (int k = 0; k < N_OUT; ++k)
    if (od[k].host) need += (8 * (size_t)ncol_total * od[k].rows * (od[k].kind == 2 ? (size_t)(nlev + 1) : 1) + 255) & ~(size_t)255;
  CK(h, h->blocked.reserve(need + 512));
  char* base = (char*)h->blocked.p;
  double* d_in = (double*)base; base += 8 * nin;
  double* d_out = (double*)base; base += 8 * nout;
  base = (char*)(((uintptr_t)base + 255) & ~(uintptr_t)255);
  void* ipd[N_IN]; void* opd[N_OUT];
  BlockJobs Ji, Jo; Ji.n = Jo.n = 0;
  long long max_in = 0, max_out = 0;
  for (int k = 0; k < N_IN; ++k) {
    ipd[k] = nullptr;
    if (!id[k].host || id[k].rows <= 0) continue;
    if (lay->in_field[k] + id[k].rows > lay->nfields_in) return fail(h, "ecrad_b200_radiation_blocked: input %d does not fit nfields_in", k);
    ipd[k] = base; base += ((size_t)id[k].elem * ncol_total * id[k].rows + 255) & ~(size_t)255;
    Ji.j[Ji.n++] = {ipd[k], lay->in_field[k], id[k].rows, 0, id[k].elem == 4};
    if ((long long)ncol_total * id[k].rows > max_in) max_in = (long long)ncol_total * id[k].rows;
  }
  for (int k = 0; k < N_OUT; ++k) {
    opd[k] = nullptr;
    if (!od[k].host) continue;
    const int nrows = od[k].kind == 2 ? od[k].rows * (nlev + 1) : od[k].rows;
    if (lay->out_field[k] + nrows > lay->nfields_out) return fail(h, "ecrad_b200_radiation_blocked: output %d does not fit nfields_out", k);
    opd[k] = base; base += (8 * (size_t)ncol_total * nrows + 255) & ~(size_t)255;
    Jo.j[Jo.n++] = {opd[k], lay->out_field[k], od[k].rows, od[k].kind, 0};
    if ((long long)ncol_total * nrows > max_out) max_out = (long long)ncol_total * nrows;
  }
  cudaStream_t st = h->s_comp[0];
  if (h->register_host) { register_range(h, zrgp_in, 8 * nin); register_range(h, zrgp_out, 8 * nout); }
  CKD(h, cudaMemcpyAsync(d_in, zrgp_in, 8 * nin, cudaMemcpyHostToDevice, st));
  CKD(h, cudaMemcpyAsync(d_out, zrgp_out, 8 * nout, cudaMemcpyHostToDevice, st));   // outputs the kernels leave alone keep the caller's values
  block_unpack_kernel<<<dim3((unsigned)((max_in + 255) / 256), Ji.n), 256, 0, st>>>(Ji, d_in, nproma, lay->nfields_in, ncol_total, nlev + 1);
  // every output array starts from the caller's values: what the kernels leave alone (cloud_cover_sw of night columns, components
  // the configuration does not produce) comes back unchanged, as with the column-layout entry
  block_unpack_kernel<<<dim3((unsigned)((max_out + 255) / 256), Jo.n), 256, 0, st>>>(Jo, d_out, nproma, lay->nfields_out, ncol_total, nlev + 1);
  h->launches += 2;
  // all columns, leading dimension ncol_total: the device entry does the tiling and the launches
  ecrad_b200_inputs din = in; ecrad_b200_outputs dout = out;
  {
    const void** ip = (const void**)&din.cos_sza;
    for (int k = 0; k < N_IN; ++k) ip[k] = ipd[k];
    double** op = (double**)&dout.lw_up;
    for (int k = 0; k < N_OUT; ++k) op[k] = (double*)opd[k];
  }
  if (device_entry_locked(h, ncol_total, nlev, ncol_total, ncol_total, &din, &dout, (void*)st)) { drain(h); return 1; }
  block_pack_kernel<<<dim3((unsigned)((max_out + 255) / 256), Jo.n), 256, 0, st>>>(Jo, d_out, nproma, lay->nfields_out, ncol_total, nlev + 1);
  h->launches += 1;
  CKD(h, cudaMemcpyAsync(zrgp_out, d_out, 8 * nout, cudaMemcpyDeviceToHost, st));
  CKD(h, cudaStreamSynchronize(st));
  CKD(h, cudaGetLastError());
  return 0;
}

}  // extern "C"
