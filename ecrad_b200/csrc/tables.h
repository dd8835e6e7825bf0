// tables.h -- host-side named-array directory (what setup_radiation leaves in Fortran module storage) and the
// packer that turns it into the GPU layout: per band one table of rows, g-point fastest (the transpose of the
// reference's ABSA(row, ig) layout, so that lanes = g-points read contiguous memory).
//
// Host code only (runs once in ecrad_b200_setup).  No device code here.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include <math.h>

#include "cloud_core.h"
#include "ecckd_core.h"
#include "gas_core.h"

struct ecrad_b200_tables {
  struct Arr { int dtype; int ndim; int64_t dims[4]; std::vector<char> data; };
  std::map<std::string, Arr> a;

  int add(const char* name, int dtype, int ndim, const int64_t* dims, const void* data) {
    if (!name || ndim < 1 || ndim > 4 || (dtype != 0 && dtype != 1) || !data || !dims) return 1;
    Arr x; x.dtype = dtype; x.ndim = ndim;
    size_t n = 1;
    for (int i = 0; i < 4; ++i) {
      x.dims[i] = i < ndim ? dims[i] : 1;
      if (x.dims[i] < 1 || (uint64_t)x.dims[i] > (uint64_t)1 << 31 || n > ((size_t)1 << 34) / (size_t)x.dims[i]) return 1;   // no table is anywhere near 2^34 elements
      n *= (size_t)x.dims[i];
    }
    x.data.assign((const char*)data, (const char*)data + n * (dtype == 0 ? 8 : 4));
    a[name] = std::move(x);
    return 0;
  }
  // "ETB1" blob written by tools/extract_rrtmg_tables.py (ecrad_b200/tables.py)
  int load_file(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) return 1;
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    if (sz < 8) { fclose(f); return 2; }   // (ftell failure is -1)
    std::vector<char> buf((size_t)sz);
    if (fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) { fclose(f); return 2; }
    fclose(f);
    return load_memory(buf.data(), (size_t)sz);
  }
  int load_memory(const char* buf, size_t sz) {
    if (sz < 8 || memcmp(buf, "ETB1", 4)) return 3;
    uint32_t n; memcpy(&n, buf + 4, 4);
    struct Entry { char name[48]; int32_t dtype, ndim; int64_t dims[4]; int64_t offset; };
    static_assert(sizeof(Entry) == 96, "ETB1 entry layout");
    if (8 + (size_t)n * sizeof(Entry) > sz) return 3;
    for (uint32_t i = 0; i < n; ++i) {
      Entry e; memcpy(&e, buf + 8 + (size_t)i * sizeof(Entry), sizeof(Entry));
      char nm[49]; memcpy(nm, e.name, 48); nm[48] = 0;
      // the payload must lie behind the directory and inside the blob: validate dims and size BEFORE anything is copied
      if (e.ndim < 1 || e.ndim > 4 || (e.dtype != 0 && e.dtype != 1)) return 4;
      size_t cnt = 1;
      for (int k = 0; k < e.ndim; ++k) {
        if (e.dims[k] < 1 || (uint64_t)e.dims[k] > (uint64_t)sz || cnt > sz / (size_t)e.dims[k]) return 4;
        cnt *= (size_t)e.dims[k];
      }
      const size_t nbytes = cnt * (e.dtype == 0 ? 8 : 4);
      if (nbytes / (e.dtype == 0 ? 8 : 4) != cnt) return 4;
      if (e.offset < (int64_t)(8 + (size_t)n * sizeof(Entry)) || (size_t)e.offset > sz || nbytes > sz - (size_t)e.offset) return 4;
      if (add(nm, e.dtype, e.ndim, e.dims, buf + e.offset)) return 4;
    }
    return 0;
  }
  const Arr* find(const std::string& n) const { auto it = a.find(n); return it == a.end() ? nullptr : &it->second; }
  const Arr& req(const std::string& n) const {
    const Arr* x = find(n);
    if (!x) throw std::runtime_error("table '" + n + "' missing");
    return *x;
  }
  const double* d(const std::string& n) const { const Arr& x = req(n); if (x.dtype != 0) throw std::runtime_error(n + ": not float64"); return (const double*)x.data.data(); }
  const int32_t* i(const std::string& n) const { const Arr& x = req(n); if (x.dtype != 1) throw std::runtime_error(n + ": not int32"); return (const int32_t*)x.data.data(); }
};

namespace ecb {

struct PackedTables {
  GasMeta meta;
  std::vector<double> lwtab, swtab;   // packed band tables
  CloudMeta cloud;
  std::vector<double> pdf_val;        // (ncdf, nfsd) Fortran order, as the reference stores it
  std::vector<double> sw_albedo_weights;  // (n_albedo_sw, n_bands_sw)
  std::vector<int32_t> i_albedo_from_band_sw;  // (n_bands_sw) 1-based; do_nearest_spectral_sw_albedo
  std::vector<int32_t> i_emiss_from_band_lw;  // (n_bands_lw) 1-based
  std::vector<double> lw_emiss_weights;   // (n_emiss_lw, n_bands_lw)
  int n_emiss_lw = 0;
  bool is_ecckd = false;              // both spectra on ecCKD
  bool ckd_lw = false, ckd_sw = false;   // per spectrum (mixed gas models: one of the two)
  int ng_lw = NG_LW, ng_sw = NG_SW, nb_lw = NB_LW, nb_sw = NB_SW;
  CkdMeta ckd{};                      // ecCKD gas optics + generalised cloud optics
  std::vector<double> ckdtab;
  AerMeta aer;                        // aer.ntype == 0: no aerosol tables
  std::vector<double> aertab;
  int n_albedo_sw = 0;
};

namespace detail {
struct BandPacker {
  const ecrad_b200_tables& T;
  std::vector<double>& tab;
  BandMeta& B;
  std::string prefix;
  // Append `rows` rows of B.ng values; returns element offset of the first row.
  int reserve(int rows) { int off = (int)tab.size(); tab.resize(tab.size() + (size_t)rows * B.ng, 0.0); return off; }
  // array stored (rows..., ng) with rows fastest -> transpose
  int rows_fast(const std::string& name) {
    const auto& x = T.req(prefix + name);
    size_t total = 1; for (int k = 0; k < x.ndim; ++k) total *= (size_t)x.dims[k];
    if (x.dims[x.ndim - 1] != B.ng) throw std::runtime_error(prefix + name + ": last dim != ng");
    int rows = (int)(total / B.ng);
    int off = reserve(rows);
    const double* s = (const double*)x.data.data();
    for (int g = 0; g < B.ng; ++g) for (int r = 0; r < rows; ++r) tab[(size_t)off + (size_t)r * B.ng + g] = s[(size_t)g * rows + r];
    return off;
  }
  // array stored (ng, rows) or (ng) -> as is
  int g_fast(const std::string& name) {
    const auto& x = T.req(prefix + name);
    size_t total = 1; for (int k = 0; k < x.ndim; ++k) total *= (size_t)x.dims[k];
    if (x.dims[0] != B.ng) throw std::runtime_error(prefix + name + ": first dim != ng");
    int rows = (int)(total / B.ng);
    int off = reserve(rows);
    memcpy(&tab[off], x.data.data(), total * 8);
    return off;
  }
  bool has(const std::string& name) const { return T.find(prefix + name) != nullptr; }
};
}  // namespace detail

inline void pack_common(const ecrad_b200_tables& T, PackedTables& P);

// ecCKD blob of tools/extract_ecckd_tables.py: "ckd_{lw,sw}_*" (read_ckd_model, radiation_ecckd.F90:128-226) and
// "gco_{lw,sw}_{0,1}_*" (setup_general_cloud_optics, radiation_general_cloud_optics_data.F90:71-243)
// one spectrum's ecCKD model + its per-g-point cloud look-up tables; bands == g-points in that spectrum
inline void pack_ckd_spectrum(const ecrad_b200_tables& T, PackedTables& P, int sw) {
  auto put = [&](const std::string& nm, size_t expect) {
    const auto& x = T.req(nm);
    if (x.dtype != 0 || x.data.size() != expect * 8) throw std::runtime_error(nm + ": unexpected size");
    size_t off = P.ckdtab.size();
    P.ckdtab.insert(P.ckdtab.end(), (const double*)x.data.data(), (const double*)x.data.data() + expect);
    return off;
  };
  static const int slot_of_code[13] = {-1, 0, 1, 8, 3, -1, 2, -1, 4, 5, 6, 7, -1};   // DevIn::gas order: h2o co2 ch4 n2o cfc11 cfc12 hcfc22 ccl4 o3
  // GasMolarMass(0:12), radiation_gas_constants.F90:43-56: AirMolarMass / GasMolarMass turns a mass mixing ratio into a mole fraction
  static const double gas_molar_mass[13] = {0.0, 18.0152833, 44.011, 47.9982, 44.013, 28.0101, 16.043, 31.9988, 137.3686, 120.914, 86.469, 153.823, 46.0055};
  {
    const std::string pre = sw ? "ckd_sw_" : "ckd_lw_";
    CkdModel& m = sw ? P.ckd.sw : P.ckd.lw;
    const double* meta = T.d(pre + "meta");
    m.ng = (int)meta[0]; m.npress = (int)meta[1]; m.ntemp = (int)meta[2]; m.nplanck = (int)meta[3]; m.ngas = (int)meta[4];
    m.log_pressure1 = meta[5]; m.d_log_pressure = meta[6]; m.d_temperature = meta[7];
    m.temperature1_planck = meta[8]; m.d_temperature_planck = meta[9];
    if ((meta[10] != 0.0) != (sw != 0)) throw std::runtime_error(pre + "meta: longwave/shortwave model mismatch");
    if (m.ngas > CKD_MAXGAS) throw std::runtime_error("more than 12 ecCKD gases");
    if (m.ng != 32 && m.ng != 64 && m.ng != 96) throw std::runtime_error("ecCKD models with 32, 64 or 96 g-points are built in");
    m.off_temperature1 = put(pre + "temperature1", (size_t)m.npress);
    if (sw) {
      m.off_solar = put(pre + "norm_solar_irradiance", (size_t)m.ng);
      m.off_rayleigh = put(pre + "rayleigh_molar_scat", (size_t)m.ng);
      m.off_solar_amp = T.find(pre + "norm_amplitude_solar_irradiance") ? (long long)put(pre + "norm_amplitude_solar_irradiance", (size_t)m.ng) : -1;
    } else {
      m.off_planck = put(pre + "planck_function", (size_t)m.ng * m.nplanck);
    }
    const double* gm = T.d(pre + "gas_meta");
    for (int j = 0; j < m.ngas; ++j) {
      CkdGas& g = m.gas[j];
      const int code = (int)gm[6 * j];
      g.dep = (int)gm[6 * j + 1]; g.reference_mole_frac = gm[6 * j + 2]; g.n_mole_frac = (int)gm[6 * j + 3];
      g.log_mole_frac1 = gm[6 * j + 4]; g.d_log_mole_frac = gm[6 * j + 5];
      g.mole_frac1 = exp(g.log_mole_frac1);
      g.lut = g.dep == CKD_CONC_LUT ? m.nlut++ : -1;
      g.slot = (code >= 0 && code <= 12) ? slot_of_code[code] : -1;
      g.mmr_scaling = (code >= 1 && code <= 12) ? 1.0 * 28.970 / gas_molar_mass[code] : 1.0;   // gas%get_scaling, radiation_gas.F90:471-486
      const size_t n = (size_t)m.ng * m.npress * m.ntemp * (g.dep == CKD_CONC_LUT ? g.n_mole_frac : 1);
      g.off = put(pre + "gas" + std::to_string(j) + "_molar_abs", n);
    }
    for (int jt = 0; jt < 2; ++jt) {
      GcoType& c = sw ? P.ckd.gco_sw[jt] : P.ckd.gco_lw[jt];
      const std::string gp = std::string("gco_") + (sw ? "sw_" : "lw_") + std::to_string(jt) + "_";
      const double* cm = T.d(gp + "meta");
      c.nre = (int)cm[0]; c.re0 = cm[1]; c.dre = cm[2];
      c.off_me = put(gp + "mass_ext", (size_t)m.ng * c.nre);
      c.off_ssa = put(gp + "ssa", (size_t)m.ng * c.nre);
      c.off_g = put(gp + "asymmetry", (size_t)m.ng * c.nre);
    }
    // bands == g-points (radiation_ecckd_interface.F90:60-63, :95-98)
    if (sw) {
      P.ckd_sw = true; P.ng_sw = P.nb_sw = m.ng;
      memset(P.meta.band_of_g_sw, 0, sizeof(P.meta.band_of_g_sw)); memset(P.meta.sw, 0, sizeof(P.meta.sw));
      for (int g = 0; g < m.ng; ++g) { P.meta.band_of_g_sw[g] = g; P.meta.sw[g].ng = 1; P.meta.sw[g].g0 = g; }
    } else {
      P.ckd_lw = true; P.ng_lw = P.nb_lw = m.ng;
      memset(P.meta.band_of_g_lw, 0, sizeof(P.meta.band_of_g_lw)); memset(P.meta.lw, 0, sizeof(P.meta.lw));
      for (int g = 0; g < m.ng; ++g) { P.meta.band_of_g_lw[g] = g; P.meta.lw[g].ng = 1; P.meta.lw[g].g0 = g; }
    }
  }
}

inline void pack_ecckd(const ecrad_b200_tables& T, PackedTables& P) {
  memset(&P.meta, 0, sizeof(P.meta));
  memset(&P.ckd, 0, sizeof(P.ckd));
  pack_ckd_spectrum(T, P, 0);
  pack_ckd_spectrum(T, P, 1);
  P.is_ecckd = true;
  memset(&P.cloud, 0, sizeof(P.cloud));
  pack_common(T, P);
}

// liq_model / ice_model: config%i_liq_model / i_ice_model (RRTMG-band cloud optics; ignored by the ecCKD tables)
// general_cloud: config%use_general_cloud_optics with RRTMG-IFS (the look-up tables "gco_*" per RRTMG band instead of the coefficients)
inline void pack_tables(const ecrad_b200_tables& T, PackedTables& P, int liq_model = LIQ_SOCRATES, int ice_model = ICE_FU, bool general_cloud = false) {
  // a spectrum runs ecCKD when its model is in the directory; one ecCKD spectrum next to the RRTMG tables = mixed gas models
  // (radiation_interface.F90:333-355, test/ifs/configCY49R1_mixed.nam)
  const bool has_ckd_lw = T.find("ckd_lw_meta") != nullptr, has_ckd_sw = T.find("ckd_sw_meta") != nullptr;
  if (has_ckd_lw && has_ckd_sw) { pack_ecckd(T, P); return; }
  const bool mixed = has_ckd_lw || has_ckd_sw;
  if (mixed && !general_cloud) throw std::runtime_error("an ecCKD spectrum needs use_general_cloud_optics");
  GasMeta& M = P.meta;
  memset(&M, 0, sizeof(M));
  auto copy = [&](double* dst, const char* name, size_t n) {
    const auto& x = T.req(name);
    if (x.data.size() != n * 8) throw std::runtime_error(std::string(name) + ": unexpected size");
    memcpy(dst, x.data.data(), n * 8);
  };
  copy(M.preflog_lw, "lw_PREFLOG", 59); copy(M.tref_lw, "lw_TREF", 59); copy(M.chi_mls, "lw_CHI_MLS", 7 * 59);
  copy(M.preflog_sw, "sw_PREFLOG", 59); copy(M.tref_sw, "sw_TREF", 59);
  copy(M.totplnk, "lw_TOTPLNK", 181 * 16); copy(M.delwave, "lw_DELWAVE", 16);
  {
    static const int pr[6][2] = {{1, 2}, {1, 3}, {1, 4}, {1, 6}, {4, 2}, {3, 2}};
    for (int k = 0; k < 6; ++k)
      for (int j = 0; j < 59; ++j) M.chi_rat[k][j] = M.chi_mls[j * 7 + pr[k][0] - 1] / M.chi_mls[j * 7 + pr[k][1] - 1];
  }
  const int32_t* ngc_lw = T.i("lw_NGC");
  const int32_t* ngc_sw = T.i("sw_NGC");
  const int32_t* ngb_lw = T.i("lw_NGB");
  const int32_t* ngb_sw = T.i("sw_NGBSW");
  for (int g = 0; g < NG_LW; ++g) M.band_of_g_lw[g] = ngb_lw[g] - 1;
  for (int g = 0; g < NG_SW; ++g) M.band_of_g_sw[g] = ngb_sw[g] - 16;

  // ---- LW bands ----
  static const char* lw_minor[16][5] = {
      {"KA_MN2", "KB_MN2", 0, 0, 0}, {0, 0, 0, 0, 0}, {"KA_MN2O", "KB_MN2O", 0, 0, 0}, {0, 0, 0, 0, 0},
      {"KA_MO3", 0, 0, 0, 0}, {"KA_MCO2", 0, 0, 0, 0}, {"KA_MCO2", "KB_MCO2", 0, 0, 0},
      {"KA_MCO2", "KA_MO3", "KA_MN2O", "KB_MCO2", "KB_MN2O"}, {"KA_MN2O", "KB_MN2O", 0, 0, 0}, {0, 0, 0, 0, 0},
      {"KA_MO2", "KB_MO2", 0, 0, 0}, {0, 0, 0, 0, 0}, {"KA_MCO2", 0, "KB_MO3", 0, 0}, {0, 0, 0, 0, 0},
      {"KA_MN2", 0, 0, 0, 0}, {0, 0, 0, 0, 0}};
  static const char* lw_const[16][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}, {"CCL4", 0}, {"CFC11ADJ", "CFC12"}, {0, 0},
                                        {"CFC12", "CFC22ADJ"}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}};
  int g0 = 0;
  for (int b = 0; b < NB_LW; ++b) {
    BandMeta& B = M.lw[b];
    B.ng = ngc_lw[b]; B.g0 = g0; g0 += B.ng;
    for (int s = 0; s < 16; ++s) B.sec[s] = -1;
    if (P.lwtab.size() & 1) P.lwtab.push_back(0.0);   // bands start 16-byte aligned (bulk copies)
    detail::BandPacker pk{T, P.lwtab, B, "lw" + std::to_string(b + 1) + "_"};
    if (pk.has("ABSA")) B.sec[L_ABSA] = pk.rows_fast("ABSA");
    if (pk.has("ABSB")) B.sec[L_ABSB] = pk.rows_fast("ABSB");
    B.sec[SEC_SMALL] = (int)P.lwtab.size();
    M.lw_rows[b][0] = B.sec[L_ABSA] >= 0 ? ((B.sec[L_ABSB] >= 0 ? B.sec[L_ABSB] : B.sec[SEC_SMALL]) - B.sec[L_ABSA]) / B.ng : 0;
    M.lw_rows[b][1] = B.sec[L_ABSB] >= 0 ? (B.sec[SEC_SMALL] - B.sec[L_ABSB]) / B.ng : 0;
    B.sec[L_SELF] = pk.rows_fast("SELFREF");
    B.sec[L_FOR] = pk.rows_fast("FORREF");
    B.sec[L_FRACA] = pk.g_fast("FRACREFA");
    if (pk.has("FRACREFB")) B.sec[L_FRACB] = pk.g_fast("FRACREFB");
    for (int m = 0; m < 5; ++m) if (lw_minor[b][m]) B.sec[L_M0 + m] = pk.rows_fast(lw_minor[b][m]);
    for (int m = 0; m < 2; ++m) if (lw_const[b][m]) B.sec[L_C0 + m] = pk.g_fast(lw_const[b][m]);
    if (b + 1 == 4 || b + 1 == 7) {
      int off = pk.reserve(1);
      for (int g = 0; g < B.ng; ++g) P.lwtab[off + g] = 1.0;
      if (b + 1 == 4) {  // rrtm_taumol4.F90:283-289 (single-precision literals)
        static const float f[7] = {0.92f, 0.88f, 1.07f, 1.1f, 0.99f, 0.88f, 0.943f};
        for (int k = 0; k < 7; ++k) P.lwtab[off + 7 + k] = (double)f[k];
      } else {           // rrtm_taumol7.F90 (double-precision literals)
        static const double f[6] = {0.92, 0.88, 1.07, 1.1, 0.99, 0.855};
        for (int k = 0; k < 6; ++k) P.lwtab[off + 5 + k] = f[k];
      }
      B.sec[L_POST] = off;
    }
    B.sec[SEC_END] = (int)P.lwtab.size();
    {
      // rows of each section = distance to the next section that starts after it
      for (int k = 0; k < L_NSEC; ++k) {
        M.lw_sec_rows[b][k] = 0;
        if (B.sec[k] < 0) continue;
        int nxt = B.sec[SEC_END];
        for (int j = 0; j < L_NSEC; ++j) if (B.sec[j] > B.sec[k] && B.sec[j] < nxt) nxt = B.sec[j];
        M.lw_sec_rows[b][k] = (short)((nxt - B.sec[k]) / B.ng);
      }
      // which small sections the band routine (gas_core.h lw_build_list) reads in the lower / upper atmosphere
      unsigned lo = (1u << L_SELF) | (1u << L_FOR) | (1u << L_FRACA) | (1u << L_C0) | (1u << L_C1);
      unsigned hi = (1u << L_FOR) | (1u << L_FRACA) | (1u << L_FRACB) | (1u << L_C0) | (1u << L_C1) | (1u << L_POST);
      for (int m = 0; m < 5; ++m)
        if (lw_minor[b][m]) { if (lw_minor[b][m][1] == 'A') lo |= 1u << (L_M0 + m); else hi |= 1u << (L_M0 + m); }
      M.lw_sec_low[b] = (unsigned short)lo; M.lw_sec_high[b] = (unsigned short)hi;
    }
    if (B.ng != kNgLwBand[b]) throw std::runtime_error("lw_NGC is not the 140-g-point reduction this build is compiled for");
  }
  if (g0 != NG_LW) throw std::runtime_error("lw_NGC does not sum to 140");

  // ---- SW bands ----
  g0 = 0;
  for (int b = 0; b < NB_SW; ++b) {
    BandMeta& B = M.sw[b];
    const int jb = b + 16;
    B.ng = ngc_sw[b]; B.g0 = g0; g0 += B.ng;
    for (int s = 0; s < 16; ++s) B.sec[s] = -1;
    if (P.swtab.size() & 1) P.swtab.push_back(0.0);
    detail::BandPacker pk{T, P.swtab, B, "sw" + std::to_string(jb) + "_"};
    if (pk.has("ABSA")) B.sec[S_ABSA] = pk.rows_fast("ABSA");
    if (pk.has("ABSB")) B.sec[S_ABSB] = pk.rows_fast("ABSB");
    B.sec[SEC_SMALL] = (int)P.swtab.size();
    M.sw_rows[b][0] = B.sec[S_ABSA] >= 0 ? ((B.sec[S_ABSB] >= 0 ? B.sec[S_ABSB] : B.sec[SEC_SMALL]) - B.sec[S_ABSA]) / B.ng : 0;
    M.sw_rows[b][1] = B.sec[S_ABSB] >= 0 ? (B.sec[SEC_SMALL] - B.sec[S_ABSB]) / B.ng : 0;
    if (pk.has("SELFREFC")) B.sec[S_SELF] = pk.rows_fast("SELFREFC");
    if (pk.has("FORREFC")) { B.sec[S_FOR] = pk.rows_fast("FORREFC"); M.nfor_sw[b] = (int)T.req(pk.prefix + "FORREFC").dims[0]; }
    B.sec[S_SFLUX] = pk.g_fast("SFLUXREFC");
    if (pk.has("RAYLC")) B.sec[S_RAYA] = pk.g_fast("RAYLC");
    if (pk.has("RAYLAC")) B.sec[S_RAYA] = pk.g_fast("RAYLAC");
    if (pk.has("RAYLBC")) B.sec[S_RAYB] = pk.g_fast("RAYLBC");
    if (jb == 20) B.sec[S_X0] = pk.g_fast("ABSCH4C");
    if (jb == 24 || jb == 25) { B.sec[S_X0] = pk.g_fast("ABSO3AC"); B.sec[S_X1] = pk.g_fast("ABSO3BC"); }
    if (jb == 29) { B.sec[S_X0] = pk.g_fast("ABSCO2C"); B.sec[S_X1] = pk.g_fast("ABSH2OC"); }
    int off = pk.reserve(1);
    for (int g = 0; g < B.ng; ++g) P.swtab[off + g] = 1.0;
    B.sec[S_ONES] = off;
    const auto* sr = T.find(pk.prefix + (jb == 16 ? "STRRAT1" : "STRRAT"));
    M.strrat_sw[b] = sr ? ((const double*)sr->data.data())[0] : 0.0;
    const auto* rl = T.find(pk.prefix + "RAYL");
    M.rayl_sw[b] = rl ? ((const double*)rl->data.data())[0] : 0.0;
    const auto* lr = T.find(pk.prefix + "LAYREFFR");
    M.layreffr_sw[b] = lr ? ((const int32_t*)lr->data.data())[0] : 0;
    B.sec[SEC_END] = (int)P.swtab.size();
    for (int k = 0; k < S_NSEC; ++k) {
      M.sw_sec_rows[b][k] = 0;
      if (B.sec[k] < 0) continue;
      int nxt = B.sec[SEC_END];
      for (int j = 0; j < S_NSEC; ++j) if (B.sec[j] > B.sec[k] && B.sec[j] < nxt) nxt = B.sec[j];
      M.sw_sec_rows[b][k] = (short)((nxt - B.sec[k]) / B.ng);
    }
    if (B.ng != kNgSwBand[b]) throw std::runtime_error("sw_NGC is not the 112-g-point reduction this build is compiled for");
  }
  if (g0 != NG_SW) throw std::runtime_error("sw_NGC does not sum to 112");
  M.givfac_23 = T.d("sw23_GIVFAC")[0];
  M.scalekur_27 = T.d("sw27_SCALEKUR")[0];

  // ---- cloud optics coefficients, PDF look-up table, surface mappings ----
  CloudMeta& C = P.cloud;
  memset(&C, 0, sizeof(C));
  memset(&P.ckd, 0, sizeof(P.ckd));
  if (general_cloud) {
    // generalised cloud optics on the RRTMG bands (radiation_general_cloud_optics.F90:39-110 with use_bands = .true.): the same
    // look-up tables the ecCKD path uses per g-point, one row per band here; read by general_cloud_optics_kernel through T.ckd / T.ckdtab
    for (int sw = 0; sw < 2; ++sw) {
      if (sw ? has_ckd_sw : has_ckd_lw) { pack_ckd_spectrum(T, P, sw); continue; }   // mixed: this spectrum per g-point
      for (int jt = 0; jt < 2; ++jt) {
        const int nb = sw ? NB_SW : NB_LW;
        GcoType& c = sw ? P.ckd.gco_sw[jt] : P.ckd.gco_lw[jt];
        const std::string gp = std::string("gco_") + (sw ? "sw_" : "lw_") + std::to_string(jt) + "_";
        const double* cm = T.d(gp + "meta");
        c.nre = (int)cm[0]; c.re0 = cm[1]; c.dre = cm[2];
        auto put = [&](const std::string& nm) {
          const auto& x = T.req(nm);
          if (x.dtype != 0 || x.data.size() != (size_t)nb * c.nre * 8) throw std::runtime_error(nm + ": not a (n_bands, n_effective_radius) table");
          const size_t off = P.ckdtab.size();
          P.ckdtab.insert(P.ckdtab.end(), (const double*)x.data.data(), (const double*)x.data.data() + (size_t)nb * c.nre);
          return off;
        };
        c.off_me = put(gp + "mass_ext"); c.off_ssa = put(gp + "ssa"); c.off_g = put(gp + "asymmetry");
      }
    }
    memset(&C, 0, sizeof(C));   // (the band parameterisations are not used)
    C.liq_model = liq_model; C.ice_model = ice_model;
  } else
  // Coefficients of the configured parameterisations (radiation_cloud_optics.F90:46-216 checks the same coefficient counts).  A host
  // model registers what it loaded as liq_coeff_* / ice_coeff_* (/ ice_coeff_gen); the stand-alone blob also holds the files of the
  // other parameterisations under "<name>.<model>".
  {
    static const char* liq_tag[] = {"", "", "slingo"};
    static const int liq_n[][2] = {{0, 0}, {16, 16}, {13, 6}};                       // coefficients per band: longwave, shortwave
    static const char* ice_tag[] = {"", "", "baran", "baran2016", "baran2017", "yi"};
    static const int ice_n[][2] = {{0, 0}, {11, 10}, {9, 9}, {5, 5}, {9, 9}, {69, 69}};
    if (liq_model < LIQ_SOCRATES || liq_model > LIQ_SLINGO) throw std::runtime_error("liquid optics model not available (SOCRATES and Slingo are)");
    if (ice_model < ICE_FU || ice_model > ICE_YI) throw std::runtime_error("ice optics model not available (Fu-IFS, Baran, Baran2016, Baran2017 and Yi are)");
    auto pick = [&](double* dst, const char* base, const char* tag, int nb, int ncoef) {
      const std::string tagged = std::string(base) + "." + tag;
      const std::string name = (tag[0] && T.find(tagged)) ? tagged : std::string(base);
      const auto& x = T.req(name);
      if (x.dtype != 0 || x.data.size() != (size_t)nb * ncoef * 8)
        throw std::runtime_error(name + ": not the (" + std::to_string(nb) + ", " + std::to_string(ncoef) + ") coefficient array of the configured cloud optics model");
      memcpy(dst, x.data.data(), x.data.size());
    };
    pick(C.liq_lw, "liq_coeff_lw", liq_tag[liq_model], 16, liq_n[liq_model][0]);
    pick(C.liq_sw, "liq_coeff_sw", liq_tag[liq_model], 14, liq_n[liq_model][1]);
    pick(C.ice_lw, "ice_coeff_lw", ice_tag[ice_model], 16, ice_n[ice_model][0]);
    pick(C.ice_sw, "ice_coeff_sw", ice_tag[ice_model], 14, ice_n[ice_model][1]);
    if (ice_model == ICE_BARAN2017) pick(C.ice_gen, "ice_coeff_gen", ice_tag[ice_model], 5, 1);
    C.liq_model = liq_model; C.ice_model = ice_model;
  }
  pack_common(T, P);
}

// tables both gas models share: McICA PDF look-up table, surface albedo / emissivity mappings, aerosol optics
inline void pack_common(const ecrad_b200_tables& T, PackedTables& P) {
  CloudMeta& C = P.cloud;
  const int NB_SW = P.nb_sw, NB_LW = P.nb_lw;   // shadow the RRTMG constants: bands of this configuration
  const auto& pv = T.req("pdf_val");
  const double* fsd = T.d("pdf_fsd");
  C.pdf_ncdf = (int)pv.dims[0]; C.pdf_nfsd = (int)pv.dims[1];
  C.pdf_fsd1 = fsd[0]; C.pdf_inv_fsd_interval = 1.0 / (fsd[1] - fsd[0]);  // radiation_pdf_sampler.F90:83-93
  P.pdf_val.assign((const double*)pv.data.data(), (const double*)pv.data.data() + (size_t)C.pdf_ncdf * C.pdf_nfsd);
  if (const auto* w = T.find("sw_albedo_weights")) {
    P.n_albedo_sw = (int)w->dims[0];
    if (w->dims[1] != NB_SW) throw std::runtime_error("sw_albedo_weights: second dim != number of shortwave bands");
    P.sw_albedo_weights.assign((const double*)w->data.data(), (const double*)w->data.data() + (size_t)P.n_albedo_sw * NB_SW);
  }
  // do_nearest_spectral_sw_albedo: config%i_albedo_from_band_sw(n_bands_sw), 1-based (radiation_config.F90:1994-1997)
  if (const auto* ia = T.find("i_albedo_from_band_sw")) {
    if (ia->dims[0] != NB_SW) throw std::runtime_error("i_albedo_from_band_sw: size != number of shortwave bands");
    P.i_albedo_from_band_sw.assign((const int32_t*)ia->data.data(), (const int32_t*)ia->data.data() + NB_SW);
  }
  // ---- aerosol optics (only if the host registered the type map) ----
  memset(&P.aer, 0, sizeof(P.aer));
  if (T.find("aerosol_iclass") && T.find("aerosol_itype")) {
    AerMeta& A = P.aer;
    const auto& ic = T.req("aerosol_iclass");
    A.ntype = (int)ic.dims[0];
    if (A.ntype > 32) throw std::runtime_error("more than 32 aerosol types");
    memcpy(A.iclass, T.i("aerosol_iclass"), 4 * A.ntype);
    memcpy(A.itype, T.i("aerosol_itype"), 4 * A.ntype);
    const auto& rh = T.req("aer_rh_lower");
    A.nrh = (int)rh.dims[0];
    if (A.nrh > 16) throw std::runtime_error("more than 16 aerosol humidity bins");
    memcpy(A.rh_lower, rh.data.data(), 8 * A.nrh);
    A.n_phobic = (int)T.req("aer_mass_ext_sw_phobic").dims[1];
    A.n_philic = (int)T.req("aer_mass_ext_sw_philic").dims[2];
    auto put = [&](const char* nm, size_t expect) {
      const auto& x = T.req(nm);
      if (x.data.size() != expect * 8) throw std::runtime_error(std::string(nm) + ": unexpected size");
      int off = (int)P.aertab.size();
      P.aertab.insert(P.aertab.end(), (const double*)x.data.data(), (const double*)x.data.data() + expect);
      return off;
    };
    const size_t pb_sw = (size_t)NB_SW * A.n_phobic, pb_lw = (size_t)NB_LW * A.n_phobic;
    const size_t pl_sw = (size_t)NB_SW * A.nrh * A.n_philic, pl_lw = (size_t)NB_LW * A.nrh * A.n_philic;
    A.me_sw_phobic = put("aer_mass_ext_sw_phobic", pb_sw); A.ssa_sw_phobic = put("aer_ssa_sw_phobic", pb_sw); A.g_sw_phobic = put("aer_g_sw_phobic", pb_sw);
    A.me_lw_phobic = put("aer_mass_ext_lw_phobic", pb_lw); A.ssa_lw_phobic = put("aer_ssa_lw_phobic", pb_lw);
    A.g_lw_phobic = T.find("aer_g_lw_phobic") ? put("aer_g_lw_phobic", pb_lw) : -1;   // (only read with do_lw_aerosol_scattering)
    A.g_lw_philic = T.find("aer_g_lw_philic") ? put("aer_g_lw_philic", pl_lw) : -1;
    A.me_sw_philic = put("aer_mass_ext_sw_philic", pl_sw); A.ssa_sw_philic = put("aer_ssa_sw_philic", pl_sw); A.g_sw_philic = put("aer_g_sw_philic", pl_sw);
    A.me_lw_philic = put("aer_mass_ext_lw_philic", pl_lw); A.ssa_lw_philic = put("aer_ssa_lw_philic", pl_lw);
    for (int k = 0; k < A.ntype; ++k) {
      if (A.iclass[k] == 1 && (A.itype[k] < 1 || A.itype[k] > A.n_phobic)) throw std::runtime_error("hydrophobic aerosol type out of range");
      if (A.iclass[k] == 2 && (A.itype[k] < 1 || A.itype[k] > A.n_philic)) throw std::runtime_error("hydrophilic aerosol type out of range");
    }
  }
  if (const auto* e = T.find("i_emiss_from_band_lw")) {
    if (e->dims[0] != NB_LW) throw std::runtime_error("i_emiss_from_band_lw: size != number of longwave bands");
    P.i_emiss_from_band_lw.assign((const int32_t*)e->data.data(), (const int32_t*)e->data.data() + NB_LW);
  }
  if (const auto* e = T.find("lw_emiss_weights")) {
    if (e->dims[1] != NB_LW) throw std::runtime_error("lw_emiss_weights: second dim != number of longwave bands");
    P.n_emiss_lw = (int)e->dims[0];
    P.lw_emiss_weights.assign((const double*)e->data.data(), (const double*)e->data.data() + (size_t)P.n_emiss_lw * NB_LW);
  }
}

}  // namespace ecb
