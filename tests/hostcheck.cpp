// hostcheck.cpp -- TEST-ONLY harness: replays the math of ecrad_b200/csrc/*_core.h on the CPU for one column, so the
// table packer and the stencil builders can be checked against the oracle without a GPU (pytest -m "not gpu").
// It is not part of the product library and is never loaded by ecrad_b200.
#include <stdlib.h>

#include <vector>

#include "../ecrad_b200/csrc/tables.h"
#include "../ecrad_b200/csrc/sp_core.h"

using namespace ecb;

extern "C" {

void* hc_load(const char* path) {
  auto* T = new ecrad_b200_tables();
  if (T->load_file(path)) { delete T; return nullptr; }
  auto* P = new PackedTables();
  try { pack_tables(*T, *P); } catch (const std::exception& e) { fprintf(stderr, "hostcheck: %s\n", e.what()); delete T; delete P; return nullptr; }
  delete T;
  return P;
}
// the same with the coefficient arrays of another liquid / ice optics model (config%i_liq_model, i_ice_model)
void* hc_load_models(const char* path, int liq_model, int ice_model) {
  auto* T = new ecrad_b200_tables();
  if (T->load_file(path)) { delete T; return nullptr; }
  auto* P = new PackedTables();
  try { pack_tables(*T, *P, liq_model, ice_model); } catch (const std::exception& e) { fprintf(stderr, "hostcheck: %s\n", e.what()); delete T; delete P; return nullptr; }
  delete T;
  return P;
}
// cloud_core.h on one column: od / ssa / g per band of every cloudy layer, [nlev][nb] (zero where there is no cloud)
void hc_cloud_optics(void* p, int nlev, const double* p_hl, const double* t_hl, const double* frac, const double* q_liq, const double* q_ice,
                     const double* re_liq, const double* re_ice, int lw_scattering, int fu_bug, int delta_with_gases, double* od_lw,
                     double* ssa_lw, double* g_lw, double* od_sw, double* ssa_sw, double* g_sw) {
  const CloudMeta& C = ((PackedTables*)p)->cloud;
  for (int l = 0; l < nlev; ++l) {
    for (int b = 0; b < 16; ++b) { od_lw[l * 16 + b] = 0.0; ssa_lw[l * 16 + b] = 0.0; g_lw[l * 16 + b] = 0.0; }
    for (int b = 0; b < 14; ++b) { od_sw[l * 14 + b] = 0.0; ssa_sw[l * 14 + b] = 0.0; g_sw[l * 14 + b] = 0.0; }
    if (!(frac[l] > 0.0)) continue;
    const double factor = (p_hl[l + 1] - p_hl[l]) / (9.80665 * frac[l]);
    CloudLayerIn L;
    L.q_ice = q_ice[l]; L.lwp = factor * q_liq[l]; L.iwp = factor * q_ice[l];
    L.re_liq = re_liq[l]; L.re_ice = re_ice[l]; L.temperature = 0.5 * (t_hl[l] + t_hl[l + 1]);
    for (int b = 0; b < 16; ++b) { CloudBandOut r = cloud_optics_lw(C, b, L, lw_scattering != 0, fu_bug != 0); od_lw[l * 16 + b] = r.od; ssa_lw[l * 16 + b] = r.ssa; g_lw[l * 16 + b] = r.g; }
    for (int b = 0; b < 14; ++b) { CloudBandOut r = cloud_optics_sw(C, b, L, delta_with_gases != 0); od_sw[l * 14 + b] = r.od; ssa_sw[l * 14 + b] = r.ssa; g_sw[l * 14 + b] = r.g; }
  }
}
void hc_free(void* p) { delete (PackedTables*)p; }
int hc_table_sizes(void* p, int* lw, int* sw) { auto* P = (PackedTables*)p; *lw = (int)P->lwtab.size(); *sw = (int)P->swtab.size(); return (int)sizeof(GasMeta); }

// One column, ecRad level order (0 = top).  gas[9][nlev]: h2o co2 ch4 n2o cfc11 cfc12 hcfc22 ccl4 o3 (mmr).
// Outputs: od_lw[nlev][140] (unclamped), pfrac[nlev][140], planck_hl[(nlev+1)][140], od_sw, ssa_sw [nlev][112], incsol[112]
// plus max list lengths for the sizing asserts.
int hc_gas_column(void* p, int nlev, const double* p_hl, const double* t_hl, const double* gas, double skin_t,
                  double* od_lw, double* pfrac, double* planck_hl, double* planck_surf, double* od_sw, double* ssa_sw,
                  double* incsol, int* kmax_lw, int* kmax_sw) {
  auto* P = (PackedTables*)p;
  const GasMeta& M = P->meta;
  std::vector<LevGas> G(nlev);
  std::vector<LwLev> LL(nlev);
  std::vector<SwLev> SL(nlev);
  int laytrop_lw = 0, laytrop_sw = 0;
  for (int jl = 0; jl < nlev; ++jl) {
    const double* g = gas;
    lev_prepare(p_hl[jl], p_hl[jl + 1], t_hl[jl], t_hl[jl + 1], g[0 * nlev + jl], g[1 * nlev + jl], g[2 * nlev + jl],
                g[3 * nlev + jl], g[4 * nlev + jl], g[5 * nlev + jl], g[6 * nlev + jl], g[7 * nlev + jl], g[8 * nlev + jl], G[jl]);
    lw_setcoef(M, G[jl], LL[jl]);
    sw_setcoef(M, G[jl], SL[jl]);
    laytrop_lw += LL[jl].tropo;
    laytrop_sw += SL[jl].tropo;
  }
  *kmax_lw = 0; *kmax_sw = 0;
  Term tt[64];
  for (int jl = 0; jl < nlev; ++jl) {
    const int il = nlev - jl;  // RRTMG layer index (1 = bottom)
    for (int b = 0; b < NB_LW; ++b) {
      ListOut out{tt, 0, 1};
      int post;
      PlanckFrac pf = lw_build_list(M, M.lw[b], LL[jl], b, il <= laytrop_lw, out, &post);
      if (out.n > *kmax_lw) *kmax_lw = out.n;
      const BandMeta& B = M.lw[b];
      for (int ig = 0; ig < B.ng; ++ig) {
        double tau = 0.0;
        for (int k = 0; k < out.n; ++k) tau += tt[k].c * P->lwtab[tt[k].o + ig];
        if (post >= 0) tau *= P->lwtab[post + ig];
        od_lw[(size_t)jl * NG_LW + B.g0 + ig] = tau;
        pfrac[(size_t)jl * NG_LW + B.g0 + ig] = pf.c0 * P->lwtab[pf.o0 + ig] + pf.c1 * P->lwtab[pf.o1 + ig];
      }
    }
  }
  for (int jh = 0; jh <= nlev; ++jh) {
    int jl = jh == 0 ? 0 : jh - 1;  // layer whose Planck fraction is used (radiation_ifs_rrtm.F90:741-743)
    for (int g = 0; g < NG_LW; ++g)
      planck_hl[(size_t)jh * NG_LW + g] = planck_band(M, t_hl[jh], M.band_of_g_lw[g]) * pfrac[(size_t)jl * NG_LW + g];
  }
  for (int g = 0; g < NG_LW; ++g) planck_surf[g] = planck_band(M, skin_t, M.band_of_g_lw[g]) * pfrac[(size_t)(nlev - 1) * NG_LW + g];
  // SW
  for (int b = 0; b < NB_SW; ++b) {
    const BandMeta& B = M.sw[b];
    int lsol = sw_solar_layer(M, b, nlev, laytrop_sw, [&](int il) { return SL[nlev - il].jp; });
    for (int ig = 0; ig < B.ng; ++ig) incsol[B.g0 + ig] = 0.0;
    for (int jl = 0; jl < nlev; ++jl) {
      const int il = nlev - jl;
      ListOut out{tt, 0, 1};
      SwAux aux;
      sw_build_list(M, M.sw[b], SL[jl], b, il <= laytrop_sw, out, aux);
      if (out.n > *kmax_sw) *kmax_sw = out.n;
      for (int ig = 0; ig < B.ng; ++ig) {
        double taug = 0.0;
        for (int k = 0; k < out.n; ++k) taug += tt[k].c * P->swtab[tt[k].o + ig];
        double taur = aux.rc0 * P->swtab[aux.ro0 + ig] + aux.rc1 * P->swtab[aux.ro1 + ig];
        double od = taur + taug;
        od_sw[(size_t)jl * NG_SW + B.g0 + ig] = od;
        ssa_sw[(size_t)jl * NG_SW + B.g0 + ig] = taur / od;
        if (il == lsol) incsol[B.g0 + ig] = aux.sc0 * P->swtab[aux.so0 + ig] + aux.sc1 * P->swtab[aux.so1 + ig];
      }
    }
  }
  return 0;
}

// McICA generator check: returns od_scaling[nlev][ng] and total cloud cover for one column, using RngMix and the walk
// exactly as the CUDA generator kernel does (see kernels.cu cloud_generator_kernel).
#include "../ecrad_b200/csrc/cloudgen_walk.h"
int hc_cloud_generator(void* p, int ng, int nlev, int scheme, int32_t iseed, double frac_threshold, const double* frac,
                       const double* overlap_param, double decorr_scaling, const double* fsd, int use_beta,
                       double* od_scaling, double* tcc_out) {
  auto* P = (PackedTables*)p;
  std::vector<double> cum(nlev), pair(nlev), opi(nlev);
  std::vector<uint32_t> code((size_t)ng * nlev, 0u);
  std::vector<int32_t> ix(JPQ + 1), rtop(ng), rcloud(nlev), ri1(nlev);
  GenColumn gc;
  gc.nlev = nlev; gc.stride = 1; gc.fstride = 1; gc.frac = frac; gc.cum = cum.data(); gc.pair = pair.data(); gc.opi = opi.data();
  double tcc = gen_prepare(scheme, nlev, 1, 1, frac, overlap_param, use_beta != 0, decorr_scaling, frac_threshold, cum.data(),
                           pair.data(), opi.data(), &gc.ibegin, &gc.iend);
  *tcc_out = tcc;
  for (size_t i = 0; i < (size_t)ng * nlev; ++i) od_scaling[i] = 0.0;
  if (tcc > 0.0) {
    RngMix rs; rs.ix = ix.data();
    gen_walk(gc, rs, iseed, ng, tcc, rtop.data(), rcloud.data(), ri1.data(), code.data(), nlev, scheme == 2);
    for (int g = 0; g < ng; ++g)
      for (int l = 0; l < nlev; ++l) {
        uint32_t cd = code[(size_t)g * nlev + l];
        if (cd) od_scaling[(size_t)l * ng + g] = pdf_sample(P->cloud, P->pdf_val.data(), fsd[l], (double)(cd & 0x3FFFFFFFu) * (1.0 / 1073741824.0));
      }
  }
  return 0;
}

// SPARTACUS matrix routines of sp_core.h (what the CUDA kernels inline), one matrix in C row-major order
void hc_expm(int m, double* a, int sw_pattern) {
  double W[5 * 81];
  if (m == 9 && sw_pattern) sp_expm<9, true>(a, W);
  else if (m == 9) sp_expm<9, false>(a, W);
  else if (m == 6) sp_expm<6, false>(a, W);
}
void hc_fast_expm_exchange_3(double a, double b, double c, double d, double* r) { sp_fast_expm_exchange_3(a, b, c, d, r); }
void hc_m3_solve_mat(const double* A, const double* B, double* X) { m3_solve_mat(A, B, X); }

}  // extern "C"
