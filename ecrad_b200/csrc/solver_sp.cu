// solver_sp.cu -- SPARTACUS solvers: three regions per layer coupled by lateral ("3D") transfer through cloud sides.
//
// Reference: radiation/radiation_spartacus_sw.F90:64-1721, radiation_spartacus_lw.F90:50-1085, radiation_matrix.F90,
// radiation_lw_derivatives.F90:138-193 (calc_lw_derivatives_matrix); region fractions and overlap matrices are those of the
// Tripleclouds path (tc_prep_kernel).  nregions = 3.
//
// Two kernels per spectrum:
//   sp_{sw,lw}_layer_kernel   grid (layer, column), thread = g-point.  Clear-sky Meador-Weaver solution of every layer; in cloudy
//                             layers the g-points whose gas optical depth allows 3D transfer (g < ng3D) exponentiate the 9x9 (SW:
//                             diffuse up/down + direct) or 6x6 (LW) matrix Gamma*dz of the layer -- scaling and squaring with a
//                             degree-7 Pade approximant, matrices in thread-local memory, register-blocked products (sp_core.h) --
//                             and derive the 3x3 reflection / transmission matrices and sources from it; the others use
//                             Meador-Weaver per region.  Results: [column][layer][k][g], g fastest (coalesced).
//   sp_{sw,lw}_sweep_kernel   CTA = column, thread = g-point: upward sweep of the albedo (and source) matrices with the overlap /
//                             entrapment rules, downward sweep of the fluxes; g-point sums through the shared-memory tile.
// This is the first correct version of the path; nothing here is tuned yet.
#include "solver_common.cuh"
#include "sp_core.h"
#include "sp_coop.cuh"
#include "tc_shared.cuh"

namespace ecb {

enum { SP_LCH = 4, SP_LCH_LW = 8 };
enum { SP_SW_NE = 63, SP_LW_NE = 36 };   // entries of a g-point's layer matrix kept in shared memory (shortwave pattern / dense 6x6)
enum { SP_SW_CLR = 5, SP_SW_MAT = 45, SP_SW_ALB = 20, SP_LW_CLR = 4, SP_LW_MAT = 24, SP_LW_ALB = 14 };

__host__ __device__ inline size_t sp_doubles_sw(int nlev, int ng) { return (size_t)ng * ((size_t)(SP_SW_CLR + SP_SW_MAT) * nlev + (size_t)SP_SW_ALB * (nlev + 1)); }
__host__ __device__ inline size_t sp_doubles_lw(int nlev, int ng) { return (size_t)ng * ((size_t)(SP_LW_CLR + SP_LW_MAT) * nlev + (size_t)SP_LW_ALB * (nlev + 1)); }
size_t sp_scratch_doubles_sw(int nlev, int ng) { return sp_doubles_sw(nlev, ng); }
size_t sp_scratch_doubles_lw(int nlev, int ng) { return sp_doubles_lw(nlev, ng); }

// correctly rounded reciprocal (one instruction sequence without the slow path the fp64 division takes for zero numerators)
__device__ __forceinline__ double sp_inv(double x) { return __drcp_rn(x); }

struct SpGeom { double tan_sza, one_over_mu0; };
__device__ __forceinline__ SpGeom sp_geometry(const SpCfg& sc, double mu0) {
  SpGeom q;
  const double min_mu0_3d = 0.004625;
  q.one_over_mu0 = 1.0 / mu0;
  if (mu0 < min_mu0_3d) q.tan_sza = sqrt(1.0 / (min_mu0_3d * min_mu0_3d) - 1.0);
  else if (q.one_over_mu0 > 1.0) q.tan_sza = sqrt(q.one_over_mu0 * q.one_over_mu0 - 1.0 + sc.overhead_sun_factor);
  else q.tan_sza = sqrt(sc.overhead_sun_factor);
  return q;
}

// Position (in the sequence the reference scans, `rank`) of the first g-point whose clear-region optical depth exceeds
// max_gas_od_3d; ng if none (radiation_spartacus_sw.F90:462-478, :486-493).  Block-wide: contains barriers.
__device__ __forceinline__ int sp_first_thick_g(int* s_first, bool act, int rank, int ng, bool thick) {
  if (threadIdx.x == 0) *s_first = ng;
  __syncthreads();
  if (act && thick) atomicMin(s_first, rank);
  __syncthreads();
  return *s_first;
}

// =========================================================================================================
// SW: layer properties (radiation_spartacus_sw.F90:420-835)
// =========================================================================================================
// mode 0: cloud-free layers (clear-sky arrays only; no shared memory needed); mode 1: the layers that need the region matrices
// (cloudy ones, or all with use_expm_everywhere).  Two launches of the same grid, so that the many cheap clear-layer CTAs are not
// throttled by the shared memory the matrix exponentials of the few cloudy ones reserve.
template <class SD, int MINB>
__global__ void __launch_bounds__(SD::THREADS, MINB)
sp_sw_layer_kernel(DevTables T, DevCfg cfg, DevIn in, Work w, int nlev, int mode) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_first;
  const int l = blockIdx.x, c = blockIdx.y, g = threadIdx.x;
  const SpCfg& sc = cfg.sp;
  const double frac = LD_IN(in.frac, c, l);
  const bool cloudy = frac > 0.0;
  const bool heavy = cloudy || sc.use_expm_everywhere;
  if (heavy != (mode == 1)) return;
  const double mu0 = in.cos_sza[c];
  if (mu0 < 1.0e-10) return;
  const bool act = g < SD::NG;
  const int gg = act ? g : 0;
  const size_t n = (size_t)nlev * SD::NG;
  const size_t i = (size_t)l * SD::NG + gg;
  const double odg = w.od_sw[(size_t)c * n + i], ssag = w.ssa_sw[(size_t)c * n + i];
  const double gas_g = (cfg.use_aerosols && w.g_sw) ? w.g_sw[(size_t)c * n + i] : 0.0;
  double* clr = w.scr_sw + (size_t)c * sp_doubles_sw(nlev, SD::NG);
  double* mats = clr + (size_t)SP_SW_CLR * n + (size_t)l * SP_SW_MAT * SD::NG;
  // clear-sky arrays, all g-points (:772-781)
  const SwLayer Lc = sw_ref_trans_cloudless(mu0, odg, ssag, gas_g);
  if (act) {
    clr[0 * n + i] = Lc.ref; clr[1 * n + i] = Lc.trans; clr[2 * n + i] = Lc.ref_dir; clr[3 * n + i] = Lc.trans_dir_diff; clr[4 * n + i] = Lc.trans_dir_dir;
  }
  // clear-sky layer: the (1,1) elements are the clear-sky values, unless use_expm_everywhere asks for the matrix exponential there too
  if (!heavy) return;
  const int nra = cloudy ? 3 : 1;   // nregactive
  // ---- cloudy layer (or a clear one treated with the matrix exponential) ----
  const double* reg = w.tc_reg + ((size_t)c * nlev + l) * 3;
  const double* ods = w.tc_ods + ((size_t)c * nlev + l) * 3;
  double edge[3] = {0.0, 0.0, 0.0}, rate_dir[9], rate_dif[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { rate_dir[k] = 0.0; rate_dif[k] = 0.0; }
  // (two regions: no lateral transfer in an overcast layer, radiation_spartacus_sw.F90:497-500)
  const bool overcast2 = sc.two_regions && frac > 1.0 - cfg.cloud_fraction_threshold;
  const bool has3d = cloudy && !overcast2 && in.inv_cloud_size && sp_edge_lengths(sc, reg, LD_IN(in.inv_cloud_size, c, l), in.inv_inhom_size != nullptr, in.inv_inhom_size ? LD_IN(in.inv_inhom_size, c, l) : 0.0, edge);
  if (has3d) {
    const SpGeom q = sp_geometry(sc, mu0);
    const double dz = sp_layer_depth(LD_IN(in.p_hl, c, l), LD_IN(in.p_hl, c, l + 1), LD_IN(in.t_hl, c, l), LD_IN(in.t_hl, c, l + 1));
    sp_transfer_rates(sc, dz, edge, reg, q.tan_sza, rate_dir);
    sp_transfer_rates(sc, dz, edge, reg, SP_PI * 0.5, rate_dif);
  }
  const int rank = T.meta->rank_sw[gg];
  const int ng3d = (has3d || sc.use_expm_everywhere) ? sp_first_thick_g(&s_first, act, rank, SD::NG, odg > sc.max_gas_od_3d) : 0;
  // optical properties of the regions (:606-655)
  const int b = T.meta->band_of_g_sw[gg];
  const double* clb = w.cl_sw + ((size_t)c * nlev + l) * 3 * SD::NB;
  double od_r[3], ssa_r[3], g_r[3];
  od_r[0] = odg; ssa_r[0] = ssag; g_r[0] = gas_g;
  const double scat_od = odg * ssag;
#pragma unroll
  for (int jr = 1; jr < 3; ++jr) {
    if (!cloudy) { od_r[jr] = 0.0; ssa_r[jr] = 0.0; g_r[jr] = 0.0; continue; }
    const double scat_od_cloud = clb[b] * clb[SD::NB + b] * ods[jr];
    od_r[jr] = odg + clb[b] * ods[jr];
    ssa_r[jr] = (scat_od + scat_od_cloud) / od_r[jr];
    g_r[jr] = (scat_od * gas_g + scat_od_cloud * clb[2 * SD::NB + b]) / (scat_od + scat_od_cloud);
    if (od_r[jr] > sc.max_cloud_od) od_r[jr] = sc.max_cloud_od;
  }
  const bool need = act && rank < ng3d;   // this g-point's layer matrices come from the matrix exponential
  if (act && !need) {
    // Meador-Weaver per region: diagonal matrices (:783-832)
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int jr = k / 4;
      SwLayer L = {0.0, 0.0, 0.0, 0.0, 0.0};
      if (k % 4 == 0 && jr < nra) L = jr == 0 ? Lc : sw_ref_trans_cloudless(mu0, od_r[jr], ssa_r[jr], g_r[jr]);
      mats[(size_t)(0 + k) * SD::NG + g] = L.ref;
      mats[(size_t)(9 + k) * SD::NG + g] = L.trans;
      mats[(size_t)(18 + k) * SD::NG + g] = L.ref_dir;
      mats[(size_t)(27 + k) * SD::NG + g] = L.trans_dir_diff;
      mats[(size_t)(36 + k) * SD::NG + g] = L.trans_dir_dir;
    }
  }
  const unsigned need_mask = __ballot_sync(0xffffffffu, need);
  if (!need_mask) return;   // (warp-uniform)
  // ---- 9x9 matrix exponential (:658-770), warp-cooperative (sp_coop.cuh) ----
  const int warp = threadIdx.x >> 5;
  double* stage = reinterpret_cast<double*>(smem_raw) + (size_t)warp * Coop<9>::PER_WARP;
  double Gl[SP_SW_NE];   // this g-point's matrix: the 63 entries of the shortwave pattern
#define G_(r, cc) Gl[coop_entry<9, true>(r, cc)]
  if (need) {
    const double one_over_mu0 = 1.0 / mu0;
#pragma unroll 1   // (a run-time index keeps Gl in thread-local memory: it only waits there for its round)
    for (int k = 0; k < SP_SW_NE; ++k) Gl[k] = 0.0;
#pragma unroll
    for (int jr = 0; jr < 3; ++jr) {
      if (jr >= nra) continue;
      const double factor = 0.75 * g_r[jr];   // calc_two_stream_gammas_sw
      const double gamma1 = 2.0 - ssa_r[jr] * (1.25 + factor), gamma2 = ssa_r[jr] * (0.75 - factor), gamma3 = 0.5 - mu0 * factor;
      G_(jr, jr) = od_r[jr] * gamma1;
      G_(jr + 3, jr) = od_r[jr] * gamma2;
      G_(jr, jr + 6) = -od_r[jr] * ssa_r[jr] * gamma3;
      G_(jr + 3, jr + 6) = od_r[jr] * ssa_r[jr] * (1.0 - gamma3);
      G_(jr + 6, jr + 6) = -od_r[jr] * one_over_mu0;
    }
#pragma unroll
    for (int jr = 0; jr < 2; ++jr) {
      if (jr + 1 >= nra) continue;
      G_(jr, jr) = G_(jr, jr) + rate_dif[jr * 3 + jr + 1];
      G_(jr + 1, jr + 1) = G_(jr + 1, jr + 1) + rate_dif[(jr + 1) * 3 + jr];
      G_(jr + 1, jr) = -rate_dif[jr * 3 + jr + 1];
      G_(jr, jr + 1) = -rate_dif[(jr + 1) * 3 + jr];
      G_(jr + 6, jr + 6) = G_(jr + 6, jr + 6) - rate_dir[jr * 3 + jr + 1];
      G_(jr + 7, jr + 7) = G_(jr + 7, jr + 7) - rate_dir[(jr + 1) * 3 + jr];
      G_(jr + 7, jr + 6) = rate_dir[jr * 3 + jr + 1];
      G_(jr + 6, jr + 7) = rate_dir[(jr + 1) * 3 + jr];
    }
    if (edge[2] > 0.0) {
      G_(0, 0) = G_(0, 0) + rate_dif[2];
      G_(2, 2) = G_(2, 2) + rate_dif[6];
      G_(2, 0) = -rate_dif[2];
      G_(0, 2) = -rate_dif[6];
      G_(6, 6) = G_(6, 6) - rate_dir[2];
      G_(8, 8) = G_(8, 8) - rate_dir[6];
      G_(8, 6) = rate_dir[2];
      G_(6, 8) = rate_dir[6];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int bb = 0; bb < 3; ++bb) if (a < nra && bb < nra) G_(3 + a, 3 + bb) = -G_(a, bb);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int bb = 0; bb < 3; ++bb) if (a < nra && bb < nra) G_(a, 3 + bb) = -G_(3 + a, bb);
  }
  coop_expm_warp<9, true>(Gl, need_mask, stage);
  if (!need) return;
  double E11[9], E21[9], X[9], R[9];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int bb = 0; bb < 3; ++bb) { E11[a * 3 + bb] = G_(a, bb); E21[a * 3 + bb] = G_(3 + a, bb); }
  // direct transmission
#pragma unroll
  for (int k = 0; k < 9; ++k) mats[(size_t)(36 + k) * SD::NG + g] = dmin(1.0, dmax(0.0, G_(6 + k / 3, 6 + k % 3)));
  // diffuse reflectance and transmittance
#pragma unroll
  for (int k = 0; k < 9; ++k) X[k] = G_(k / 3, 3 + k % 3);
  m3_solve_mat(E11, X, X);
#pragma unroll
  for (int k = 0; k < 9; ++k) { R[k] = dmin(1.0, dmax(0.0, -X[k])); mats[(size_t)k * SD::NG + g] = R[k]; }
  m3_x_m3(E21, R, X);
#pragma unroll
  for (int k = 0; k < 9; ++k) mats[(size_t)(9 + k) * SD::NG + g] = dmin(1.0, dmax(0.0, X[k] + G_(3 + k / 3, 3 + k % 3)));
  // direct -> diffuse up / down
#pragma unroll
  for (int k = 0; k < 9; ++k) X[k] = G_(k / 3, 6 + k % 3);
  m3_solve_mat(E11, X, X);
#pragma unroll
  for (int k = 0; k < 9; ++k) { R[k] = dmin(mu0, dmax(0.0, -X[k])); mats[(size_t)(18 + k) * SD::NG + g] = R[k]; }
  m3_x_m3(E21, R, X);
#pragma unroll
  for (int k = 0; k < 9; ++k) mats[(size_t)(27 + k) * SD::NG + g] = dmin(mu0, dmax(0.0, X[k] + G_(3 + k / 3, 6 + k % 3)));
#undef G_
}

// per-column shared data of the sweeps on top of the Tripleclouds region data: layer depths and cloud edge lengths
struct SpGeomShared { double* edge; double* depth; int i_cloud_top; };
static size_t sp_shared_bytes(int nlev) { return tc_shared_bytes(nlev) + sizeof(double) * 4 * nlev + 16; }
// block-wide, ends with a barrier; S: result of tc_load_shared (already complete)
__device__ __forceinline__ SpGeomShared sp_load_geometry(double* base, const TcShared& S, const DevCfg& cfg, const DevIn& in, int c, int nlev, int nthreads) {
  SpGeomShared P;
  P.edge = base;
  P.depth = base + 3 * nlev;
  for (int l = threadIdx.x; l < nlev; l += nthreads) {
    P.depth[l] = sp_layer_depth(LD_IN(in.p_hl, c, l), LD_IN(in.p_hl, c, l + 1), LD_IN(in.t_hl, c, l), LD_IN(in.t_hl, c, l + 1));
    double e[3] = {0.0, 0.0, 0.0};
    if (!S.clear[l + 1] && in.inv_cloud_size && !(cfg.sp.two_regions && LD_IN(in.frac, c, l) > 1.0 - cfg.cloud_fraction_threshold)) {
      const double r[3] = {S.reg[l * 3], S.reg[l * 3 + 1], S.reg[l * 3 + 2]};
      sp_edge_lengths(cfg.sp, r, LD_IN(in.inv_cloud_size, c, l), in.inv_inhom_size != nullptr, in.inv_inhom_size ? LD_IN(in.inv_inhom_size, c, l) : 0.0, e);
    }
    P.edge[l * 3] = e[0]; P.edge[l * 3 + 1] = e[1]; P.edge[l * 3 + 2] = e[2];
  }
  __syncthreads();
  int ict = nlev + 1;
  for (int jl = nlev; jl >= 1; --jl) if (!S.clear[jl]) ict = jl;
  P.i_cloud_top = ict;
  return P;
}
// the 3x3 overlap matrix of one half-level from global memory (uniform across the CTA)
__device__ __forceinline__ void sp_load_m3(const double* __restrict__ p, double* m) {
#pragma unroll
  for (int k = 0; k < 9; ++k) m[k] = __ldg(p + k);
}

// =========================================================================================================
// SW: albedo matrices upward, fluxes downward (radiation_spartacus_sw.F90:837-1590)
// =========================================================================================================
template <class SD, int MINB>
__global__ void __launch_bounds__(SD::THREADS, MINB)
sp_sw_sweep_kernel(DevTables T, DevCfg cfg, DevIn in, DevOut out, Work w, int nlev) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int c = blockIdx.x, g = threadIdx.x, nl1 = nlev + 1;
  const bool act = g < SD::NG;
  const int gg = act ? g : 0;
  const double mu0 = in.cos_sza[c];
  const SpCfg& sc = cfg.sp;
  if (g == 0 && out.cloud_cover_sw) out.cloud_cover_sw[c] = w.tc_cc[c];
  if (mu0 < 1.0e-10) { sw_night_column<SD>(cfg, out, c, g, act, nl1, SD::THREADS); return; }
  double* sums = reinterpret_cast<double*>(smem_raw);     // [6][nl1]: up, dn_dif, dn_dir, up_c, dn_dif_c, dn_dir_c
  double* tile = sums + 6 * nl1;                           // [6][SP_LCH][SD::RS]
  double* bandv = tile + 6 * SP_LCH * SD::RS;            // [2][NB]
  double* geom = bandv + 2 * SD::NB;                     // [4][nlev]
  const TcShared S = tc_load_shared(reinterpret_cast<unsigned char*>(geom + 4 * nlev), w, in, c, nlev, SD::THREADS);
  const SpGeomShared P = sp_load_geometry(geom, S, cfg, in, c, nlev, SD::THREADS);
  const double* Ug = w.tc_u + (size_t)c * (nlev + 1) * 9;
  const double* Vg = w.tc_v + (size_t)c * (nlev + 1) * 9;
  if (g < SD::NB) {   // get_albedos, radiation_single_level.F90:216-365
    double bd = 0.0, bdir = 0.0;
    for (int ja = 0; ja < cfg.n_albedo_sw; ++ja) {
      const double wgt = T.sw_albedo_weights[g * cfg.n_albedo_sw + ja];
      if (wgt != 0.0) { bd = bd + wgt * LD_IN(in.sw_albedo, c, ja); if (in.sw_albedo_direct) bdir = bdir + wgt * LD_IN(in.sw_albedo_direct, c, ja); }
    }
    bandv[g] = bd; bandv[SD::NB + g] = in.sw_albedo_direct ? bdir : bd;
  }
  __syncthreads();
  const size_t n = (size_t)nlev * SD::NG;
  const double* clr = w.scr_sw + (size_t)c * sp_doubles_sw(nlev, SD::NG);
  const double* mats = clr + (size_t)SP_SW_CLR * n;
  double* alb = w.scr_sw + (size_t)c * sp_doubles_sw(nlev, SD::NG) + (size_t)(SP_SW_CLR + SP_SW_MAT) * n;
#define MAT(l, k) mats[((size_t)(l) * SP_SW_MAT + (k)) * SD::NG + g]
#define ALB(hl, k) alb[((size_t)(hl) * SP_SW_ALB + (k)) * SD::NG + g]
  const int b = T.meta->band_of_g_sw[gg];
  const double alb_diff = bandv[b], alb_dir = bandv[SD::NB + b];
  const double inc = w.incoming[(size_t)c * SD::NG + gg];
  const SpGeom q = sp_geometry(sc, mu0);
  const double tan_diffuse_angle_3d = SP_PI * 0.5;
  const bool explicit_entr = sc.entrapment == SP_ENTR_EXPLICIT || sc.entrapment == SP_ENTR_NON_FRACTAL;

  // ---- upward sweep ----
  double TA[9], TAD[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { TA[k] = 0.0; TAD[k] = 0.0; }
  TA[0] = TA[4] = TA[8] = alb_diff;
  TAD[0] = TAD[4] = TAD[8] = mu0 * alb_dir;
  double tac = TA[0], tdc = TAD[0];
  if (act) {
    double xdif[3] = {0.0, 0.0, 0.0}, xdir[3] = {0.0, 0.0, 0.0};
    for (int l = nlev - 1; l >= 0; --l) {
      const int jl = l + 1;
      const size_t i = (size_t)l * SD::NG + g;
      const bool clear_l = S.clear[jl];
      // what the downward sweep needs of the albedos below this layer
      if (clear_l) { ALB(jl, 0) = TA[0]; ALB(jl, 9) = TAD[0]; }
      else {
#pragma unroll
        for (int k = 0; k < 9; ++k) { ALB(jl, k) = TA[k]; ALB(jl, 9 + k) = TAD[k]; }
      }
      ALB(jl, 18) = tac; ALB(jl, 19) = tdc;
      const double rc = clr[i], trc = clr[n + i], rdc = clr[2 * n + i], tddc = clr[3 * n + i], tdrc = clr[4 * n + i];
      {   // clear-sky column (:879-893)
        const double inv_denom = sp_inv(1.0 - tac * rc);
        const double tac_new = rc + trc * trc * tac * inv_denom;
        tdc = rdc + (tdrc * tdc + tddc * tac) * trc * inv_denom;
        tac = tac_new;
      }
      double below[9], belowd[9], R[9], Tm[9], RD[9], TDD[9], TD[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) { below[k] = 0.0; belowd[k] = 0.0; }
      if (clear_l) {
#pragma unroll
        for (int k = 0; k < 9; ++k) { R[k] = 0.0; Tm[k] = 0.0; RD[k] = 0.0; TDD[k] = 0.0; TD[k] = 0.0; }
        if (sc.use_expm_everywhere) { R[0] = MAT(l, 0); Tm[0] = MAT(l, 9); RD[0] = MAT(l, 18); TDD[0] = MAT(l, 27); TD[0] = MAT(l, 36); }
        else { R[0] = rc; Tm[0] = trc; RD[0] = rdc; TDD[0] = tddc; TD[0] = tdrc; }
        const double inv_denom = sp_inv(1.0 - TA[0] * R[0]);
        below[0] = R[0] + Tm[0] * Tm[0] * TA[0] * inv_denom;
        belowd[0] = RD[0] + (TD[0] * TAD[0] + TDD[0] * TA[0]) * Tm[0] * inv_denom;
      } else {
#pragma unroll
        for (int k = 0; k < 9; ++k) { R[k] = MAT(l, k); Tm[k] = MAT(l, 9 + k); RD[k] = MAT(l, 18 + k); TDD[k] = MAT(l, 27 + k); TD[k] = MAT(l, 36 + k); }
        double den[9], X[9], Y[9];
        m3_identity_minus_product(TA, R, den);
        m3_x_m3(TA, Tm, X);
        m3_solve_mat(den, X, X);
        m3_x_m3(Tm, X, X);
#pragma unroll
        for (int k = 0; k < 9; ++k) below[k] = R[k] + X[k];
        m3_x_m3(TAD, TD, X);
        m3_x_m3(TA, TDD, Y);
#pragma unroll
        for (int k = 0; k < 9; ++k) X[k] = X[k] + Y[k];
        m3_solve_mat(den, X, X);
        m3_x_m3(Tm, X, X);
#pragma unroll
        for (int k = 0; k < 9; ++k) belowd[k] = RD[k] + X[k];
      }
      if (explicit_entr && jl >= P.i_cloud_top)
        sp_step_migrations(LD_IN(in.frac, c, l), P.depth[l], tan_diffuse_angle_3d, q.tan_sza, R, Tm, RD, TD, TDD, TA, TAD, xdif, xdir);
      double U[9], V[9];
      sp_load_m3(Ug + l * 9, U);   // u_matrix(:,:,jlev)
      sp_load_m3(Vg + l * 9, V);   // v_matrix(:,:,jlev)
      const bool clear_above = S.clear[jl - 1];
      if (clear_l && clear_above) {
#pragma unroll
        for (int k = 0; k < 9; ++k) { TA[k] = 0.0; TAD[k] = 0.0; }
        TA[0] = below[0]; TAD[0] = belowd[0];
      } else if (sc.entrapment == SP_ENTR_MAXIMUM || clear_above) {
        m3_u_a_v(U, below, V, TA);
        m3_u_a_v(U, belowd, V, TAD);
      } else if (sc.entrapment == SP_ENTR_ZERO) {
#pragma unroll
        for (int k = 0; k < 9; ++k) { TA[k] = 0.0; TAD[k] = 0.0; }
#pragma unroll
        for (int jr = 0; jr < 3; ++jr)
#pragma unroll
          for (int jr2 = 0; jr2 < 3; ++jr2) {
            double s = 0.0, sd = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) { s = s + below[k * 3 + jr2]; sd = sd + belowd[k * 3 + jr2]; }
            TA[jr * 4] = TA[jr * 4] + s * V[jr2 * 3 + jr];
            TAD[jr * 4] = TAD[jr * 4] + sd * V[jr2 * 3 + jr];
          }
      } else {
        double part[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) part[k] = (k % 4 == 0) ? 0.0 : below[k];
        m3_u_a_v(U, part, V, TA);
#pragma unroll
        for (int k = 0; k < 9; ++k) part[k] = (k % 4 == 0) ? 0.0 : belowd[k];
        m3_u_a_v(U, part, V, TAD);
        if (sc.entrapment == SP_ENTR_EDGE_ONLY || !sc.do_3d_effects) {
#pragma unroll
          for (int jr = 0; jr < 3; ++jr)
#pragma unroll
            for (int jr2 = 0; jr2 < 3; ++jr2) {
              TA[jr * 4] = TA[jr * 4] + below[jr2 * 4] * V[jr2 * 3 + jr];
              TAD[jr * 4] = TAD[jr * 4] + belowd[jr2 * 4] * V[jr2 * 3 + jr];
            }
        } else {   // explicit entrapment (:1080-1265); never reached for jlev = 1 (the pseudo-layer above is clear)
          const double inv_effective_size = dmin(LD_IN(in.inv_cloud_size, c, l - 1), 1.0 / sc.min_cloud_effective_size);
          const double op = LD_IN(in.overlap, c, l - 1);
#pragma unroll 1
          for (int jr2 = 0; jr2 < 3; ++jr2) {
            double rate[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) rate[k] = 0.0;
            const double transfer_scaling = 1.0 - (1.0 - sc.overhang_factor) * op * dmin(S.reg[l * 3 + jr2], S.reg[(l - 1) * 3 + jr2]) /
                                                      dmax(cfg.cloud_fraction_threshold, S.reg[l * 3 + jr2]);
#pragma unroll
            for (int jr = 0; jr < 2; ++jr) {
              rate[jr * 3 + jr + 1] = transfer_scaling * P.edge[(l - 1) * 3 + jr] * sp_inv(dmax(U[jr * 3 + jr2], 1.0e-5));
              rate[(jr + 1) * 3 + jr] = transfer_scaling * P.edge[(l - 1) * 3 + jr] * sp_inv(dmax(U[(jr + 1) * 3 + jr2], 1.0e-5));
            }
            sp_entrapment_part(sc, rate, xdif[jr2], inv_effective_size, part);
#pragma unroll
            for (int jr3 = 0; jr3 < 3; ++jr3)
#pragma unroll
              for (int jr = 0; jr < 3; ++jr) TA[jr3 * 3 + jr] = TA[jr3 * 3 + jr] + part[jr3 * 3 + jr] * V[jr2 * 3 + jr] * below[jr2 * 4];
            sp_entrapment_part(sc, rate, xdir[jr2], inv_effective_size, part);
#pragma unroll
            for (int jr3 = 0; jr3 < 3; ++jr3)
#pragma unroll
              for (int jr = 0; jr < 3; ++jr) TAD[jr3 * 3 + jr] = TAD[jr3 * 3 + jr] + part[jr3 * 3 + jr] * V[jr2 * 3 + jr] * belowd[jr2 * 4];
          }
        }
      }
      if (explicit_entr && !(clear_l && clear_above)) {
        double xd[3] = {0.0, 0.0, 0.0}, xf[3] = {0.0, 0.0, 0.0};
        const int nra = clear_l ? 1 : 3;
#pragma unroll
        for (int jr = 0; jr < 3; ++jr)
#pragma unroll
          for (int jr2 = 0; jr2 < 3; ++jr2)
            if (jr2 < nra) { xd[jr] = xd[jr] + xdir[jr2] * V[jr2 * 3 + jr]; xf[jr] = xf[jr] + xdif[jr2] * V[jr2 * 3 + jr]; }
#pragma unroll
        for (int r = 0; r < 3; ++r) { xdir[r] = xd[r]; xdif[r] = xf[r]; }
      }
    }
  }
  // ---- downward sweep (:1324-1590) ----
  double ddn[3], fdn[3] = {0.0, 0.0, 0.0}, fup[3];
#pragma unroll
  for (int jr = 0; jr < 3; ++jr) ddn[jr] = inc * S.reg[jr];
  m3_x_vec(TAD, ddn, fup);
  double ddc = inc, fdc = 0.0, fuc = ddc * tdc;
  const double toa_a = fup[0] + fup[1] + fup[2], toa_c = fuc;
  double* dst[6] = {sums, sums + nl1, sums + 2 * nl1, sums + 3 * nl1, sums + 4 * nl1, sums + 5 * nl1};
  const bool bands = cfg.do_save_spectral_flux && (out.sw_up_band || out.sw_dn_band || out.sw_dn_direct_band);
  const BandOut bo[3] = {{out.sw_up_band, out.ld, 0, -1, 1.0, 0.0, nullptr, 0}, {out.sw_dn_direct_band, out.ld, 2, -1, mu0, 0.0, nullptr, 0},
                         {out.sw_dn_band, out.ld, 2, 1, mu0, 1.0, nullptr, 0}};
  int slot = 0, lfirst = 0;
#define PUT_ROWS()                                                                                        \
  if (act) {                                                                                              \
    tile[(0 * SP_LCH + slot) * SD::RS + g] = fup[0] + fup[1] + fup[2];                                  \
    tile[(1 * SP_LCH + slot) * SD::RS + g] = fdn[0] + fdn[1] + fdn[2];                                  \
    tile[(2 * SP_LCH + slot) * SD::RS + g] = ddn[0] + ddn[1] + ddn[2];                                  \
    tile[(3 * SP_LCH + slot) * SD::RS + g] = fuc;                                                       \
    tile[(4 * SP_LCH + slot) * SD::RS + g] = fdc;                                                       \
    tile[(5 * SP_LCH + slot) * SD::RS + g] = ddc;                                                       \
  }                                                                                                       \
  ++slot;
  PUT_ROWS();
  double dif_a = 0.0, dir_a = 0.0;
  for (int l = 0; l < nlev; ++l) {
    const int jl = l + 1;
    if (act) {
      const size_t i = (size_t)l * SD::NG + g;
      const bool clear_l = S.clear[jl];
      const double rc = clr[i], trc = clr[n + i], tddc = clr[3 * n + i], tdrc = clr[4 * n + i];
      const double tacb = ALB(jl, 18), tdcb = ALB(jl, 19);
      {
        const double source_dn_clear = tddc * ddc;
        ddc = tdrc * ddc;
        fdc = (trc * fdc + rc * tdcb * ddc + source_dn_clear) * sp_inv(1.0 - rc * tacb);   // (the reciprocal does not wait for the flux recurrence)
        fuc = tdcb * ddc + tacb * fdc;
      }
      if (clear_l) {
        const double ta0 = ALB(jl, 0), tad0 = ALB(jl, 9);
        const bool ev = sc.use_expm_everywhere != 0;   // all-sky (1,1) elements from the matrix exponential also in clear layers
        const double r0 = ev ? MAT(l, 0) : rc, t0 = ev ? MAT(l, 9) : trc, tdd0 = ev ? MAT(l, 27) : tddc, td0 = ev ? MAT(l, 36) : tdrc;
        const double source_dn = tdd0 * ddn[0];
        const double dabove = td0 * ddn[0];
        fdn[0] = (t0 * fdn[0] + r0 * tad0 * dabove + source_dn) * sp_inv(1.0 - r0 * ta0);
        fup[0] = tad0 * dabove + ta0 * fdn[0];
        ddn[0] = dabove;
        fdn[1] = 0.0; fdn[2] = 0.0; fup[1] = 0.0; fup[2] = 0.0; ddn[1] = 0.0; ddn[2] = 0.0;
      } else {
        double R[9], Tm[9], M[9], TAb[9], TADb[9], src[3], tot[3], a[3], bb[3];
#pragma unroll
        for (int k = 0; k < 9; ++k) { R[k] = MAT(l, k); Tm[k] = MAT(l, 9 + k); TAb[k] = ALB(jl, k); TADb[k] = ALB(jl, 9 + k); }
#pragma unroll
        for (int k = 0; k < 9; ++k) M[k] = MAT(l, 27 + k);
        m3_x_vec(M, ddn, src);            // source_dn = trans_dir_diff * direct_dn_below
#pragma unroll
        for (int k = 0; k < 9; ++k) M[k] = MAT(l, 36 + k);
        m3_x_vec(M, ddn, ddn);            // direct_dn_above = trans_dir_dir * direct_dn_below
        m3_identity_minus_product(R, TAb, M);
        m3_x_vec(TADb, ddn, tot);         // total_source
        m3_x_vec(Tm, fdn, a);
        m3_x_vec(R, tot, bb);
#pragma unroll
        for (int r = 0; r < 3; ++r) a[r] = a[r] + bb[r] + src[r];
        m3_solve_vec(M, a, fdn);
        m3_x_vec(TAb, fdn, fup);
#pragma unroll
        for (int r = 0; r < 3; ++r) fup[r] = fup[r] + tot[r];
      }
    }
    PUT_ROWS();
    if (l == nlev - 1) { dif_a = fdn[0] + fdn[1] + fdn[2]; dir_a = mu0 * (ddn[0] + ddn[1] + ddn[2]); }
    if (act && l < nlev - 1 && !(S.clear[jl] && S.clear[jl + 1])) {
      double V[9];
      sp_load_m3(Vg + jl * 9, V);   // v_matrix(:,:,jlev+1)
      m3_x_vec(V, fdn, fdn);
      m3_x_vec(V, ddn, ddn);
    }
    if (slot == SP_LCH || l == nlev - 1) {
      if (bands) flush_bands(tile, SD::RS, SP_LCH, slot, bo, 3, lfirst, 1, c, SD::NB, T.meta->sw);
      flush_tile(tile, SD::RS, SD::NG, 6, slot, dst, lfirst, 1, SP_LCH); lfirst += slot; slot = 0;
    }
  }
#undef PUT_ROWS
#undef MAT
#undef ALB
  for (int l = g; l < nl1; l += SD::THREADS) {
    const double dir = mu0 * sums[2 * nl1 + l], dirc = mu0 * sums[5 * nl1 + l];
    const size_t o = (size_t)l * out.ld + c;
    if (out.sw_up) out.sw_up[o] = sums[l];
    if (out.sw_dn) out.sw_dn[o] = l == 0 ? dir : dir + sums[nl1 + l];
    if (out.sw_dn_direct) out.sw_dn_direct[o] = dir;
    if (out.sw_up_clear) out.sw_up_clear[o] = sums[3 * nl1 + l];
    if (out.sw_dn_clear) out.sw_dn_clear[o] = l == 0 ? dirc : dirc + sums[4 * nl1 + l];
    if (out.sw_dn_direct_clear) out.sw_dn_direct_clear[o] = dirc;
  }
  const double dif_c = fdc, dir_c = mu0 * ddc;
  if (act) {
    const size_t i = (size_t)c * SD::NG + T.meta->rank_sw[g];   // flux_type's g-point arrays are in the solver's (reordered) sequence
    if (out.sw_dn_diffuse_surf_g) out.sw_dn_diffuse_surf_g[i] = dif_a;
    if (out.sw_dn_direct_surf_g) out.sw_dn_direct_surf_g[i] = dir_a;
    if (out.sw_dn_diffuse_surf_clear_g) out.sw_dn_diffuse_surf_clear_g[i] = dif_c;
    if (out.sw_dn_direct_surf_clear_g) out.sw_dn_direct_surf_clear_g[i] = dir_c;
    if (out.sw_up_toa_g) out.sw_up_toa_g[i] = toa_a;
    if (out.sw_up_toa_clear_g) out.sw_up_toa_clear_g[i] = toa_c;
  }
  sw_surface_spectral<SD>(T, cfg, out, c, g, act, tile, SD::RS, dir_a, dif_a, dir_c, dif_c);
}

// =========================================================================================================
// LW: layer properties (radiation_spartacus_lw.F90:349-780); no LW aerosol scattering: clear-region ssa = g = 0
// =========================================================================================================
template <class SD, int MINB>
__global__ void __launch_bounds__(SD::THREADS, MINB)
sp_lw_layer_kernel(DevTables T, DevCfg cfg, DevIn in, Work w, int nlev, int mode) {   // mode: see sp_sw_layer_kernel
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_first;
  const int l = blockIdx.x, c = blockIdx.y, g = threadIdx.x;
  const SpCfg& sc = cfg.sp;
  const double frac = LD_IN(in.frac, c, l);
  const bool cloudy = frac > 0.0;
  const bool heavy = cloudy || sc.use_expm_everywhere;
  if (heavy != (mode == 1)) return;
  const bool act = g < SD::NG;
  const int gg = act ? g : 0;
  const size_t n = (size_t)nlev * SD::NG;
  const size_t i = (size_t)l * SD::NG + gg;
  const double odg = w.od_lw[(size_t)c * n + i];
  const double pt = w.planck[(size_t)c * (nlev + 1) * SD::NG + i], pb = w.planck[(size_t)c * (nlev + 1) * SD::NG + i + SD::NG];
  double* clr = w.scr_lw + (size_t)c * sp_doubles_lw(nlev, SD::NG);
  double* mats = clr + (size_t)SP_LW_CLR * n + (size_t)l * SP_LW_MAT * SD::NG;
  // gas + aerosol single-scattering albedo and asymmetry factor of the clear region (do_lw_aerosol_scattering,
  // radiation_spartacus_lw.F90:366-371); zero otherwise
  const bool lwscat = cfg.do_lw_aerosol_scattering && w.ssa_lw;
  const double ssa0 = lwscat ? w.ssa_lw[(size_t)c * n + i] : 0.0, g0 = lwscat ? w.g_lw[(size_t)c * n + i] : 0.0;
  const LwLayer Lc = lw_ref_trans(odg, ssa0, g0, pt, pb);
  if (act) { clr[i] = Lc.ref; clr[n + i] = Lc.trans; clr[2 * n + i] = Lc.source_up; clr[3 * n + i] = Lc.source_dn; }
  if (!heavy) return;
  const int nra = cloudy ? 3 : 1;   // nregActive
  const double* reg = w.tc_reg + ((size_t)c * nlev + l) * 3;
  const double* ods = w.tc_ods + ((size_t)c * nlev + l) * 3;
  double edge[3] = {0.0, 0.0, 0.0}, rate[9], dz = 1.0;
#pragma unroll
  for (int k = 0; k < 9; ++k) rate[k] = 0.0;
  const double inv_size = in.inv_cloud_size ? LD_IN(in.inv_cloud_size, c, l) : 0.0;
  const bool overcast2 = sc.two_regions && frac > 1.0 - cfg.cloud_fraction_threshold;   // radiation_spartacus_lw.F90:423-426
  const bool has3d = cloudy && !overcast2 && in.inv_cloud_size && sp_edge_lengths(sc, reg, inv_size, in.inv_inhom_size != nullptr, in.inv_inhom_size ? LD_IN(in.inv_inhom_size, c, l) : 0.0, edge);
  if (has3d) {
    dz = sp_layer_depth(LD_IN(in.p_hl, c, l), LD_IN(in.p_hl, c, l + 1), LD_IN(in.t_hl, c, l), LD_IN(in.t_hl, c, l + 1));
    sp_transfer_rates(sc, dz, edge, reg, SP_PI * 0.5, rate);
  }
  const int rank = T.meta->rank_lw[gg];
  const int ng3d = (has3d || sc.use_expm_everywhere) ? sp_first_thick_g(&s_first, act, rank, SD::NG, odg > sc.max_gas_od_3d) : 0;
  const int b = T.meta->band_of_g_lw[gg];
  const double* clb = w.cl_lw + ((size_t)c * nlev + l) * 3 * SD::NB;
  double od_r[3], ssa_r[3] = {ssa0, 0.0, 0.0}, g_r[3] = {g0, 0.0, 0.0};
  od_r[0] = odg;
  const double scat_od = odg * ssa0;   // scattering optical depth of the clear region (:541)
#pragma unroll
  for (int jr = 1; jr < 3; ++jr) {
    if (!cloudy) { od_r[jr] = 0.0; continue; }
    od_r[jr] = odg + clb[b] * ods[jr];
    if (cfg.do_lw_cloud_scattering) {
      const double scat_od_cloud = clb[b] * clb[SD::NB + b] * ods[jr];
      ssa_r[jr] = (scat_od + scat_od_cloud) / od_r[jr];
      if (scat_od + scat_od_cloud > 0.0) g_r[jr] = (scat_od * g0 + scat_od_cloud * clb[2 * SD::NB + b]) / (scat_od + scat_od_cloud);
    }
    if (od_r[jr] > sc.max_cloud_od) od_r[jr] = sc.max_cloud_od;
  }
  const bool need = act && rank < ng3d;
  if (act && !need) {
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int jr = k / 4;
      LwLayer L = {0.0, 0.0, 0.0, 0.0};
      if (k % 4 == 0) {
        if (jr == 0) { L = Lc; L.source_up = reg[0] * Lc.source_up; L.source_dn = reg[0] * Lc.source_dn; }
        else if (jr < nra) L = lw_ref_trans(od_r[jr], ssa_r[jr], g_r[jr], reg[jr] * pt, reg[jr] * pb);
        mats[(size_t)(18 + jr) * SD::NG + g] = L.source_up;
        mats[(size_t)(21 + jr) * SD::NG + g] = L.source_dn;
      }
      mats[(size_t)k * SD::NG + g] = L.ref;
      mats[(size_t)(9 + k) * SD::NG + g] = L.trans;
    }
  }
  const unsigned need_mask = __ballot_sync(0xffffffffu, need);
  if (!need_mask) return;   // (warp-uniform)
  // ---- 6x6 matrix exponential (:596-727), warp-cooperative (sp_coop.cuh) ----
  // the lanes' matrices wait in shared memory, entry e of lane t at Gs[e * SP_LD + t]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* Gs = reinterpret_cast<double*>(smem_raw) + (size_t)warp * (SP_LW_NE * SP_LD + Coop<6>::PER_WARP);
  double* stage = Gs + SP_LW_NE * SP_LD;
#define G_(r, cc) Gs[((r) * 6 + (cc)) * SP_LD + lane]
  double solution0[6], solution_diff[6], planck_top[6], planck_diff[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) { planck_top[k] = 0.0; planck_diff[k] = 0.0; solution0[k] = 0.0; solution_diff[k] = 0.0; }
  if (need) {
    for (int k = 0; k < 36; ++k) Gs[k * SP_LD + lane] = 0.0;
#pragma unroll
    for (int jr = 0; jr < 3; ++jr) {
      if (jr >= nra) continue;
      const double factor = (ECB_LW_DIFFUSIVITY * 0.5) * ssa_r[jr];   // calc_two_stream_gammas_lw
      const double gamma1 = ECB_LW_DIFFUSIVITY - factor * (1.0 + g_r[jr]), gamma2 = factor * (1.0 - g_r[jr]);
      G_(jr, jr) = od_r[jr] * gamma1;
      G_(jr + 3, jr) = od_r[jr] * gamma2;
      planck_top[3 + jr] = od_r[jr] * (1.0 - ssa_r[jr]) * reg[jr] * pt * ECB_LW_DIFFUSIVITY;
      planck_top[jr] = -planck_top[3 + jr];
      planck_diff[3 + jr] = od_r[jr] * (1.0 - ssa_r[jr]) * reg[jr] * (pb - pt) * ECB_LW_DIFFUSIVITY;
      planck_diff[jr] = -planck_diff[3 + jr];
    }
    // non-zeros in the empty cloudy regions of a clear layer, to avoid NaNs (:622-630)
#pragma unroll
    for (int jr = 1; jr < 3; ++jr) if (jr >= nra) { G_(jr, jr) = G_(0, 0); G_(3 + jr, jr) = G_(3, 0); }
    double side_emiss = 1.0;
    if (sc.do_lw_side_emissivity && reg[0] > 0.0 && reg[1] > 0.0 && sc.do_3d_effects && inv_size > 0.0) {
      const double aspect_ratio = 1.0 / (dmin(inv_size, 1.0 / sc.min_cloud_effective_size) * reg[0] * dz);
      double sm = 0.0;
#pragma unroll
      for (int jr = 1; jr < 3; ++jr) sm = sm + od_r[jr] * (1.0 - ssa_r[jr]);
      const double lateral_od = (aspect_ratio / (3 - 1.0)) * sm;
      const double sqrt_1_minus_ssa = sqrt(1.0 - ssa_r[1]);
      const double side_emiss_thick = 2.0 * sqrt_1_minus_ssa / (sqrt_1_minus_ssa + sqrt(1.0 - ssa_r[1] * g_r[1]));
      side_emiss = (1.4107 - side_emiss_thick) / (lateral_od + 1.0) + side_emiss_thick;
    }
#pragma unroll
    for (int jr = 0; jr < 2; ++jr) {
      if (jr + 1 >= nra) continue;
      G_(jr, jr) = G_(jr, jr) + rate[jr * 3 + jr + 1];
      G_(jr + 1, jr) = -rate[jr * 3 + jr + 1];
      if (jr > 0) {
        G_(jr + 1, jr + 1) = G_(jr + 1, jr + 1) + rate[(jr + 1) * 3 + jr];
        G_(jr, jr + 1) = -rate[(jr + 1) * 3 + jr];
      } else {
        G_(jr + 1, jr + 1) = G_(jr + 1, jr + 1) + side_emiss * rate[(jr + 1) * 3 + jr];
        G_(jr, jr + 1) = -side_emiss * rate[(jr + 1) * 3 + jr];
      }
    }
    if (edge[2] > 0.0) {
      G_(0, 0) = G_(0, 0) + rate[2];
      G_(2, 0) = -rate[2];
      G_(2, 2) = G_(2, 2) + side_emiss * rate[6];
      G_(0, 2) = -side_emiss * rate[6];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int bb = 0; bb < 3; ++bb) G_(3 + a, 3 + bb) = -G_(a, bb);
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int bb = 0; bb < 3; ++bb) G_(a, 3 + bb) = -G_(3 + a, bb);
  }
  __syncwarp();
  // exponential of Gamma and, with the same column distribution, the two solves for the particular solution (:697-707)
  coop_expm_warp_shared_lw<6>(Gs, need_mask, stage, planck_top, planck_diff, solution0, solution_diff);
  if (!need) return;
  double E11[9], E12[9], E21[9], E22[9], R[9], X[9];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int bb = 0; bb < 3; ++bb) {
      E11[a * 3 + bb] = G_(a, bb); E12[a * 3 + bb] = G_(a, 3 + bb);
      E21[a * 3 + bb] = G_(3 + a, bb); E22[a * 3 + bb] = G_(3 + a, 3 + bb);
    }
#undef G_
  m3_solve_mat(E11, E12, X);
#pragma unroll
  for (int k = 0; k < 9; ++k) { R[k] = -X[k]; mats[(size_t)k * SD::NG + g] = R[k]; }
  m3_x_m3(E21, R, X);
#pragma unroll
  for (int k = 0; k < 9; ++k) mats[(size_t)(9 + k) * SD::NG + g] = X[k] + E22[k];
  double tmp[3], a3[3], b3[3], su[3];
  m3_x_vec(E12, solution0 + 3, a3);
#pragma unroll
  for (int r = 0; r < 3; ++r) tmp[r] = solution0[r] + solution_diff[r] - a3[r];
  m3_solve_vec(E11, tmp, a3);
#pragma unroll
  for (int r = 0; r < 3; ++r) { su[r] = solution0[r] - a3[r]; tmp[r] = su[r] - solution0[r]; mats[(size_t)(18 + r) * SD::NG + g] = su[r]; }
  m3_x_vec(E21, tmp, a3);
  m3_x_vec(E22, solution0 + 3, b3);
#pragma unroll
  for (int r = 0; r < 3; ++r) mats[(size_t)(21 + r) * SD::NG + g] = a3[r] + solution0[3 + r] - b3[r] + solution_diff[3 + r];
}

// =========================================================================================================
// LW: albedo/source upward, fluxes downward, derivatives (radiation_spartacus_lw.F90:782-1060)
// =========================================================================================================
template <class SD, int MINB>
__global__ void __launch_bounds__(SD::THREADS, MINB)
sp_lw_sweep_kernel(DevTables T, DevCfg cfg, DevIn in, DevOut out, Work w, int nlev) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int c = blockIdx.x, g = threadIdx.x, nl1 = nlev + 1;
  const bool act = g < SD::NG;
  const int gg = act ? g : 0;
  const SpCfg& sc = cfg.sp;
  double* sums = reinterpret_cast<double*>(smem_raw);     // [5][nl1]: up, dn, up_clear, dn_clear, deriv
  double* tile = sums + 5 * nl1;                           // [4][SP_LCH_LW][SD::RS]
  const TcShared S = tc_load_shared(reinterpret_cast<unsigned char*>(tile + 4 * SP_LCH_LW * SD::RS), w, in, c, nlev, SD::THREADS);
  const double* Ug = w.tc_u + (size_t)c * (nlev + 1) * 9;
  const double* Vg = w.tc_v + (size_t)c * (nlev + 1) * 9;
  if (g == 0 && out.cloud_cover_lw) out.cloud_cover_lw[c] = w.tc_cc[c];
  const size_t n = (size_t)nlev * SD::NG;
  const double* clr = w.scr_lw + (size_t)c * sp_doubles_lw(nlev, SD::NG);
  const double* mats = clr + (size_t)SP_LW_CLR * n;
  double* alb = w.scr_lw + (size_t)c * sp_doubles_lw(nlev, SD::NG) + (size_t)(SP_LW_CLR + SP_LW_MAT) * n;
#define MAT(l, k) mats[((size_t)(l) * SP_LW_MAT + (k)) * SD::NG + g]
#define ALB(hl, k) alb[((size_t)(hl) * SP_LW_ALB + (k)) * SD::NG + g]
  const double emission = w.emission[(size_t)c * SD::NG + gg], albedo = w.lw_albedo[(size_t)c * SD::NG + gg];
  const bool matrix_adding = sc.do_3d_effects || sc.do_3d_lw_multilayer_effects;

  // ---- upward sweep ----
  double TA[9], TS[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) TA[k] = 0.0;
  TA[0] = TA[4] = TA[8] = albedo;
#pragma unroll
  for (int jr = 0; jr < 3; ++jr) TS[jr] = S.reg[(nlev - 1) * 3 + jr] * emission;
  double tac = albedo, tsc = emission;
  if (act) {
    for (int l = nlev - 1; l >= 0; --l) {
      const int jl = l + 1;
      const size_t i = (size_t)l * SD::NG + g;
      const bool clear_l = S.clear[jl];
      if (clear_l) { ALB(jl, 0) = TA[0]; ALB(jl, 9) = TS[0]; }
      else {
#pragma unroll
        for (int k = 0; k < 9; ++k) ALB(jl, k) = TA[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) ALB(jl, 9 + k) = TS[k];
      }
      ALB(jl, 12) = tac; ALB(jl, 13) = tsc;
      const double rc = clr[i], trc = clr[n + i], suc = clr[2 * n + i], sdc = clr[3 * n + i];
      {
        const double inv_denom = sp_inv(1.0 - tac * rc);
        const double tac_new = rc + trc * trc * tac * inv_denom;
        tsc = suc + trc * (tsc + tac * sdc) * inv_denom;
        tac = tac_new;
      }
      double below[9], sbelow[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int k = 0; k < 9; ++k) below[k] = 0.0;
      if (clear_l) {
        const bool ev = sc.use_expm_everywhere != 0;
        const double r0 = ev ? MAT(l, 0) : rc, t0 = ev ? MAT(l, 9) : trc;
        const double su0 = ev ? MAT(l, 18) : S.reg[l * 3] * suc, sd0 = ev ? MAT(l, 21) : S.reg[l * 3] * sdc;
        const double inv_denom = sp_inv(1.0 - TA[0] * r0);
        below[0] = r0 + t0 * t0 * TA[0] * inv_denom;
        sbelow[0] = su0 + t0 * (TS[0] + TA[0] * sd0) * inv_denom;
      } else {
        double R[9], Tm[9], SU[3], SDn[3];
#pragma unroll
        for (int k = 0; k < 9; ++k) { R[k] = MAT(l, k); Tm[k] = MAT(l, 9 + k); }
#pragma unroll
        for (int k = 0; k < 3; ++k) { SU[k] = MAT(l, 18 + k); SDn[k] = MAT(l, 21 + k); }
        if (matrix_adding) {
          double den[9], X[9], a[3];
          m3_identity_minus_product(TA, R, den);
          m3_x_m3(TA, Tm, X);
          m3_solve_mat(den, X, X);
          m3_x_m3(Tm, X, X);
#pragma unroll
          for (int k = 0; k < 9; ++k) below[k] = R[k] + X[k];
          m3_x_vec(TA, SDn, a);
#pragma unroll
          for (int r = 0; r < 3; ++r) a[r] = TS[r] + a[r];
          m3_solve_vec(den, a, a);
          m3_x_vec(Tm, a, a);
#pragma unroll
          for (int r = 0; r < 3; ++r) sbelow[r] = SU[r] + a[r];
        } else {
#pragma unroll
          for (int jr = 0; jr < 3; ++jr) {
            const double inv_denom = sp_inv(1.0 - TA[jr * 4] * R[jr * 4]);
            below[jr * 4] = R[jr * 4] + Tm[jr * 4] * Tm[jr * 4] * TA[jr * 4] * inv_denom;
            sbelow[jr] = SU[jr] + Tm[jr * 4] * (TS[jr] + TA[jr * 4] * SDn[jr]) * inv_denom;
          }
        }
      }
      double U[9], V[9];
      sp_load_m3(Ug + l * 9, U);
      sp_load_m3(Vg + l * 9, V);
#pragma unroll
      for (int k = 0; k < 9; ++k) TA[k] = 0.0;
      if (clear_l && S.clear[jl - 1]) {
        TA[0] = below[0];
        TS[0] = sbelow[0]; TS[1] = 0.0; TS[2] = 0.0;
      } else {
        m3_x_vec(U, sbelow, TS);
        if (sc.do_3d_lw_multilayer_effects) {
          m3_u_a_v(U, below, V, TA);
        } else {
#pragma unroll
          for (int jr = 0; jr < 3; ++jr)
#pragma unroll
            for (int jr2 = 0; jr2 < 3; ++jr2) TA[jr * 4] = TA[jr * 4] + below[jr2 * 4] * V[jr2 * 3 + jr];
        }
      }
    }
  }
  // ---- downward sweep ----
  double fdn[3] = {0.0, 0.0, 0.0}, fup[3] = {TS[0], TS[1], TS[2]};
  double fdc = 0.0, fuc = tsc;
  const double toa_a = TS[0] + TS[1] + TS[2], toa_c = tsc;
  double* dst[4] = {sums, sums + nl1, sums + 2 * nl1, sums + 3 * nl1};
  const BandOut bo[2] = {{cfg.do_save_spectral_flux ? out.lw_up_band : nullptr, out.ld, 0, -1, 1.0, 0.0, nullptr, 0},
                         {cfg.do_save_spectral_flux ? out.lw_dn_band : nullptr, out.ld, 1, -1, 1.0, 0.0, nullptr, 0}};
  const bool bands = cfg.do_save_spectral_flux && (out.lw_up_band || out.lw_dn_band);
  int slot = 0, lfirst = 0;
#define PUT_ROWS()                                                            \
  if (act) {                                                                  \
    tile[(0 * SP_LCH_LW + slot) * SD::RS + g] = fup[0] + fup[1] + fup[2];   \
    tile[(1 * SP_LCH_LW + slot) * SD::RS + g] = fdn[0] + fdn[1] + fdn[2];   \
    tile[(2 * SP_LCH_LW + slot) * SD::RS + g] = fuc;                        \
    tile[(3 * SP_LCH_LW + slot) * SD::RS + g] = fdc;                        \
  }                                                                           \
  ++slot;
  PUT_ROWS();
  double dn_surf_g = 0.0;
  for (int l = 0; l < nlev; ++l) {
    const int jl = l + 1;
    if (act) {
      const size_t i = (size_t)l * SD::NG + g;
      const bool clear_l = S.clear[jl];
      const double rc = clr[i], trc = clr[n + i], sdc = clr[3 * n + i];
      const double tacb = ALB(jl, 12), tscb = ALB(jl, 13);
      fdc = (trc * fdc + rc * tscb + sdc) * sp_inv(1.0 - rc * tacb);   // (the reciprocal does not wait for the flux recurrence)
      fuc = tscb + tacb * fdc;
      if (clear_l) {
        const bool ev = sc.use_expm_everywhere != 0;
        const double r0 = ev ? MAT(l, 0) : rc, t0 = ev ? MAT(l, 9) : trc;
        const double ta0 = ALB(jl, 0), ts0 = ALB(jl, 9), sd0 = ev ? MAT(l, 21) : S.reg[l * 3] * sdc;
        fdn[0] = (t0 * fdn[0] + r0 * ts0 + sd0) * sp_inv(1.0 - r0 * ta0);
        fup[0] = ts0 + ta0 * fdn[0];
        fdn[1] = 0.0; fdn[2] = 0.0; fup[1] = 0.0; fup[2] = 0.0;
      } else {
        double R[9], Tm[9], TAb[9], TSb[3], SDn[3];
#pragma unroll
        for (int k = 0; k < 9; ++k) { R[k] = MAT(l, k); Tm[k] = MAT(l, 9 + k); TAb[k] = ALB(jl, k); }
#pragma unroll
        for (int k = 0; k < 3; ++k) { SDn[k] = MAT(l, 21 + k); TSb[k] = ALB(jl, 9 + k); }
        if (matrix_adding) {
          double den[9], a[3], bb[3];
          m3_identity_minus_product(R, TAb, den);
          m3_x_vec(Tm, fdn, a);
          m3_x_vec(R, TSb, bb);
#pragma unroll
          for (int r = 0; r < 3; ++r) a[r] = a[r] + bb[r] + SDn[r];
          m3_solve_vec(den, a, fdn);
          m3_x_vec(TAb, fdn, fup);
#pragma unroll
          for (int r = 0; r < 3; ++r) fup[r] = fup[r] + TSb[r];
        } else {
#pragma unroll
          for (int jr = 0; jr < 3; ++jr) {
            fdn[jr] = (Tm[jr * 4] * fdn[jr] + R[jr * 4] * TSb[jr] + SDn[jr]) * sp_inv(1.0 - R[jr * 4] * TAb[jr * 4]);
            fup[jr] = TSb[jr] + TAb[jr * 4] * fdn[jr];
          }
        }
      }
    }
    PUT_ROWS();
    if (l == nlev - 1) dn_surf_g = fdn[0] + fdn[1] + fdn[2];
    if (act && l < nlev - 1 && !(S.clear[jl] && S.clear[jl + 1])) {
      double V[9];
      sp_load_m3(Vg + jl * 9, V);
      m3_x_vec(V, fdn, fdn);
    }
    if (slot == SP_LCH_LW || l == nlev - 1) {
      if (bands) flush_bands(tile, SD::RS, SP_LCH_LW, slot, bo, 2, lfirst, 1, c, SD::NB, T.meta->lw);
      flush_tile(tile, SD::RS, SD::NG, 4, slot, dst, lfirst, 1, SP_LCH_LW); lfirst += slot; slot = 0;
    }
  }
#undef PUT_ROWS
  // ---- derivatives (calc_lw_derivatives_matrix): rate of change of the upwelling flux with the surface value ----
  const bool want_dv = cfg.do_lw_derivatives && out.lw_derivatives;
  if (want_dv) {
    double d[3] = {(fup[0] + fup[1] + fup[2]) / sums[nlev], 0.0, 0.0};
    double* dv = sums + 4 * nl1;
    int slot2 = 0, lf = nlev - 1;
    for (int l = nlev - 1; l >= 0; --l) {
      const int jl = l + 1;
      if (act) {
        double U[9];
        sp_load_m3(Ug + jl * 9, U);   // u_matrix(:,:,jlev+1)
        m3_x_vec(U, d, d);
        if (S.clear[jl]) {
          const double trc = sc.use_expm_everywhere ? MAT(l, 9) : clr[n + (size_t)l * SD::NG + g];
          d[0] = trc * d[0]; d[1] = 0.0; d[2] = 0.0;
        } else {
          double Tm[9];
#pragma unroll
          for (int k = 0; k < 9; ++k) Tm[k] = MAT(l, 9 + k);
          m3_x_vec(Tm, d, d);
        }
        tile[slot2 * SD::RS + g] = d[0] + d[1] + d[2];
      }
      ++slot2;
      if (slot2 == SP_LCH_LW || l == 0) {
        double* dst1[1] = {dv};
        flush_tile(tile, SD::RS, SD::NG, 1, slot2, dst1, lf, -1, SP_LCH_LW); lf -= slot2; slot2 = 0;
      }
    }
  }
#undef MAT
#undef ALB
  __syncthreads();
  for (int l = g; l < nl1; l += SD::THREADS) {
    const size_t o = (size_t)l * out.ld + c;
    if (out.lw_up) out.lw_up[o] = sums[l];
    if (out.lw_dn) out.lw_dn[o] = l == 0 ? 0.0 : sums[nl1 + l];
    if (out.lw_up_clear) out.lw_up_clear[o] = sums[2 * nl1 + l];
    if (out.lw_dn_clear) out.lw_dn_clear[o] = l == 0 ? 0.0 : sums[3 * nl1 + l];
    if (want_dv) out.lw_derivatives[o] = l == nlev ? 1.0 : sums[4 * nl1 + l];
  }
  if (act) {
    const size_t i = (size_t)c * SD::NG + T.meta->rank_lw[g];   // flux_type's g-point arrays are in the solver's (reordered) sequence
    if (out.lw_dn_surf_clear_g) out.lw_dn_surf_clear_g[i] = fdc;
    if (out.lw_up_toa_clear_g) out.lw_up_toa_clear_g[i] = toa_c;
    if (out.lw_dn_surf_g) out.lw_dn_surf_g[i] = dn_surf_g;
    if (out.lw_up_toa_g) out.lw_up_toa_g[i] = toa_a;
  }
  lw_surface_canopy<SD>(T, cfg, out, c, g, act, tile, dn_surf_g);
}

// =========================================================================================================
// resident CTAs per SM asked of the compiler (register budget = 65536 / (MINB * threads)); tuning knobs for the measurement scripts
static int sp_minb(const char* env, int dflt) {
  const char* s = getenv(env);
  const int v = s ? atoi(s) : dflt;
  return v <= 2 ? 2 : v == 3 ? 3 : 4;
}
// measured on the B200 (tools/sp_minb_sweep.sh, 20 000 columns): 2/2 -> 97 k, 4/2 -> 104 k, 2/4 -> 107 k, 4/4 -> 115 k columns/s
#define SP_DISPATCH_MINB(minb, CALL) \
  switch (minb) { case 2: { constexpr int MB = 2; CALL; } break; case 3: { constexpr int MB = 3; CALL; } break; default: { constexpr int MB = 4; CALL; } break; }

template <class SD>
static int launch_sp_sw_t(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, const Work& w, int nc, int nlev, cudaStream_t st) {
  static const int mb_layer = sp_minb("ECRAD_B200_SP_MINB_LAYER_SW", 4), mb_sweep = sp_minb("ECRAD_B200_SP_MINB_SWEEP", 4);
  // cloud-free layers first (no shared memory), then the layers whose g-points need the matrix exponential
  const size_t sml = sizeof(double) * (size_t)(SD::THREADS / 32) * Coop<9>::PER_WARP;
  SP_DISPATCH_MINB(mb_layer, (cudaFuncSetAttribute(sp_sw_layer_kernel<SD, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sml),
                              sp_sw_layer_kernel<SD, MB><<<dim3(nlev, nc), SD::THREADS, 0, st>>>(T, cfg, in, w, nlev, 0),
                              sp_sw_layer_kernel<SD, MB><<<dim3(nlev, nc), SD::THREADS, sml, st>>>(T, cfg, in, w, nlev, 1)));
  const size_t sm = sizeof(double) * (6 * (nlev + 1) + 6 * SP_LCH * SD::RS + 2 * SD::NB) + sp_shared_bytes(nlev);
  SP_DISPATCH_MINB(mb_sweep, (cudaFuncSetAttribute(sp_sw_sweep_kernel<SD, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm),
                              sp_sw_sweep_kernel<SD, MB><<<nc, SD::THREADS, sm, st>>>(T, cfg, in, out, w, nlev)));
  return 3;
}
template <class SD>
static int launch_sp_lw_t(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, const Work& w, int nc, int nlev, cudaStream_t st) {
  static const int mb_layer = sp_minb("ECRAD_B200_SP_MINB_LAYER_LW", 3), mb_sweep = sp_minb("ECRAD_B200_SP_MINB_SWEEP", 4);
  const size_t sml = sizeof(double) * (size_t)(SD::THREADS / 32) * (SP_LW_NE * SP_LD + Coop<6>::PER_WARP);
  SP_DISPATCH_MINB(mb_layer, (cudaFuncSetAttribute(sp_lw_layer_kernel<SD, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sml),
                              sp_lw_layer_kernel<SD, MB><<<dim3(nlev, nc), SD::THREADS, 0, st>>>(T, cfg, in, w, nlev, 0),
                              sp_lw_layer_kernel<SD, MB><<<dim3(nlev, nc), SD::THREADS, sml, st>>>(T, cfg, in, w, nlev, 1)));
  const size_t sm = sizeof(double) * (5 * (nlev + 1) + 4 * SP_LCH_LW * SD::RS) + tc_shared_bytes(nlev);
  SP_DISPATCH_MINB(mb_sweep, (cudaFuncSetAttribute(sp_lw_sweep_kernel<SD, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm),
                              sp_lw_sweep_kernel<SD, MB><<<nc, SD::THREADS, sm, st>>>(T, cfg, in, out, w, nlev)));
  return 3;
}
int launch_sp_sw(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, const Work& w, int nc, int nlev, cudaStream_t st) {
  switch (cfg.ng_sw) {
    case NG_SW: return launch_sp_sw_t<SwRrtmg>(T, cfg, in, out, w, nc, nlev, st);
    case 32: return launch_sp_sw_t<Ckd32>(T, cfg, in, out, w, nc, nlev, st);
    case 64: return launch_sp_sw_t<Ckd64>(T, cfg, in, out, w, nc, nlev, st);
    case 96: return launch_sp_sw_t<Ckd96>(T, cfg, in, out, w, nc, nlev, st);
  }
  return -1;
}
int launch_sp_lw(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, const Work& w, int nc, int nlev, cudaStream_t st) {
  switch (cfg.ng_lw) {
    case NG_LW: return launch_sp_lw_t<LwRrtmg>(T, cfg, in, out, w, nc, nlev, st);
    case 32: return launch_sp_lw_t<Ckd32>(T, cfg, in, out, w, nc, nlev, st);
    case 64: return launch_sp_lw_t<Ckd64>(T, cfg, in, out, w, nc, nlev, st);
    case 96: return launch_sp_lw_t<Ckd96>(T, cfg, in, out, w, nc, nlev, st);
  }
  return -1;
}

}  // namespace ecb
