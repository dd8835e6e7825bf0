// solver_tc.cu -- Tripleclouds solvers (3 regions: clear, optically thin cloud, optically thick cloud).
//
// Reference: radiation/radiation_tripleclouds_sw.F90:42-661, radiation_tripleclouds_lw.F90:38-605,
// radiation_regions.F90, radiation_overlap.F90, radiation_lw_derivatives.F90:200-290 (calc_lw_derivatives_region).
// One CTA per column, one thread per g-point holding the three regions' albedo/source (upward sweep) and fluxes
// (downward sweep) in registers; the 3x3 overlap matrices of every half-level sit in shared memory.  What the downward
// sweep needs from the upward one goes to the per-column scratch, pre-combined per region:
//   SW:  a = T/(1-R*A), b = (Tdir*Adir*R + Tdirdif)/(1-R*A), Tdir, A, Adir       (A, Adir: albedos of everything below)
//   LW:  a = T/(1-R*A), b = (R*S + src_dn)/(1-R*A), A, S, T                       (S: source of everything below)
// First version: one kernel per spectrum (not yet split/tuned like the McICA path).
//
// The Homogeneous solvers (radiation_homogeneous_sw.F90, radiation_homogeneous_lw.F90: every cloudy layer is overcast with the
// gridbox-mean cloud) run on the same kernels: tc_prep_kernel then puts the whole layer into region 2 with unit optical-depth
// scaling and maximum overlap, so every overlap matrix is an exact permutation of 0s and 1s and the unused regions carry exact
// zeros; the SW two-stream solution is the gammas + calc_reflectance_transmittance_sw pair that solver uses.
#include "solver_common.cuh"
#include "tc_core.h"
#include "tc_shared.cuh"

namespace ecb {
enum { TC_PF_DIST = 4 };

// layers between two g-point reductions: rows per flush (6 x 4 in the SW, 2 x 16 in the LW) ~ the 32-40 row slots of flush_tile;
// the tile competes with resident CTAs for shared memory
enum { TC_LCH = 4, TC_LCH_LW = 16 };
enum { TC_SW_ARRAYS = 20, TC_LW_ARRAYS = 15 };

// ---------------------------------------------------------------------------------------------------------
// region fractions, optical-depth scalings and overlap matrices (calc_region_properties + calc_overlap_matrices): one CTA
// per column, one thread per half-level -- an interface only needs the regions of the two layers it separates, which each
// thread recomputes (tc_region is a handful of operations); the total cloud cover 1 - prod v_matrix(1,1,:) is then
// multiplied up in interface order by one thread, as the reference does.
// ---------------------------------------------------------------------------------------------------------
enum { TC_PREP_THREADS = 160 };
__global__ void __launch_bounds__(TC_PREP_THREADS)
tc_prep_kernel(DevCfg cfg, DevIn in, Work w, int nc, int nlev) {
  __shared__ double v11[264];   // nlev + 1 <= 257 (ecrad_b200_radiation refuses more levels)
  const int c = blockIdx.x;
  double* reg = w.tc_reg + (size_t)c * nlev * 3;
  double* ods = w.tc_ods + (size_t)c * nlev * 3;
  const double expo = 1.0 / cfg.cloud_inhom_decorr_scaling, thr = cfg.cloud_fraction_threshold;
  for (int jlev = 1 + (int)threadIdx.x; jlev <= nlev + 1; jlev += TC_PREP_THREADS) {   // interface above layer jlev (1-based)
    double fu[3] = {1.0, 0.0, 0.0}, fl[3] = {1.0, 0.0, 0.0}, o_[3], M[3][3];
    const bool homog = cfg.is_homogeneous != 0;
    if (jlev > 1) { if (homog) tc_region_homogeneous(LD_IN(in.frac, c, jlev - 2), thr, fu, o_); else tc_region(LD_IN(in.frac, c, jlev - 2), LD_IN(in.fsd, c, jlev - 2), thr, fu, o_, cfg.pdf_gamma != 0); }
    if (jlev <= nlev) {
      if (homog) tc_region_homogeneous(LD_IN(in.frac, c, jlev - 1), thr, fl, o_); else tc_region(LD_IN(in.frac, c, jlev - 1), LD_IN(in.fsd, c, jlev - 1), thr, fl, o_, cfg.pdf_gamma != 0);
      for (int r = 0; r < 3; ++r) { reg[(jlev - 1) * 3 + r] = fl[r]; ods[(jlev - 1) * 3 + r] = o_[r]; }
    }
    if (cfg.sp.two_regions) {
      // config%nregions = 2 (radiation_regions.F90:105-110): clear sky + one homogeneous cloudy region.  Carried as three regions
      // with an empty third one: every formula of the three-region path (overlap matrices, radiation_overlap.F90:169-209; edge
      // lengths; exchange terms) then reduces to the two-region one, the third region exchanges nothing and carries no flux.
      if (jlev > 1) { const double f = LD_IN(in.frac, c, jlev - 2); fu[0] = 1.0 - f; fu[1] = f; fu[2] = 0.0; }
      if (jlev <= nlev) {
        const double f = LD_IN(in.frac, c, jlev - 1);
        fl[0] = 1.0 - f; fl[1] = f; fl[2] = 0.0;
        for (int r = 0; r < 3; ++r) { reg[(jlev - 1) * 3 + r] = fl[r]; ods[(jlev - 1) * 3 + r] = r == 0 ? 0.0 : 1.0; }
      }
    }
    double op1 = 1.0, op2 = 1.0;
    if (jlev > 1 && jlev <= nlev && !homog) {
      op1 = LD_IN(in.overlap, c, jlev - 2);
      op2 = op1 >= 0.0 ? (expo == 2.0 ? mul_rn(op1, op1) : pow(op1, expo)) : op1;
    }
    if (cfg.use_beta_overlap && !homog) { const double op[3] = {op1, op2, op2}; tc_beta_overlap_matrix(op, fu, fl, thr, M); }
    else tc_alpha_overlap_matrix(op1, op2, fu, fl, M);
    double* u = w.tc_u + ((size_t)c * (nlev + 1) + (jlev - 1)) * 9;
    double* v = w.tc_v + ((size_t)c * (nlev + 1) + (jlev - 1)) * 9;
    for (int ju = 0; ju < 3; ++ju)
      for (int jw = 0; jw < 3; ++jw) {
        u[ju * 3 + jw] = fl[jw] >= thr ? M[ju][jw] / fl[jw] : 0.0;
        v[jw * 3 + ju] = fu[ju] >= thr ? M[ju][jw] / fu[ju] : 0.0;
      }
    if (jlev - 1 < 264) v11[jlev - 1] = v[0];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double prod = 1.0;
    for (int k = 0; k <= nlev; ++k) prod = prod * v11[k];
    w.tc_cc[c] = 1.0 - prod;
  }
}

// =========================================================================================================
// SW
// =========================================================================================================
template <class SD>
__global__ void __launch_bounds__(SD::THREADS, scaled_min_blocks(SD::THREADS, 128, 5))
tc_sw_kernel(DevTables T, DevCfg cfg, DevIn in, DevOut out, Work w, int nlev) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int c = blockIdx.x, g = threadIdx.x, nl1 = nlev + 1;
  const bool act = g < SD::NG;
  const int gg = act ? g : 0;
  const double mu0 = in.cos_sza[c];
  const bool homog = cfg.is_homogeneous != 0;   // Homogeneous: no cloud cover output, night is cos_sza <= 0 (radiation_homogeneous_sw.F90:118)
  if (g == 0 && out.cloud_cover_sw && !homog) out.cloud_cover_sw[c] = w.tc_cc[c];   // set for every column, also at night
  if (homog ? !(mu0 > 0.0) : mu0 < 1.0e-10) { sw_night_column<SD>(cfg, out, c, g, act, nl1, SD::THREADS); return; }
  double* sums = reinterpret_cast<double*>(smem_raw);     // [6][nl1]: up, dn_dif, dn_dir, up_c, dn_dif_c, dn_dir_c
  double* tile = sums + 6 * nl1;                           // [6][TC_LCH][SD::RS]
  double* bandv = tile + 6 * TC_LCH * SD::RS;            // [2][14]
  const TcShared S = tc_load_shared(reinterpret_cast<unsigned char*>(bandv + 2 * SD::NB), w, in, c, nlev, SD::THREADS);
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
  const double* od = w.od_sw + (size_t)c * n;
  const double* ssa = w.ssa_sw + (size_t)c * n;
  const double* gas_g = (cfg.use_aerosols && w.g_sw) ? w.g_sw + (size_t)c * n : nullptr;
  const double* cl = w.cl_sw + (size_t)c * nlev * 3 * SD::NB;
  double* scr = w.scr_sw + (size_t)c * TC_SW_ARRAYS * n;
#define SCR(set, f, i) scr[(size_t)((set) * 5 + (f)) * n + (i)]   // sets: 0 clear-sky, 1..3 regions; fields: a b tdir talb talbdir
  const int b = T.meta->band_of_g_sw[gg];
  const double alb_diff = bandv[b], alb_dir = bandv[SD::NB + b];
  const double inc = w.incoming[(size_t)c * SD::NG + gg];

  // ---- upward sweep: total albedos of everything below each half-level (radiation_tripleclouds_sw.F90:322-420) ----
  double ta[3] = {alb_diff, 0.0, 0.0}, td[3] = {mu0 * alb_dir, 0.0, 0.0};
  if (!S.clear[nlev]) { ta[1] = ta[0]; ta[2] = ta[0]; td[1] = td[0]; td[2] = td[0]; }
  double tac = ta[0], tdc = td[0];
  if (act) {
    // software pipeline: the gas optical properties of layer l-1 are loaded before the two-stream arithmetic of layer l
    size_t in_ = (size_t)(nlev - 1) * SD::NG + g;
    double od_n = od[in_], ssa_n = ssa[in_], gg_n = gas_g ? gas_g[in_] : 0.0;
    for (int l = nlev - 1; l >= 0; --l) {
      const int jl = l + 1;
      const size_t i = (size_t)l * SD::NG + g;
      const double odg = od_n, ssag = ssa_n, gg_gas = gg_n;
      if (l > 0) { od_n = od[i - SD::NG]; ssa_n = ssa[i - SD::NG]; if (gas_g) gg_n = gas_g[i - SD::NG]; }
      if (l > TC_PF_DIST) {   // (rows a few layers further up: into L2 now, no register held)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(od + i - (size_t)TC_PF_DIST * SD::NG));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(ssa + i - (size_t)TC_PF_DIST * SD::NG));
      }
      // do_sw_delta_scaling_with_gases: the Homogeneous solver scales the clear-sky mixture too (radiation_homogeneous_sw.F90:145-175),
      // Tripleclouds only its cloudy regions (radiation_tripleclouds_sw.F90:269 vs :298-302)
      double odc = odg, ssac = ssag, gc = gg_gas;
      if (homog && cfg.do_sw_delta_scaling_with_gases) sw_delta_eddington(odc, ssac, gc);
      const SwLayer Lc = homog ? sw_ref_trans_cloudless(mu0, odc, ssac, gc) : sw_ref_trans(mu0, odg, ssag, gg_gas);
      {   // clear-sky column
        const double id = 1.0 / (1.0 - tac * Lc.ref);
        SCR(0, 0, i) = Lc.trans * id;
        SCR(0, 1, i) = (Lc.trans_dir_dir * tdc * Lc.ref + Lc.trans_dir_diff) * id;
        SCR(0, 2, i) = Lc.trans_dir_dir; SCR(0, 3, i) = tac; SCR(0, 4, i) = tdc;
        const double tac_new = Lc.ref + Lc.trans * Lc.trans * tac * id;
        tdc = Lc.ref_dir + (Lc.trans_dir_dir * tdc + Lc.trans_dir_diff * tac) * Lc.trans * id;
        tac = tac_new;
      }
      double below[3] = {0.0, 0.0, 0.0}, belowd[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int jr = 0; jr < 3; ++jr) {
        if (jr > 0 && S.clear[jl]) continue;
        SwLayer L = Lc;
        if (jr > 0) {   // cloudy region: gas/aerosol + scaled cloud (radiation_tripleclouds_sw.F90:286-302)
          const double* clb = cl + (size_t)l * 3 * SD::NB;
          const double scal = S.ods[l * 3 + jr];
          const double scat_od = odg * ssag;
          const double scat_od_cloud = clb[b] * clb[SD::NB + b] * scal;
          const double od_total = odg + clb[b] * scal;
          const double ssa_total = (scat_od + scat_od_cloud) / od_total;
          const double g_total = (scat_od * gg_gas + scat_od_cloud * clb[2 * SD::NB + b]) / (scat_od + scat_od_cloud);
          double odt = od_total, ssat = ssa_total, gt = g_total;
          if (cfg.do_sw_delta_scaling_with_gases) sw_delta_eddington(odt, ssat, gt);
          L = homog ? sw_ref_trans_cloudless(mu0, odt, ssat, gt) : sw_ref_trans(mu0, odt, ssat, gt);
        }
        const double id = 1.0 / (1.0 - ta[jr] * L.ref);
        SCR(1 + jr, 0, i) = L.trans * id;
        SCR(1 + jr, 1, i) = (L.trans_dir_dir * td[jr] * L.ref + L.trans_dir_diff) * id;
        SCR(1 + jr, 2, i) = L.trans_dir_dir; SCR(1 + jr, 3, i) = ta[jr]; SCR(1 + jr, 4, i) = td[jr];
        below[jr] = L.ref + L.trans * L.trans * ta[jr] * id;
        belowd[jr] = L.ref_dir + (L.trans_dir_dir * td[jr] + L.trans_dir_diff * ta[jr]) * L.trans * id;
      }
      if (S.clear[jl] && S.clear[jl - 1]) {
#pragma unroll
        for (int jr = 0; jr < 3; ++jr) { ta[jr] = below[jr]; td[jr] = belowd[jr]; }
      } else {
        const double* V = S.V + l * 9;   // v_matrix(:,:,jlev)
#pragma unroll
        for (int jr = 0; jr < 3; ++jr) {
          double a = 0.0, d = 0.0;
#pragma unroll
          for (int jr2 = 0; jr2 < 3; ++jr2) { a = a + below[jr2] * V[jr2 * 3 + jr]; d = d + belowd[jr2] * V[jr2 * 3 + jr]; }
          ta[jr] = a; td[jr] = d;
        }
      }
    }
  }
  // ---- downward sweep: fluxes (radiation_tripleclouds_sw.F90:424-645) ----
  double ddn[3], fdn[3] = {0.0, 0.0, 0.0}, fup[3];
#pragma unroll
  for (int jr = 0; jr < 3; ++jr) { ddn[jr] = inc * S.reg[jr]; fup[jr] = ddn[jr] * td[jr]; }
  double ddc = inc, fdc = 0.0, fuc = ddc * tdc;
  const double toa_a = fup[0] + fup[1] + fup[2], toa_c = fuc;
  double* dst[6] = {sums, sums + nl1, sums + 2 * nl1, sums + 3 * nl1, sums + 4 * nl1, sums + 5 * nl1};
  // do_save_spectral_flux: per-band all-sky profiles (radiation_tripleclouds_sw.F90:604-624)
  const bool bands = cfg.do_save_spectral_flux && (out.sw_up_band || out.sw_dn_band || out.sw_dn_direct_band);
  const BandOut bo[3] = {{out.sw_up_band, out.ld, 0, -1, 1.0, 0.0, nullptr, 0}, {out.sw_dn_direct_band, out.ld, 2, -1, mu0, 0.0, nullptr, 0},
                         {out.sw_dn_band, out.ld, 2, 1, mu0, 1.0, nullptr, 0}};
  int slot = 0, lfirst = 0;
#define PUT_ROWS()                                                                                        \
  if (act) {                                                                                              \
    tile[(0 * TC_LCH + slot) * SD::RS + g] = fup[0] + fup[1] + fup[2];                                  \
    tile[(1 * TC_LCH + slot) * SD::RS + g] = fdn[0] + fdn[1] + fdn[2];                                  \
    tile[(2 * TC_LCH + slot) * SD::RS + g] = ddn[0] + ddn[1] + ddn[2];                                  \
    tile[(3 * TC_LCH + slot) * SD::RS + g] = fuc;                                                       \
    tile[(4 * TC_LCH + slot) * SD::RS + g] = fdc;                                                       \
    tile[(5 * TC_LCH + slot) * SD::RS + g] = ddc;                                                       \
  }                                                                                                       \
  ++slot;
  PUT_ROWS();
  // software pipeline: the ten values of the clear-sky column and of region 1 for layer l+1 are requested before the arithmetic of
  // layer l (the recurrences are two multiply-adds per layer behind a global load: profiles/r2u_tc_sw_kernel_lines.txt)
  double nx[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) nx[k] = act ? SCR(k / 5, k % 5, g) : 0.0;
  for (int l = 0; l < nlev; ++l) {
    const int jl = l + 1;
    if (act) {
      const size_t i = (size_t)l * SD::NG + g;
      double cu[10];
#pragma unroll
      for (int k = 0; k < 10; ++k) cu[k] = nx[k];
      if (l + 1 < nlev) {
#pragma unroll
        for (int k = 0; k < 10; ++k) nx[k] = SCR(k / 5, k % 5, i + SD::NG);
      }
      if (l + TC_PF_DIST < nlev) {   // (the same ten rows a few layers further down: into L2 now)
#pragma unroll
        for (int k = 0; k < 10; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(&SCR(k / 5, k % 5, i + (size_t)TC_PF_DIST * SD::NG)));
      }
      fdc = cu[0] * fdc + cu[1] * ddc;
      ddc = cu[2] * ddc;
      fuc = ddc * cu[4] + fdc * cu[3];
      fdn[0] = cu[5] * fdn[0] + cu[6] * ddn[0];
      ddn[0] = cu[7] * ddn[0];
      fup[0] = ddn[0] * cu[9] + fdn[0] * cu[8];
#pragma unroll
      for (int jr = 1; jr < 3; ++jr) {
        if (S.clear[jl]) { fdn[jr] = 0.0; fup[jr] = 0.0; ddn[jr] = 0.0; continue; }
        fdn[jr] = SCR(1 + jr, 0, i) * fdn[jr] + SCR(1 + jr, 1, i) * ddn[jr];
        ddn[jr] = SCR(1 + jr, 2, i) * ddn[jr];
        fup[jr] = ddn[jr] * SCR(1 + jr, 4, i) + fdn[jr] * SCR(1 + jr, 3, i);
      }
      if (!(S.clear[jl] && S.clear[jl + 1])) {
        const double* V = S.V + jl * 9;   // v_matrix(:,:,jlev+1)
        mat3_x_vec(V, fdn);
        mat3_x_vec(V, ddn);
      }
    }
    PUT_ROWS();
    if (slot == TC_LCH || l == nlev - 1) {
      if (bands) flush_bands(tile, SD::RS, TC_LCH, slot, bo, 3, lfirst, 1, c, SD::NB, T.meta->sw);
      flush_tile(tile, SD::RS, SD::NG, 6, slot, dst, lfirst, 1, TC_LCH); lfirst += slot; slot = 0;
    }
  }
#undef PUT_ROWS
#undef SCR
  // ---- outputs ----
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
  const double dif_a = fdn[0] + fdn[1] + fdn[2], dir_a = mu0 * (ddn[0] + ddn[1] + ddn[2]), dif_c = fdc, dir_c = mu0 * ddc;
  if (act) {
    const size_t i = (size_t)c * SD::NG + g;
    if (out.sw_dn_diffuse_surf_g) out.sw_dn_diffuse_surf_g[i] = dif_a;
    if (out.sw_dn_direct_surf_g) out.sw_dn_direct_surf_g[i] = dir_a;
    if (out.sw_dn_diffuse_surf_clear_g) out.sw_dn_diffuse_surf_clear_g[i] = dif_c;
    if (out.sw_dn_direct_surf_clear_g) out.sw_dn_direct_surf_clear_g[i] = dir_c;
    if (out.sw_up_toa_g) out.sw_up_toa_g[i] = toa_a;
    if (out.sw_up_toa_clear_g) out.sw_up_toa_clear_g[i] = toa_c;
    if (out.sw_dn_toa_g && !homog) out.sw_dn_toa_g[i] = inc * mu0;   // radiation_tripleclouds_sw.F90:444
  }
  sw_surface_spectral<SD>(T, cfg, out, c, g, act, tile, SD::RS, dir_a, dif_a, dir_c, dif_c);
}

// =========================================================================================================
// LW (after lw_down_kernel: clear-sky flux_dn sums, flux_dn at cloud top and at the surface per g-point)
// =========================================================================================================
template <class SD>
__global__ void __launch_bounds__(SD::THREADS, scaled_min_blocks(SD::THREADS, 160, 4))
tc_lw_kernel(DevTables T, DevCfg cfg, DevIn in, DevOut out, Work w, int nlev) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int c = blockIdx.x, g = threadIdx.x, nl1 = nlev + 1;
  const bool act = g < SD::NG;
  const int gg = act ? g : 0;
  double* sums = reinterpret_cast<double*>(smem_raw);     // [4][nl1]: up_clear, up, dn, deriv
  double* tile = sums + 4 * nl1;                           // [2][TC_LCH_LW][SD::RS]
  double* red = tile + 2 * TC_LCH_LW * SD::RS;              // [8] block reduction scratch
  const TcShared S = tc_load_shared(reinterpret_cast<unsigned char*>(red + 8), w, in, c, nlev, SD::THREADS);
  const size_t n = (size_t)nlev * SD::NG;
  const double* od = w.od_lw + (size_t)c * n;
  const double* pl = w.planck + (size_t)c * nl1 * SD::NG;
  const double* cl = w.cl_lw + (size_t)c * nlev * 3 * SD::NB;
  const double* gsum = w.lw_sums + (size_t)c * 6 * nl1;     // row 0: clear-sky flux_dn sums (lw_down_kernel)
  const double* carry = w.lw_carry + (size_t)c * 4 * SD::NG;
  double* scr = w.scr_lw + (size_t)c * TC_LW_ARRAYS * n;
#define SCR(jr, f, i) scr[(size_t)((jr) * 5 + (f)) * n + (i)]   // fields: a b talb tsrc trans
  const int ict = w.ict[c];                                  // first cloudy layer (0-based), nlev if none
  const int b = T.meta->band_of_g_lw[gg];
  const double emission = w.emission[(size_t)c * SD::NG + gg], albedo = w.lw_albedo[(size_t)c * SD::NG + gg];
  const double fd_surf_clear = carry[SD::NG + gg];
  const double fd_ict = ict < nlev ? carry[gg] : fd_surf_clear;   // clear-sky flux_dn at the cloud-top half-level
  double* s_up_c = sums, *s_up = sums + nl1, *s_dn = sums + 2 * nl1, *s_dv = sums + 3 * nl1;

  const bool keep_trans = cfg.do_lw_derivatives && out.lw_derivatives;
  // ---- upward sweep (radiation_tripleclouds_lw.F90:215-420) ----
  double ta[3] = {albedo, albedo, albedo}, ts[3];
#pragma unroll
  for (int jr = 0; jr < 3; ++jr) ts[jr] = S.reg[(nlev - 1) * 3 + jr] * emission;
  double fuc = emission + albedo * fd_surf_clear;           // clear-sky flux_up at the surface
  double fu = ts[0] + ta[0] * fd_ict;                       // all-sky flux_up at cloud top (valid as is if no cloud)
  const double fuc_surf = fuc;
  {
    double* dst[2] = {s_up_c, s_up};
    const BandOut bo[1] = {{cfg.do_save_spectral_flux ? out.lw_up_band : nullptr, out.ld, 1, -1, 1.0, 0.0, nullptr, 0}};
    int slot = 0, lfirst = nlev;
    if (act) { tile[slot * SD::RS + g] = fuc; tile[(TC_LCH_LW + slot) * SD::RS + g] = fu; }
    ++slot;
    double pb = act ? pl[(size_t)nlev * SD::NG + g] : 0.0;
    double od_n = act ? od[(size_t)(nlev - 1) * SD::NG + g] : 0.0, pt_n = act ? pl[(size_t)(nlev - 1) * SD::NG + g] : 0.0;   // software pipeline
    // (the band's cloud properties of the next layer up are requested with them when that layer is cloudy)
    double cl_n[3] = {0.0, 0.0, 0.0};
    if (act && !S.clear[nlev]) { const double* q = cl + (size_t)(nlev - 1) * 3 * SD::NB; cl_n[0] = q[b]; cl_n[1] = q[SD::NB + b]; cl_n[2] = q[2 * SD::NB + b]; }
    for (int l = nlev - 1; l >= 0; --l) {
      const int jl = l + 1;
      if (act) {
        const size_t i = (size_t)l * SD::NG + g;
        const double odg = od_n, pt = pt_n;
        const double clb0 = cl_n[0], clb1 = cl_n[1], clb2 = cl_n[2];
        if (SD::NG < 128 && l > TC_PF_DIST) {   // (measured: helps the 64-term ecCKD spectra by 4 %, costs the 140 RRTMG g-points 4 %)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(od + i - (size_t)TC_PF_DIST * SD::NG));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pl + i - (size_t)TC_PF_DIST * SD::NG));
        }
        if (l > 0) {
          od_n = od[i - SD::NG]; pt_n = pl[i - SD::NG];
          if (l - 1 >= ict && !S.clear[jl - 1]) { const double* q = cl + (size_t)(l - 1) * 3 * SD::NB; cl_n[0] = q[b]; cl_n[1] = q[SD::NB + b]; cl_n[2] = q[2 * SD::NB + b]; }
        }
        const LwLayer Lc = lw_no_scat(odg, pt, pb);
        fuc = Lc.trans * fuc + Lc.source_up;
        if (l >= ict) {
          const bool cloudy_layer = !S.clear[jl];
          double below[3] = {0.0, 0.0, 0.0}, sbelow[3] = {0.0, 0.0, 0.0};
#pragma unroll
          for (int jr = 0; jr < 3; ++jr) {
            if (jr > 0 && !cloudy_layer) continue;   // (nothing reads the slots of regions 2 and 3 in a clear layer)
            LwLayer L = Lc;
            if (jr > 0) {   // radiation_tripleclouds_lw.F90:247-300
              const double od_cloud_new = clb0 * S.ods[l * 3 + jr];
              const double od_total = odg + od_cloud_new;
              if (cfg.do_lw_cloud_scattering) {
                double ssa_total = 0.0, g_total = 0.0;
                if (od_total > 0.0) ssa_total = clb1 * od_cloud_new / od_total;
                if (ssa_total > 0.0 && od_total > 0.0) g_total = clb2 * clb1 * od_cloud_new / (ssa_total * od_total);
                L = lw_ref_trans(od_total, ssa_total, g_total, pt, pb);
              } else {
                L = lw_no_scat(od_total, pt, pb);
              }
            }
            double su = L.source_up, sd = L.source_dn;
            if (cloudy_layer) { su = S.reg[l * 3 + jr] * su; sd = S.reg[l * 3 + jr] * sd; }
            const double id = 1.0 / (1.0 - ta[jr] * L.ref);
            SCR(jr, 0, i) = L.trans * id;
            SCR(jr, 1, i) = (L.ref * ts[jr] + sd) * id;
            SCR(jr, 2, i) = ta[jr]; SCR(jr, 3, i) = ts[jr]; SCR(jr, 4, i) = L.trans;
            below[jr] = L.ref + L.trans * L.trans * ta[jr] * id;
            sbelow[jr] = su + L.trans * (ts[jr] + ta[jr] * sd) * id;
          }
          if (S.clear[jl] && S.clear[jl - 1]) {
#pragma unroll
            for (int jr = 0; jr < 3; ++jr) { ta[jr] = below[jr]; ts[jr] = sbelow[jr]; }
          } else {
            const double* U = S.U + l * 9; const double* V = S.V + l * 9;
#pragma unroll
            for (int j1 = 0; j1 < 3; ++j1) {
              double a = 0.0, sacc = 0.0;
#pragma unroll
              for (int j2 = 0; j2 < 3; ++j2) { sacc = sacc + U[j1 * 3 + j2] * sbelow[j2]; a = a + below[j2] * V[j2 * 3 + j1]; }
              ts[j1] = sacc; ta[j1] = a;
            }
          }
          if (l == ict) fu = ts[0] + ta[0] * fd_ict;   // flux_up(:,1) at cloud top
        } else {
          fu = Lc.trans * fu + Lc.source_up;            // above cloud top
          if (keep_trans) SCR(0, 4, i) = Lc.trans;       // (the LW derivatives multiply by it again)
        }
        pb = pt;
        tile[slot * SD::RS + g] = fuc; tile[(TC_LCH_LW + slot) * SD::RS + g] = l <= ict ? fu : 0.0;
      }
      ++slot;
      if (slot == TC_LCH_LW || l == 0) {
        if (bo[0].dst) flush_bands(tile, SD::RS, TC_LCH_LW, slot, bo, 1, lfirst, -1, c, SD::NB, T.meta->lw);
        flush_tile(tile, SD::RS, SD::NG, 2, slot, dst, lfirst, -1, TC_LCH_LW); lfirst -= slot; slot = 0;
      }
    }
  }
  const double fu_toa = fu, fuc_toa = fuc;
  // ---- downward sweep from cloud top (radiation_tripleclouds_lw.F90:470-540) ----
  double fdn[3], fup[3] = {fu, 0.0, 0.0};
#pragma unroll
  for (int jr = 0; jr < 3; ++jr) fdn[jr] = S.V[ict * 9 + jr * 3 + 0] * fd_ict;
  if (ict < nlev) {
    double* dst[2] = {s_up, s_dn};
    const BandOut bo[2] = {{cfg.do_save_spectral_flux ? out.lw_up_band : nullptr, out.ld, 0, -1, 1.0, 0.0, nullptr, 0},
                           {cfg.do_save_spectral_flux ? out.lw_dn_band : nullptr, out.ld, 1, -1, 1.0, 0.0, nullptr, 0}};
    int slot = 0, lfirst = ict + 1;
    // software pipeline: the four values per region of the next layer are requested one layer ahead (regions 2 and 3 only if that
    // layer is cloudy: their slots are not written otherwise)
    double nx[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) nx[k] = (act && (k < 4 || !S.clear[ict + 1])) ? SCR(k / 4, k % 4, (size_t)ict * SD::NG + g) : 0.0;
    for (int l = ict; l < nlev; ++l) {
      const int jl = l + 1;
      if (act) {
        const size_t i = (size_t)l * SD::NG + g;
        double cu[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) cu[k] = nx[k];
        if (l + 1 < nlev) {
          const bool cloudy_next = !S.clear[jl + 1];
#pragma unroll
          for (int k = 0; k < 12; ++k) if (k < 4 || cloudy_next) nx[k] = SCR(k / 4, k % 4, i + SD::NG);
        }
        if (l + TC_PF_DIST < nlev) {   // (region 1's four rows a few layers further down: into L2 now)
#pragma unroll
          for (int k = 0; k < 4; ++k) asm volatile("prefetch.global.L2 [%0];" ::"l"(&SCR(0, k, i + (size_t)TC_PF_DIST * SD::NG)));
        }
#pragma unroll
        for (int jr = 0; jr < 3; ++jr) {
          if (jr > 0 && S.clear[jl]) { fdn[jr] = 0.0; fup[jr] = 0.0; continue; }
          fdn[jr] = cu[jr * 4] * fdn[jr] + cu[jr * 4 + 1];
          fup[jr] = cu[jr * 4 + 3] + fdn[jr] * cu[jr * 4 + 2];
        }
        if (!(S.clear[jl] && S.clear[jl + 1])) mat3_x_vec(S.V + jl * 9, fdn);
        tile[slot * SD::RS + g] = fup[0] + fup[1] + fup[2];
        tile[(TC_LCH_LW + slot) * SD::RS + g] = fdn[0] + fdn[1] + fdn[2];
      }
      ++slot;
      if (slot == TC_LCH_LW || l == nlev - 1) {
        if (bo[0].dst || bo[1].dst) flush_bands(tile, SD::RS, TC_LCH_LW, slot, bo, 2, lfirst, 1, c, SD::NB, T.meta->lw);
        flush_tile(tile, SD::RS, SD::NG, 2, slot, dst, lfirst, 1, TC_LCH_LW); lfirst += slot; slot = 0;
      }
    }
  }
  const double dn_surf_g = fdn[0] + fdn[1] + fdn[2];
  // ---- derivatives: calc_lw_derivatives_region (weights = last flux_up, i.e. the surface one unless the column is cloud free) ----
  const bool want_dv = cfg.do_lw_derivatives && out.lw_derivatives;
  if (want_dv) {
    // (Tripleclouds: in a cloud-free column the reference's flux_up still holds the top-of-atmosphere values at this point and the
    // oracle follows it; the Homogeneous solver weights with the surface flux in every column, calc_lw_derivatives_ica)
    const double fus = (cfg.is_homogeneous && ict >= nlev) ? fuc_surf : fup[0] + fup[1] + fup[2];
    // block sum of fus (same order of partial sums as flush_tile is not required: only a normalisation)
    __syncthreads();
    if (act) tile[g] = fus;
    __syncthreads();
    if (g == 0) { double s = 0.0; for (int k = 0; k < SD::NG; ++k) s = s + tile[k]; red[0] = s; }
    __syncthreads();
    double d[3] = {fus / red[0], 0.0, 0.0};
    double* dst[1] = {s_dv};
    int slot = 0, lfirst = nlev - 1;
    // Transmittances of the three regions in layer l: exp(-D od) of the clear-sky region kept by the upward sweep; regions 2 and 3 exist
    // in cloudy layers only and transmit everything elsewhere (x * 1.0 is exact).  They do not depend on d: next layer's values first.
    auto trans = [&](int l, double* t) {
      const size_t i = (size_t)l * SD::NG + g;
      t[0] = SCR(0, 4, i);
      const bool cloudy = l >= ict && !S.clear[l + 1];
      t[1] = cloudy ? SCR(1, 4, i) : 1.0; t[2] = cloudy ? SCR(2, 4, i) : 1.0;
    };
    double tn[3] = {0.0, 1.0, 1.0};
    if (act) trans(nlev - 1, tn);
    for (int l = nlev - 1; l >= 0; --l) {
      const int jl = l + 1;
      if (act) {
        const double t[3] = {tn[0], tn[1], tn[2]};
        if (l > 0) trans(l - 1, tn);
        if (l > TC_PF_DIST) asm volatile("prefetch.global.L2 [%0];" ::"l"(&SCR(0, 4, (size_t)(l - TC_PF_DIST) * SD::NG + g)));
        if (jl >= ict) mat3_x_vec(S.U + jl * 9, d);   // u_matrix(:,:,jlev+1); the identity between two layers above cloud top (1*d + 0 + 0 is exact)
        d[0] = d[0] * t[0]; d[1] = d[1] * t[1]; d[2] = d[2] * t[2];
        tile[slot * SD::RS + g] = d[0] + d[1] + d[2];
      }
      ++slot;
      if (slot == TC_LCH_LW || l == 0) { flush_tile(tile, SD::RS, SD::NG, 1, slot, dst, lfirst, -1, TC_LCH_LW); lfirst -= slot; slot = 0; }
    }
  }
#undef SCR
  __syncthreads();
  // ---- outputs ----
  for (int l = g; l < nl1; l += SD::THREADS) {
    const size_t o = (size_t)l * out.ld + c;
    const double dnc = gsum[l];
    if (out.lw_up_clear) out.lw_up_clear[o] = s_up_c[l];
    if (out.lw_dn_clear) out.lw_dn_clear[o] = dnc;
    if (out.lw_up) out.lw_up[o] = s_up[l];
    if (out.lw_dn) out.lw_dn[o] = l <= ict ? dnc : s_dn[l];
    if (want_dv) out.lw_derivatives[o] = l == nlev ? 1.0 : s_dv[l];
  }
  if (g == 0 && out.cloud_cover_lw && !cfg.is_homogeneous) out.cloud_cover_lw[c] = w.tc_cc[c];
  if (act) {
    const size_t i = (size_t)c * SD::NG + g;
    if (out.lw_dn_surf_clear_g) out.lw_dn_surf_clear_g[i] = fd_surf_clear;
    if (out.lw_up_toa_clear_g) out.lw_up_toa_clear_g[i] = fuc_toa;
    if (out.lw_dn_surf_g) out.lw_dn_surf_g[i] = dn_surf_g;
    if (out.lw_up_toa_g) out.lw_up_toa_g[i] = fu_toa;
  }
  (void)fuc_surf;
  lw_surface_canopy<SD>(T, cfg, out, c, g, act, tile, dn_surf_g);
}

// =========================================================================================================
size_t tc_scratch_doubles_lw(int nlev, int ng) { return (size_t)TC_LW_ARRAYS * nlev * ng; }
size_t tc_scratch_doubles_sw(int nlev, int ng) { return (size_t)TC_SW_ARRAYS * nlev * ng; }

int launch_tc_prep(const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st) {
  tc_prep_kernel<<<nc, TC_PREP_THREADS, 0, st>>>(cfg, in, w, nc, nlev);
  return 1;
}
template <class SD>
static int launch_tc_sw_t(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, const Work& w, int nc, int nlev, cudaStream_t st) {
  const size_t sm = sizeof(double) * (6 * (nlev + 1) + 6 * TC_LCH * SD::RS + 2 * SD::NB) + tc_shared_bytes(nlev);
  cudaFuncSetAttribute(tc_sw_kernel<SD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  tc_sw_kernel<SD><<<nc, SD::THREADS, sm, st>>>(T, cfg, in, out, w, nlev);
  return 1;
}
template <class SD>
static int launch_tc_lw_t(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, const Work& w, int nc, int nlev, cudaStream_t st) {
  const size_t sm = sizeof(double) * (4 * (nlev + 1) + 2 * TC_LCH_LW * SD::RS + 8) + tc_shared_bytes(nlev);
  cudaFuncSetAttribute(tc_lw_kernel<SD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  tc_lw_kernel<SD><<<nc, SD::THREADS, sm, st>>>(T, cfg, in, out, w, nlev);
  return 1;
}
int launch_tc_sw(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, const Work& w, int nc, int nlev, cudaStream_t st) {
  switch (cfg.ng_sw) {
    case NG_SW: return launch_tc_sw_t<SwRrtmg>(T, cfg, in, out, w, nc, nlev, st);
    case 32: return launch_tc_sw_t<Ckd32>(T, cfg, in, out, w, nc, nlev, st);
    case 64: return launch_tc_sw_t<Ckd64>(T, cfg, in, out, w, nc, nlev, st);
    case 96: return launch_tc_sw_t<Ckd96>(T, cfg, in, out, w, nc, nlev, st);
  }
  return -1;
}
int launch_tc_lw(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, const Work& w, int nc, int nlev, cudaStream_t st) {
  switch (cfg.ng_lw) {
    case NG_LW: return launch_tc_lw_t<LwRrtmg>(T, cfg, in, out, w, nc, nlev, st);
    case 32: return launch_tc_lw_t<Ckd32>(T, cfg, in, out, w, nc, nlev, st);
    case 64: return launch_tc_lw_t<Ckd64>(T, cfg, in, out, w, nc, nlev, st);
    case 96: return launch_tc_lw_t<Ckd96>(T, cfg, in, out, w, nc, nlev, st);
  }
  return -1;
}

}  // namespace ecb
