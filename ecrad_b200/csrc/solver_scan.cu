// solver_scan.cu -- McICA and Cloudless solvers (shortwave: radiation_mcica_sw.F90:41-408, radiation_cloudless_sw.F90;
// longwave: radiation_mcica_lw.F90:39-419, radiation_cloudless_lw.F90) with the adding method done as warp scans
// (scan_core.cuh).
//
// One CTA per column, SCAN_NW warps; a warp owns one (column, g-point) at a time and loops over the g-points g = warp,
// warp + SCAN_NW, ...; its 32 lanes hold LPL consecutive layers each.  Per g-point:
//   1. every lane evaluates the two-stream solution of its own layers (calc_ref_trans_sw / calc_no_scattering_transmittance_lw):
//      LPL independent evaluations per lane = the instruction-level parallelism the fp64 pipe needs;
//   2. clear-sky sub-column: upward scans (albedo, then direct albedo / source), downward scan (direct beam + diffuse flux),
//      fluxes at the lane's half-levels added to the warp's g-point sums in shared memory;
//   3. cloudy sub-column (McICA, columns with cloud): the cloudy layers of the column are dealt out to the lanes (one layer per
//      lane, so a warp spends one or two two-stream evaluations on ~40 cloudy layers instead of LPL at a quarter utilisation),
//      the results go back to the lanes that own those layers through a small shared-memory exchange, then the same scans.
// The layer solutions live in registers from the upward to the downward pass, so nothing of the adding method is written to
// memory: the kernels read the gas optical properties (laid out [column][g][layer] by the gas-optics kernels, one contiguous
// row per warp) and the generator's code words, and write the flux profiles.  The g-point sums of the SCAN_NW warps are added in
// a fixed order (bit-reproducible runs).
#include "scan_core.cuh"
#include "solver_common.cuh"

namespace ecb {

enum { SCAN_NW = 8, SCAN_THREADS = SCAN_NW * 32, SCAN_XCH = 64 };

// level bookkeeping of a lane: layers l = L0 + j (j < LPL, L0 = lane*LPL), half-level below layer l is l + 1; lane 0 also owns
// the top-of-atmosphere half-level 0.  Per-warp sums: [q][LPL + 1][32], slot LPL = top of atmosphere (lane 0).
template <int LPL>
__device__ __forceinline__ double* sum_slot(double* sums, int q, int j, int lane) { return sums + ((q * (LPL + 1) + j) * 32 + lane); }

// ---------------------------------------------------------------------------------------------------------
// shortwave
// ---------------------------------------------------------------------------------------------------------
// One sub-column: L[j] = two-stream solution of the lane's layers, As / Ds = surface albedo to diffuse / direct (x cos_sza)
// radiation, inc = incoming flux at the top of atmosphere (per unit area normal to the beam).  Adds dir, diffuse-down and up fluxes
// at the lane's half-levels to sums[q0 .. q0+2]; returns the per-g-point surface and top-of-atmosphere values.
template <int LPL>
__device__ __forceinline__ void sw_subcolumn(const SwLayer (&L)[LPL], double As, double Ds, double inc, int lane, int nlev, double* sums, int q0,
                                             double& dir_surf, double& dn_surf, double& up_toa) {
  // ---- albedo of everything below each half-level: radiation_adding_ica_sw.F90:97-121 ----
  Mob own = mob_id();
#pragma unroll
  for (int j = 0; j < LPL; ++j) own = mob_push_below(own, L[j].ref, L[j].trans);
  const Mob below = suffix_exclusive(own, lane, mob_id(), [](const Mob& x, const Mob& y) { return mob_mul(x, y); });
  double A = mob_apply(below, As);   // albedo below the lane's bottom layer
  double Ab[LPL], inv[LPL];
  Aff ownD = aff_id();
#pragma unroll
  for (int j = LPL - 1; j >= 0; --j) {
    Ab[j] = A;
    const double inv_den = 1.0 / (1.0 - A * L[j].ref);
    inv[j] = inv_den;
    // direct albedo D(l) = Rdir + (Tdir D(l+1) + Tdirdif A(l+1)) T / (1 - A R): affine in D(l+1)
    Aff f;
    f.al = L[j].ref_dir + L[j].trans_dir_diff * A * L[j].trans * inv_den;
    f.be = L[j].trans_dir_dir * L[j].trans * inv_den;
    ownD = aff_mul(f, ownD);
    A = L[j].ref + L[j].trans * L[j].trans * A * inv_den;
  }
  const Aff belowD = suffix_exclusive(ownD, lane, aff_id(), [](const Aff& x, const Aff& y) { return aff_mul(x, y); });
  double D = belowD.al + belowD.be * Ds;
  double Db[LPL];
  Tri ownF = tri_id();
#pragma unroll
  for (int j = LPL - 1; j >= 0; --j) {
    Db[j] = D;
    D = L[j].ref_dir + (L[j].trans_dir_dir * D + L[j].trans_dir_diff * Ab[j]) * L[j].trans * inv[j];
  }
  // ---- fluxes downwards: :85-88, :134-146 ----
#pragma unroll
  for (int j = 0; j < LPL; ++j) {
    Tri m;
    m.t = L[j].trans_dir_dir;
    m.b = (L[j].trans_dir_dir * Db[j] * L[j].ref + L[j].trans_dir_diff) * inv[j];
    m.a = L[j].trans * inv[j];
    ownF = tri_mul(m, ownF);
  }
  const Tri above = prefix_exclusive(ownF, lane, tri_id(), [](const Tri& x, const Tri& y) { return tri_mul(x, y); });
  double dir = above.t * inc, dn = above.b * inc;
  if (lane == 0) {   // top of atmosphere: D is now the direct albedo of the whole atmosphere + surface
    up_toa = inc * D;
    *sum_slot<LPL>(sums, q0, LPL, 0) += inc;
    *sum_slot<LPL>(sums, q0 + 2, LPL, 0) += up_toa;
  }
  const int L0 = lane * LPL;
#pragma unroll
  for (int j = 0; j < LPL; ++j) {
    dn = L[j].trans * inv[j] * dn + (L[j].trans_dir_dir * Db[j] * L[j].ref + L[j].trans_dir_diff) * inv[j] * dir;
    dir = L[j].trans_dir_dir * dir;
    const double up = dir * Db[j] + dn * Ab[j];
    if (L0 + j < nlev) {
      *sum_slot<LPL>(sums, q0, j, lane) += dir;
      *sum_slot<LPL>(sums, q0 + 1, j, lane) += dn;
      *sum_slot<LPL>(sums, q0 + 2, j, lane) += up;
      if (L0 + j == nlev - 1) { dir_surf = dir; dn_surf = dn; }
    }
  }
}

__device__ __forceinline__ SwLayer sw_identity_layer() {
  SwLayer r; r.ref = 0.0; r.trans = 1.0; r.ref_dir = 0.0; r.trans_dir_diff = 0.0; r.trans_dir_dir = 1.0; return r;
}

template <class SD, int LPL>
__global__ void __launch_bounds__(SCAN_THREADS, 2)
sw_scan_kernel(DevTables T, DevCfg cfg, DevIn in, DevOut out, Work w, int nlev, int ls, int nlevp, int cloudless, int aer, int delta) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int c = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nl1 = nlev + 1;
  const double mu0 = in.cos_sza[c];
  if (!(mu0 > 0.0)) { sw_night_column<SD>(cfg, out, c, tid, tid < SD::NG, nl1, SCAN_THREADS); return; }
  // ---- shared memory ----
  double* sums = reinterpret_cast<double*>(smem_raw);            // [SCAN_NW][6][LPL+1][32]
  double* xch = sums + SCAN_NW * 6 * (LPL + 1) * 32;             // [SCAN_NW][5][SCAN_XCH]
  double* gt = xch + SCAN_NW * 5 * SCAN_XCH;                     // [6][NG] per-g-point surface / TOA values
  double* bandv = gt + 6 * SD::NG;                               // [2][NB] band albedos
  double* fsds = bandv + 2 * SD::NB;                             // [nlev]
  double* tile = fsds + nlev;                                    // [4*RS + 2*NB] surface spectral fluxes
  short* clidx = reinterpret_cast<short*>(tile + 4 * SD::RS + 2 * SD::NB);   // [nlev] cloudy layers, top-down
  short* clrank = clidx + nlev;                                  // [nlev] position of a layer in clidx, -1: cloud-free
  __shared__ int s_ncl;
  const double tcc = cfg.solver_sw == 2 ? w.tcc[c] : 0.0;
  const bool cloudy = tcc > 0.0;
  for (int i = tid; i < SCAN_NW * 6 * (LPL + 1) * 32; i += SCAN_THREADS) sums[i] = 0.0;
  for (int l = tid; l < nlev; l += SCAN_THREADS) {
    const bool cl = cloudy && LD_IN(in.frac, c, l) >= cfg.cloud_fraction_threshold;
    fsds[l] = cl ? LD_IN(in.fsd, c, l) : 0.0;
    clrank[l] = cl ? 0 : -1;
  }
  // get_albedos, radiation_single_level.F90:216-365 (weighted-interval mapping to bands)
  if (tid < SD::NB) {
    double bd = 0.0, bdir = 0.0;
    for (int ja = 0; ja < cfg.n_albedo_sw; ++ja) {
      const double wgt = T.sw_albedo_weights[tid * cfg.n_albedo_sw + ja];
      if (wgt != 0.0) {
        bd = bd + wgt * LD_IN(in.sw_albedo, c, ja);
        if (in.sw_albedo_direct) bdir = bdir + wgt * LD_IN(in.sw_albedo_direct, c, ja);
      }
    }
    bandv[tid] = bd; bandv[SD::NB + tid] = in.sw_albedo_direct ? bdir : bd;
  }
  __syncthreads();
  if (tid == 0) {
    int n = 0;
    for (int l = 0; l < nlev; ++l) if (clrank[l] == 0) { clrank[l] = (short)n; clidx[n] = (short)l; ++n; }
    s_ncl = n;
  }
  __syncthreads();
  const int ncl = s_ncl;
  const CloudMeta& C = *T.cloud;
  double* wsums = sums + warp * 6 * (LPL + 1) * 32;
  double* wx = xch + warp * 5 * SCAN_XCH;
  const int L0 = lane * LPL;
  const double* cl = w.cl_sw + (size_t)c * nlev * 3 * SD::NB;

  for (int g = warp; g < SD::NG; g += SCAN_NW) {
    const int b = T.meta->band_of_g_sw[g];
    const double* odp = w.od_sw + ((size_t)c * SD::NG + g) * ls;
    const double* ssap = w.ssa_sw + ((size_t)c * SD::NG + g) * ls;
    const double* ggp = aer ? w.g_sw + ((size_t)c * SD::NG + g) * ls : nullptr;
    const double inc = w.incoming[(size_t)c * SD::NG + g];
    // ---- two-stream solutions of the lane's layers, clear sky ----
    SwLayer L[LPL];
#pragma unroll
    for (int j = 0; j < LPL; ++j) {
      const int l = L0 + j;
      if (l < nlev) {
        double od = odp[l], ssa = ssap[l], gg = aer ? ggp[l] : 0.0;
        if (delta && aer) sw_delta_eddington(od, ssa, gg);   // radiation_mcica_sw.F90:165-180 (a no-op when g = 0, i.e. without aerosols)
        L[j] = cloudless ? sw_ref_trans_cloudless(mu0, od, ssa, gg) : sw_ref_trans(mu0, od, ssa, gg);
      } else {
        L[j] = sw_identity_layer();
      }
    }
    const double As = bandv[b], Ds = mu0 * bandv[SD::NB + b];   // radiation_adding_ica_sw.F90:95-96
    double dir_c = 0.0, dn_c = 0.0, toa_c = 0.0;
    sw_subcolumn<LPL>(L, As, Ds, inc, lane, nlev, wsums, 0, dir_c, dn_c, toa_c);
    if (L0 <= nlev - 1 && nlev - 1 < L0 + LPL) { gt[g] = dir_c; gt[SD::NG + g] = dn_c; }
    if (lane == 0) gt[2 * SD::NG + g] = toa_c;
    if (cloudy) {
      // ---- cloudy layers: one layer per lane (radiation_mcica_sw.F90:249-278), results handed to the owning lanes ----
      const uint32_t* codep = w.code_sw + ((size_t)c * SD::NG + g) * nlevp;
      for (int k0 = 0; k0 < ncl; k0 += SCAN_XCH) {
        const int kend = imin(ncl, k0 + SCAN_XCH);
        for (int k = k0 + lane; k < kend; k += 32) {
          const int l = clidx[k];
          const double od = odp[l], ssa = ssap[l], gg = aer ? ggp[l] : 0.0;
          double odt, ssat, gtot;
          sw_cloudy_props<SD>(C, T.pdf_val, codep[l], fsds[l], cl + (size_t)l * 3 * SD::NB, b, od, ssa, gg, odt, ssat, gtot);
          if (delta) sw_delta_eddington(odt, ssat, gtot);   // :274-278
          const SwLayer La = sw_ref_trans(mu0, odt, ssat, gtot);
          const int o = k - k0;
          wx[o] = La.ref; wx[SCAN_XCH + o] = La.trans; wx[2 * SCAN_XCH + o] = La.ref_dir; wx[3 * SCAN_XCH + o] = La.trans_dir_diff;
          wx[4 * SCAN_XCH + o] = La.trans_dir_dir;
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < LPL; ++j) {
          const int l = L0 + j;
          if (l < nlev) {
            const int r = clrank[l];
            if (r >= k0 && r < kend) {
              const int o = r - k0;
              L[j].ref = wx[o]; L[j].trans = wx[SCAN_XCH + o]; L[j].ref_dir = wx[2 * SCAN_XCH + o]; L[j].trans_dir_diff = wx[3 * SCAN_XCH + o];
              L[j].trans_dir_dir = wx[4 * SCAN_XCH + o];
            }
          }
        }
        __syncwarp();
      }
      double dir_a = 0.0, dn_a = 0.0, toa_a = 0.0;
      sw_subcolumn<LPL>(L, As, Ds, inc, lane, nlev, wsums, 3, dir_a, dn_a, toa_a);
      if (L0 <= nlev - 1 && nlev - 1 < L0 + LPL) { gt[3 * SD::NG + g] = dir_a; gt[4 * SD::NG + g] = dn_a; }
      if (lane == 0) gt[5 * SD::NG + g] = toa_a;
    }
  }
  __syncthreads();
  // ---- g-point sums of the warps (fixed order), blending, flux_type outputs: radiation_mcica_sw.F90:330-378 ----
#define OUT2(p, l) ((p)[(size_t)(l) * out.ld + c])
  const double wc = tcc, w1 = 1.0 - tcc;
  for (int h = tid; h < nl1; h += SCAN_THREADS) {
    const int ln = h == 0 ? 0 : (h - 1) / LPL, j = h == 0 ? LPL : (h - 1) - ln * LPL;
    double s[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      double acc = 0.0;
      if (q < 3 || cloudy)
        for (int ww = 0; ww < SCAN_NW; ++ww) acc += *sum_slot<LPL>(sums + ww * 6 * (LPL + 1) * 32, q, j, ln);
      s[q] = acc;
    }
    const double dirc = s[0] * mu0, upc = s[2], dnc = s[1] + dirc;
    if (out.sw_up_clear) OUT2(out.sw_up_clear, h) = upc;
    if (out.sw_dn_clear) OUT2(out.sw_dn_clear, h) = dnc;
    if (out.sw_dn_direct_clear) OUT2(out.sw_dn_direct_clear, h) = dirc;
    double up = upc, dn = dnc, dir = dirc;
    if (cloudy) {
      const double dira = s[3] * mu0;
      up = wc * s[5] + w1 * upc;
      dn = wc * (s[4] + dira) + w1 * dnc;
      dir = wc * dira + w1 * dirc;
    }
    if (out.sw_up) OUT2(out.sw_up, h) = up;
    if (out.sw_dn) OUT2(out.sw_dn, h) = dn;
    if (out.sw_dn_direct) OUT2(out.sw_dn_direct, h) = dir;
  }
#undef OUT2
  if (tid == 0 && out.cloud_cover_sw && cfg.solver_sw == 2) out.cloud_cover_sw[c] = tcc;
  // per-g-point surface / top-of-atmosphere fluxes
  const bool act = tid < SD::NG;
  double dir_as = 0.0, dif_a = 0.0, dir_cs = 0.0, dif_c = 0.0;
  if (act) {
    const int g = tid;
    dif_c = gt[SD::NG + g]; dir_cs = gt[g] * mu0;
    const double toa_c = gt[2 * SD::NG + g];
    dif_a = dif_c; dir_as = dir_cs;
    double toa_a = toa_c;
    if (cloudy) {
      dif_a = wc * gt[4 * SD::NG + g] + w1 * dif_c;
      dir_as = wc * (gt[3 * SD::NG + g] * mu0) + w1 * dir_cs;
      toa_a = wc * gt[5 * SD::NG + g] + w1 * toa_c;
    }
    const size_t i = (size_t)c * SD::NG + g;
    if (out.sw_dn_diffuse_surf_clear_g) out.sw_dn_diffuse_surf_clear_g[i] = dif_c;
    if (out.sw_dn_direct_surf_clear_g) out.sw_dn_direct_surf_clear_g[i] = dir_cs;
    if (out.sw_up_toa_clear_g) out.sw_up_toa_clear_g[i] = toa_c;
    if (out.sw_dn_diffuse_surf_g) out.sw_dn_diffuse_surf_g[i] = dif_a;
    if (out.sw_dn_direct_surf_g) out.sw_dn_direct_surf_g[i] = dir_as;
    if (out.sw_up_toa_g) out.sw_up_toa_g[i] = toa_a;
  }
  sw_surface_spectral<SD>(T, cfg, out, c, tid, act, tile, SD::RS, dir_as, dif_a, dir_cs, dif_c);
}

// ---------------------------------------------------------------------------------------------------------
// longwave
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ LwLayer lw_identity_layer() { LwLayer r; r.ref = 0.0; r.trans = 1.0; r.source_up = 0.0; r.source_dn = 0.0; return r; }

// Clear-sky sub-column (calc_fluxes_no_scattering_lw, radiation_adding_ica_lw.F90:266-330; derivative products of
// calc_lw_derivatives_ica, radiation_lw_derivatives.F90:43-91).  sums q0: down, q0+1: up, q0+2: flux_up(surface) x prod(trans).
template <int LPL>
__device__ __forceinline__ void lw_clear_subcolumn(const LwLayer (&L)[LPL], double emission, double albedo, int lane, int nlev, double* sums,
                                                   double& dn_surf, double& up_toa) {
  Aff ownD = aff_id();
#pragma unroll
  for (int j = 0; j < LPL; ++j) { Aff f; f.al = L[j].source_dn; f.be = L[j].trans; ownD = aff_mul(f, ownD); }
  const Aff above = prefix_exclusive(ownD, lane, aff_id(), [](const Aff& x, const Aff& y) { return aff_mul(x, y); });
  double fd = above.al;   // flux_dn at the top of atmosphere is zero
  double fdv[LPL];
#pragma unroll
  for (int j = 0; j < LPL; ++j) { fd = L[j].trans * fd + L[j].source_dn; fdv[j] = fd; }
  dn_surf = __shfl_sync(ECB_FULL, fd, 31);   // layers beyond nlev are identities: lane 31 ends with the surface value
  const double fu_surf = emission + albedo * dn_surf;
  Aff ownU = aff_id();
#pragma unroll
  for (int j = 0; j < LPL; ++j) { Aff f; f.al = L[j].source_up; f.be = L[j].trans; ownU = aff_mul(ownU, f); }
  const Aff below = suffix_exclusive(ownU, lane, aff_id(), [](const Aff& x, const Aff& y) { return aff_mul(x, y); });
  double fu = below.al + below.be * fu_surf, prod = below.be * fu_surf;
  const int L0 = lane * LPL;
#pragma unroll
  for (int j = LPL - 1; j >= 0; --j) {
    if (L0 + j < nlev) {   // half-level below layer j
      *sum_slot<LPL>(sums, 0, j, lane) += fdv[j];
      *sum_slot<LPL>(sums, 1, j, lane) += fu;
      *sum_slot<LPL>(sums, 2, j, lane) += prod;
    }
    fu = L[j].trans * fu + L[j].source_up;
    prod = prod * L[j].trans;
  }
  if (lane == 0) {
    up_toa = fu;
    *sum_slot<LPL>(sums, 1, LPL, 0) += fu;
    *sum_slot<LPL>(sums, 2, LPL, 0) += prod;
  }
}

// Cloudy sub-column with or without scattering (fast_adding_ica_lw, radiation_adding_ica_lw.F90:137-263, written for all layers:
// where the reference takes its cloud-free shortcut above cloud top the layer reflectance is zero and the general step reduces to
// the same operations; with do_lw_aerosol_scattering it is also the clear-sky sub-column, adding_ica_lw :24-130).
// sums q0: down, q0+1: up, q0+2: flux_up(surface) x prod(trans).
template <int LPL>
__device__ __forceinline__ void lw_cloudy_subcolumn(const LwLayer (&L)[LPL], double emission, double albedo, int lane, int nlev, double* sums, int q0,
                                                    double& dn_surf, double& up_toa) {
  Mob own = mob_id();
#pragma unroll
  for (int j = 0; j < LPL; ++j) own = mob_push_below(own, L[j].ref, L[j].trans);
  const Mob below = suffix_exclusive(own, lane, mob_id(), [](const Mob& x, const Mob& y) { return mob_mul(x, y); });
  double A = mob_apply(below, albedo);
  double Ab[LPL], inv[LPL];
  Aff ownS = aff_id();
#pragma unroll
  for (int j = LPL - 1; j >= 0; --j) {
    Ab[j] = A;
    const double inv_den = 1.0 / (1.0 - A * L[j].ref);
    inv[j] = inv_den;
    // S(l) = S_up + T (S(l+1) + A(l+1) S_dn) / (1 - A R)
    Aff f;
    f.al = L[j].source_up + L[j].trans * A * L[j].source_dn * inv_den;
    f.be = L[j].trans * inv_den;
    ownS = aff_mul(f, ownS);
    A = L[j].ref + L[j].trans * L[j].trans * A * inv_den;
  }
  const Aff belowS = suffix_exclusive(ownS, lane, aff_id(), [](const Aff& x, const Aff& y) { return aff_mul(x, y); });
  double S = belowS.al + belowS.be * emission;
  double Sb[LPL];
#pragma unroll
  for (int j = LPL - 1; j >= 0; --j) {
    Sb[j] = S;
    S = L[j].source_up + L[j].trans * (S + Ab[j] * L[j].source_dn) * inv[j];
  }
  // downward: flux_dn(l+1) = (T flux_dn(l) + R S(l+1) + S_dn) / (1 - A R)
  Aff ownD = aff_id();
#pragma unroll
  for (int j = 0; j < LPL; ++j) { Aff f; f.al = (L[j].ref * Sb[j] + L[j].source_dn) * inv[j]; f.be = L[j].trans * inv[j]; ownD = aff_mul(f, ownD); }
  const Aff above = prefix_exclusive(ownD, lane, aff_id(), [](const Aff& x, const Aff& y) { return aff_mul(x, y); });
  double fd = above.al;
  double fdv[LPL];
#pragma unroll
  for (int j = 0; j < LPL; ++j) { fd = L[j].trans * inv[j] * fd + (L[j].ref * Sb[j] + L[j].source_dn) * inv[j]; fdv[j] = fd; }
  dn_surf = __shfl_sync(ECB_FULL, fd, 31);
  const double fu_surf = albedo * dn_surf + emission;
  // product of the sub-column's layer transmittances from the surface up (calc_lw_derivatives_ica)
  double pt = 1.0;
#pragma unroll
  for (int j = 0; j < LPL; ++j) pt = pt * L[j].trans;
  double pin = pt;   // suffix product over the lanes below
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const double q = __shfl_down_sync(ECB_FULL, pin, d); if (lane + d < 32) pin = pin * q; }
  double prod = __shfl_down_sync(ECB_FULL, pin, 1);
  if (lane == 31) prod = 1.0;
  prod = prod * fu_surf;
  const int L0 = lane * LPL;
#pragma unroll
  for (int j = LPL - 1; j >= 0; --j) {
    if (L0 + j < nlev) {
      *sum_slot<LPL>(sums, q0, j, lane) += fdv[j];
      *sum_slot<LPL>(sums, q0 + 1, j, lane) += Ab[j] * fdv[j] + Sb[j];
      *sum_slot<LPL>(sums, q0 + 2, j, lane) += prod;
    }
    prod = prod * L[j].trans;
  }
  if (lane == 0) {
    up_toa = S;   // flux_dn(TOA) = 0
    *sum_slot<LPL>(sums, q0 + 1, LPL, 0) += S;
    *sum_slot<LPL>(sums, q0 + 2, LPL, 0) += prod;
  }
}

template <class SD, int LPL>
__global__ void __launch_bounds__(SCAN_THREADS, 2)
lw_scan_kernel(DevTables T, DevCfg cfg, DevIn in, DevOut out, Work w, int nlev, int ls, int nlevp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int c = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nl1 = nlev + 1;
  double* sums = reinterpret_cast<double*>(smem_raw);            // [SCAN_NW][6][LPL+1][32]
  double* xch = sums + SCAN_NW * 6 * (LPL + 1) * 32;             // [SCAN_NW][4][SCAN_XCH]
  double* gt = xch + SCAN_NW * 4 * SCAN_XCH;                     // [4][NG]
  double* fsds = gt + 4 * SD::NG;                                // [nlev]
  double* tile = fsds + nlev;                                    // [NG] canopy fluxes
  short* clidx = reinterpret_cast<short*>(tile + SD::NG);        // [nlev]
  short* clrank = clidx + nlev;                                  // [nlev]
  __shared__ int s_ncl;
  const bool mcica = cfg.solver_lw == 2;
  const double tcc = mcica ? w.tcc[c] : 0.0;
  const bool cloudy = tcc > 0.0;
  for (int i = tid; i < SCAN_NW * 6 * (LPL + 1) * 32; i += SCAN_THREADS) sums[i] = 0.0;
  for (int l = tid; l < nlev; l += SCAN_THREADS) {
    const bool cl = cloudy && LD_IN(in.frac, c, l) >= cfg.cloud_fraction_threshold;
    fsds[l] = cl ? LD_IN(in.fsd, c, l) : 0.0;
    clrank[l] = cl ? 0 : -1;
  }
  __syncthreads();
  if (tid == 0) {
    int n = 0;
    for (int l = 0; l < nlev; ++l) if (clrank[l] == 0) { clrank[l] = (short)n; clidx[n] = (short)l; ++n; }
    s_ncl = n;
  }
  __syncthreads();
  const int ncl = s_ncl;
  const CloudMeta& C = *T.cloud;
  double* wsums = sums + warp * 6 * (LPL + 1) * 32;
  double* wx = xch + warp * 4 * SCAN_XCH;
  const int L0 = lane * LPL;
  const double* cl = w.cl_lw + (size_t)c * nlev * 3 * SD::NB;

  for (int g = warp; g < SD::NG; g += SCAN_NW) {
    const int b = T.meta->band_of_g_lw[g];
    const double* odp = w.od_lw + ((size_t)c * SD::NG + g) * ls;
    const double* plp = w.planck + ((size_t)c * SD::NG + g) * ls;
    const double emission = w.emission[(size_t)c * SD::NG + g], albedo = w.lw_albedo[(size_t)c * SD::NG + g];
    const bool lwscat = cfg.do_lw_aerosol_scattering != 0;
    const double* ssp = lwscat ? w.ssa_lw + ((size_t)c * SD::NG + g) * ls : nullptr;
    const double* ggp = lwscat ? w.g_lw + ((size_t)c * SD::NG + g) * ls : nullptr;
    LwLayer L[LPL];
    {
      double pt = L0 < nl1 ? plp[L0] : 0.0;
#pragma unroll
      for (int j = 0; j < LPL; ++j) {
        const int l = L0 + j;
        if (l < nlev) {
          const double pb = plp[l + 1];
          // radiation_two_stream.F90:342-409, or with aerosol scattering :246-333 (radiation_mcica_lw.F90:160-165)
          L[j] = lwscat ? lw_ref_trans(odp[l], ssp[l], ggp[l], pt, pb) : lw_no_scat(odp[l], pt, pb);
          pt = pb;
        } else {
          L[j] = lw_identity_layer();
        }
      }
    }
    double dn_c = 0.0, toa_c = 0.0;
    if (lwscat) lw_cloudy_subcolumn<LPL>(L, emission, albedo, lane, nlev, wsums, 0, dn_c, toa_c);
    else lw_clear_subcolumn<LPL>(L, emission, albedo, lane, nlev, wsums, dn_c, toa_c);
    if (lane == 0) { gt[g] = dn_c; gt[SD::NG + g] = toa_c; }
    if (cloudy) {
      const uint32_t* codep = w.code_lw + ((size_t)c * SD::NG + g) * nlevp;
      for (int k0 = 0; k0 < ncl; k0 += SCAN_XCH) {
        const int kend = imin(ncl, k0 + SCAN_XCH);
        for (int k = k0 + lane; k < kend; k += 32) {
          const int l = clidx[k];
          // radiation_mcica_lw.F90:248-294: gas + scaled cloud
          const double odg = odp[l], pt = plp[l], pb = plp[l + 1];
          const double scal = od_scaling_from_code(C, T.pdf_val, codep[l], fsds[l]);
          const double* clb = cl + (size_t)l * 3 * SD::NB;
          const double od_cloud_new = scal * clb[b];
          const double od_total = odg + od_cloud_new;
          LwLayer La;
          if (cfg.do_lw_cloud_scattering) {
            double ssa_total = 0.0, g_total = 0.0;
            if (od_total > 0.0) {
              const double ssac = clb[SD::NB + b];
              if (lwscat) {   // :260-280
                const double ssag = ssp[l], scat_od_total = ssag * odg + ssac * od_cloud_new;
                ssa_total = scat_od_total / od_total;
                if (scat_od_total > 0.0) g_total = (ggp[l] * ssag * odg + clb[2 * SD::NB + b] * ssac * od_cloud_new) / scat_od_total;
              } else {
                const double scat_od = ssac * od_cloud_new;
                ssa_total = scat_od / od_total;
                if (scat_od > 0.0) g_total = clb[2 * SD::NB + b] * ssac * od_cloud_new / scat_od;
              }
            }
            La = lw_ref_trans(od_total, ssa_total, g_total, pt, pb);
          } else {
            La = lw_no_scat(od_total, pt, pb);
          }
          const int o = k - k0;
          wx[o] = La.ref; wx[SCAN_XCH + o] = La.trans; wx[2 * SCAN_XCH + o] = La.source_up; wx[3 * SCAN_XCH + o] = La.source_dn;
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < LPL; ++j) {
          const int l = L0 + j;
          if (l < nlev) {
            const int r = clrank[l];
            if (r >= k0 && r < kend) {
              const int o = r - k0;
              L[j].ref = wx[o]; L[j].trans = wx[SCAN_XCH + o]; L[j].source_up = wx[2 * SCAN_XCH + o]; L[j].source_dn = wx[3 * SCAN_XCH + o];
            }
          }
        }
        __syncwarp();
      }
      double dn_a = 0.0, toa_a = 0.0;
      lw_cloudy_subcolumn<LPL>(L, emission, albedo, lane, nlev, wsums, 3, dn_a, toa_a);
      if (lane == 0) { gt[2 * SD::NG + g] = dn_a; gt[3 * SD::NG + g] = toa_a; }
    }
  }
  __syncthreads();
  // ---- outputs: radiation_mcica_lw.F90:300-400 ----
#define OUT2(p, l) ((p)[(size_t)(l) * out.ld + c])
  const double wc = tcc, w1 = 1.0 - tcc;
  const bool want_dv = cfg.do_lw_derivatives && out.lw_derivatives;
  // flux_up at the surface summed over g (denominator of the derivatives): half-level nlev
  const int lns = (nlev - 1) / LPL, js = (nlev - 1) - lns * LPL;
  double up_surf_c = 0.0, up_surf_a = 0.0;
  for (int ww = 0; ww < SCAN_NW; ++ww) {
    up_surf_c += *sum_slot<LPL>(sums + ww * 6 * (LPL + 1) * 32, 1, js, lns);
    if (cloudy) up_surf_a += *sum_slot<LPL>(sums + ww * 6 * (LPL + 1) * 32, 4, js, lns);
  }
  for (int h = tid; h < nl1; h += SCAN_THREADS) {
    const int ln = h == 0 ? 0 : (h - 1) / LPL, j = h == 0 ? LPL : (h - 1) - ln * LPL;
    double s[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      double acc = 0.0;
      if (q < 3 || cloudy)
        for (int ww = 0; ww < SCAN_NW; ++ww) acc += *sum_slot<LPL>(sums + ww * 6 * (LPL + 1) * 32, q, j, ln);
      s[q] = acc;
    }
    const double dnc = s[0], upc = s[1];   // (the TOA slot of the downward sums is never written: flux_dn(TOA) = 0)
    if (out.lw_up_clear) OUT2(out.lw_up_clear, h) = upc;
    if (out.lw_dn_clear) OUT2(out.lw_dn_clear, h) = dnc;
    double up = upc, dn = dnc;
    if (cloudy) { up = wc * s[4] + w1 * upc; dn = wc * s[3] + w1 * dnc; }
    if (out.lw_up) OUT2(out.lw_up, h) = up;
    if (out.lw_dn) OUT2(out.lw_dn, h) = dn;
    if (want_dv) {
      const double dclear = h == nlev ? 1.0 : s[2] / up_surf_c;
      double d = dclear;
      if (cloudy) {
        d = h == nlev ? 1.0 : s[5] / up_surf_a;
        if (tcc < 1.0 - cfg.cloud_fraction_threshold) d = h == nlev ? 1.0 : (1.0 - w1) * d + w1 * dclear;   // modify_lw_derivatives_ica
      }
      OUT2(out.lw_derivatives, h) = d;
    }
  }
#undef OUT2
  if (tid == 0 && out.cloud_cover_lw && mcica) out.cloud_cover_lw[c] = tcc;
  const bool act = tid < SD::NG;
  double dn_surf_g = 0.0;
  if (act) {
    const int g = tid;
    const double fd_surf_clear = gt[g], fu_toa_clear = gt[SD::NG + g];
    dn_surf_g = cloudy ? wc * gt[2 * SD::NG + g] + w1 * fd_surf_clear : fd_surf_clear;
    const size_t i = (size_t)c * SD::NG + g;
    if (out.lw_dn_surf_clear_g) out.lw_dn_surf_clear_g[i] = fd_surf_clear;
    if (out.lw_up_toa_clear_g) out.lw_up_toa_clear_g[i] = fu_toa_clear;
    if (out.lw_dn_surf_g) out.lw_dn_surf_g[i] = dn_surf_g;
    if (out.lw_up_toa_g) out.lw_up_toa_g[i] = cloudy ? wc * gt[3 * SD::NG + g] + w1 * fu_toa_clear : fu_toa_clear;
  }
  lw_surface_canopy<SD>(T, cfg, out, c, tid, act, tile, dn_surf_g);
}

// ---------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------
template <class SD, int LPL>
static int launch_sw_scan_t(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, const Work& w, int nc, int nlev, cudaStream_t st) {
  const int nlevp = (nlev + 3) & ~3;
  const size_t sm = sizeof(double) * (SCAN_NW * 6 * (LPL + 1) * 32 + SCAN_NW * 5 * SCAN_XCH + 6 * SD::NG + 2 * SD::NB + nlev + 4 * SD::RS + 2 * SD::NB) +
                    sizeof(short) * 2 * nlev + 16;
  cudaFuncSetAttribute(sw_scan_kernel<SD, LPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  sw_scan_kernel<SD, LPL><<<nc, SCAN_THREADS, sm, st>>>(T, cfg, in, out, w, nlev, w.ls, nlevp, cfg.solver_sw == 0, cfg.use_aerosols && w.g_sw,
                                                       cfg.do_sw_delta_scaling_with_gases != 0);
  return 1;
}
template <class SD, int LPL>
static int launch_lw_scan_t(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, const Work& w, int nc, int nlev, cudaStream_t st) {
  const int nlevp = (nlev + 3) & ~3;
  const size_t sm = sizeof(double) * (SCAN_NW * 6 * (LPL + 1) * 32 + SCAN_NW * 4 * SCAN_XCH + 4 * SD::NG + nlev + SD::NG) + sizeof(short) * 2 * nlev + 16;
  cudaFuncSetAttribute(lw_scan_kernel<SD, LPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  lw_scan_kernel<SD, LPL><<<nc, SCAN_THREADS, sm, st>>>(T, cfg, in, out, w, nlev, w.ls, nlevp);
  return 1;
}

int scan_max_levels() { return 32 * 6; }

int launch_solver_sw_scan(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, const Work& w, int nc, int nlev, cudaStream_t st) {
  if (cfg.ng_sw != NG_SW) return -1;
  if (nlev <= 32 * 5) return launch_sw_scan_t<SwRrtmg, 5>(T, cfg, in, out, w, nc, nlev, st);
  return launch_sw_scan_t<SwRrtmg, 6>(T, cfg, in, out, w, nc, nlev, st);
}
int launch_solver_lw_scan(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, const Work& w, int nc, int nlev, cudaStream_t st) {
  if (cfg.ng_lw != NG_LW) return -1;
  if (nlev <= 32 * 5) return launch_lw_scan_t<LwRrtmg, 5>(T, cfg, in, out, w, nc, nlev, st);
  return launch_lw_scan_t<LwRrtmg, 6>(T, cfg, in, out, w, nc, nlev, st);
}

}  // namespace ecb
