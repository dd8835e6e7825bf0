// solver_sw.cu -- shortwave McICA (radiation_mcica_sw.F90:41-408) and Cloudless (radiation_cloudless_sw.F90) solvers.
//
// One CTA per sunlit column, one thread per g-point.  The clear-sky and the cloudy (McICA sub-column) solutions of
// the adding method (radiation_adding_ica_sw.F90:24-151) are advanced TOGETHER by the same thread: layers in which
// this column has no cloud share one two-stream evaluation (in the reference the cloudy arrays are copies of the
// clear-sky ones there, radiation_mcica_sw.F90:287-297), and the two independent recurrences give the instruction-level
// parallelism that hides the fp64 latencies.  Three kernels, so that each runs at the occupancy it needs:
//   sw_direct_kernel   top-down: direct beam (1 exp per layer)                      -> fdir per layer, g-sums of fdir
//   sw_adding_kernel   bottom-up: two-stream (calc_ref_trans_sw) + albedo/source     -> a, b, albedo, source per layer
//   sw_flux_kernel     top-down: flux recurrence, g-point sums, flux_type outputs    (pure streaming)
// State between the kernels lives in the per-column scratch ([layer][g], coalesced): 2 + 8 arrays.
#include "bulk_pipe.cuh"
#include "solver_common.cuh"

namespace ecb {

enum { SW_LCH_FLUX = 2, SW_BATCH = 2, SW_NST = 2 };   // sw_flux_kernel: layers per g-point reduction, layers per TMA stage, stages in the ring
typedef BulkRing<SW_NST, SW_BATCH, 10> SwRing;

struct SwColumn {
  int c, g, gg, b; bool act, cloudy; double mu0, tcc, thr;
  size_t n;
  const double *od, *ssa, *gas_g, *cl; const uint4* codep;   // gas_g: asymmetry factor of gas + aerosol, NULL = 0
  double* scr;
};

template <class SD>
__device__ __forceinline__ SwColumn sw_column(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, int nlev, int nlevp) {
  SwColumn s;
  s.c = blockIdx.x; s.g = threadIdx.x; s.act = s.g < SD::NG; s.gg = s.act ? s.g : 0;
  s.mu0 = in.cos_sza[s.c];
  s.tcc = cfg.solver_sw == 2 ? w.tcc[s.c] : 0.0;
  s.cloudy = s.tcc > 0.0;
  s.thr = cfg.cloud_fraction_threshold;
  s.n = (size_t)nlev * SD::NG;
  s.od = w.od_sw + (size_t)s.c * s.n;
  s.ssa = w.ssa_sw + (size_t)s.c * s.n;
  s.gas_g = (cfg.use_aerosols && w.g_sw) ? w.g_sw + (size_t)s.c * s.n : nullptr;
  s.cl = w.cl_sw + (size_t)s.c * nlev * 3 * SD::NB;
  s.b = T.meta->band_of_g_sw[s.gg];
  s.codep = reinterpret_cast<const uint4*>(w.code_sw + ((size_t)s.c * SD::NG + s.gg) * nlevp);
  s.scr = w.scr_sw + (size_t)s.c * SW_SCR_ARRAYS * s.n;
  return s;
}

// ---------------------------------------------------------------------------------------------------------
// A: two-stream layer solutions and the upward sweep (radiation_adding_ica_sw.F90:90-121).
// The reference carries the albedo A and a source S = (direct albedo) x (direct flux at that half-level), which needs the
// direct beam first (its loop :85-88).  S is linear in the direct flux, so the sweep here carries the direct albedo D
// itself (S = D * Fdir, the form radiation_tripleclouds_sw.F90:322-420 uses): D' = Rdir + (Tdir*D + Tdirdif*A) * T / (1 - A*R),
// and the direct beam is marched by the downward kernel together with the fluxes.  No separate direct-beam pass, no second
// evaluation of the cloudy layers' optical properties.  Stored per layer and sub-column (clear / cloudy):
//   a = T/(1-A*R), b = (Tdir*D*R + Tdirdif)/(1-A*R), t = Tdir, A, D      (A, D: of everything below the layer)
// ---------------------------------------------------------------------------------------------------------
enum { SW_PF_DIST = 4 };
template <class SD, bool CLOUDLESS, bool AER, bool DELTA>   // AER: gas + aerosol asymmetry factor g_sw is non-zero and read from memory; DELTA: do_sw_delta_scaling_with_gases
__global__ void __launch_bounds__(SD::THREADS, scaled_min_blocks(SD::THREADS, 128, 6))
sw_adding_kernel(DevTables T, DevCfg cfg, DevIn in, Work w, int nlev, int nlevp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SwColumn s = sw_column<SD>(T, cfg, in, w, nlev, nlevp);
  if (!(s.mu0 > 0.0)) return;
  double* fracs = reinterpret_cast<double*>(smem_raw);   // [nlev]
  double* fsds = fracs + nlev;                           // [nlev]
  double* bandv = fsds + nlev;                           // [2][nb] band albedos
  const int g = s.g, c = s.c;
  for (int l = g; l < nlev; l += SD::THREADS) {
    fracs[l] = s.cloudy ? LD_IN(in.frac, c, l) : 0.0;
    fsds[l] = s.cloudy ? LD_IN(in.fsd, c, l) : 0.0;
  }
  // get_albedos, radiation_single_level.F90:216-365 (weighted-interval mapping to bands)
  if (g < SD::NB) {
    double bd = 0.0, bdir = 0.0;
    for (int ja = 0; ja < cfg.n_albedo_sw; ++ja) {
      const double wgt = T.sw_albedo_weights[g * cfg.n_albedo_sw + ja];
      if (wgt != 0.0) {
        bd = bd + wgt * LD_IN(in.sw_albedo, c, ja);
        if (in.sw_albedo_direct) bdir = bdir + wgt * LD_IN(in.sw_albedo_direct, c, ja);
      }
    }
    bandv[g] = bd; bandv[SD::NB + g] = in.sw_albedo_direct ? bdir : bd;
  }
  __syncthreads();
  if (!s.act) return;
  const CloudMeta& C = *T.cloud;
  const size_t n = s.n;
  double *ac = s.scr, *bc = s.scr + n, *tc = s.scr + 2 * n, *Ac = s.scr + 3 * n, *Dc = s.scr + 4 * n;
  double *aa = s.scr + 5 * n, *ba = s.scr + 6 * n, *ta = s.scr + 7 * n, *Aa = s.scr + 8 * n, *Da = s.scr + 9 * n;
  double* carry = w.sw_carry + (size_t)c * 4 * SD::NG;
  const double mu0 = s.mu0;
  double A_c = bandv[s.b], D_c = mu0 * bandv[SD::NB + s.b];   // surface: diffuse albedo, direct albedo x cos_sza (:95-96)
  double A_a = A_c, D_a = D_c;
  uint4 cq = make_uint4(0, 0, 0, 0);
  // software pipeline: the loads of layer l-1 are issued before the arithmetic of layer l
  size_t i = (size_t)(nlev - 1) * SD::NG + g;
  double od_n = s.od[i], ssa_n = s.ssa[i], gg_n = AER ? s.gas_g[i] : 0.0;
  for (int l = nlev - 1; l >= 0; --l) {
    const double odg = od_n, ssag = ssa_n, gg_gas = AER ? gg_n : 0.0;
    i = (size_t)l * SD::NG + g;
    if (l > 0) {
      const size_t ip = i - SD::NG;
      od_n = s.od[ip]; ssa_n = s.ssa[ip];
      if (AER) gg_n = s.gas_g[ip];
      if (l > SW_PF_DIST) {   // (the rows of a few layers further up: into L2 now, no register held)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(s.od + i - (size_t)SW_PF_DIST * SD::NG));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(s.ssa + i - (size_t)SW_PF_DIST * SD::NG));
      }
    }
    SwLayer Lc;
    if (DELTA && AER) {   // radiation_mcica_sw.F90:165-180 (a no-op when g = 0, i.e. without aerosols)
      double odc = odg, ssac = ssag, gc = gg_gas;
      sw_delta_eddington(odc, ssac, gc);
      Lc = CLOUDLESS ? sw_ref_trans_cloudless(mu0, odc, ssac, gc) : sw_ref_trans(mu0, odc, ssac, gc);
    } else {
      Lc = CLOUDLESS ? sw_ref_trans_cloudless(mu0, odg, ssag, gg_gas) : sw_ref_trans(mu0, odg, ssag, gg_gas);
    }
    {
      const double inv_den = 1.0 / (1.0 - A_c * Lc.ref);
      ac[i] = Lc.trans * inv_den;
      bc[i] = (Lc.trans_dir_dir * D_c * Lc.ref + Lc.trans_dir_diff) * inv_den;
      tc[i] = Lc.trans_dir_dir; Ac[i] = A_c; Dc[i] = D_c;   // albedos of everything below the half-level under layer l
      const double A_new = Lc.ref + Lc.trans * Lc.trans * A_c * inv_den;
      D_c = Lc.ref_dir + (Lc.trans_dir_dir * D_c + Lc.trans_dir_diff * A_c) * Lc.trans * inv_den;
      A_c = A_new;
    }
    if (s.cloudy) {
      if (l == nlev - 1 || (l & 3) == 3) cq = __ldg(s.codep + (l >> 2));
      SwLayer La = Lc;
      if (fracs[l] >= s.thr) {
        double odt, ssat, gt;
        sw_cloudy_props<SD>(C, T.pdf_val, pick4(cq, l & 3), fsds[l], s.cl + (size_t)l * 3 * SD::NB, s.b, odg, ssag, gg_gas, odt, ssat, gt);
        if (DELTA) sw_delta_eddington(odt, ssat, gt);   // radiation_mcica_sw.F90:274-278
        La = sw_ref_trans(mu0, odt, ssat, gt);
      }
      const double inv_den = 1.0 / (1.0 - A_a * La.ref);
      aa[i] = La.trans * inv_den;
      ba[i] = (La.trans_dir_dir * D_a * La.ref + La.trans_dir_diff) * inv_den;
      ta[i] = La.trans_dir_dir; Aa[i] = A_a; Da[i] = D_a;
      const double A_new = La.ref + La.trans * La.trans * A_a * inv_den;
      D_a = La.ref_dir + (La.trans_dir_dir * D_a + La.trans_dir_diff * A_a) * La.trans * inv_den;
      A_a = A_new;
    }
  }
  carry[g] = D_c;             // direct albedo of the whole atmosphere + surface: flux_up(TOA) = D * incoming
  carry[SD::NG + g] = D_a;
}

// ---------------------------------------------------------------------------------------------------------
// B: direct beam and fluxes top-down (radiation_adding_ica_sw.F90:85-88, :134-146), g-point sums, blending, flux_type outputs
// ---------------------------------------------------------------------------------------------------------
template <class SD>
__global__ void __launch_bounds__(SD::THREADS, scaled_min_blocks(SD::THREADS, 128, 5))
sw_flux_kernel(DevTables T, DevCfg cfg, DevIn in, DevOut out, Work w, int nlev, int nlevp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SwColumn s = sw_column<SD>(T, cfg, in, w, nlev, nlevp);
  const int c = s.c, g = s.g, nl1 = nlev + 1;
  const bool act = s.act;
  const double mu0 = s.mu0;
#define OUT2(p, l) ((p)[(size_t)(l) * out.ld + c])
  if (!(mu0 > 0.0)) {
    // night column: radiation_mcica_sw.F90:380-401
    for (int l = g; l < nl1; l += SD::THREADS) {
      if (out.sw_up) OUT2(out.sw_up, l) = 0.0;
      if (out.sw_dn) OUT2(out.sw_dn, l) = 0.0;
      if (out.sw_dn_direct) OUT2(out.sw_dn_direct, l) = 0.0;
      if (out.sw_up_clear) OUT2(out.sw_up_clear, l) = 0.0;
      if (out.sw_dn_clear) OUT2(out.sw_dn_clear, l) = 0.0;
      if (out.sw_dn_direct_clear) OUT2(out.sw_dn_direct_clear, l) = 0.0;
    }
    if (cfg.solver_sw == 0 && cfg.do_save_spectral_flux) {   // Cloudless: per-band profiles (radiation_cloudless_sw.F90)
      double* pb[3] = {out.sw_up_band, out.sw_dn_band, out.sw_dn_direct_band};
      for (int k = 0; k < 3; ++k)
        if (pb[k]) for (int i = g; i < nl1 * SD::NB; i += SD::THREADS) pb[k][((size_t)(i / SD::NB) * out.ld + c) * SD::NB + (i % SD::NB)] = 0.0;
    }
    if (act) {
      const size_t i = (size_t)c * SD::NG + g;
      double* gs[6] = {out.sw_dn_diffuse_surf_g, out.sw_dn_direct_surf_g, out.sw_up_toa_g,
                       out.sw_dn_diffuse_surf_clear_g, out.sw_dn_direct_surf_clear_g, out.sw_up_toa_clear_g};
      for (int k = 0; k < 6; ++k) if (gs[k]) gs[k][i] = 0.0;
    }
    if (g < SD::NB) {
      double* bs[4] = {out.sw_dn_surf_band, out.sw_dn_direct_surf_band, out.sw_dn_surf_clear_band, out.sw_dn_direct_surf_clear_band};
      for (int k = 0; k < 4; ++k) if (bs[k] && cfg.do_surface_sw_spectral_flux) bs[k][(size_t)c * SD::NB + g] = 0.0;
    }
    if (g < cfg.n_canopy_bands_sw && cfg.do_canopy_fluxes_sw) {
      if (out.sw_dn_diffuse_surf_canopy) out.sw_dn_diffuse_surf_canopy[(size_t)c * cfg.n_canopy_bands_sw + g] = 0.0;
      if (out.sw_dn_direct_surf_canopy) out.sw_dn_direct_surf_canopy[(size_t)c * cfg.n_canopy_bands_sw + g] = 0.0;
    }
    return;
  }
  double* ssum = reinterpret_cast<double*>(smem_raw);    // [6][nl1]: dir_c, dn_c, up_c, dir_a, dn_a, up_a  (dir: per unit area normal to the beam)
  double* tile = ssum + 6 * nl1;                          // [6][SW_LCH_FLUX][SD::RS]
  double *s_dir_c = ssum, *s_dn_c = ssum + nl1, *s_up_c = ssum + 2 * nl1, *s_dir = ssum + 3 * nl1, *s_dn = ssum + 4 * nl1, *s_up = ssum + 5 * nl1;
  const size_t n = s.n;
  const double *ac = s.scr, *bc = s.scr + n, *tc = s.scr + 2 * n, *Ac = s.scr + 3 * n, *Dc = s.scr + 4 * n;
  const double *aa = s.scr + 5 * n, *ba = s.scr + 6 * n, *ta = s.scr + 7 * n, *Aa = s.scr + 8 * n, *Da = s.scr + 9 * n;
  const double* carry = w.sw_carry + (size_t)c * 4 * SD::NG;
  const bool cloudy = s.cloudy;
  const double inc = w.incoming[(size_t)c * SD::NG + s.gg];
  double dir_c = inc, dir_a = inc, fdd_c = 0.0, fdd_a = 0.0;
  const double toa_c = dir_c * carry[s.gg], toa_a0 = dir_a * carry[SD::NG + s.gg];
  {
    double* dst[6] = {s_dir_c, s_dn_c, s_up_c, s_dir, s_dn, s_up};
    const int nf = cloudy ? 6 : 3;
    // Cloudless + do_save_spectral_flux (radiation_cloudless_sw.F90): per-band up, direct (x mu0) and total down profiles
    const bool bands = cfg.solver_sw == 0 && cfg.do_save_spectral_flux && (out.sw_up_band || out.sw_dn_band || out.sw_dn_direct_band);
    const BandOut bo[3] = {{out.sw_up_band, out.ld, 2, -1, 1.0, 0.0, nullptr, 0}, {out.sw_dn_direct_band, out.ld, 0, -1, mu0, 0.0, nullptr, 0},
                           {out.sw_dn_band, out.ld, 0, 1, mu0, 1.0, nullptr, 0}};
    int slot = 0, lfirst = 0;
    if (act) {
      tile[slot * SD::RS + g] = dir_c; tile[(SW_LCH_FLUX + slot) * SD::RS + g] = 0.0; tile[(2 * SW_LCH_FLUX + slot) * SD::RS + g] = toa_c;
      tile[(3 * SW_LCH_FLUX + slot) * SD::RS + g] = dir_a; tile[(4 * SW_LCH_FLUX + slot) * SD::RS + g] = 0.0; tile[(5 * SW_LCH_FLUX + slot) * SD::RS + g] = toa_a0;
    }
    ++slot;
    // The 5 (clear) or 10 (clear + cloudy) scratch arrays are streamed through a ring of shared-memory stages filled by the TMA
    // unit, SW_BATCH layers per stage, requested two stages ahead by thread 0: the copies run while the CTA reduces.
    SwRing ring;
    ring.carve(reinterpret_cast<unsigned char*>(tile + 6 * SW_LCH_FLUX * SD::RS), SD::NG);
    ring.init(SD::THREADS);
    const double* src[10] = {ac, bc, tc, Ac, Dc, aa, ba, ta, Aa, Da};
    const int narr = cloudy ? 10 : 5, nstage = (nlev + SW_BATCH - 1) / SW_BATCH;
    if (threadIdx.x == 0)
      for (int j = 0; j < SW_NST && j < nstage; ++j) ring.issue(j, src, narr, j * SW_BATCH, imin((int)SW_BATCH, nlev - j * SW_BATCH));
    for (int j = 0; j < nstage; ++j) {
      const int l0 = j * SW_BATCH, st = j % SW_NST;
      double ca[SW_BATCH], cb[SW_BATCH], ct[SW_BATCH], cA[SW_BATCH], cD[SW_BATCH], da[SW_BATCH], db[SW_BATCH], dt[SW_BATCH], dA[SW_BATCH], dD[SW_BATCH];
      ring.wait_full(j);
#pragma unroll
      for (int k = 0; k < SW_BATCH; ++k)
        if (act && l0 + k < nlev) {
          const int o = k * SD::NG + g;
          ca[k] = ring.stage(st, 0)[o]; cb[k] = ring.stage(st, 1)[o]; ct[k] = ring.stage(st, 2)[o]; cA[k] = ring.stage(st, 3)[o]; cD[k] = ring.stage(st, 4)[o];
          if (cloudy) { da[k] = ring.stage(st, 5)[o]; db[k] = ring.stage(st, 6)[o]; dt[k] = ring.stage(st, 7)[o]; dA[k] = ring.stage(st, 8)[o]; dD[k] = ring.stage(st, 9)[o]; }
        }
#pragma unroll
      for (int k = 0; k < SW_BATCH; ++k) {
        const int l = l0 + k;
        if (l < nlev) {
          if (act) {
            fdd_c = ca[k] * fdd_c + cb[k] * dir_c;
            dir_c = ct[k] * dir_c;
            const double fu_c = dir_c * cD[k] + fdd_c * cA[k];
            tile[slot * SD::RS + g] = dir_c; tile[(SW_LCH_FLUX + slot) * SD::RS + g] = fdd_c; tile[(2 * SW_LCH_FLUX + slot) * SD::RS + g] = fu_c;
            if (cloudy) {
              fdd_a = da[k] * fdd_a + db[k] * dir_a;
              dir_a = dt[k] * dir_a;
              const double fu_a = dir_a * dD[k] + fdd_a * dA[k];
              tile[(3 * SW_LCH_FLUX + slot) * SD::RS + g] = dir_a; tile[(4 * SW_LCH_FLUX + slot) * SD::RS + g] = fdd_a;
              tile[(5 * SW_LCH_FLUX + slot) * SD::RS + g] = fu_a;
            }
          }
          ++slot;
          if (k == SW_BATCH - 1 || l == nlev - 1) {
            // every value of this stage has gone through the arithmetic above (so its shared-memory loads have completed): hand
            // the stage back and request the one after next before the reduction starts
            ring.release(j);
            if (threadIdx.x == 0 && j + SW_NST < nstage)
              ring.issue(j + SW_NST, src, narr, (j + SW_NST) * SW_BATCH, imin((int)SW_BATCH, nlev - (j + SW_NST) * SW_BATCH));
          }
          if (slot == SW_LCH_FLUX || l == nlev - 1) {
            if (bands) flush_bands(tile, SD::RS, SW_LCH_FLUX, slot, bo, 3, lfirst, 1, c, SD::NB, T.meta->sw);
            flush_tile(tile, SD::RS, SD::NG, nf, slot, dst, lfirst, 1, SW_LCH_FLUX); lfirst += slot; slot = 0;
          }
        }
      }
    }
  }
  const double tcc = s.tcc;
  const double wc = tcc, w1 = 1.0 - tcc;
  for (int l = g; l < nl1; l += SD::THREADS) {
    const double dirc = s_dir_c[l] * mu0, upc = s_up_c[l], dnc = s_dn_c[l] + dirc;
    if (out.sw_up_clear) OUT2(out.sw_up_clear, l) = upc;
    if (out.sw_dn_clear) OUT2(out.sw_dn_clear, l) = dnc;
    if (out.sw_dn_direct_clear) OUT2(out.sw_dn_direct_clear, l) = dirc;
    double up = upc, dn = dnc, dir = dirc;
    if (cloudy) {
      const double dira = s_dir[l] * mu0;
      up = wc * s_up[l] + w1 * upc;
      dn = wc * (s_dn[l] + dira) + w1 * dnc;
      dir = wc * dira + w1 * dirc;
    }
    if (out.sw_up) OUT2(out.sw_up, l) = up;
    if (out.sw_dn) OUT2(out.sw_dn, l) = dn;
    if (out.sw_dn_direct) OUT2(out.sw_dn_direct, l) = dir;
  }
  if (g == 0 && out.cloud_cover_sw && cfg.solver_sw == 2) out.cloud_cover_sw[c] = tcc;
  // per-g surface / TOA fluxes
  const double dif_c = fdd_c, dir_cs = dir_c * mu0;
  double dif_a = dif_c, dir_as = dir_cs, toa_a = toa_c;
  if (cloudy) {
    dif_a = wc * fdd_a + w1 * dif_c;
    dir_as = wc * (dir_a * mu0) + w1 * dir_cs;
    toa_a = wc * toa_a0 + w1 * toa_c;
  }
  if (act) {
    const size_t i = (size_t)c * SD::NG + g;
    if (out.sw_dn_diffuse_surf_clear_g) out.sw_dn_diffuse_surf_clear_g[i] = dif_c;
    if (out.sw_dn_direct_surf_clear_g) out.sw_dn_direct_surf_clear_g[i] = dir_cs;
    if (out.sw_up_toa_clear_g) out.sw_up_toa_clear_g[i] = toa_c;
    if (out.sw_dn_diffuse_surf_g) out.sw_dn_diffuse_surf_g[i] = dif_a;
    if (out.sw_dn_direct_surf_g) out.sw_dn_direct_surf_g[i] = dir_as;
    if (out.sw_up_toa_g) out.sw_up_toa_g[i] = toa_a;
  }
  sw_surface_spectral<SD>(T, cfg, out, c, g, act, tile, SD::RS, dir_as, dif_a, dir_cs, dif_c);
#undef OUT2
}

template <class SD>
static int launch_solver_sw_t(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, const Work& w, int nc, int nlev, cudaStream_t st) {
  const int nlevp = (nlev + 3) & ~3;
  const size_t smB = sizeof(double) * (2 * nlev + 2 * SD::NB) + 16;
  const size_t smC = sizeof(double) * (6 * (nlev + 1) + 6 * SW_LCH_FLUX * SD::RS) + sizeof(double) * SW_NST * 10 * SW_BATCH * SD::NG + 2 * SW_NST * sizeof(uint64_t) + 16;
  const bool aer = cfg.use_aerosols && w.g_sw;
  const bool delta = cfg.do_sw_delta_scaling_with_gases != 0;
#define SW_ADDING(CL, AE, DE) sw_adding_kernel<SD, CL, AE, DE><<<nc, SD::THREADS, smB, st>>>(T, cfg, in, w, nlev, nlevp)
  if (cfg.solver_sw == 2) {
    if (delta) { if (aer) SW_ADDING(false, true, true); else SW_ADDING(false, false, true); }
    else { if (aer) SW_ADDING(false, true, false); else SW_ADDING(false, false, false); }
  } else {   // Cloudless: only the gas-aerosol mixture can need the scaling
    if (aer && delta) SW_ADDING(true, true, true);
    else if (aer) SW_ADDING(true, true, false);
    else SW_ADDING(true, false, false);
  }
#undef SW_ADDING
  cudaFuncSetAttribute(sw_flux_kernel<SD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smC);
  sw_flux_kernel<SD><<<nc, SD::THREADS, smC, st>>>(T, cfg, in, out, w, nlev, nlevp);
  return 2;
}

int launch_solver_sw(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, const Work& w, int nc, int nlev, cudaStream_t st) {
  if (cfg.solver_sw == 4 || cfg.solver_sw == 1) return launch_tc_sw(T, cfg, in, out, w, nc, nlev, st);   // Tripleclouds, Homogeneous
  if (cfg.solver_sw == 3) return launch_sp_sw(T, cfg, in, out, w, nc, nlev, st);   // SPARTACUS
  if (w.layout_b_sw) return launch_solver_sw_scan(T, cfg, in, out, w, nc, nlev, st);   // McICA / Cloudless as warp scans (solver_scan.cu)
  switch (cfg.ng_sw) {
    case NG_SW: return launch_solver_sw_t<SwRrtmg>(T, cfg, in, out, w, nc, nlev, st);
    case 32: return launch_solver_sw_t<Ckd32>(T, cfg, in, out, w, nc, nlev, st);
    case 64: return launch_solver_sw_t<Ckd64>(T, cfg, in, out, w, nc, nlev, st);
    case 96: return launch_solver_sw_t<Ckd96>(T, cfg, in, out, w, nc, nlev, st);
  }
  return -1;   // check_config refuses other spectral sizes
}

}  // namespace ecb
