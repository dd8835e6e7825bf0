// solver_lw.cu -- longwave McICA (radiation_mcica_lw.F90:39-419) and Cloudless (radiation_cloudless_lw.F90) solvers.
//
// One CTA per column, one thread per g-point.  Clear-sky layers are no-scattering (radiation_two_stream.F90:342-409:
// one exp and one division per layer and g-point), so their transmittance/sources are RECOMPUTED in every sweep
// instead of being stored: the fp64 pipe is idle in these memory-bound sweeps and 32 bytes per element are saved.
//   lw_down_kernel   top-down: clear-sky downward flux (calc_fluxes_no_scattering_lw, first loop)
//   lw_up_kernel     bottom-up: clear-sky upward flux + derivative products, and -- in the same sweep -- the cloudy
//                    sub-column: two-stream of cloudy layers (calc_ref_trans_lw), albedo/source up to cloud top
//                    (fast_adding_ica_lw), then the upward flux above cloud top
//   lw_flux_kernel   top-down from cloud top: cloudy fluxes; derivative sums (calc_lw_derivatives_ica); outputs
#include "bulk_pipe.cuh"
#include "solver_common.cuh"

namespace ecb {
enum { LW_PF_DIST = 4 };

enum { LW_LCH_FLUX = 8, LW_LCH_UP = 8, LW_FLUX_NL = 2, LW_FLUX_NST = 2 };   // lw_flux_kernel: layers per TMA stage, stages in the ring
typedef BulkRing<LW_FLUX_NST, LW_FLUX_NL, 4> LwRing;
enum { LW_UP_NL = 4, LW_UP_NST = 2 };   // lw_up_kernel: optical depth + Planck function streamed bottom-up, 4 layers per stage
typedef BulkRing<LW_UP_NST, LW_UP_NL, 2> LwUpRing;
enum { LWS_DN_C = 0, LWS_UP_C = 1, LWS_DV_C = 2, LWS_UP_A = 3, LWS_DN_A = 4, LWS_DV_A = 5 };

struct LwColumn {
  int c, g, gg, ict; bool act, mcica, cloudy; double tcc, thr;
  size_t n;
  const double *od, *pl;
  double *scr, *sums, *carry;
};

template <class SD>
__device__ __forceinline__ LwColumn lw_column(const DevCfg& cfg, const Work& w, int nlev) {
  LwColumn s;
  s.c = blockIdx.x; s.g = threadIdx.x; s.act = s.g < SD::NG; s.gg = s.act ? s.g : 0;
  s.mcica = cfg.solver_lw == 2;
  s.tcc = s.mcica ? w.tcc[s.c] : 0.0;
  s.cloudy = s.tcc > 0.0;
  s.ict = (s.cloudy || cfg.solver_lw == 4 || cfg.solver_lw == 1) ? w.ict[s.c] : nlev;   // 4 = Tripleclouds (1 = Homogeneous runs on its kernels): needs the clear-sky flux_dn at cloud top too
  s.thr = cfg.cloud_fraction_threshold;
  s.n = (size_t)nlev * SD::NG;
  s.od = w.od_lw + (size_t)s.c * s.n;
  s.pl = w.planck + (size_t)s.c * (nlev + 1) * SD::NG;
  s.scr = w.scr_lw + (size_t)s.c * LW_SCR_ARRAYS * s.n;
  s.sums = w.lw_sums + (size_t)s.c * 6 * (nlev + 1);
  s.carry = w.lw_carry + (size_t)s.c * 4 * SD::NG;
  return s;
}

// ---------------------------------------------------------------------------------------------------------
// clear-sky downward flux
// ---------------------------------------------------------------------------------------------------------
template <class SD>
__global__ void __launch_bounds__(SD::THREADS, scaled_min_blocks(SD::THREADS, 160, 6))
lw_down_kernel(DevTables T, DevCfg cfg, DevOut out, Work w, int nlev) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const LwColumn s = lw_column<SD>(cfg, w, nlev);
  double* tile = reinterpret_cast<double*>(smem_raw);   // [1][LCH][SD::RS]
  const int g = s.g, nl1 = nlev + 1;
  double* dst[1] = {s.sums + LWS_DN_C * nl1};
  // per-band profile of flux_dn: clear-sky = all-sky for Cloudless; Tripleclouds overwrites the levels below cloud top
  const BandOut bo[1] = {{(cfg.do_save_spectral_flux && !s.mcica) ? out.lw_dn_band : nullptr, out.ld, 0, -1, 1.0, 0.0, nullptr, 0}};
  double fd = 0.0, fd_ict = 0.0;
  int slot = 0, lfirst = 0;
  if (s.act) tile[g] = 0.0;   // flux_dn at TOA
  ++slot;
  double pt = s.act ? s.pl[g] : 0.0;
  // software pipeline: the loads of layer l+1 are issued before the exp/div of layer l
  double od_n = s.act ? s.od[g] : 0.0, pb_n = s.act ? s.pl[SD::NG + g] : 0.0;
  for (int l = 0; l < nlev; ++l) {
    if (s.act) {
      if (l == s.ict) fd_ict = fd;
      const double odg = od_n, pb = pb_n;
      if (l + 1 < nlev) { const size_t i1 = (size_t)(l + 1) * SD::NG + g; od_n = s.od[i1]; pb_n = s.pl[i1 + SD::NG]; }
      if (l + LW_PF_DIST < nlev) {   // (rows a few layers further down: into L2 now, no register held)
        const size_t ipf = (size_t)(l + LW_PF_DIST) * SD::NG + g;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(s.od + ipf));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(s.pl + ipf + SD::NG));
      }
      const LwLayer L = lw_no_scat(odg, pt, pb);
      pt = pb;
      fd = L.trans * fd + L.source_dn;
      tile[slot * SD::RS + g] = fd;
    }
    ++slot;
    if (slot == LCH || l == nlev - 1) {
      if (bo[0].dst) flush_bands(tile, SD::RS, LCH, slot, bo, 1, lfirst, 1, s.c, SD::NB, T.meta->lw);
      flush_tile(tile, SD::RS, SD::NG, 1, slot, dst, lfirst, 1); lfirst += slot; slot = 0;
    }
  }
  if (s.act) { s.carry[g] = fd_ict; s.carry[SD::NG + g] = fd; }
}

// ---------------------------------------------------------------------------------------------------------
// upward sweep: clear-sky flux_up and derivative products; cloudy albedo/source below cloud top, flux_up above
// ---------------------------------------------------------------------------------------------------------
template <class SD>
__global__ void __launch_bounds__(SD::THREADS, scaled_min_blocks(SD::THREADS, 160, 4))
lw_up_kernel(DevTables T, DevCfg cfg, DevIn in, DevOut out, Work w, int nlev, int nlevp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const LwColumn s = lw_column<SD>(cfg, w, nlev);
  double* tile = reinterpret_cast<double*>(smem_raw);   // [3][LW_LCH_UP][SD::RS]
  double* fracs = tile + 3 * LW_LCH_UP * SD::RS;               // [nlev]
  double* fsds = fracs + nlev;                          // [nlev]
  const int c = s.c, g = s.g, nl1 = nlev + 1;
  for (int l = g; l < nlev; l += SD::THREADS) {
    fracs[l] = s.cloudy ? LD_IN(in.frac, c, l) : 0.0;
    fsds[l] = s.cloudy ? LD_IN(in.fsd, c, l) : 0.0;
  }
  __syncthreads();
  const CloudMeta& C = *T.cloud;
  const size_t n = s.n;
  double *sa = s.scr, *sb = s.scr + n, *sA = s.scr + 2 * n, *sS = s.scr + 3 * n, *sP = s.scr + 4 * n;
  const double emission = w.emission[(size_t)c * SD::NG + s.gg], albedo = w.lw_albedo[(size_t)c * SD::NG + s.gg];
  const int b = T.meta->band_of_g_lw[s.gg];
  const uint4* codep = reinterpret_cast<const uint4*>(w.code_lw + ((size_t)c * SD::NG + s.gg) * nlevp);
  const double* cl = w.cl_lw + (size_t)c * nlev * 3 * SD::NB;
  const bool cloudy = s.cloudy;
  const int ict = s.ict;
  const double fd_ict = s.carry[s.gg], fd_surf_clear = s.carry[SD::NG + s.gg];
  double* dst[3] = {s.sums + LWS_UP_C * nl1, s.sums + LWS_DV_C * nl1, s.sums + LWS_UP_A * nl1};
  const int nf = cloudy ? 3 : 2;
  const BandOut bo[1] = {{(cfg.do_save_spectral_flux && cfg.solver_lw == 0) ? out.lw_up_band : nullptr, out.ld, 0, -1, 1.0, 0.0, nullptr, 0}};   // Cloudless
  // surface
  double fu = emission + albedo * fd_surf_clear;   // clear-sky flux_up at the surface
  double prod = fu;                                 // flux_up(surface) * prod(trans) : calc_lw_derivatives_ica
  double A = albedo, S = emission;                  // cloudy sub-column: albedo / source of everything below
  double fu_a = 0.0, pa = 1.0;                      // cloudy flux_up (above cloud top), product of transmittances
  int slot = 0, lfirst = nlev;
  if (s.act) { tile[g] = fu; tile[LW_LCH_UP * SD::RS + g] = prod; tile[2 * LW_LCH_UP * SD::RS + g] = 0.0; }
  ++slot;
  uint4 cq = make_uint4(0, 0, 0, 0);
  double cl_n[3] = {0.0, 0.0, 0.0};
  if (s.act && cloudy && fracs[nlev - 1] >= s.thr) { const double* q = cl + (size_t)(nlev - 1) * 3 * SD::NB; cl_n[0] = q[b]; cl_n[1] = q[SD::NB + b]; cl_n[2] = q[2 * SD::NB + b]; }
  double pb = s.act ? s.pl[(size_t)nlev * SD::NG + g] : 0.0;
  // optical depth and Planck function (top of layer) come through a ring of shared-memory stages filled by the TMA unit
  // (bulk_pipe.cuh), LW_UP_NL layers per stage from the surface upwards, requested two stages ahead
  LwUpRing ring;
  ring.carve(reinterpret_cast<unsigned char*>(fsds + nlev), SD::NG);
  ring.init(SD::THREADS);
  const double* src[2] = {s.od, s.pl};
  const int nstage = (nlev + LW_UP_NL - 1) / LW_UP_NL;
  auto stage_l0 = [&](int j) { const int hi = nlev - j * LW_UP_NL; return hi - LW_UP_NL > 0 ? hi - LW_UP_NL : 0; };
  if (threadIdx.x == 0)
    for (int j = 0; j < LW_UP_NST && j < nstage; ++j) ring.issue(j, src, 2, stage_l0(j), nlev - j * LW_UP_NL - stage_l0(j));
  for (int js = 0; js < nstage; ++js) {
   const int sl0 = stage_l0(js), snl = nlev - js * LW_UP_NL - sl0, sst = js % LW_UP_NST;
   ring.wait_full(js);
   for (int l = sl0 + snl - 1; l >= sl0; --l) {
    if (s.act) {
      const size_t i = (size_t)l * SD::NG + g;
      const double odg = ring.stage(sst, 0)[(l - sl0) * SD::NG + g], pt = ring.stage(sst, 1)[(l - sl0) * SD::NG + g];
      const LwLayer Lc = lw_no_scat(odg, pt, pb);
      fu = Lc.trans * fu + Lc.source_up;
      prod = prod * Lc.trans;
      tile[slot * SD::RS + g] = fu; tile[(LW_LCH_UP + slot) * SD::RS + g] = prod;
      if (cloudy) {
        if (l >= ict) {
          if (l == nlev - 1 || (l & 3) == 3) cq = __ldg(codep + (l >> 2));
          double a_, b_, t_;
          // the band's cloud properties: this layer's were requested one layer ago, the next layer's are requested now
          const double clb0 = cl_n[0], clb1 = cl_n[1], clb2 = cl_n[2];
          if (l > ict && fracs[l - 1] >= s.thr) { const double* q = cl + (size_t)(l - 1) * 3 * SD::NB; cl_n[0] = q[b]; cl_n[1] = q[SD::NB + b]; cl_n[2] = q[2 * SD::NB + b]; }
          if (fracs[l] >= s.thr) {
            // radiation_mcica_lw.F90:248-294: gas + scaled cloud
            const double scal = od_scaling_from_code(C, T.pdf_val, pick4(cq, l & 3), fsds[l]);
            const double od_cloud_new = scal * clb0;
            const double od_total = odg + od_cloud_new;
            LwLayer L;
            if (cfg.do_lw_cloud_scattering) {
              double ssa_total = 0.0, g_total = 0.0;
              if (od_total > 0.0) {
                const double ssac = clb1;
                const double scat_od = ssac * od_cloud_new;
                ssa_total = scat_od / od_total;
                if (scat_od > 0.0) g_total = clb2 * ssac * od_cloud_new / scat_od;
              }
              L = lw_ref_trans(od_total, ssa_total, g_total, pt, pb);
            } else {
              L = lw_no_scat(od_total, pt, pb);
            }
            const double inv_den = 1.0 / (1.0 - A * L.ref);
            a_ = L.trans * inv_den;
            b_ = (L.ref * S + L.source_dn) * inv_den;
            t_ = L.trans;
            const double A_new = L.ref + L.trans * L.trans * A * inv_den;
            const double S_new = L.source_up + L.trans * (S + A * L.source_dn) * inv_den;
            sA[i] = A; sS[i] = S;   // albedo/source at the half-level below layer l
            A = A_new; S = S_new;
          } else {
            a_ = Lc.trans; b_ = Lc.source_dn; t_ = Lc.trans;
            sA[i] = A; sS[i] = S;
            const double A_new = t_ * t_ * A;
            const double S_new = Lc.source_up + t_ * (S + A * Lc.source_dn);
            A = A_new; S = S_new;
          }
          sa[i] = a_; sb[i] = b_;
          pa = pa * t_;
          if (l == ict) fu_a = S + A * fd_ict;   // flux_up at cloud top
        } else {
          fu_a = Lc.trans * fu_a + Lc.source_up;
          pa = pa * Lc.trans;
        }
        sP[i] = pa;
        tile[(2 * LW_LCH_UP + slot) * SD::RS + g] = l <= ict ? fu_a : 0.0;
      }
      pb = pt;
    }
    ++slot;
    if (l == sl0) {   // the stage's values have all been used: hand it back, request the stage after next
      ring.release(js);
      if (threadIdx.x == 0 && js + LW_UP_NST < nstage)
        ring.issue(js + LW_UP_NST, src, 2, stage_l0(js + LW_UP_NST), nlev - (js + LW_UP_NST) * LW_UP_NL - stage_l0(js + LW_UP_NST));
    }
    if (slot == LW_LCH_UP || l == 0) {
      if (bo[0].dst) flush_bands(tile, SD::RS, LW_LCH_UP, slot, bo, 1, lfirst, -1, c, SD::NB, T.meta->lw);
      flush_tile(tile, SD::RS, SD::NG, nf, slot, dst, lfirst, -1, LW_LCH_UP); lfirst -= slot; slot = 0;
    }
   }
  }
  if (s.act) { s.carry[2 * SD::NG + g] = fu; s.carry[3 * SD::NG + g] = fu_a; }
}

// ---------------------------------------------------------------------------------------------------------
// cloudy fluxes from cloud top down, derivative sums, and the flux_type outputs
// ---------------------------------------------------------------------------------------------------------
template <class SD>
__global__ void __launch_bounds__(SD::THREADS, scaled_min_blocks(SD::THREADS, 160, 6))
lw_flux_kernel(DevTables T, DevCfg cfg, DevOut out, Work w, int nlev) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const LwColumn s = lw_column<SD>(cfg, w, nlev);
  double* tile = reinterpret_cast<double*>(smem_raw);   // [2][LW_LCH_FLUX][SD::RS]
  const int c = s.c, g = s.g, nl1 = nlev + 1;
  const bool act = s.act, cloudy = s.cloudy;
  const int ict = s.ict;
  const size_t n = s.n;
  const double *sa = s.scr, *sb = s.scr + n, *sA = s.scr + 2 * n, *sS = s.scr + 3 * n, *sP = s.scr + 4 * n;
  const double fd_ict = s.carry[s.gg], fd_surf_clear = s.carry[SD::NG + s.gg];
  const double fu_toa_clear = s.carry[2 * SD::NG + s.gg], fu_toa_a = s.carry[3 * SD::NG + s.gg];
  double fd_surf = fd_surf_clear;
  const bool want_dv = cfg.do_lw_derivatives && out.lw_derivatives;
  if (cloudy) {
    double fd = fd_ict, fu = 0.0;
    // scratch arrays streamed through a ring of shared-memory stages filled by the TMA unit (bulk_pipe.cuh), two stages ahead
    LwRing ring;
    ring.carve(reinterpret_cast<unsigned char*>(tile + 2 * LW_LCH_FLUX * SD::RS), SD::NG);
    ring.init(SD::THREADS);
    int jbase = 0;
    {
      double* dst[2] = {s.sums + LWS_DN_A * nl1, s.sums + LWS_UP_A * nl1};
      const double* src[4] = {sa, sb, sA, sS};
      const int nstage = (nlev - ict + LW_FLUX_NL - 1) / LW_FLUX_NL;
      if (threadIdx.x == 0)
        for (int j = 0; j < LW_FLUX_NST && j < nstage; ++j) ring.issue(j, src, 4, ict + j * LW_FLUX_NL, imin((int)LW_FLUX_NL, nlev - ict - j * LW_FLUX_NL));
      int slot = 0, lfirst = ict + 1;
      for (int j = 0; j < nstage; ++j) {
        const int l0 = ict + j * LW_FLUX_NL, st = j % LW_FLUX_NST;
        double va[LW_FLUX_NL], vb[LW_FLUX_NL], vA[LW_FLUX_NL], vS[LW_FLUX_NL];
        ring.wait_full(j);
#pragma unroll
        for (int k = 0; k < LW_FLUX_NL; ++k)
          if (act && l0 + k < nlev) {
            const int o = k * SD::NG + g;
            va[k] = ring.stage(st, 0)[o]; vb[k] = ring.stage(st, 1)[o]; vA[k] = ring.stage(st, 2)[o]; vS[k] = ring.stage(st, 3)[o];
          }
#pragma unroll
        for (int k = 0; k < LW_FLUX_NL; ++k) {
          const int l = l0 + k;
          if (l < nlev) {
            if (act) {
              fd = va[k] * fd + vb[k];
              fu = vA[k] * fd + vS[k];
              tile[slot * SD::RS + g] = fd; tile[(LW_LCH_FLUX + slot) * SD::RS + g] = fu;
            }
            ++slot;
            if (k == LW_FLUX_NL - 1 || l == nlev - 1) {   // all values of the stage consumed (loads complete): hand it back, request the next
              ring.release(j);
              if (threadIdx.x == 0 && j + LW_FLUX_NST < nstage)
                ring.issue(j + LW_FLUX_NST, src, 4, ict + (j + LW_FLUX_NST) * LW_FLUX_NL, imin((int)LW_FLUX_NL, nlev - ict - (j + LW_FLUX_NST) * LW_FLUX_NL));
            }
            if (slot == LW_LCH_FLUX || l == nlev - 1) { flush_tile(tile, SD::RS, SD::NG, 2, slot, dst, lfirst, 1, LW_LCH_FLUX); lfirst += slot; slot = 0; }
          }
        }
      }
      jbase = nstage;
    }
    fd_surf = fd;
    if (want_dv) {
      // derivative of flux_up at each half-level w.r.t. the surface emission: sum_g flux_up_surf(g) * prod(trans)
      double* dst[1] = {s.sums + LWS_DV_A * nl1};
      const double* src[1] = {sP};
      const int nstage = (nlev + LW_FLUX_NL - 1) / LW_FLUX_NL;
      if (threadIdx.x == 0)
        for (int j = 0; j < LW_FLUX_NST && j < nstage; ++j) ring.issue(jbase + j, src, 1, j * LW_FLUX_NL, imin((int)LW_FLUX_NL, nlev - j * LW_FLUX_NL));
      int slot = 0, lfirst = 0;
      for (int j = 0; j < nstage; ++j) {
        const int l0 = j * LW_FLUX_NL, st = (jbase + j) % LW_FLUX_NST;
        double vP[LW_FLUX_NL];
        ring.wait_full(jbase + j);
#pragma unroll
        for (int k = 0; k < LW_FLUX_NL; ++k)
          if (act && l0 + k < nlev) vP[k] = ring.stage(st, 0)[k * SD::NG + g];
#pragma unroll
        for (int k = 0; k < LW_FLUX_NL; ++k) {
          const int l = l0 + k;
          if (l < nlev) {
            if (act) tile[slot * SD::RS + g] = fu * vP[k];
            ++slot;
            if (k == LW_FLUX_NL - 1 || l == nlev - 1) {
              ring.release(jbase + j);
              if (threadIdx.x == 0 && j + LW_FLUX_NST < nstage)
                ring.issue(jbase + j + LW_FLUX_NST, src, 1, (j + LW_FLUX_NST) * LW_FLUX_NL, imin((int)LW_FLUX_NL, nlev - (j + LW_FLUX_NST) * LW_FLUX_NL));
            }
            if (slot == LW_LCH_FLUX || l == nlev - 1) { flush_tile(tile, SD::RS, SD::NG, 1, slot, dst, lfirst, 1, LW_LCH_FLUX); lfirst += slot; slot = 0; }
          }
        }
      }
    }
  }
  __syncthreads();
  // ---- outputs ----
#define OUT2(p, l) ((p)[(size_t)(l) * out.ld + c])
  const double *s_dn_clear = s.sums + LWS_DN_C * nl1, *s_up_clear = s.sums + LWS_UP_C * nl1, *s_dv_clear = s.sums + LWS_DV_C * nl1;
  const double *s_up = s.sums + LWS_UP_A * nl1, *s_dn = s.sums + LWS_DN_A * nl1, *s_dv = s.sums + LWS_DV_A * nl1;
  const double tcc = s.tcc, wc = tcc, w1 = 1.0 - tcc;
  for (int l = g; l < nl1; l += SD::THREADS) {
    const double upc = s_up_clear[l], dnc = s_dn_clear[l];
    if (out.lw_up_clear) OUT2(out.lw_up_clear, l) = upc;
    if (out.lw_dn_clear) OUT2(out.lw_dn_clear, l) = dnc;
    double up = upc, dn = dnc;
    if (cloudy) {
      up = wc * s_up[l] + w1 * upc;
      dn = wc * (l <= ict ? dnc : s_dn[l]) + w1 * dnc;
    }
    if (out.lw_up) OUT2(out.lw_up, l) = up;
    if (out.lw_dn) OUT2(out.lw_dn, l) = dn;
    if (want_dv) {
      const double dclear = l == nlev ? 1.0 : s_dv_clear[l] / s_up_clear[nlev];
      double d = dclear;
      if (cloudy) {
        d = l == nlev ? 1.0 : s_dv[l] / s_up[nlev];
        if (tcc < 1.0 - s.thr) d = l == nlev ? 1.0 : (1.0 - w1) * d + w1 * dclear;
      }
      OUT2(out.lw_derivatives, l) = d;
    }
  }
  if (g == 0 && out.cloud_cover_lw && s.mcica) out.cloud_cover_lw[c] = tcc;
  const double dn_surf_g = cloudy ? wc * fd_surf + w1 * fd_surf_clear : fd_surf_clear;
  if (act) {
    const size_t i = (size_t)c * SD::NG + g;
    if (out.lw_dn_surf_clear_g) out.lw_dn_surf_clear_g[i] = fd_surf_clear;
    if (out.lw_up_toa_clear_g) out.lw_up_toa_clear_g[i] = fu_toa_clear;
    if (out.lw_dn_surf_g) out.lw_dn_surf_g[i] = dn_surf_g;
    if (out.lw_up_toa_g) out.lw_up_toa_g[i] = cloudy ? wc * fu_toa_a + w1 * fu_toa_clear : fu_toa_clear;
  }
  lw_surface_canopy<SD>(T, cfg, out, c, g, act, tile, dn_surf_g);
#undef OUT2
}

template <class SD>
static int launch_solver_lw_t(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, const Work& w, int nc, int nlev, cudaStream_t st) {
  const int nlevp = (nlev + 3) & ~3;
  const size_t sm1 = sizeof(double) * (LCH * SD::RS) + 16;
  const size_t sm2 = sizeof(double) * (3 * LW_LCH_UP * SD::RS + 2 * nlev) + sizeof(double) * LW_UP_NST * 2 * LW_UP_NL * SD::NG + 2 * LW_UP_NST * sizeof(uint64_t) + 32;
  const size_t sm3 = sizeof(double) * (2 * LW_LCH_FLUX * SD::RS + 2 * SD::NB) + sizeof(double) * LW_FLUX_NST * 4 * LW_FLUX_NL * SD::NG + 2 * LW_FLUX_NST * sizeof(uint64_t) + 32;
  lw_down_kernel<SD><<<nc, SD::THREADS, sm1, st>>>(T, cfg, out, w, nlev);
  if (cfg.solver_lw == 4 || cfg.solver_lw == 1) return 1 + launch_tc_lw(T, cfg, in, out, w, nc, nlev, st);   // Tripleclouds, Homogeneous
  cudaFuncSetAttribute(lw_up_kernel<SD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
  lw_up_kernel<SD><<<nc, SD::THREADS, sm2, st>>>(T, cfg, in, out, w, nlev, nlevp);
  cudaFuncSetAttribute(lw_flux_kernel<SD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3);
  lw_flux_kernel<SD><<<nc, SD::THREADS, sm3, st>>>(T, cfg, out, w, nlev);
  return 3;
}

int launch_solver_lw(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, const Work& w, int nc, int nlev, cudaStream_t st) {
  if (cfg.solver_lw == 3) return launch_sp_lw(T, cfg, in, out, w, nc, nlev, st);   // SPARTACUS
  if (w.layout_b_lw) return launch_solver_lw_scan(T, cfg, in, out, w, nc, nlev, st);   // McICA / Cloudless as warp scans (solver_scan.cu)
  switch (cfg.ng_lw) {
    case NG_LW: return launch_solver_lw_t<LwRrtmg>(T, cfg, in, out, w, nc, nlev, st);
    case 32: return launch_solver_lw_t<Ckd32>(T, cfg, in, out, w, nc, nlev, st);
    case 64: return launch_solver_lw_t<Ckd64>(T, cfg, in, out, w, nc, nlev, st);
    case 96: return launch_solver_lw_t<Ckd96>(T, cfg, in, out, w, nc, nlev, st);
  }
  return -1;   // check_config refuses other spectral sizes
}

}  // namespace ecb
