// kernels.cu -- sm_100a kernels of the ecRad hot path (RRTMG gas optics -> cloud optics / McICA generator ->
// McICA & Cloudless two-stream solvers).  Double precision throughout.
//
// Mapping:
//   gas_lw_kernel / gas_sw_kernel   one CTA per column; stage A = one thread per (layer, band) builds the
//                                   interpolation stencil into shared memory, stage B = one lane per g-point
//                                   evaluates it against the packed k-tables (rows contiguous in g -> coalesced).
//   cloud_prep/optics/gen kernels   one thread per column (column-fastest inputs -> coalesced).
//   solver_lw_kernel/solver_sw_kernel  one CTA per column, one thread per g-point marching the layers; layer
//                                   two-stream solutions are computed in registers, the adding-method state that
//                                   must survive between the upward and downward sweeps goes to a per-column
//                                   scratch in global memory ([layer][g], coalesced), g-point sums are done through
//                                   a shared-memory tile every LCH layers.
#include <stdio.h>

#include "cloudgen_walk.h"
#include "kernels.cuh"
#include "solver_common.cuh"

namespace ecb {

// ---------------------------------------------------------------------------------------------------------
// tight packing of the per-(layer, band) stencil slots in shared memory
// ---------------------------------------------------------------------------------------------------------
// slots per band = max number of terms rounded up to a multiple of 4 (lists are zero-padded so stage B unrolls by 4)
//                                 band:  1   2   3   4   5   6   7   8   9  10  11  12  13  14  15  16
__constant__ int c_lw_koff[NB_LW] = {0, 12, 20, 40, 56, 80, 92, 112, 128, 148, 156, 168, 184, 204, 212, 232};
enum { LW_KTOT = 248 };
//                                 band: 16  17  18  19  20  21  22  23  24  25  26  27  28  29
__constant__ int c_sw_koff[NB_SW] = {0, 12, 24, 36, 48, 60, 72, 88, 96, 112, 120, 120, 124, 132};
enum { SW_KTOT = 144 };

enum { GAS_LC = 16, GAS_THREADS = 256 };

// sum_k coef_k * tab_g[off_k] over a zero-padded list of n (multiple of 4) packed terms in shared memory.
// tab_g = table base + in-band g index (hoisted: one 64-bit pointer per item; offsets are unsigned 32-bit, so each
// address is a single IMAD.WIDE.U32).
__device__ __forceinline__ double stencil_dot(const Term* tt, int n, const double* __restrict__ tab_g) {
  // term k of this (layer, band) list is tt[k * GAS_LC]: the lists of the chunk's 16 layers are interleaved, so stage A's
  // 128-bit stores (lanes = layers) and stage B's 128-bit loads (a quarter-warp reads <= 2 distinct terms) are conflict-free
  double acc0 = 0.0, acc1 = 0.0;
  const uint4* q = reinterpret_cast<const uint4*>(tt);
  for (int k = 0; k < n; k += 4) {
    const uint4 t0 = q[k * GAS_LC], t1 = q[(k + 1) * GAS_LC], t2 = q[(k + 2) * GAS_LC], t3 = q[(k + 3) * GAS_LC];
    const double v0 = __ldg(tab_g + t0.z), v1 = __ldg(tab_g + t1.z), v2 = __ldg(tab_g + t2.z), v3 = __ldg(tab_g + t3.z);
    acc0 = fma(__hiloint2double((int)t0.y, (int)t0.x), v0, acc0);
    acc1 = fma(__hiloint2double((int)t1.y, (int)t1.x), v1, acc1);
    acc0 = fma(__hiloint2double((int)t2.y, (int)t2.x), v2, acc0);
    acc1 = fma(__hiloint2double((int)t3.y, (int)t3.x), v3, acc1);
  }
  return acc0 + acc1;
}

// =========================================================================================================
// per-layer state of both spectra (rrtm_prepare_gases + rrtm_setcoef_140gp + srtm_setcoef): one thread per (column, layer),
// column fastest so that the reference-layout inputs are read coalesced
// =========================================================================================================
// A CTA takes a tile of GP_T columns x GP_T layers: inputs are read in runs of GP_T columns (their layout), the state leaves in runs
// of GP_T layers (its layout) through a shared-memory transpose.
enum { GP_T = 16 };
__global__ void __launch_bounds__(GP_T * GP_T)
gas_prep_kernel(DevTables T, DevCfg cfg, DevIn in, Work w, int nc, int nlev) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* slw = reinterpret_cast<double*>(smem_raw);            // [GP_T columns][LWLEV_NF][GP_T layers]
  double* ssw = slw + GP_T * LWLEV_NF * GP_T;                   // [GP_T columns][SWLEV_NF][GP_T layers]
  const int tc = threadIdx.x % GP_T, tl = threadIdx.x / GP_T;
  const int c0 = blockIdx.x * GP_T, l0 = blockIdx.y * GP_T;
  const int c = c0 + tc, l = l0 + tl;
  const GasMeta& M = *T.meta;
  const bool valid = c < nc && l < nlev;
  bool sunlit = false;
  if (valid) {
    LevGas G;
    lev_prepare(LD_IN(in.p_hl, c, l), LD_IN(in.p_hl, c, l + 1), LD_IN(in.t_hl, c, l), LD_IN(in.t_hl, c, l + 1),
                LD_IN(in.gas[0], c, l), LD_IN(in.gas[1], c, l), LD_IN(in.gas[2], c, l), LD_IN(in.gas[3], c, l),
                LD_IN(in.gas[4], c, l), LD_IN(in.gas[5], c, l), LD_IN(in.gas[6], c, l), LD_IN(in.gas[7], c, l),
                LD_IN(in.gas[8], c, l), G);
    LwLev L; lw_setcoef(M, G, L);
    L.t_top = LD_IN(in.t_hl, c, l); L.t_bot = LD_IN(in.t_hl, c, l + 1); L.pad_ = 0.0;
    lwlev_store(slw + (size_t)tc * LWLEV_NF * GP_T, GP_T, tl, L);
    w.gas_jp[(size_t)c * nlev + l] = (uint8_t)(L.jp | (L.tropo << 7));   // (jp is the same expression in srtm_setcoef)
    sunlit = cfg.do_sw && in.cos_sza[c] > 0.0;
    if (sunlit) { SwLev S; sw_setcoef(M, G, S); swlev_store(ssw + (size_t)tc * SWLEV_NF * GP_T, GP_T, tl, S); }
  }
  __syncthreads();
  const int nl = imin((int)GP_T, nlev - l0);
  if (cfg.do_lw)
    for (int e = threadIdx.x; e < GP_T * LWLEV_NF * GP_T; e += GP_T * GP_T) {
      const int k = e % GP_T, f = (e / GP_T) % LWLEV_NF, cc = e / (GP_T * LWLEV_NF);
      if (c0 + cc < nc && k < nl) w.lev_lw[((size_t)(c0 + cc) * LWLEV_NF + f) * nlev + l0 + k] = slw[e];
    }
  if (cfg.do_sw)
    for (int e = threadIdx.x; e < GP_T * SWLEV_NF * GP_T; e += GP_T * GP_T) {
      const int k = e % GP_T, f = (e / GP_T) % SWLEV_NF, cc = e / (GP_T * SWLEV_NF);
      if (c0 + cc < nc && k < nl && in.cos_sza[c0 + cc] > 0.0) w.lev_sw[((size_t)(c0 + cc) * SWLEV_NF + f) * nlev + l0 + k] = ssw[e];
    }
}

// =========================================================================================================
// aerosol optics per band: add_aerosol_optics, radiation_aerosol_optics.F90:487-750 (band-wise properties, no LW aerosol
// scattering).  One thread per (column, layer); the merge into the g-point arrays happens in stage B of the gas kernels.
// =========================================================================================================
__global__ void aerosol_optics_kernel(DevTables T, DevCfg cfg, DevIn in, Work w, int nc, int nlev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc * nlev) return;
  const int c = i % nc, l = i / nc;
  const AerMeta& A = *T.aer;
  const double* tab = T.aertab;
  const double rh = LD_IN(in.gas[0], c, l) / LD_IN(in.h2o_sat_liq, c, l);
  int irh;   // calc_rh_index (1-based), radiation_aerosol_optics_data.F90:640-664
  if (rh > A.rh_lower[A.nrh - 1]) irh = A.nrh;
  else { irh = 1; while (rh > A.rh_lower[irh]) ++irh; }
  const double factor = (LD_IN(in.p_hl, c, l + 1) - LD_IN(in.p_hl, c, l)) * (1.0 / 9.80665);
  const bool lwscat = cfg.do_lw_aerosol_scattering != 0;
  const bool psw = !cfg.ckd_sw, plw = !cfg.ckd_lw;   // mixed gas models: an ecCKD spectrum merges its aerosols per g-point in its own gas kernel
  double od_sw[NB_SW], sc_sw[NB_SW], sg_sw[NB_SW], od_lw[NB_LW], sc_lw[NB_LW], sg_lw[NB_LW];
#pragma unroll
  for (int b = 0; b < NB_SW; ++b) { od_sw[b] = 0.0; sc_sw[b] = 0.0; sg_sw[b] = 0.0; }
#pragma unroll
  for (int b = 0; b < NB_LW; ++b) { od_lw[b] = 0.0; sc_lw[b] = 0.0; sg_lw[b] = 0.0; }
  for (int jt = 0; jt < A.ntype; ++jt) {
    const int iclass = A.iclass[jt];
    if (iclass == 0) continue;
    const int itype = A.itype[jt] - 1;
    const double mr = in.aerosol_mmr[((size_t)jt * nlev + l) * in.ld + c];
    const int isw = iclass == 1 ? itype * NB_SW : (itype * A.nrh + (irh - 1)) * NB_SW;
    const int ilw = iclass == 1 ? itype * NB_LW : (itype * A.nrh + (irh - 1)) * NB_LW;
    const double* me_sw = tab + (iclass == 1 ? A.me_sw_phobic : A.me_sw_philic) + isw;
    const double* ss_sw = tab + (iclass == 1 ? A.ssa_sw_phobic : A.ssa_sw_philic) + isw;
    const double* gg_sw = tab + (iclass == 1 ? A.g_sw_phobic : A.g_sw_philic) + isw;
    const double* me_lw = tab + (iclass == 1 ? A.me_lw_phobic : A.me_lw_philic) + ilw;
    const double* ss_lw = tab + (iclass == 1 ? A.ssa_lw_phobic : A.ssa_lw_philic) + ilw;
    const double* gg_lw = tab + (iclass == 1 ? A.g_lw_phobic : A.g_lw_philic) + ilw;
    if (psw) {
#pragma unroll
      for (int b = 0; b < NB_SW; ++b) {
        const double local_od = factor * mr * me_sw[b];
        od_sw[b] = od_sw[b] + local_od;
        sc_sw[b] = sc_sw[b] + local_od * ss_sw[b];
        sg_sw[b] = sg_sw[b] + local_od * ss_sw[b] * gg_sw[b];
      }
    }
    if (!plw) continue;
    if (lwscat) {   // radiation_aerosol_optics.F90:657-670, :697-710
#pragma unroll
      for (int b = 0; b < NB_LW; ++b) {
        const double local_od = factor * mr * me_lw[b];
        od_lw[b] = od_lw[b] + local_od;
        sc_lw[b] = sc_lw[b] + local_od * ss_lw[b];
        sg_lw[b] = sg_lw[b] + local_od * ss_lw[b] * gg_lw[b];
      }
    } else {
#pragma unroll
      for (int b = 0; b < NB_LW; ++b) od_lw[b] = od_lw[b] + factor * mr * me_lw[b] * (1.0 - ss_lw[b]);
    }
  }
  double* osw = w.aer_sw + ((size_t)c * nlev + l) * 3 * NB_SW;
  if (psw) {
#pragma unroll
  for (int b = 0; b < NB_SW; ++b) {
    double od = od_sw[b], sc = sc_sw[b], sg = sg_sw[b];
    if (!cfg.do_sw_delta_scaling_with_gases) {   // delta_eddington_extensive_vec, radiation_delta_eddington.h:74-96
      const double g = sg / dmax(sc, (double)1.0e-24f);
      const double f = g * g;
      od = od - sc * f;
      sc = sc * (1.0 - f);
      sg = sc * g / (1.0 + g);
    }
    osw[b] = od; osw[NB_SW + b] = sc; osw[2 * NB_SW + b] = sg;
  }
  }
  if (!plw) return;
  if (lwscat) {   // [c][l][3][16]: od, scattering od, scattering od x g after delta_eddington_extensive_vec (:778-779)
    double* olw = w.aer_lw + ((size_t)c * nlev + l) * 3 * NB_LW;
#pragma unroll
    for (int b = 0; b < NB_LW; ++b) {
      double od = od_lw[b], sc = sc_lw[b], sg = sg_lw[b];
      const double g = sg / dmax(sc, (double)1.0e-24f);
      const double f = g * g;
      od = od - sc * f;
      sc = sc * (1.0 - f);
      sg = sc * g / (1.0 + g);
      olw[b] = od; olw[NB_LW + b] = sc; olw[2 * NB_LW + b] = sg;
    }
  } else {
    double* olw = w.aer_lw + ((size_t)c * nlev + l) * NB_LW;
#pragma unroll
    for (int b = 0; b < NB_LW; ++b) olw[b] = od_lw[b];
  }
}

// =========================================================================================================
// LW gas optics
// =========================================================================================================
struct GasLwSmem {
  LwLev lev[1];  // [nlev] followed by the arrays below (carved manually)
};

template <bool WEIGHTED_EMISS>   // !do_nearest_spectral_lw_emiss (a template so that the default instantiation is the code that was tuned)
__global__ void __launch_bounds__(GAS_THREADS, 2)
gas_lw_kernel(DevTables T, DevCfg cfg, DevIn in, Work w, int nlev) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int c = blockIdx.x, tid = threadIdx.x;
  const GasMeta& M = *T.meta;
  // carve shared memory
  Term* lt = reinterpret_cast<Term*>(smem_raw);                      // [LW_KTOT][GAS_LC] (16-byte aligned)
  double* pfc = reinterpret_cast<double*>(lt + GAS_LC * LW_KTOT);    // [GAS_LC][16][2]
  double* plk = pfc + GAS_LC * NB_LW * 2;                            // [GAS_LC+1][16]
  double* plk_surf = plk + (GAS_LC + 1) * NB_LW;                     // [16]
  int* ln = reinterpret_cast<int*>(plk_surf + NB_LW);                // [GAS_LC][16]
  int* lpost = ln + GAS_LC * NB_LW;                                  // [GAS_LC][16]
  int* pfo = lpost + GAS_LC * NB_LW;                                 // [GAS_LC][16][2]
  int* bog = pfo + GAS_LC * NB_LW * 2;                               // [140] band of g
  int* g0b = bog + NG_LW;                                            // [140] in-band index of g
  int* sng = g0b + NG_LW;                                            // [16] g-points per band
  float* srn = reinterpret_cast<float*>(sng + NB_LW);                // [16] 1 / g-points per band

  // per-layer state was prepared by gas_prep_kernel; LAYTROP = number of layers with plog > 4.56
  const double* lev = w.lev_lw + (size_t)c * LWLEV_NF * nlev;
  const int tropo = tid < nlev ? (int)lev[6 * nlev + tid] : 0;
  for (int g = tid; g < NG_LW; g += GAS_THREADS) { int b = M.band_of_g_lw[g]; bog[g] = b; g0b[g] = g - M.lw[b].g0; }
  if (tid < NB_LW) { sng[tid] = M.lw[tid].ng; srn[tid] = 1.0f / (float)M.lw[tid].ng; }
  if (tid < NB_LW) plk_surf[tid] = planck_band(M, in.skin_t[c], tid);
  const int laytrop = __syncthreads_count(tropo);

  // output layout (kernels.cuh, Work): [layer][g] or [g][ls]
  const bool lb = w.layout_b_lw != 0;
  const size_t sg = lb ? (size_t)w.ls : 1, sl = lb ? 1 : (size_t)NG_LW;
  double* od_out = w.od_lw + (size_t)c * (lb ? (size_t)NG_LW * w.ls : (size_t)nlev * NG_LW);
  double* pl_out = w.planck + (size_t)c * (lb ? (size_t)NG_LW * w.ls : (size_t)(nlev + 1) * NG_LW);

  __shared__ LwLev s_lev[GAS_LC];
  for (int l0 = 0; l0 < nlev; l0 += GAS_LC) {
    const int nl = imin((int)GAS_LC, nlev - l0);
    if (tid < nl) s_lev[tid] = lwlev_load(lev, nlev, l0 + tid);   // (every band's builder of a layer reads the same state)
    __syncthreads();
    // ---- stage A: stencils of (layer, band) ----
    {
      const int b = tid >> 4, ll = tid & 15;   // GAS_LC == 16
      if (ll < nl) {
        const int l = l0 + ll;
        const int il = nlev - l;               // RRTMG layer index (1 = bottom)
        ListOut out;
        out.t = lt + c_lw_koff[b] * GAS_LC + ll;
        out.n = 0; out.stride = GAS_LC;
        int post;
        PlanckFrac pf = lw_build_list(M, M.lw[b], s_lev[ll], b, il <= laytrop, out, &post);
        out.pad4();
        ln[ll * NB_LW + b] = out.n;
        lpost[ll * NB_LW + b] = post;
        pfc[(ll * NB_LW + b) * 2] = pf.c0; pfc[(ll * NB_LW + b) * 2 + 1] = pf.c1;
        pfo[(ll * NB_LW + b) * 2] = pf.o0; pfo[(ll * NB_LW + b) * 2 + 1] = pf.o1;
      }
    }
    for (int i = tid; i < (nl + 1) * NB_LW; i += GAS_THREADS) {
      int h = i >> 4, b = i & 15;
      plk[i] = planck_band(M, LD_IN(in.t_hl, c, l0 + h), b);
    }
    __syncthreads();
    // ---- stage B: one lane per (layer, g-point), items ordered band-major: [band][layer][g in band], so that the
    //      whole CTA works on one or two bands at a time (table rows stay in L1, lanes of a warp share the term count)
    const int items = nl * NG_LW;
    for (int it = tid; it < items; it += GAS_THREADS) {
      const int gq = nl == GAS_LC ? it >> 4 : it / nl;
      const int b = bog[gq];
      const int g0 = gq - g0b[gq];
      const int j = it - nl * g0;
      const int ll = __float2int_rz(((float)j + 0.5f) * srn[b]);
      const int igb = j - ll * sng[b];
      const int g = g0 + igb;
      const int l = l0 + ll;
      const int n = ln[ll * NB_LW + b];
      const double* tab_g = T.lwtab + igb;
      double tau = stencil_dot(lt + c_lw_koff[b] * GAS_LC + ll, n, tab_g);
      const int post = lpost[ll * NB_LW + b];
      if (post >= 0) tau *= __ldg(tab_g + (unsigned)post);
      const double pf = pfc[(ll * NB_LW + b) * 2] * __ldg(tab_g + (unsigned)pfo[(ll * NB_LW + b) * 2]) +
                        pfc[(ll * NB_LW + b) * 2 + 1] * __ldg(tab_g + (unsigned)pfo[(ll * NB_LW + b) * 2 + 1]);
      double odv = dmax(tau, cfg.min_gas_od_lw);                           // radiation_ifs_rrtm.F90:506-511
      if (cfg.use_aerosols) odv = odv + w.aer_lw[((size_t)c * nlev + l) * NB_LW + b];   // radiation_aerosol_optics.F90:806-812
      od_out[l * sl + g * sg] = odv;
      pl_out[(l + 1) * sl + g * sg] = plk[(ll + 1) * NB_LW + b] * pf;      // half-level below uses this layer's PFRAC
      if (l == 0) pl_out[g * sg] = plk[b] * pf;                            // TOA half-level: PFRAC of the top layer
      if (l == nlev - 1) {
        // surface: planck_function_surf :757-852, lw_emission = planck_surf * (1 - lw_albedo) :466
        double alb;
        if (!WEIGHTED_EMISS) {
          alb = 1.0 - LD_IN(in.lw_emissivity, c, T.i_emiss_from_band_lw[b] - 1);
        } else {   // weighted emissivity intervals (get_albedos, radiation_single_level.F90:330-352)
          alb = 0.0;
          for (int ja = 0; ja < cfg.n_emiss_lw; ++ja) {
            const double wgt = T.lw_emiss_weights[b * cfg.n_emiss_lw + ja];
            if (wgt != 0.0) alb = alb + wgt * (1.0 - LD_IN(in.lw_emissivity, c, ja));
          }
        }
        w.lw_albedo[(size_t)c * NG_LW + g] = alb;
        double em = plk_surf[b] * pf;
        w.emission[(size_t)c * NG_LW + g] = em * (1.0 - alb);
      }
    }
    __syncthreads();
  }
}

// =========================================================================================================
// SW gas optics (sunlit columns only)
// =========================================================================================================
__global__ void __launch_bounds__(GAS_THREADS, 3)
gas_sw_kernel(DevTables T, DevCfg cfg, DevIn in, Work w, int nlev) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int c = blockIdx.x, tid = threadIdx.x;
  if (!(in.cos_sza[c] > 0.0)) return;   // srtm only runs for sunlit columns (radiation_ifs_rrtm.F90:518-542)
  const GasMeta& M = *T.meta;
  Term* lt = reinterpret_cast<Term*>(smem_raw);                      // [SW_KTOT][GAS_LC] (16-byte aligned)
  double* rc = reinterpret_cast<double*>(lt + GAS_LC * SW_KTOT);     // [GAS_LC][14][2] Rayleigh coefficients
  double* sc = rc + GAS_LC * NB_SW * 2;                              // [14][2]  solar-source coefficients
  double* inc = sc + NB_SW * 2;                                      // [112]
  int* ln = reinterpret_cast<int*>(inc + NG_SW);                     // [GAS_LC][14]
  int* ro = ln + GAS_LC * NB_SW;                                     // [GAS_LC][14][2]
  int* so = ro + GAS_LC * NB_SW * 2;                                 // [14][2]
  int* lsol = so + NB_SW * 2;                                        // [14] ecRad layer index supplying the solar source (-1: none)
  int* bog = lsol + NB_SW;                                           // [112]
  int* g0b = bog + NG_SW;                                            // [112]
  int* sng = g0b + NG_SW;                                            // [14]
  float* srn = reinterpret_cast<float*>(sng + NB_SW);                // [14]
  int* jps = reinterpret_cast<int*>(srn + NB_SW);                    // [nlev] JP per layer (ecRad order)
  __shared__ double s_scale;

  const double* lev = w.lev_sw + (size_t)c * SWLEV_NF * nlev;
  int tropo = 0;
  if (tid < nlev) { jps[tid] = (int)lev[tid]; tropo = (int)lev[5 * nlev + tid]; }
  for (int g = tid; g < NG_SW; g += GAS_THREADS) { int b = M.band_of_g_sw[g]; bog[g] = b; g0b[g] = g - M.sw[b].g0; inc[g] = 0.0; }
  if (tid < NB_SW) { sng[tid] = M.sw[tid].ng; srn[tid] = 1.0f / (float)M.sw[tid].ng; }
  const int laytrop = __syncthreads_count(tropo);
  if (tid < NB_SW) {
    int il = sw_solar_layer(M, tid, nlev, laytrop, [&](int i) { return jps[nlev - i]; });
    lsol[tid] = il > 0 ? nlev - il : -1;
  }
  __syncthreads();

  const bool lb = w.layout_b_sw != 0;
  const size_t sg = lb ? (size_t)w.ls : 1, sl = lb ? 1 : (size_t)NG_SW;
  const size_t cbase = (size_t)c * (lb ? (size_t)NG_SW * w.ls : (size_t)nlev * NG_SW);
  double* od_out = w.od_sw + cbase;
  double* ssa_out = w.ssa_sw + cbase;

  __shared__ SwLev s_lev[GAS_LC];
  for (int l0 = 0; l0 < nlev; l0 += GAS_LC) {
    const int nl = imin((int)GAS_LC, nlev - l0);
    if (tid < nl) s_lev[tid] = swlev_load(lev, nlev, l0 + tid);
    __syncthreads();
    {
      const int b = tid >> 4, ll = tid & 15;
      if (b < NB_SW && ll < nl) {
        const int l = l0 + ll;
        const int il = nlev - l;
        ListOut out;
        out.t = lt + c_sw_koff[b] * GAS_LC + ll;
        out.n = 0; out.stride = GAS_LC;
        SwAux aux;
        sw_build_list(M, M.sw[b], s_lev[ll], b, il <= laytrop, out, aux);
        out.pad4();
        ln[ll * NB_SW + b] = out.n;
        rc[(ll * NB_SW + b) * 2] = aux.rc0; rc[(ll * NB_SW + b) * 2 + 1] = aux.rc1;
        ro[(ll * NB_SW + b) * 2] = aux.ro0; ro[(ll * NB_SW + b) * 2 + 1] = aux.ro1;
        if (l == lsol[b]) { sc[b * 2] = aux.sc0; sc[b * 2 + 1] = aux.sc1; so[b * 2] = aux.so0; so[b * 2 + 1] = aux.so1; }
      }
    }
    __syncthreads();
    const int items = nl * NG_SW;   // band-major item order, as in gas_lw_kernel
    for (int it = tid; it < items; it += GAS_THREADS) {
      const int gq = nl == GAS_LC ? it >> 4 : it / nl;
      const int b = bog[gq];
      const int g0 = gq - g0b[gq];
      const int j = it - nl * g0;
      const int ll = __float2int_rz(((float)j + 0.5f) * srn[b]);
      const int igb = j - ll * sng[b];
      const int g = g0 + igb;
      const int l = l0 + ll;
      const int n = ln[ll * NB_SW + b];
      const double* tab_g = T.swtab + igb;
      const double taug = stencil_dot(lt + c_sw_koff[b] * GAS_LC + ll, n, tab_g);
      const double taur = rc[(ll * NB_SW + b) * 2] * __ldg(tab_g + (unsigned)ro[(ll * NB_SW + b) * 2]) +
                          rc[(ll * NB_SW + b) * 2 + 1] * __ldg(tab_g + (unsigned)ro[(ll * NB_SW + b) * 2 + 1]);
      const double od = taur + taug;                                        // srtm_gas_optical_depth.F90:314-320
      double odv = dmax(od, cfg.min_gas_od_sw), ssav = taur / od;          // radiation_ifs_rrtm.F90:593
      if (cfg.use_aerosols) {
        // merge aerosol and gas per g-point: radiation_aerosol_optics.F90:765-781
        const double* a = w.aer_sw + ((size_t)c * nlev + l) * 3 * NB_SW;
        const double od_a = a[b], local_od = odv + od_a;
        double gv = 0.0;
        if (local_od > 0.0 && od_a > 0.0) {
          const double local_scat = ssav * odv + a[NB_SW + b];
          if (local_scat > 0.0) gv = a[2 * NB_SW + b] / local_scat;
          ssav = local_scat / local_od;
          odv = local_od;
        }
        w.g_sw[cbase + l * sl + g * sg] = gv;
      }
      od_out[l * sl + g * sg] = odv;
      ssa_out[l * sl + g * sg] = ssav;
      if (l == lsol[b]) inc[g] = sc[b * 2] * __ldg(tab_g + (unsigned)so[b * 2]) + sc[b * 2 + 1] * __ldg(tab_g + (unsigned)so[b * 2 + 1]);
    }
    __syncthreads();
  }
  // incoming_sw = ZINCSOL * solar_irradiance / sum(ZINCSOL): radiation_ifs_rrtm.F90:557-605
  if (tid == 0) {
    double s = 0.0;
    for (int g = 0; g < NG_SW; ++g) s = s + inc[g];
    s_scale = in.solar_irradiance / s;
  }
  __syncthreads();
  for (int g = tid; g < NG_SW; g += GAS_THREADS) w.incoming[(size_t)c * NG_SW + g] = s_scale * inc[g];
}

// =========================================================================================================
// clouds: crop, cumulative cover, band optics, McICA generator
// =========================================================================================================
__global__ void cloud_prep_kernel(DevCfg cfg, DevIn in, Work w, int nc, int nlev) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  // crop_cloud_fraction, radiation_cloud.F90:700-740 (in place)
  int ict = nlev;
  for (int l = 0; l < nlev; ++l) {
    double f = LD_IN(in.frac, c, l);
    double sum_mr = 0.0;
    sum_mr = sum_mr + LD_IN(in.q_liq, c, l);
    sum_mr = sum_mr + LD_IN(in.q_ice, c, l);
    if (f < cfg.cloud_fraction_threshold || sum_mr < cfg.cloud_mixing_ratio_threshold) { f = 0.0; LD_IN(in.frac, c, l) = 0.0; }
    if (f >= cfg.cloud_fraction_threshold && ict == nlev) ict = l;
  }
  int ibegin, iend;
  double tcc = gen_prepare(cfg.overlap_scheme, nlev, in.ld, nc, in.frac + c, in.overlap + c, cfg.use_beta_overlap != 0,
                           cfg.cloud_inhom_decorr_scaling, cfg.cloud_fraction_threshold, w.cum + c, w.pair + c, w.opi + c,
                           &ibegin, &iend);
  w.tcc[c] = tcc;
  w.ibegin[c] = ibegin; w.iend[c] = iend; w.ict[c] = ict;
}

__global__ void cloud_optics_kernel(DevTables T, DevCfg cfg, DevIn in, Work w, int nc, int nlev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc * nlev) return;
  const int c = i % nc, l = i / nc;
  const double frac = LD_IN(in.frac, c, l);
  if (!(frac > 0.0)) return;
  const CloudMeta& C = *T.cloud;
  // config%is_homogeneous (radiation_cloud_optics.F90:318-327): gridbox-mean water path for the Homogeneous solvers
  const double factor = cfg.is_homogeneous ? (LD_IN(in.p_hl, c, l + 1) - LD_IN(in.p_hl, c, l)) / 9.80665
                                           : (LD_IN(in.p_hl, c, l + 1) - LD_IN(in.p_hl, c, l)) / (9.80665 * frac);
  CloudLayerIn L;
  L.q_ice = LD_IN(in.q_ice, c, l);
  L.lwp = factor * LD_IN(in.q_liq, c, l); L.iwp = factor * L.q_ice;
  L.re_liq = LD_IN(in.re_liq, c, l); L.re_ice = LD_IN(in.re_ice, c, l);
  L.temperature = 0.5 * (LD_IN(in.t_hl, c, l) + LD_IN(in.t_hl, c, l + 1));
  if (cfg.do_lw) {
    double* o = w.cl_lw + ((size_t)c * nlev + l) * 3 * NB_LW;
    for (int b = 0; b < NB_LW; ++b) {
      CloudBandOut r = cloud_optics_lw(C, b, L, cfg.do_lw_cloud_scattering != 0, cfg.do_fu_lw_ice_optics_bug != 0);
      o[b] = r.od; o[NB_LW + b] = r.ssa; o[2 * NB_LW + b] = r.g;
    }
  }
  if (cfg.do_sw) {
    double* o = w.cl_sw + ((size_t)c * nlev + l) * 3 * NB_SW;
    for (int b = 0; b < NB_SW; ++b) {
      CloudBandOut r = cloud_optics_sw(C, b, L, cfg.do_sw_delta_scaling_with_gases != 0);
      o[b] = r.od; o[NB_SW + b] = r.ssa; o[2 * NB_SW + b] = r.g;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// McICA cloud generator, warp-cooperative.  One CTA (2 warps) per column: warp 0 = SW stream (seed iseed), warp 1 =
// LW stream (seed iseed + 997, radiation_mcica_lw.F90:223).  Same random stream, same decisions and same stream
// consumption as radiation_cloud_generator.F90:202-390 + utilities/radiation_random_numbers_mix.F90 (the sequential
// specification is gen_walk() in cloudgen_walk.h, which tests/ replays against the oracle), restructured so that
// the 32 lanes work on one sub-column together:
//   * seeding: the 29 x 604 serial LFSR steps of initialize_random_numbers are independent per bit plane once the
//     LFSR (u' = u*x mod x^32+x^7+x^5+x^3+x^2+x+1 over GF(2)) is jumped ahead -> lane b fills bit plane b+1, a ballot
//     assembles the words;
//   * the lagged-Fibonacci stream s[t] = (s[t-607] + s[t-273]) mod 2^30 is produced 32 numbers per step into a
//     1024-entry ring in shared memory and addressed by absolute stream position;
//   * the cloudy/clear walk down the layers is a 2-state automaton: each layer's transition is an affine map over
//     GF(2), composed with a warp prefix scan; run starts/ends, the offsets of the per-run random draws and the
//     "reuse the previous layer's number" chain are further warp scans.
// Output: code[g][layer] = 0 (clear) or 0x80000000 | rand30, consumed lane-parallel by the solvers (pdf_sample).
// ---------------------------------------------------------------------------------------------------------
enum { GW_MAXLEV = 256, GW_LPL = 8 };   // 32 lanes x up to 8 layers

__constant__ uint32_t c_lfsr_jump[29];   // x^(604 b) mod P, b = 0..28

__host__ __device__ __forceinline__ uint32_t lfsr_mul_x(uint32_t u) { return (u << 1) ^ ((u >> 31) ? 175u : 0u); }
__host__ __device__ __forceinline__ uint32_t gf2_mulmod(uint32_t a, uint32_t b) {
  uint32_t r = 0;
  for (int i = 31; i >= 0; --i) { r = lfsr_mul_x(r); if ((b >> i) & 1u) r ^= a; }
  return r;
}
void init_generator_constants() {
  uint32_t x604 = 1u;
  for (int i = 0; i < 604; ++i) x604 = lfsr_mul_x(x604);
  uint32_t c[29];
  c[0] = 1u;
  for (int b = 1; b < 29; ++b) c[b] = gf2_mulmod(c[b - 1], x604);
  cudaMemcpyToSymbol(c_lfsr_jump, c, sizeof(c));
}

// One sub-column (g-point) of the walk, the 32 lanes covering the cloudy range Lb..Le with LPL consecutive layers each
// (LPL = ceil((Le-Lb+1)/32): the median cloudy range of IFS columns is ~40 layers, so most columns run with LPL = 1 or 2
// instead of the 5 that cover a whole 137-level profile).
__device__ __forceinline__ void gen_to(int32_t* ring, int& tgen, int lane, int target) {
  while (tgen < target) {
    const int t = tgen + lane;
    ring[t & 1023] = 0x3FFFFFFF & (ring[(t - 607) & 1023] + ring[(t - 273) & 1023]);
    __syncwarp();
    tgen += 32;
  }
}

template <int LPL>
__device__ __forceinline__ void gen_walk_warp(const DevCfg& cfg, int32_t* ring, const int32_t* rtop, uint32_t* code, int ng, int nlevp, int lane,
                                              int Lb, int Le, double tcc, const double* sA1, const double* sT1, const double* sA2,
                                              const double* sT2, const double* sCUM, const double* sOPI, int& tgen, int& pos) {
  const unsigned FULL = 0xffffffffu;
  const double RM = 1.0 / 1073741824.0;
  const uint32_t MASK = (1u << LPL) - 1u;
  const int L0 = Lb + lane * LPL;   // first layer of this lane
  for (int g = 0; g < ng; ++g) {
    gen_to(ring, tgen, lane, pos + 3 * GW_MAXLEV);   // one sub-column consumes at most 3 numbers per layer
    // ---- cloud-top trigger: first layer whose cumulative cover reaches rand_top*total_cloud_cover ----
    const double trigger = mul_rn((double)rtop[g] * RM, tcc);
    int first = 1 << 30;
#pragma unroll
    for (int m = LPL - 1; m >= 0; --m) {
      const int L = L0 + m;
      if (L >= Lb && L <= Le && !(trigger > sCUM[L])) first = L;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) first = min(first, __shfl_xor_sync(FULL, first, d));
    const int Lt = min(first, Le);
    const int N = Le - Lt + 1;
    // ---- transition maps s' = a*s ^ b of every layer (s = 1: cloudy) ----
    uint32_t am = 0, bm = 0;
#pragma unroll
    for (int m = 0; m < LPL; ++m) {
      const int L = L0 + m;
      uint32_t a = 0, b = 0;
      if (L == Lt) b = 1;
      else if (L > Lt && L <= Le) {
        const double r = (double)ring[(pos + (L - Lt - 1)) & 1023] * RM;
        const bool c1 = mul_rn(r, sA1[L]) < sT1[L];   // cloudy above -> stays cloudy
        const bool c2 = mul_rn(r, sA2[L]) < sT2[L];   // clear above  -> becomes cloudy
        a = (uint32_t)(c1 != c2); b = (uint32_t)c2;
      }
      am |= a << m; bm |= b << m;
    }
    uint32_t A = 1, B = 0;
#pragma unroll
    for (int m = 0; m < LPL; ++m) { const uint32_t a = (am >> m) & 1u, b = (bm >> m) & 1u; B = (a & B) ^ b; A = a & A; }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t Ap = __shfl_up_sync(FULL, A, d), Bp = __shfl_up_sync(FULL, B, d);
      if (lane >= d) { B = (A & Bp) ^ B; A = A & Ap; }
    }
    uint32_t sin = __shfl_up_sync(FULL, B, 1);
    if (lane == 0) sin = 0;
    uint32_t cl = 0;
    {
      uint32_t s = sin;
#pragma unroll
      for (int m = 0; m < LPL; ++m) { s = (((am >> m) & 1u) & s) ^ ((bm >> m) & 1u); cl |= s << m; }
    }
    // ---- number of cloudy layers above each layer ----
    const int cnt = __popc(cl);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += t; }
    const int C = __shfl_sync(FULL, incl, 31);
    const int cntb_lane = incl - cnt;
    // ---- first and last layer of the contiguous cloudy run each layer belongs to ----
    const uint32_t isstart = cl & ~((cl << 1) | sin) & MASK;
    int sc = isstart ? L0 + (31 - __clz((int)isstart)) : -1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULL, sc, d); if (lane >= d) sc = max(sc, t); }
    int rs_in = __shfl_up_sync(FULL, sc, 1);
    if (lane == 0) rs_in = -1;
    uint32_t nextfirst = __shfl_down_sync(FULL, cl & 1u, 1);
    if (lane == 31) nextfirst = 0;
    const uint32_t isend = cl & ~((cl >> 1) | (nextfirst << (LPL - 1))) & MASK;
    int ec = isend ? L0 + (__ffs((int)isend) - 1) : (1 << 30);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_down_sync(FULL, ec, d); if (lane + d < 32) ec = min(ec, t); }
    int re_in = __shfl_down_sync(FULL, ec, 1);
    if (lane == 31) re_in = 1 << 30;
    int rsv[LPL], rev[LPL];
    {
      int cur = rs_in;
#pragma unroll
      for (int m = 0; m < LPL; ++m) { if ((isstart >> m) & 1u) cur = L0 + m; rsv[m] = cur; }
      cur = re_in;
#pragma unroll
      for (int m = LPL - 1; m >= 0; --m) { if ((isend >> m) & 1u) cur = L0 + m; rev[m] = cur; }
    }
    // ---- per-run random draws: run r takes n numbers (rand_inhom1) then n numbers (rand_inhom2), runs in order ----
    // (Exp-Exp, generate_column_exp_exp: one "run" = the whole range Lt..Le whether cloudy or not: N + N numbers)
    const bool exp_exp = cfg.overlap_scheme == 2;
    int offv[LPL];
    uint32_t fresh = 0;
#pragma unroll
    for (int m = 0; m < LPL; ++m) {
      offv[m] = 0;
      if (exp_exp) {
        const int L = L0 + m;
        if (L >= Lt && L <= Le) {
          rsv[m] = Lt;
          offv[m] = pos + N;
          const double r2 = (double)ring[(pos + 2 * N + (L - Lt)) & 1023] * RM;
          if (L == Lt || !(r2 < sOPI[L])) fresh |= 1u << m;
        }
      } else if ((cl >> m) & 1u) {
        const int L = L0 + m;
        const int p = L - rsv[m], n = rev[m] - rsv[m] + 1;
        const int cb = cntb_lane + __popc(cl & ((1u << m) - 1u));
        const int off = pos + N + 2 * (cb - p);
        offv[m] = off;
        const double r2 = (double)ring[(off + n + p) & 1023] * RM;
        if (p == 0 || !(r2 < sOPI[L])) fresh |= 1u << m;   // else: reuse the number of the layer above
      }
    }
    int qc = fresh ? L0 + (31 - __clz((int)fresh)) : -1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULL, qc, d); if (lane >= d) qc = max(qc, t); }
    int q_in = __shfl_up_sync(FULL, qc, 1);
    if (lane == 0) q_in = -1;
    uint32_t* row = code + (size_t)g * nlevp;
    {
      int cur = q_in;
#pragma unroll
      for (int m = 0; m < LPL; ++m) {
        const int L = L0 + m;
        uint32_t word = 0;
        if ((fresh >> m) & 1u) cur = L;
        if ((cl >> m) & 1u) word = 0x80000000u | (uint32_t)ring[(offv[m] + (cur - rsv[m])) & 1023];
        if (L < nlevp) row[L] = word;
      }
    }
    pos += exp_exp ? 3 * N : N + 2 * C;
    __syncwarp();
    // layers outside the range the lanes cover are clear
    for (int L = lane; L < nlevp; L += 32) if (L < Lb || L >= Lb + 32 * LPL) row[L] = 0u;
  }
}

__global__ void __launch_bounds__(64)
cloud_gen_warp_kernel(DevCfg cfg, DevIn in, Work w, int nc, int nlev, int nlevp, int nlev8) {
  extern __shared__ __align__(16) unsigned char gen_smem[];
  double* sA1 = reinterpret_cast<double*>(gen_smem);   // six arrays of nlev8 = nlev rounded up to a multiple of 8 (<= GW_MAXLEV)
  double *sT1 = sA1 + nlev8, *sA2 = sT1 + nlev8, *sT2 = sA2 + nlev8, *sCUM = sT2 + nlev8, *sOPI = sCUM + nlev8;
  __shared__ int32_t sRing[2][1024];
  __shared__ int32_t sTop[2][NG_LW];
  const unsigned FULL = 0xffffffffu;
  const int c = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double tcc = w.tcc[c];
  if (!(tcc > 0.0)) return;
  const int Lb = w.ibegin[c] - 1, Le = w.iend[c] - 1;   // first / last cloudy layer, 0-based
  // per-layer constants of the two transition tests (same operations, same order as the reference's expressions)
  for (int L = threadIdx.x; L < nlev8; L += 64) {
    double a1 = 0.0, t1 = 0.0, a2 = 0.0, t2 = 0.0, cu = 0.0, op = 0.0;
    if (L < nlev) {
      cu = w.cum[(size_t)L * nc + c];
      if (L >= 1) {
        const double f = LD_IN(in.frac, c, L), fp = LD_IN(in.frac, c, L - 1);
        const double pr = w.pair[(size_t)(L - 1) * nc + c], cup = w.cum[(size_t)(L - 1) * nc + c];
        a1 = fp;                                            // rand*frac(jlev-1) < frac(jlev)+frac(jlev-1)-pair(jlev-1)
        t1 = sub_rn(add_rn(f, fp), pr);
        a2 = sub_rn(cup, fp);                               // rand*(cum(jlev-1)-frac(jlev-1)) < pair(jlev-1)-overhang(jlev-1)-frac(jlev-1)
        t2 = sub_rn(sub_rn(pr, sub_rn(cu, cup)), fp);
        op = w.opi[(size_t)(L - 1) * nc + c];               // overlap_param_inhom between layers L-1 and L
      }
    }
    sA1[L] = a1; sT1[L] = t1; sA2[L] = a2; sT2[L] = t2; sCUM[L] = cu; sOPI[L] = op;
  }
  __syncthreads();
  const int spec = warp;
  const bool mine = spec == 0 ? (cfg.do_sw && cfg.solver_sw == 2 && in.cos_sza[c] > 0.0) : (cfg.do_lw && cfg.solver_lw == 2);
  if (!mine) return;
  const int ng = spec ? cfg.ng_lw : cfg.ng_sw;
  uint32_t* code = (spec ? w.code_lw : w.code_sw) + (size_t)c * ng * nlevp;
  int32_t* ring = sRing[warp];
  int32_t* rtop = sTop[warp];
  const double RM = 1.0 / 1073741824.0;

  // ---- initialize_random_numbers (radiation_random_numbers_mix.F90:142-231) ----
  {
    const int32_t JPMASK = 123459876;
    int32_t idum = (in.iseed[c] + (spec ? 997 : 0)) ^ JPMASK;
    if (idum < 0) idum = (idum == INT32_MIN) ? idum : -idum;
    if (idum == 0) idum = JPMASK;
    uint32_t u = (uint32_t)idum;
    for (int k = 0; k < 64; ++k) u = lfsr_mul_x(u);
    for (int i = lane; i < 1024; i += 32) ring[i] = 0;
    __syncwarp();
    // state word ix(jj) is stream element s[jj-608]; ring index = stream index mod 1024
    if (lane == 0) {
      ring[(2 - 608) & 1023] = (int32_t)((u & ((1u << (JPMM - 1)) - 1u)) << 1);
      ring[(607 - 608) & 1023] = (int32_t)((u >> (JPMM - 1)) & 7u);
    }
    uint32_t v = lane < 29 ? gf2_mulmod(u, c_lfsr_jump[lane]) : 0u;
    for (int jj = 3; jj <= JPQ - 1; ++jj) {
      const uint32_t word = __ballot_sync(FULL, v >> 31);     // bit b = top bit of plane b+1 before the step
      v = lfsr_mul_x(v);
      if (lane == (jj & 31)) ring[(jj - 608) & 1023] = (int32_t)((word & 0x1FFFFFFFu) << 1);
    }
    __syncwarp();
    if (lane == 0) ring[(JPQ - JPS - 608) & 1023] |= 1;
    __syncwarp();
  }
  int tgen = 0;   // stream elements [tgen-607, tgen) are in the ring
  gen_to(ring, tgen, lane, 999 + ng);                                   // 999 warm-up numbers, then rand_top(1:ng)
  for (int g = lane; g < ng; g += 32) rtop[g] = ring[(999 + g) & 1023];
  __syncwarp();
  int pos = 999 + ng;

  const int lpl = (Le - Lb + 1 + 31) / 32;
  switch (lpl) {
    case 1: gen_walk_warp<1>(cfg, ring, rtop, code, ng, nlevp, lane, Lb, Le, tcc, sA1, sT1, sA2, sT2, sCUM, sOPI, tgen, pos); break;
    case 2: gen_walk_warp<2>(cfg, ring, rtop, code, ng, nlevp, lane, Lb, Le, tcc, sA1, sT1, sA2, sT2, sCUM, sOPI, tgen, pos); break;
    case 3: gen_walk_warp<3>(cfg, ring, rtop, code, ng, nlevp, lane, Lb, Le, tcc, sA1, sT1, sA2, sT2, sCUM, sOPI, tgen, pos); break;
    case 4: gen_walk_warp<4>(cfg, ring, rtop, code, ng, nlevp, lane, Lb, Le, tcc, sA1, sT1, sA2, sT2, sCUM, sOPI, tgen, pos); break;
    case 5: gen_walk_warp<5>(cfg, ring, rtop, code, ng, nlevp, lane, Lb, Le, tcc, sA1, sT1, sA2, sT2, sCUM, sOPI, tgen, pos); break;
    case 6: gen_walk_warp<6>(cfg, ring, rtop, code, ng, nlevp, lane, Lb, Le, tcc, sA1, sT1, sA2, sT2, sCUM, sOPI, tgen, pos); break;
    default: gen_walk_warp<8>(cfg, ring, rtop, code, ng, nlevp, lane, Lb, Le, tcc, sA1, sT1, sA2, sT2, sCUM, sOPI, tgen, pos); break;
  }
}

// ---------------------------------------------------------------------------------------------------------
// "Vectorizable" McICA generator (use_vectorizable_generator): radiation_cloud_generator.F90:587-734 with the vector
// MINSTD generator of radiation_random_numbers.F90 -- one independent Lehmer stream s' = 48271 s mod (2^31-1) per g-point,
// and a fixed consumption pattern: 1 number for the cloud-top trigger, one block per cloudy layer (rand_cloud), one
// block per layer of ibegin-1..iend (rand_inhom), one block per cloudy layer (rand_inhom2).  On a GPU that is one thread per
// (spectrum, g-point): the three blocks a layer needs are three positions of the thread's own stream, reached by jump-ahead
// (A^k mod M), so the walk down the layers needs no communication at all.  Output code word = the 31-bit stream state
// of the number that sets the optical-depth scaling (never 0), 0 = clear.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t minstd_next(uint32_t s) { return (uint32_t)((48271ull * s) % 2147483647ull); }
__device__ __forceinline__ uint32_t minstd_mulmod(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) % 2147483647ull); }
__device__ __forceinline__ uint32_t minstd_pow(int k) {   // 48271^k mod (2^31-1)
  uint32_t r = 1u, a = 48271u;
  while (k > 0) { if (k & 1) r = minstd_mulmod(r, a); a = minstd_mulmod(a, a); k >>= 1; }
  return r;
}

__global__ void __launch_bounds__(320)
cloud_gen_vec_kernel(DevCfg cfg, DevIn in, Work w, int nc, int nlev, int nlevp, int threads_sw) {
  __shared__ double sA1[GW_MAXLEV], sT1[GW_MAXLEV], sA2[GW_MAXLEV], sT2[GW_MAXLEV], sCUM[GW_MAXLEV], sOPI[GW_MAXLEV];
  __shared__ unsigned char sCLD[GW_MAXLEV];
  __shared__ int sNcl;
  const int c = blockIdx.x;
  const double tcc = w.tcc[c];
  if (!(tcc > 0.0)) return;
  const int Lb = w.ibegin[c] - 1, Le = w.iend[c] - 1;   // first / last layer with cloud, 0-based
  for (int L = threadIdx.x; L < GW_MAXLEV; L += blockDim.x) {   // same per-layer constants as cloud_gen_warp_kernel
    double a1 = 0.0, t1 = 0.0, a2 = 0.0, t2 = 0.0, cu = 0.0, op = 0.0;
    unsigned char cld = 0;
    if (L < nlev) {
      cu = w.cum[(size_t)L * nc + c];
      const double f = LD_IN(in.frac, c, L);
      cld = (L >= Lb && L <= Le && f >= cfg.cloud_fraction_threshold) ? 1 : 0;
      if (L >= 1) {
        const double fp = LD_IN(in.frac, c, L - 1);
        const double pr = w.pair[(size_t)(L - 1) * nc + c], cup = w.cum[(size_t)(L - 1) * nc + c];
        a1 = fp; t1 = sub_rn(add_rn(f, fp), pr);
        a2 = sub_rn(cup, fp); t2 = sub_rn(sub_rn(pr, sub_rn(cu, cup)), fp);
        op = w.opi[(size_t)(L - 1) * nc + c];
      }
    }
    sA1[L] = a1; sT1[L] = t1; sA2[L] = a2; sT2[L] = t2; sCUM[L] = cu; sOPI[L] = op; sCLD[L] = cld;
  }
  __syncthreads();
  if (threadIdx.x == 0) { int n = 0; for (int L = Lb; L <= Le; ++L) n += sCLD[L]; sNcl = n; }
  __syncthreads();
  const int spec = (int)threadIdx.x >= threads_sw;   // 0: SW stream (seed iseed), 1: LW stream (seed iseed + 997)
  const int g = spec ? (int)threadIdx.x - threads_sw : (int)threadIdx.x;
  const int ng = spec ? cfg.ng_lw : cfg.ng_sw;
  const bool mine = spec == 0 ? (cfg.do_sw && cfg.solver_sw == 2 && in.cos_sza[c] > 0.0) : (cfg.do_lw && cfg.solver_lw == 2);
  if (!mine || g >= ng) return;
  uint32_t* row = (spec ? w.code_lw : w.code_sw) + ((size_t)c * ng + g) * nlevp;
  const double SCALE = 1.0 / 2147483647.0;
  // rng_type%initialize (radiation_random_numbers.F90:96-150), stream jstr = g + 1
  const int32_t seed = in.iseed[c] + (spec ? 997 : 0);
  const double rseed = fabs((double)seed);
  const int jstr = g + 1;
  const double x = mul_rn(mul_rn(mul_rn(rseed, (double)jstr), add_rn(sub_rn(1.0, mul_rn(0.05, (double)jstr)), mul_rn(0.005, (double)(jstr * jstr)))), 16807.0);
  uint32_t s = (uint32_t)llround(fmod(x, 2147483647.0));
  s = minstd_next(s);
  s = minstd_next(s);                                   // trigger(jg)
  const double trigger = mul_rn((double)s * SCALE, tcc);
  const int n = Le - Lb + 1, ncl = sNcl;
  uint32_t s_rc = s;                                    // rand_cloud: one number per cloudy layer
  uint32_t s_ri = minstd_mulmod(s, minstd_pow(ncl));    // rand_inhom: layers ibegin-1 .. iend
  uint32_t s_r2 = minstd_mulmod(s, minstd_pow(ncl + n + 1));   // rand_inhom2: one number per cloudy layer
  s_ri = minstd_next(s_ri);                             // rand_inhom(ibegin-1)
  uint32_t ri_above = s_ri;
  bool is_cloud = false, found = false;
  for (int L = 0; L < Lb && L < nlevp; ++L) row[L] = 0u;
  for (int L = Lb; L <= Le; ++L) {
    s_ri = minstd_next(s_ri);
    uint32_t ri = s_ri, word = 0u;
    if (sCLD[L]) {
      s_rc = minstd_next(s_rc); s_r2 = minstd_next(s_r2);
      const bool prev = is_cloud;
      const bool first = !(trigger > sCUM[L]) && !found;
      found = found || first;
      const double rc = (double)s_rc * SCALE;
      const bool test = prev ? (mul_rn(rc, sA1[L]) < sT1[L]) : (mul_rn(rc, sA2[L]) < sT2[L]);
      is_cloud = first || (found && L >= 1 && test);
      if (is_cloud) {
        if (L >= 1 && prev && ((double)s_r2 * SCALE < sOPI[L])) ri = ri_above;
        word = ri;
      }
    } else {
      is_cloud = false;
    }
    ri_above = is_cloud ? ri : s_ri;   // rand_inhom(jg, jlev) as left in the array (a clear layer's value is never reused)
    row[L] = word;
  }
  for (int L = Le + 1; L < nlevp; ++L) row[L] = 0u;
}

// =========================================================================================================
// launchers
// =========================================================================================================
static size_t gas_lw_smem(int) {
  return sizeof(Term) * GAS_LC * LW_KTOT + sizeof(double) * (GAS_LC * NB_LW * 2 + (GAS_LC + 1) * NB_LW + NB_LW) +
         sizeof(int) * (GAS_LC * NB_LW * 2 + GAS_LC * NB_LW * 2 + 2 * NG_LW + 2 * NB_LW) + 16;
}
static size_t gas_sw_smem(int nlev) {
  return sizeof(Term) * GAS_LC * SW_KTOT + sizeof(double) * (GAS_LC * NB_SW * 2 + NB_SW * 2 + NG_SW) +
         sizeof(int) * (GAS_LC * NB_SW + GAS_LC * NB_SW * 2 + NB_SW * 2 + NB_SW + 2 * NG_SW + 2 * NB_SW + nlev) + 16;
}
template <class K>
static void allow_smem(K kernel, size_t bytes) {
  // (kernels with the same signature share this template instantiation, so no caching by type here)
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

int launch_gas_prep(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st) {
  const size_t sm = sizeof(double) * GP_T * GP_T * (LWLEV_NF + SWLEV_NF);
  cudaFuncSetAttribute(gas_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  gas_prep_kernel<<<dim3((nc + GP_T - 1) / GP_T, (nlev + GP_T - 1) / GP_T), GP_T * GP_T, sm, st>>>(T, cfg, in, w, nc, nlev);
  return 1;
}
int launch_aerosol(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st) {
  aerosol_optics_kernel<<<(nc * nlev + 127) / 128, 128, 0, st>>>(T, cfg, in, w, nc, nlev);
  return 1;
}
int launch_gas_lw(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st) {
  size_t sm = gas_lw_smem(nlev);
  if (cfg.do_nearest_spectral_lw_emiss) {
    allow_smem(gas_lw_kernel<false>, sm);
    gas_lw_kernel<false><<<nc, GAS_THREADS, sm, st>>>(T, cfg, in, w, nlev);
  } else {
    allow_smem(gas_lw_kernel<true>, sm);
    gas_lw_kernel<true><<<nc, GAS_THREADS, sm, st>>>(T, cfg, in, w, nlev);
  }
  return 1;
}
int launch_gas_sw(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st) {
  size_t sm = gas_sw_smem(nlev);
  allow_smem(gas_sw_kernel, sm);
  gas_sw_kernel<<<nc, GAS_THREADS, sm, st>>>(T, cfg, in, w, nlev);
  return 1;
}
// flux%calc_toa_spectral (radiation_flux.F90:579-660): band sums of the per-g-point top-of-atmosphere fluxes, one thread per
// (column, band), g-points added in ascending order like indexed_sum
__global__ void toa_spectral_kernel(DevTables T, DevOut out, const double* mu0_of, int nc, int ng, int nb, int sw, int do_clear, int with_dn) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * nb) return;
  const int c = t / nb, b = t - c * nb;
  const BandMeta& B = sw ? T.meta->sw[b] : T.meta->lw[b];
  const double* src[3] = {sw ? out.sw_up_toa_g : out.lw_up_toa_g, sw ? out.sw_up_toa_clear_g : out.lw_up_toa_clear_g, sw ? out.sw_dn_toa_g : nullptr};
  double* dst[3] = {sw ? out.sw_up_toa_band : out.lw_up_toa_band, sw ? out.sw_up_toa_clear_band : out.lw_up_toa_clear_band, sw ? out.sw_dn_toa_band : nullptr};
  for (int k = 0; k < 3; ++k) {
    if (!src[k] || !dst[k] || (k == 1 && !do_clear) || (k == 2 && !with_dn)) continue;
    if (k == 2 && mu0_of && mu0_of[c] < 1.0e-10) continue;   // night column: sw_dn_toa_g was not set, leave its band sums alone
    double acc = 0.0;
    const short* rank = sw ? T.meta->rank_sw : T.meta->rank_lw;   // (SPARTACUS on RRTMG stores its g-point arrays reordered)
    for (int g = B.g0; g < B.g0 + B.ng; ++g) acc = acc + src[k][(size_t)c * ng + rank[g]];
    dst[k][(size_t)c * nb + b] = acc;
  }
}
int launch_toa_spectral(const DevTables& T, const DevCfg& cfg, const DevIn& in, const DevOut& out, int nc, bool sw, cudaStream_t st) {
  const int ng = sw ? cfg.ng_sw : cfg.ng_lw, nb = sw ? cfg.nb_sw : cfg.nb_lw;
  // sw_dn_toa_g is only set by the Tripleclouds solver (radiation_tripleclouds_sw.F90:444), so only then is its band sum defined
  toa_spectral_kernel<<<(nc * nb + 127) / 128, 128, 0, st>>>(T, out, sw ? in.cos_sza : nullptr, nc, ng, nb, sw ? 1 : 0, cfg.do_clear, sw && cfg.solver_sw == 4);
  return 1;
}
// ---------------------------------------------------------------------------------------------------------
// save_radiative_properties (radiation_interface.F90:405-425): the optics stages' scratch, whatever its layout, rewritten in the
// reference's element order (spectral index fastest, column slowest), g-points at their position in the solver's sequence
// (T.meta->rank_*: the SPARTACUS reordering on RRTMG-IFS, identity otherwise).  Block = (half-level, column), threads = spectral index.
// ---------------------------------------------------------------------------------------------------------
__global__ void radprops_gather_kernel(DevTables T, DevCfg cfg, DevIn in, Work w, DevProps p, int nc, int nlev) {
  const int l = blockIdx.x, c = blockIdx.y, t = threadIdx.x;
  const int ng_lw = cfg.ng_lw, ng_sw = cfg.ng_sw, nb_lw = cfg.nb_lw, nb_sw = cfg.nb_sw;
  const int ls = w.ls;
  if (cfg.do_lw && t < ng_lw) {
    const int g = t, gd = T.meta->rank_lw[g];
    const bool lb = w.layout_b_lw != 0;
    const size_t hl = lb ? ((size_t)c * ng_lw + g) * ls + l : ((size_t)c * (nlev + 1) + l) * ng_lw + g;   // arrays with nlev + 1 rows
    const size_t fl = lb ? hl : ((size_t)c * nlev + l) * ng_lw + g;                                        // arrays with nlev rows
    if (p.planck_hl) p.planck_hl[((size_t)c * (nlev + 1) + l) * ng_lw + gd] = w.planck[hl];
    if (l < nlev) {
      const size_t o = ((size_t)c * nlev + l) * ng_lw + gd;
      if (p.od_lw) p.od_lw[o] = w.od_lw[fl];
      const bool sc = cfg.do_lw_aerosol_scattering && w.ssa_lw && w.g_lw;
      if (p.ssa_lw) p.ssa_lw[o] = sc ? w.ssa_lw[fl] : 0.0;
      if (p.g_lw) p.g_lw[o] = sc ? w.g_lw[fl] : 0.0;
    }
    if (l == 0) {
      if (p.lw_emission) p.lw_emission[(size_t)c * ng_lw + gd] = w.emission[(size_t)c * ng_lw + g];
      if (p.lw_albedo) p.lw_albedo[(size_t)c * ng_lw + gd] = w.lw_albedo[(size_t)c * ng_lw + g];
    }
  }
  if (cfg.do_sw && t < ng_sw) {
    const int g = t, gd = T.meta->rank_sw[g];
    const bool sunlit = in.cos_sza[c] > 0.0;
    const bool lb = w.layout_b_sw != 0;
    if (l < nlev) {
      const size_t fl = lb ? ((size_t)c * ng_sw + g) * ls + l : ((size_t)c * nlev + l) * ng_sw + g;
      const size_t o = ((size_t)c * nlev + l) * ng_sw + gd;
      double odv = 0.0, ssav = 0.0, gv = 0.0;
      if (sunlit) {
        odv = w.od_sw[fl]; ssav = w.ssa_sw[fl];
        gv = w.g_sw ? w.g_sw[fl] : 0.0;   // (no aerosols: the gases' asymmetry factor is zero)
      } else {
        // Night column with RRTMG-IFS (an ecCKD spectrum arrives here with every column marked sunlit): the gas kernels skip it like
        // srtm_gas_optical_depth does, so the reference's arrays hold max(min_gas_od_sw, 0) and ssa = 0 (radiation_ifs_rrtm.F90:531-594)
        // merged with the aerosols (radiation_aerosol_optics.F90:753-790), which are computed for every column
        odv = dmax(cfg.min_gas_od_sw, 0.0);
        if (cfg.use_aerosols && T.aer && w.aer_sw) {
          const int jb = T.meta->band_of_g_sw[g];
          const double* a = w.aer_sw + ((size_t)c * nlev + l) * 3 * nb_sw;
          const double od_aer = a[jb], scat_aer = a[nb_sw + jb], scat_g_aer = a[2 * nb_sw + jb];
          const double local_od = odv + od_aer;
          if (local_od > 0.0 && od_aer > 0.0) {
            const double local_scat = ssav * odv + scat_aer;
            if (local_scat > 0.0) gv = scat_g_aer / local_scat;
            ssav = local_scat / local_od;
            odv = local_od;
          }
        }
      }
      if (p.od_sw) p.od_sw[o] = odv;
      if (p.ssa_sw) p.ssa_sw[o] = ssav;
      if (p.g_sw) p.g_sw[o] = gv;
    }
    if (l == 0) {
      if (p.incoming_sw) p.incoming_sw[(size_t)c * ng_sw + gd] = sunlit ? w.incoming[(size_t)c * ng_sw + g] : 0.0;
      // get_albedos, radiation_single_level.F90:216-301 (the nearest-interval mapping arrives as 0/1 weights)
      const int jb = T.meta->band_of_g_sw[g];
      double bd = 0.0, bdir = 0.0;
      for (int ja = 0; ja < cfg.n_albedo_sw; ++ja) {
        const double wgt = T.sw_albedo_weights[jb * cfg.n_albedo_sw + ja];
        if (wgt != 0.0) {
          bd = bd + wgt * LD_IN(in.sw_albedo, c, ja);
          if (in.sw_albedo_direct) bdir = bdir + wgt * LD_IN(in.sw_albedo_direct, c, ja);
        }
      }
      if (p.sw_albedo_diffuse) p.sw_albedo_diffuse[(size_t)c * ng_sw + gd] = bd;
      if (p.sw_albedo_direct) p.sw_albedo_direct[(size_t)c * ng_sw + gd] = in.sw_albedo_direct ? bdir : bd;
    }
  }
  if (l < nlev) {
    // cloud optics per band, zero where the (cropped) layer holds no cloud (radiation_cloud_optics.F90:263-270 initialises to zero)
    const bool cloudy = cfg.do_clouds && LD_IN(in.frac, c, l) > 0.0;
    if (cfg.do_lw && t < nb_lw) {
      const double* s = w.cl_lw + ((size_t)c * nlev + l) * 3 * nb_lw;
      const size_t o = ((size_t)c * nlev + l) * nb_lw + t;
      // (the generalised no-scattering form leaves absorption in cropped layers that hold condensate, see general_cloud_optics_kernel)
      const bool od_any = cfg.do_clouds && cfg.use_general_cloud_optics && !cfg.do_lw_cloud_scattering &&
                          (LD_IN(in.q_liq, c, l) > 0.0 || LD_IN(in.q_ice, c, l) > 0.0);
      if (p.od_lw_cloud) p.od_lw_cloud[o] = (cloudy || od_any) ? s[t] : 0.0;
      if (p.ssa_lw_cloud) p.ssa_lw_cloud[o] = cloudy ? s[nb_lw + t] : 0.0;
      if (p.g_lw_cloud) p.g_lw_cloud[o] = cloudy ? s[2 * nb_lw + t] : 0.0;
    }
    if (cfg.do_sw && t < nb_sw) {
      const double* s = w.cl_sw + ((size_t)c * nlev + l) * 3 * nb_sw;
      const size_t o = ((size_t)c * nlev + l) * nb_sw + t;
      if (p.od_sw_cloud) p.od_sw_cloud[o] = cloudy ? s[t] : 0.0;
      if (p.ssa_sw_cloud) p.ssa_sw_cloud[o] = cloudy ? s[nb_sw + t] : 0.0;
      if (p.g_sw_cloud) p.g_sw_cloud[o] = cloudy ? s[2 * nb_sw + t] : 0.0;
    }
  }
}
int launch_radprops_gather(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, const DevProps& p, int nc, int nlev, cudaStream_t st) {
  const int n = cfg.ng_lw > cfg.ng_sw ? cfg.ng_lw : cfg.ng_sw;
  radprops_gather_kernel<<<dim3(nlev + 1, nc), (n + 31) / 32 * 32, 0, st>>>(T, cfg, in, w, p, nc, nlev);
  return 1;
}
int launch_cloud(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st) {
  const int nlevp = (nlev + 3) & ~3;
  int n = 0;
  cloud_prep_kernel<<<(nc + 127) / 128, 128, 0, st>>>(cfg, in, w, nc, nlev); ++n;
  if (cfg.use_general_cloud_optics) n += launch_general_cloud_optics(T, cfg, in, w, nc, nlev, st);
  else { cloud_optics_kernel<<<(nc * nlev + 127) / 128, 128, 0, st>>>(T, cfg, in, w, nc, nlev); ++n; }
  if ((cfg.do_lw && cfg.solver_lw == 2) || (cfg.do_sw && cfg.solver_sw == 2)) {
    if (cfg.use_vectorizable_generator) {
      const int tsw = (cfg.ng_sw + 31) / 32 * 32, tlw = (cfg.ng_lw + 31) / 32 * 32;
      cloud_gen_vec_kernel<<<nc, tsw + tlw, 0, st>>>(cfg, in, w, nc, nlev, nlevp, tsw);
    } else {
      const int nlev8 = (nlev + 7) & ~7;
      cloud_gen_warp_kernel<<<nc, 64, sizeof(double) * 6 * nlev8, st>>>(cfg, in, w, nc, nlev, nlevp, nlev8);
    }
    ++n;
  }
  return n;
}
}  // namespace ecb
