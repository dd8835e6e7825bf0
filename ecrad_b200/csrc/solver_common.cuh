// solver_common.cuh -- device helpers shared by the LW and SW solver kernels.
#pragma once
#include "cloud_core.h"
#include "kernels.cuh"
#include "solver_core.h"

namespace ecb {

#define LD_IN(p, c, j) ((p)[(size_t)(j) * in.ld + (c)])

enum { LCH = 16 };   // layers between two g-point reductions

// Sum the rows of the shared-memory tile over g (4 threads per row) into sums[f][level].
// Row (f, s), s < ns, holds level lfirst + dir*s of flux f.  Ends with a barrier so the tile can be refilled.
__device__ __forceinline__ void flush_tile(const double* tile, int rs, int ng, int nf, int ns, double* const* dst, int lfirst, int dir, int lch = LCH) {
  __syncthreads();
  const int row = threadIdx.x >> 2, sub = threadIdx.x & 3;
  const int nrows = nf * ns, rows_per_round = blockDim.x >> 2;
  for (int r0 = 0; r0 < nrows; r0 += rows_per_round) {
    const int r = r0 + row;
    const bool valid = r < nrows;
    int f = 0, s = 0;
    double p = 0.0;
    if (valid) {
      f = r / ns; s = r - f * ns;
      const double* t = tile + (size_t)(f * lch + s) * rs;
      for (int g = sub; g < ng; g += 4) p += t[g];
    }
    p += __shfl_xor_sync(0xffffffffu, p, 1);
    p += __shfl_xor_sync(0xffffffffu, p, 2);
    if (valid && sub == 0) dst[f][lfirst + dir * s] = p;
  }
  __syncthreads();
}

// Per-band sums of tile rows (do_save_spectral_flux: flux%*_band(nband, ncol, nlev+1), radiation_flux.F90 indexed_sum_profile).
// Output k (k < nout) = sa[k] * sum_{g in band} row(fa[k]) [+ sb[k] * sum_{g in band} row(fb[k]) if fb[k] >= 0] [+ add[k][...]],
// written to dst[k][((level * ld) + c) * nb + band]; g-points are summed in ascending order like the reference.
// Call BEFORE flush_tile (it starts with a barrier; the tile is only read here).
struct BandOut { double* dst; int ld; int fa, fb; double sa, sb; const double* add; int add_ld; };   // add: optional addend, same layout with its own ld
__device__ __forceinline__ void flush_bands(const double* tile, int rs, int lch, int ns, const BandOut* bo, int nout, int lfirst, int dir,
                                            int c, int nb, const BandMeta* bands) {
  __syncthreads();
  const int ntask = nout * ns * nb;
  for (int t = threadIdx.x; t < ntask; t += blockDim.x) {
    const int k = t / (ns * nb), r = t - k * ns * nb, s = r / nb, b = r - s * nb;
    const BandOut& o = bo[k];
    if (!o.dst) continue;
    const int g0 = bands[b].g0, ng = bands[b].ng, level = lfirst + dir * s;
    const double* ra = tile + (size_t)(o.fa * lch + s) * rs + g0;
    double acc = 0.0;
    for (int g = 0; g < ng; ++g) acc = acc + ra[g];
    double v = o.sa * acc;
    if (o.fb >= 0) {
      const double* rb = tile + (size_t)(o.fb * lch + s) * rs + g0;
      double accb = 0.0;
      for (int g = 0; g < ng; ++g) accb = accb + rb[g];
      v = v + o.sb * accb;
    }
    if (o.add) v = v + o.add[((size_t)level * o.add_ld + c) * nb + b];
    o.dst[((size_t)level * o.ld + c) * nb + b] = v;
  }
}

__device__ __forceinline__ uint32_t pick4(const uint4& q, int k) { return k == 0 ? q.x : k == 1 ? q.y : k == 2 ? q.z : q.w; }

// optical-depth scaling of this (g, layer) from the generator's code word
__device__ __forceinline__ double od_scaling_from_code(const CloudMeta& C, const double* pdf_val, uint32_t code, double fsd) {
  if (!code) return 0.0;
  return pdf_sample(C, pdf_val, fsd, (double)(code & C.gen_mask) * C.gen_scale);
}


// total (gas + scaled cloud) optical properties of a cloudy layer for this g-point: radiation_mcica_sw.F90:249-272
template <class SD>
__device__ __forceinline__ void sw_cloudy_props(const CloudMeta& C, const double* pdf_val, uint32_t code, double fsd, const double* clb,
                                                int b, double od_gas, double ssa_gas, double g_gas, double& odt, double& ssat, double& gt) {
  const double scal = od_scaling_from_code(C, pdf_val, code, fsd);
  const double od_cloud_new = scal * clb[b];
  odt = od_gas + od_cloud_new;
  ssat = 0.0; gt = 0.0;
  if (odt > 0.0) {
    const double ssac = clb[SD::NB + b];
    const double scat_od = ssa_gas * od_gas + ssac * od_cloud_new;
    ssat = scat_od / odt;
    if (scat_od > 0.0) gt = (g_gas * ssa_gas * od_gas + clb[2 * SD::NB + b] * ssac * od_cloud_new) / scat_od;
  }
}

// surface spectral and canopy fluxes: radiation_flux.F90:397-577 calc_surface_spectral.  Block-wide (contains barriers);
// tile: >= 4*rs + 28 doubles of shared memory; dir/dif: per-g direct and diffuse surface fluxes (all-sky, clear-sky).
template <class SD>
__device__ __forceinline__ void sw_surface_spectral(const DevTables& T, const DevCfg& cfg, const DevOut& out, int c, int g, bool act,
                                                    double* tile, int rs, double dir_a, double dif_a, double dir_c, double dif_c) {
  if (!(cfg.do_surface_sw_spectral_flux || cfg.do_canopy_fluxes_sw)) return;
  __syncthreads();
  if (act) { tile[g] = dir_a; tile[rs + g] = dif_a; tile[2 * rs + g] = dir_c; tile[3 * rs + g] = dif_c; }
  __syncthreads();
  double* bdir = tile + 4 * rs;  // [14] all-sky direct band, [14] all-sky total band
  if (g < SD::NB) {
    const int g0 = T.meta->sw[g].g0, ngb = T.meta->sw[g].ng;
    double d = 0.0, t = 0.0, dcl = 0.0, tcl = 0.0;
    for (int k = g0; k < g0 + ngb; ++k) { d = d + tile[k]; t = t + tile[rs + k]; dcl = dcl + tile[2 * rs + k]; tcl = tcl + tile[3 * rs + k]; }
    t = t + d; tcl = tcl + dcl;
    bdir[g] = d; bdir[SD::NB + g] = t;
    if (cfg.do_surface_sw_spectral_flux) {
      if (out.sw_dn_direct_surf_band) out.sw_dn_direct_surf_band[(size_t)c * SD::NB + g] = d;
      if (out.sw_dn_surf_band) out.sw_dn_surf_band[(size_t)c * SD::NB + g] = t;
      if (cfg.do_clear && out.sw_dn_direct_surf_clear_band) out.sw_dn_direct_surf_clear_band[(size_t)c * SD::NB + g] = dcl;
      if (cfg.do_clear && out.sw_dn_surf_clear_band) out.sw_dn_surf_clear_band[(size_t)c * SD::NB + g] = tcl;
    }
  }
  __syncthreads();
  if (cfg.do_canopy_fluxes_sw && out.sw_dn_diffuse_surf_canopy && out.sw_dn_direct_surf_canopy && g < cfg.n_albedo_sw) {
    double dif = 0.0, dir = 0.0;
    for (int jb = 0; jb < SD::NB; ++jb) {
      const double wgt = T.sw_albedo_weights[jb * cfg.n_albedo_sw + g];
      if (wgt != 0.0) { dif = dif + wgt * bdir[SD::NB + jb]; dir = dir + wgt * bdir[jb]; }
    }
    out.sw_dn_diffuse_surf_canopy[(size_t)c * cfg.n_albedo_sw + g] = dif - dir;
    out.sw_dn_direct_surf_canopy[(size_t)c * cfg.n_albedo_sw + g] = dir;
  }
}

// LW canopy fluxes, radiation_flux.F90:529-572 calc_surface_spectral: nearest-interval mapping (i_emiss_from_band_lw) or the
// weighted emissivity intervals (lw_emiss_weights applied to the band sums of lw_dn_surf_g).  Block-wide.
template <class SD>
__device__ __forceinline__ void lw_surface_canopy(const DevTables& T, const DevCfg& cfg, const DevOut& out, int c, int g, bool act,
                                                  double* tile, double dn_surf_g) {
  if (!(cfg.do_canopy_fluxes_lw && out.lw_dn_surf_canopy)) return;
  __syncthreads();
  if (act) tile[g] = dn_surf_g;
  __syncthreads();
  if (g < cfg.n_canopy_bands_lw) {
    double sum = 0.0;
    if (cfg.do_nearest_spectral_lw_emiss) {
      for (int k = 0; k < SD::NG; ++k)
        if (T.i_emiss_from_band_lw[T.meta->band_of_g_lw[k]] - 1 == g) sum = sum + tile[k];
    } else {
      for (int jb = 0; jb < SD::NB; ++jb) {
        const double wgt = T.lw_emiss_weights[jb * cfg.n_emiss_lw + g];
        if (wgt != 0.0) {
          double band = 0.0;
          const int g0 = T.meta->lw[jb].g0, ngb = T.meta->lw[jb].ng;
          for (int k = g0; k < g0 + ngb; ++k) band = band + tile[k];
          sum = sum + wgt * band;
        }
      }
    }
    out.lw_dn_surf_canopy[(size_t)c * cfg.n_canopy_bands_lw + g] = sum;
  }
}

// night column of a SW solver: radiation_mcica_sw.F90:380-401 / radiation_tripleclouds_sw.F90:236-280
template <class SD>
__device__ __forceinline__ void sw_night_column(const DevCfg& cfg, const DevOut& out, int c, int g, bool act, int nl1, int nthreads) {
  for (int l = g; l < nl1; l += nthreads) {
    double* p[6] = {out.sw_up, out.sw_dn, out.sw_dn_direct, out.sw_up_clear, out.sw_dn_clear, out.sw_dn_direct_clear};
    for (int k = 0; k < 6; ++k) if (p[k]) p[k][(size_t)l * out.ld + c] = 0.0;
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
  if (cfg.do_save_spectral_flux && cfg.solver_sw != 2) {
    double* pb[3] = {out.sw_up_band, out.sw_dn_band, out.sw_dn_direct_band};
    for (int k = 0; k < 3; ++k)
      if (pb[k]) for (int i = g; i < nl1 * SD::NB; i += nthreads) pb[k][((size_t)(i / SD::NB) * out.ld + c) * SD::NB + (i % SD::NB)] = 0.0;
  }
  if (g < cfg.n_canopy_bands_sw && cfg.do_canopy_fluxes_sw) {
    if (out.sw_dn_diffuse_surf_canopy) out.sw_dn_diffuse_surf_canopy[(size_t)c * cfg.n_canopy_bands_sw + g] = 0.0;
    if (out.sw_dn_direct_surf_canopy) out.sw_dn_direct_surf_canopy[(size_t)c * cfg.n_canopy_bands_sw + g] = 0.0;
  }
}

}  // namespace ecb
