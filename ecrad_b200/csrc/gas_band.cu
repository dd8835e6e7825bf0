// gas_band.cu -- RRTMG gas optics, band-wise, from k-tables staged in shared memory.
//
// Reference: radiation/radiation_ifs_rrtm.F90:323-636 gas_optics (un-reverse, clamps, Planck function, solar scaling) over
// ifsrrtm/rrtm_gas_optical_depth.F90:97-183 + rrtm_taumol1..16.F90 and srtm_gas_optical_depth.F90:96-321 + srtm_taumol16..29.F90.
//
// Mapping.  A CTA owns ONE band, a chunk of GB_LCH consecutive layers and a block of GB_CBLK columns; a thread owns one
// (column, layer) at a time and loops over the column block.  What makes that shape work:
//   * every thread of the CTA runs the same band routine (no divergence over the 16 / 14 band-specific code paths);
//   * a chunk of layers sees only a narrow window of the reference-pressure index jp (the model's layers are much finer than the
//     59 reference pressures and pressure on a model level varies little between columns), and the rows of the big
//     major-species tables ABSA / ABSB are ordered by jp: the CTA stages the rows of [jp_min, jp_max + 1] plus the band's
//     small tables (self / foreign continuum, minor species, Planck fractions) -- tens of KB instead of the band's up to 257 KB
//     -- with TMA bulk copies (cp.async.bulk, one per row so that rows can be padded to a bank-conflict-free stride);
//   * the interpolation stencil of gas_core.h (<= 21 terms per layer and band) is never materialised: the builder emits each
//     {coefficient, row} into a sink that multiply-adds the row straight from shared memory into one register accumulator per
//     g-point of the band.  The number of g-points per band is a compile-time constant (kNgLwBand / kNgSwBand).
// If a chunk's window does not fit the image (coarse vertical grids, wildly different surface pressures), the CTA works
// through it in several jp passes.
// Output layout of the optical properties: Work::layout_b_* (kernels.cuh).
#include "bulk_pipe.cuh"
#include "kernels.cuh"
#include "solver_common.cuh"

namespace ecb {

#ifndef GB_IMG_KB
#define GB_IMG_KB 96
#define GB_MINB 2
#endif
enum { GB_LCH = 16, GB_CC = 16, GB_THREADS = GB_LCH * GB_CC, GB_CBLK = 128, GB_IMG_BYTES = GB_IMG_KB * 1024 };

// row stride (doubles) of the shared-memory image for rows of ng doubles: stride mod 16 in {2, 6, 10, 14}, so that consecutive rows
// start 16 bytes x (odd number) apart in the 128-byte bank line and 8 different rows can be read by a warp without conflict
__host__ __device__ constexpr int gb_row_stride(int ng) { return (ng % 4 == 2) ? ng : ng + 2; }

// ---------------------------------------------------------------------------------------------------------
// per-column bookkeeping: LAYTROP of both spectra, the layer that supplies each SW band's solar source function
// ---------------------------------------------------------------------------------------------------------
__global__ void gas_col_kernel(DevTables T, DevCfg cfg, DevIn in, Work w, int nc, int nlev) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const GasMeta& M = *T.meta;
  const uint8_t* jp = w.gas_jp + (size_t)c * nlev;
  int ntrop_lw = 0, ntrop_sw = 0;
  for (int l = 0; l < nlev; ++l) { ntrop_lw += jp[l] >> 7; ntrop_sw += (jp[l] & 127) < 13; }
  GasCol gc;
  gc.laytrop_lw = ntrop_lw; gc.laytrop_sw = ntrop_sw;
  for (int b = 0; b < NB_SW; ++b) {
    const int il = (cfg.do_sw && in.cos_sza[c] > 0.0) ? sw_solar_layer(M, b, nlev, ntrop_sw, [&](int i) { return (int)(jp[nlev - i] & 127); }) : 0;
    gc.lsol[b] = il > 0 ? nlev - il : -1;
  }
  w.gas_col[c] = gc;
  if (cfg.do_sw) {
    for (int g = 0; g < NG_SW; ++g) w.incoming[(size_t)c * NG_SW + g] = 0.0;
    // work list of the sunlit columns for the shortwave kernel (order is irrelevant: columns are independent)
    if (in.cos_sza[c] > 0.0) w.sunlit[1 + atomicAdd(w.sunlit, 1)] = c;
  }
}

// incoming_sw = ZINCSOL * solar_irradiance / sum(ZINCSOL): radiation_ifs_rrtm.F90:557-605
__global__ void sw_incoming_norm_kernel(DevIn in, Work w, int nc) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc || !(in.cos_sza[c] > 0.0)) return;
  double* inc = w.incoming + (size_t)c * NG_SW;
  double s = 0.0;
  for (int g = 0; g < NG_SW; ++g) s = s + inc[g];
  const double scale = in.solar_irradiance / s;
  for (int g = 0; g < NG_SW; ++g) inc[g] = scale * inc[g];
}

// ---------------------------------------------------------------------------------------------------------
// the evaluating sink: acc[g] += coef * row[g]
// ---------------------------------------------------------------------------------------------------------
template <int NG>
struct EvalSink {
  const double* tab;   // shared-memory image; offsets are elements, always even (16-byte aligned rows)
  double acc[NG];
  int n;               // (interface of ListOut; unused)
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int g = 0; g < NG; ++g) acc[g] = 0.0;
  }
  __device__ __forceinline__ void add(double coef, int off) {
    const double2* r = reinterpret_cast<const double2*>(tab + off);
#pragma unroll
    for (int g = 0; g < NG / 2; ++g) {
      const double2 v = r[g];
      acc[2 * g] = fma(coef, v.x, acc[2 * g]);
      acc[2 * g + 1] = fma(coef, v.y, acc[2 * g + 1]);
    }
  }
  __device__ __forceinline__ void pad4() {}
  __device__ __forceinline__ void scale_all(double f) {
#pragma unroll
    for (int g = 0; g < NG; ++g) acc[g] = f * acc[g];
  }
};

// ---------------------------------------------------------------------------------------------------------
// staging of a band's tables
// ---------------------------------------------------------------------------------------------------------
struct BandStage {
  BandMeta B;          // section offsets remapped to the image, B.ng = row stride
  int ng, rs;          // g-points of the band, row stride of the image
  int off_small, n_small;   // first element (global) / rows of the small sections
  int off_a, rows_a, blk_a; // ABSA: first element, rows, rows per reference pressure
  int off_b, rows_b, blk_b; // ABSB
  int cap_rows;        // rows the image can hold
  // per pass
  int jp_lo, jp_hi;    // reference pressures handled by this pass
  int remaining;       // items whose jp lies above jp_hi
};

// Remap `G` (offsets into the packed global table) to the image: small sections first, then the ABSA window starting at
// reference pressure ja0 and the ABSB window starting at jb0.
__device__ __forceinline__ void stage_remap(BandStage& S, const BandMeta& G, int nsec, int sec_a, int sec_b, int ja0, int jb0, int rows_wa) {
  const int rs = S.rs, ng = S.ng;
  for (int k = 0; k < 16; ++k) S.B.sec[k] = -1;
  for (int k = 0; k < nsec; ++k)
    if (k != sec_a && k != sec_b && G.sec[k] >= 0) S.B.sec[k] = (G.sec[k] - S.off_small) / ng * rs;
  // element offset such that  sec + row * rs  lands on window row (row - first_row)
  if (S.rows_a > 0) S.B.sec[sec_a] = (S.n_small - ja0 * S.blk_a) * rs;
  if (S.rows_b > 0) S.B.sec[sec_b] = (S.n_small + rows_wa - jb0 * S.blk_b) * rs;
  S.B.ng = rs; S.B.g0 = G.g0;
}

// One pass of the pressure-window schedule, run by thread 0: picks [jp_lo, jp_hi] starting at the smallest reference pressure index
// still to do and as wide as the image allows.  lo/hi[2]: smallest / largest jp among the CTA's low (ABSA) and high (ABSB) items.
// Returns the window geometry: first reference-pressure block and rows of each window.
__device__ __forceinline__ void stage_plan(BandStage& S, int start, const int* lo, const int* hi, int& ja0, int& na, int& jb0, int& nb) {
  // low items use ABSA blocks jp-1, jp (0-based: reference pressures jp, jp+1); high items ABSB blocks jp-13, jp-12
  auto need = [&](int j1, int& a0, int& ra, int& b0, int& rb) {
    a0 = 0; ra = 0; b0 = 0; rb = 0;
    if (S.rows_a > 0 && lo[0] <= hi[0]) {
      const int f = imax(start, lo[0]), t = imin(j1, hi[0]);
      if (f <= t) { a0 = f - 1; ra = imin((t - f + 2) * S.blk_a, S.rows_a - a0 * S.blk_a); }
    }
    if (S.rows_b > 0 && lo[1] <= hi[1]) {
      const int f = imax(start, lo[1]), t = imin(j1, hi[1]);
      if (f <= t) { b0 = f - 13; rb = imin((t - f + 2) * S.blk_b, S.rows_b - b0 * S.blk_b); }
    }
    return S.n_small + ra + rb;
  };
  const int last = imax(hi[0], hi[1]);
  int j1 = start;
  while (j1 < last) {
    int a0, ra, b0, rb;
    if (need(j1 + 1, a0, ra, b0, rb) > S.cap_rows) break;
    ++j1;
  }
  need(j1, ja0, na, jb0, nb);
  S.jp_lo = start; S.jp_hi = j1;
  S.remaining = j1 < last;
}

// all threads: issue the bulk copies of `rows` rows of ng doubles from gsrc (packed) to image rows [row0, row0 + rows)
__device__ __forceinline__ void stage_rows(double* img, int rs, int ng, const double* gsrc, int row0, int rows, uint64_t* bar) {
  for (int r = threadIdx.x; r < rows; r += GB_THREADS) bulk_g2s(img + (size_t)(row0 + r) * rs, gsrc + (size_t)r * ng, (uint32_t)(ng * sizeof(double)), bar);
}

// ---------------------------------------------------------------------------------------------------------
// longwave
// ---------------------------------------------------------------------------------------------------------
// V: kernel variant -- 0: od / Planck in [layer][g]; 1: in [g][ls] (scan solvers; longwave aerosol scattering possible);
// 2: [layer][g] with longwave aerosol scattering (SPARTACUS).  The default (0) carries no scattering code.
template <int IB, int V>
__device__ __forceinline__ void lw_item(const GasMeta& M, const BandMeta& Bs, const double* img, const double* tp, double pfac, const DevTables& T,
                                        const DevCfg& cfg, const DevIn& in, const Work& w, int c, int l, int nlev, bool low, const double* __restrict__ Lg, int jw) {
  constexpr int NG = kNgLwBand[IB];
  constexpr bool LAYB = V == 1;
  LwLev L = lwlev_load(Lg, nlev, l);     // (inlined field by field: only the loads of the fields this band reads survive)
  if (IB != 15 || low) L.jp = jw;
  EvalSink<NG> sink;
  sink.tab = img; sink.clear();
  int post;
  const PlanckFrac pf = lw_build_list(M, Bs, L, IB, low, sink, &post);
  const int g0 = Bs.g0;
  // Planck function of the band at the half-levels: radiation_ifs_rrtm.F90:676-699
  auto planck = [&](double temperature) {
    int ind; double frac;
    if (temperature < 339.0 && temperature >= 160.0) { ind = (int)(temperature - 159.0); frac = temperature - (int)temperature; }
    else if (temperature >= 339.0) { ind = 180; frac = temperature - 339.0; }
    else { ind = 1; frac = 0.0; }
    return pfac * (tp[ind - 1] + frac * (tp[ind] - tp[ind - 1]));
  };
  const double plk_bot = planck(L.t_bot);
  const bool lwscat = V != 0 && cfg.do_lw_aerosol_scattering != 0;
  double aer = 0.0, aer_sc = 0.0, aer_sg = 0.0;
  if (cfg.use_aerosols) {
    if (lwscat) { const double* a = w.aer_lw + ((size_t)c * nlev + l) * 3 * NB_LW; aer = a[IB]; aer_sc = a[NB_LW + IB]; aer_sg = a[2 * NB_LW + IB]; }
    else aer = w.aer_lw[((size_t)c * nlev + l) * NB_LW + IB];   // radiation_aerosol_optics.F90:806-812
  }
  const size_t sg = LAYB ? (size_t)w.ls : 1, sl = LAYB ? 1 : (size_t)NG_LW;
  double* od_out = w.od_lw + (size_t)c * (LAYB ? (size_t)NG_LW * w.ls : (size_t)nlev * NG_LW) + (size_t)g0 * sg;
  double* pl_out = w.planck + (size_t)c * (LAYB ? (size_t)NG_LW * w.ls : (size_t)(nlev + 1) * NG_LW) + (size_t)g0 * sg;
  auto pfrac = [&](int g) { return pf.c0 * img[pf.o0 + g] + pf.c1 * img[pf.o1 + g]; };
  auto odval = [&](int g) {
    double tau = sink.acc[g];
    if (post >= 0) tau *= img[post + g];
    double o = dmax(tau, cfg.min_gas_od_lw);                             // radiation_ifs_rrtm.F90:506-511
    if (cfg.use_aerosols && !lwscat) o = o + aer;
    return o;
  };
  if (lwscat) {
    // gas + aerosol with scattering (radiation_aerosol_optics.F90:781-796): gases do not scatter, so g is the aerosol's.
    // (ssa_lw / g_lw have the layout of od_lw: [g][ls] for the scan solvers, [layer][g] for SPARTACUS.)
    double* ss_out = w.ssa_lw + (size_t)c * (LAYB ? (size_t)NG_LW * w.ls : (size_t)nlev * NG_LW) + (size_t)g0 * sg;
    double* gg_out = w.g_lw + (size_t)c * (LAYB ? (size_t)NG_LW * w.ls : (size_t)nlev * NG_LW) + (size_t)g0 * sg;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      double o = odval(g), ssa = 0.0, gg = 0.0;
      const double local_od = o + aer;
      if (local_od > 0.0 && aer > 0.0) {
        if (aer_sc > 0.0) gg = aer_sg / aer_sc;
        ssa = aer_sc / local_od;
        o = local_od;
      }
      od_out[l * sl + g * sg] = o; ss_out[l * sl + g * sg] = ssa; gg_out[l * sl + g * sg] = gg;
      pl_out[(l + 1) * sl + g * sg] = plk_bot * pfrac(g);
    }
  } else
  if (LAYB) {
#pragma unroll
    for (int g = 0; g < NG; ++g) { od_out[l * sl + g * sg] = odval(g); pl_out[(l + 1) * sl + g * sg] = plk_bot * pfrac(g); }   // the half-level below uses this layer's PFRAC
  } else {   // rows of the band's g-points: 16-byte stores (g0 is even, rows are 16-byte aligned)
#pragma unroll
    for (int g = 0; g < NG; g += 2) {
      *reinterpret_cast<double2*>(od_out + l * sl + g) = make_double2(odval(g), odval(g + 1));
      *reinterpret_cast<double2*>(pl_out + (l + 1) * sl + g) = make_double2(plk_bot * pfrac(g), plk_bot * pfrac(g + 1));
    }
  }
  if (l == 0) {
    const double plk_top = planck(L.t_top);                              // top-of-atmosphere half-level: PFRAC of the top layer
#pragma unroll
    for (int g = 0; g < NG; ++g) pl_out[g * sg] = plk_top * pfrac(g);
  }
  if (l == nlev - 1) {
    // surface: planck_function_surf :757-852, lw_emission = planck_surf * (1 - lw_albedo) :466
    double alb;
    if (cfg.do_nearest_spectral_lw_emiss) {
      alb = 1.0 - LD_IN(in.lw_emissivity, c, T.i_emiss_from_band_lw[IB] - 1);
    } else {   // weighted emissivity intervals (get_albedos, radiation_single_level.F90:330-352)
      alb = 0.0;
      for (int ja = 0; ja < cfg.n_emiss_lw; ++ja) {
        const double wgt = T.lw_emiss_weights[IB * cfg.n_emiss_lw + ja];
        if (wgt != 0.0) alb = alb + wgt * (1.0 - LD_IN(in.lw_emissivity, c, ja));
      }
    }
    const double plk_surf = planck(in.skin_t[c]);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      w.lw_albedo[(size_t)c * NG_LW + g0 + g] = alb;
      w.emission[(size_t)c * NG_LW + g0 + g] = (plk_surf * pfrac(g)) * (1.0 - alb);
    }
  }
}

template <int V>
__global__ void __launch_bounds__(GB_THREADS, GB_MINB)
gas_lw_band_kernel(DevTables T, DevCfg cfg, DevIn in, Work w, int nc, int nlev) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* img = reinterpret_cast<double*>(smem_raw);   // [GB_IMG_BYTES / 8]
  double* tp = img + GB_IMG_BYTES / 8;                 // [181] TOTPLNK of this band
  __shared__ BandStage S;
  __shared__ int s_lo[2], s_hi[2];
  __shared__ __align__(8) uint64_t s_bar;
  const GasMeta& M = *T.meta;
  // band is the fastest grid dimension: the CTAs that work on the same (columns, layers) -- and read the same per-layer state
  // and write neighbouring g-points of the same rows -- run next to each other in time, so both go through the L2
  const int band = blockIdx.x, tid = threadIdx.x;
  const int l = blockIdx.z * GB_LCH + (tid % GB_LCH);
  const int c_first = blockIdx.y * GB_CBLK, cc = tid / GB_LCH;
  const bool lvalid = l < nlev;
  const BandMeta& G = M.lw[band];
  if (tid == 0) {
    S.ng = G.ng; S.rs = gb_row_stride(G.ng);
    S.off_small = G.sec[SEC_SMALL]; S.n_small = (G.sec[SEC_END] - G.sec[SEC_SMALL]) / G.ng;
    S.off_a = G.sec[L_ABSA]; S.rows_a = M.lw_rows[band][0]; S.blk_a = S.rows_a / 13;
    S.off_b = G.sec[L_ABSB]; S.rows_b = M.lw_rows[band][1]; S.blk_b = S.rows_b / 47;
    S.cap_rows = (GB_IMG_BYTES / 8) / S.rs;
    s_lo[0] = s_lo[1] = 1 << 20; s_hi[0] = s_hi[1] = -1;
    mbar_init(&s_bar, 1);
    mbar_fence_init();
  }
  for (int i = tid; i < 181; i += GB_THREADS) tp[i] = M.totplnk[band * 181 + i];
  __syncthreads();
  // ---- which reference pressures do the CTA's items touch? ----
  int jpv[GB_CBLK / GB_CC];       // per item: jp (window index), bit 8 = low, bit 9 = to do
#pragma unroll
  for (int it = 0; it < GB_CBLK / GB_CC; ++it) {
    const int c = c_first + it * GB_CC + cc;
    jpv[it] = 0;
    if (lvalid && c < nc) {
      const int jp = w.gas_jp[(size_t)c * nlev + l] & 127;
      const bool low = (nlev - l) <= w.gas_col[c].laytrop_lw;
      // rows exist for jp <= 12 (ABSA) / jp >= 13 (ABSB): a layer on the other side of its own LAYTROP (non-monotonic pressure) is
      // clamped into the table instead of reading outside it
      int jw = low ? imin(jp, 12) : imax(jp, 13);
      if (band == 15 && !low) jw = 13;   // NSPB(16) = 0: rrtm_taumol16 always reads rows 1-2 of ABSB
      jpv[it] = jw | (low ? 256 : 0) | 512;
      atomicMin(&s_lo[low ? 0 : 1], jw);
      atomicMax(&s_hi[low ? 0 : 1], jw);
    }
  }
  __syncthreads();
  const double pfac = (2.0 * asin(1.0) * 1.0e4) * M.delwave[band];
  uint32_t phase = 0;
  int start = imin(s_lo[0], s_lo[1]);
  bool first = true;
  while (start < (1 << 20)) {
    int ja0 = 0, na = 0, jb0 = 0, nb = 0;
    if (tid == 0) {
      stage_plan(S, start, s_lo, s_hi, ja0, na, jb0, nb);
      stage_remap(S, G, L_NSEC, L_ABSA, L_ABSB, ja0, jb0, na);
      S.off_a = G.sec[L_ABSA] + ja0 * S.blk_a * S.ng;    // (reused as: first element of the window)
      S.off_b = G.sec[L_ABSB] + jb0 * S.blk_b * S.ng;
      S.B.sec[SEC_SMALL] = na; S.B.sec[SEC_END] = nb;    // window rows of this pass
      const uint32_t bytes = (uint32_t)(((first ? S.n_small : 0) + na + nb) * S.ng * sizeof(double));
      mbar_arrive_expect_tx(&s_bar, bytes);
    }
    __syncthreads();
    {
      const int wa = S.B.sec[SEC_SMALL], wb = S.B.sec[SEC_END];
      if (first) stage_rows(img, S.rs, S.ng, T.lwtab + S.off_small, 0, S.n_small, &s_bar);
      if (wa > 0) stage_rows(img, S.rs, S.ng, T.lwtab + S.off_a, S.n_small, wa, &s_bar);
      if (wb > 0) stage_rows(img, S.rs, S.ng, T.lwtab + S.off_b, S.n_small + wa, wb, &s_bar);
    }
    mbar_wait(&s_bar, phase & 1);
    ++phase; first = false;
    const int jlo = S.jp_lo, jhi = S.jp_hi;
#pragma unroll 1
    for (int it = 0; it < GB_CBLK / GB_CC; ++it) {
      const int v = jpv[it], jw = v & 127;
      if (!(v & 512) || jw < jlo || jw > jhi) continue;
      jpv[it] = v & ~512;
      const int c = c_first + it * GB_CC + cc;
      const double* Lg = w.lev_lw + (size_t)c * LWLEV_NF * nlev;
      const bool low = (v & 256) != 0;
#define LWB(I) case I: lw_item<I, V>(M, S.B, img, tp, pfac, T, cfg, in, w, c, l, nlev, low, Lg, jw); break;
      switch (band) { LWB(0) LWB(1) LWB(2) LWB(3) LWB(4) LWB(5) LWB(6) LWB(7) LWB(8) LWB(9) LWB(10) LWB(11) LWB(12) LWB(13) LWB(14) LWB(15) }
#undef LWB
    }
    const int more = S.remaining, nxt = S.jp_hi + 1;
    __syncthreads();   // everyone is done with the image (and with S) before the next pass overwrites them
    start = more ? nxt : (1 << 20);
  }
}

// ---------------------------------------------------------------------------------------------------------
// shortwave (sunlit columns)
// ---------------------------------------------------------------------------------------------------------
template <int IB, bool LAYB>
__device__ __forceinline__ void sw_item(const GasMeta& M, const BandMeta& Bs, const double* img, const DevCfg& cfg, const Work& w, int c, int l,
                                        int nlev, bool low, bool solar_layer, const double* __restrict__ Lg, int jw) {
  constexpr int NG = kNgSwBand[IB];
  SwLev L = swlev_load(Lg, nlev, l);
  L.jp = jw;
  EvalSink<NG> sink;
  sink.tab = img; sink.clear();
  SwAux aux;
  sw_build_list(M, Bs, L, IB, low, sink, aux);
  const int g0 = Bs.g0;
  const size_t sg = LAYB ? (size_t)w.ls : 1, sl = LAYB ? 1 : (size_t)NG_SW;
  const size_t base = (size_t)c * (LAYB ? (size_t)NG_SW * w.ls : (size_t)nlev * NG_SW) + (size_t)g0 * sg;
  double od_a = 0.0, sc_a = 0.0, sg_a = 0.0;
  if (cfg.use_aerosols) {
    const double* a = w.aer_sw + ((size_t)c * nlev + l) * 3 * NB_SW;
    od_a = a[IB]; sc_a = a[NB_SW + IB]; sg_a = a[2 * NB_SW + IB];
  }
  auto eval = [&](int g, double& odv, double& ssav, double& gv) {
    const double taug = sink.acc[g];
    const double taur = aux.rc0 * img[aux.ro0 + g] + aux.rc1 * img[aux.ro1 + g];
    const double od = taur + taug;                                        // srtm_gas_optical_depth.F90:314-320
    odv = dmax(od, cfg.min_gas_od_sw); ssav = taur / od;                 // radiation_ifs_rrtm.F90:593
    gv = 0.0;
    if (cfg.use_aerosols) {
      // merge aerosol and gas per g-point: radiation_aerosol_optics.F90:765-781
      const double local_od = odv + od_a;
      if (local_od > 0.0 && od_a > 0.0) {
        const double local_scat = ssav * odv + sc_a;
        if (local_scat > 0.0) gv = sg_a / local_scat;
        ssav = local_scat / local_od;
        odv = local_od;
      }
    }
    if (solar_layer) w.incoming[(size_t)c * NG_SW + g0 + g] = aux.sc0 * img[aux.so0 + g] + aux.sc1 * img[aux.so1 + g];
  };
  if (LAYB) {
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      double o, ss, gv; eval(g, o, ss, gv);
      w.od_sw[base + l * sl + g * sg] = o; w.ssa_sw[base + l * sl + g * sg] = ss;
      if (cfg.use_aerosols) w.g_sw[base + l * sl + g * sg] = gv;
    }
  } else {
#pragma unroll
    for (int g = 0; g < NG; g += 2) {
      double o0, s0, v0, o1, s1, v1; eval(g, o0, s0, v0); eval(g + 1, o1, s1, v1);
      *reinterpret_cast<double2*>(w.od_sw + base + l * sl + g) = make_double2(o0, o1);
      *reinterpret_cast<double2*>(w.ssa_sw + base + l * sl + g) = make_double2(s0, s1);
      if (cfg.use_aerosols) *reinterpret_cast<double2*>(w.g_sw + base + l * sl + g) = make_double2(v0, v1);
    }
  }
}

template <bool LAYB>
__global__ void __launch_bounds__(GB_THREADS, GB_MINB)
gas_sw_band_kernel(DevTables T, DevCfg cfg, DevIn in, Work w, int nc, int nlev) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* img = reinterpret_cast<double*>(smem_raw);
  __shared__ BandStage S;
  __shared__ int s_lo[2], s_hi[2];
  __shared__ __align__(8) uint64_t s_bar;
  const GasMeta& M = *T.meta;
  // band is the fastest grid dimension: the CTAs that work on the same (columns, layers) -- and read the same per-layer state
  // and write neighbouring g-points of the same rows -- run next to each other in time, so both go through the L2
  const int band = blockIdx.x, tid = threadIdx.x;
  const int l = blockIdx.z * GB_LCH + (tid % GB_LCH);
  const int c_first = blockIdx.y * GB_CBLK, cc = tid / GB_LCH;
  const bool lvalid = l < nlev;
  const int nsun = w.sunlit[0];
  if (c_first >= nsun) return;   // the grid is sized for "every column sunlit"
  const int* sun = w.sunlit + 1;
  const BandMeta& G = M.sw[band];
  if (tid == 0) {
    S.ng = G.ng; S.rs = gb_row_stride(G.ng);
    S.off_small = G.sec[SEC_SMALL]; S.n_small = (G.sec[SEC_END] - G.sec[SEC_SMALL]) / G.ng;
    S.off_a = G.sec[S_ABSA]; S.rows_a = M.sw_rows[band][0]; S.blk_a = S.rows_a / 13;
    S.off_b = G.sec[S_ABSB]; S.rows_b = M.sw_rows[band][1]; S.blk_b = S.rows_b / 47;
    S.cap_rows = (GB_IMG_BYTES / 8) / S.rs;
    s_lo[0] = s_lo[1] = 1 << 20; s_hi[0] = s_hi[1] = -1;
    mbar_init(&s_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  int jpv[GB_CBLK / GB_CC];
  bool any = false;
#pragma unroll
  for (int it = 0; it < GB_CBLK / GB_CC; ++it) {
    const int ci = c_first + it * GB_CC + cc;
    jpv[it] = 0;
    if (lvalid && ci < nsun) {   // srtm only runs for sunlit columns (radiation_ifs_rrtm.F90:518-542)
      const int c = sun[ci];
      const int jp = w.gas_jp[(size_t)c * nlev + l] & 127;
      const bool low = (nlev - l) <= w.gas_col[c].laytrop_sw;
      const int jw = low ? imin(jp, 12) : imax(jp, 13);
      jpv[it] = jw | (low ? 256 : 0) | 512;
      atomicMin(&s_lo[low ? 0 : 1], jw);
      atomicMax(&s_hi[low ? 0 : 1], jw);
      any = true;
    }
  }
  if (!__syncthreads_or(any)) return;   // no sunlit column in this block
  uint32_t phase = 0;
  int start = imin(s_lo[0], s_lo[1]);
  bool first = true;
  while (start < (1 << 20)) {
    int ja0 = 0, na = 0, jb0 = 0, nb = 0;
    if (tid == 0) {
      stage_plan(S, start, s_lo, s_hi, ja0, na, jb0, nb);
      stage_remap(S, G, S_NSEC, S_ABSA, S_ABSB, ja0, jb0, na);
      S.off_a = G.sec[S_ABSA] + ja0 * S.blk_a * S.ng;
      S.off_b = G.sec[S_ABSB] + jb0 * S.blk_b * S.ng;
      S.B.sec[SEC_SMALL] = na; S.B.sec[SEC_END] = nb;
      const uint32_t bytes = (uint32_t)(((first ? S.n_small : 0) + na + nb) * S.ng * sizeof(double));
      mbar_arrive_expect_tx(&s_bar, bytes);
    }
    __syncthreads();
    {
      const int wa = S.B.sec[SEC_SMALL], wb = S.B.sec[SEC_END];
      if (first) stage_rows(img, S.rs, S.ng, T.swtab + S.off_small, 0, S.n_small, &s_bar);
      if (wa > 0) stage_rows(img, S.rs, S.ng, T.swtab + S.off_a, S.n_small, wa, &s_bar);
      if (wb > 0) stage_rows(img, S.rs, S.ng, T.swtab + S.off_b, S.n_small + wa, wb, &s_bar);
    }
    mbar_wait(&s_bar, phase & 1);
    ++phase; first = false;
    const int jlo = S.jp_lo, jhi = S.jp_hi;
#pragma unroll 1
    for (int it = 0; it < GB_CBLK / GB_CC; ++it) {
      const int v = jpv[it], jw = v & 127;
      if (!(v & 512) || jw < jlo || jw > jhi) continue;
      jpv[it] = v & ~512;
      const int c = sun[c_first + it * GB_CC + cc];
      const double* Lg = w.lev_sw + (size_t)c * SWLEV_NF * nlev;
      const bool low = (v & 256) != 0;
      const bool solar_layer = l == w.gas_col[c].lsol[band];
#define SWB(I) case I: sw_item<I, LAYB>(M, S.B, img, cfg, w, c, l, nlev, low, solar_layer, Lg, jw); break;
      switch (band) { SWB(0) SWB(1) SWB(2) SWB(3) SWB(4) SWB(5) SWB(6) SWB(7) SWB(8) SWB(9) SWB(10) SWB(11) SWB(12) SWB(13) }
#undef SWB
    }
    const int more = S.remaining, nxt = S.jp_hi + 1;
    __syncthreads();
    start = more ? nxt : (1 << 20);
  }
}

// ---------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------
int launch_gas_col(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st) {
  cudaMemsetAsync(w.sunlit, 0, sizeof(int), st);
  gas_col_kernel<<<(nc + 127) / 128, 128, 0, st>>>(T, cfg, in, w, nc, nlev);
  return 1;
}
int launch_gas_lw_band(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st) {
  const size_t sm = GB_IMG_BYTES + 181 * sizeof(double) + 16;
  const dim3 grid(NB_LW, (nc + GB_CBLK - 1) / GB_CBLK, (nlev + GB_LCH - 1) / GB_LCH);
  const int v = w.layout_b_lw ? 1 : cfg.do_lw_aerosol_scattering ? 2 : 0;
  switch (v) {
    case 1:
      cudaFuncSetAttribute(gas_lw_band_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      gas_lw_band_kernel<1><<<grid, GB_THREADS, sm, st>>>(T, cfg, in, w, nc, nlev);
      break;
    case 2:
      cudaFuncSetAttribute(gas_lw_band_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      gas_lw_band_kernel<2><<<grid, GB_THREADS, sm, st>>>(T, cfg, in, w, nc, nlev);
      break;
    default:
      cudaFuncSetAttribute(gas_lw_band_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      gas_lw_band_kernel<0><<<grid, GB_THREADS, sm, st>>>(T, cfg, in, w, nc, nlev);
  }
  return 1;
}
int launch_gas_sw_band(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st) {
  const size_t sm = GB_IMG_BYTES + 16;
  const dim3 grid(NB_SW, (nc + GB_CBLK - 1) / GB_CBLK, (nlev + GB_LCH - 1) / GB_LCH);
  if (w.layout_b_sw) {
    cudaFuncSetAttribute(gas_sw_band_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    gas_sw_band_kernel<true><<<grid, GB_THREADS, sm, st>>>(T, cfg, in, w, nc, nlev);
  } else {
    cudaFuncSetAttribute(gas_sw_band_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    gas_sw_band_kernel<false><<<grid, GB_THREADS, sm, st>>>(T, cfg, in, w, nc, nlev);
  }
  sw_incoming_norm_kernel<<<(nc + 127) / 128, 128, 0, st>>>(in, w, nc);
  return 2;
}

}  // namespace ecb
