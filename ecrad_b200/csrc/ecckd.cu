// ecckd.cu -- ecCKD gas optics and generalised (look-up-table) cloud optics, per g-point.
//
// Reference: radiation/radiation_ecckd_interface.F90:174-324 (gas_optics), radiation_ecckd.F90:457-654
// (calc_optical_depth_ckd_model), :900-928 (calc_planck_function), :935-964 (calc_incoming_sw);
// radiation_general_cloud_optics.F90:134-287 + radiation_general_cloud_optics_data.F90:249-330 (add_optical_properties);
// radiation_aerosol_optics.F90:487-826 (add_aerosol_optics, one "band" per g-point); radiation_single_level.F90:216-365.
// Same operations in the same order as the reference (the library is built without FMA contraction).
//
// Mapping: one CTA of 4 warps per column.  Prologue: one thread per layer reduces (p, T, mole fractions) to the
// interpolation state of that layer -- table corner, weights, one multiplier per gas -- in shared memory.  Main loop:
// warp = layer, lane = g-point, so every look-up-table access is a contiguous row of ng doubles (the tables are
// g-point fastest exactly as the reference stores them, 0.9-4.8 MB per model: L2 resident) and every store is coalesced.
#include "kernels.cuh"
#include "solver_common.cuh"

namespace ecb {

enum { CKD_THREADS = 128, CKD_WARPS = 4 };

// Interpolation state of one layer (calc_optical_depth_ckd_model :531-556), carved from shared memory as structure of arrays
// sized by the model (ngas multipliers per layer, one concentration corner per look-up-table gas):
struct CkdLayers {
  int2* corner;        // [nlev] (ip1, it1): 1-based lower corner in pressure / temperature
  double* pw2;         // [nlev] upper weights (the lower ones are 1 - w, as in the reference)
  double* tw2;         // [nlev]
  double* smult;       // [nlev] simple_multiplier: mol of dry air per m2 in the layer
  double* mult;        // [nlev][ngas] multiplier of the interpolated molar absorption of each gas
  int* ic1;            // [nlev][nlut] concentration corner of the look-up-table gases
  double* cw2;         // [nlev][nlut]
  int ngas, nlut;
};
__host__ __device__ inline size_t ckd_layers_bytes(int nlev, int ngas, int nlut) {
  return (size_t)nlev * (sizeof(int2) + 3 * sizeof(double) + ngas * sizeof(double) + nlut * (sizeof(double) + sizeof(int))) + 32;
}
__device__ __forceinline__ unsigned char* ckd_layers_carve(unsigned char* base, int nlev, const CkdModel& m, CkdLayers& L) {
  L.ngas = m.ngas; L.nlut = m.nlut;
  double* d = reinterpret_cast<double*>(base);
  L.pw2 = d; d += nlev; L.tw2 = d; d += nlev; L.smult = d; d += nlev;
  L.mult = d; d += (size_t)nlev * m.ngas;
  L.cw2 = d; d += (size_t)nlev * m.nlut;
  L.corner = reinterpret_cast<int2*>(d);
  L.ic1 = reinterpret_cast<int*>(L.corner + nlev);
  return reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(L.ic1 + (size_t)nlev * m.nlut) + 15) & ~(uintptr_t)15);
}

__device__ __forceinline__ void ckd_layer_state(const CkdModel& m, const double* __restrict__ tab, const DevIn& in, int c, int l, const CkdLayers& L, bool gas_mmr) {
  const double global_multiplier = 1.0 / (9.80665 * 0.001 * 28.970);   // 1 / (AccelDueToGravity * 0.001 * AirMolarMass)
  const double p1 = LD_IN(in.p_hl, c, l), p2 = LD_IN(in.p_hl, c, l + 1);
  const double t1 = LD_IN(in.t_hl, c, l), t2 = LD_IN(in.t_hl, c, l + 1);
  const double temperature_fl = (t1 * p1 + t2 * p2) / (p1 + p2);       // radiation_ecckd_interface.F90:239-245
  const double log_pressure_fl = log(0.5 * (p1 + p2));
  double pindex1 = (log_pressure_fl - m.log_pressure1) / m.d_log_pressure;
  pindex1 = 1.0 + dmax(0.0, dmin(pindex1, m.npress - 1.0001));
  const int ip1 = (int)pindex1;
  const double pw2 = pindex1 - ip1, pw1 = 1.0 - pw2;
  const double* tt = tab + m.off_temperature1;
  const double temperature1 = pw1 * tt[ip1 - 1] + pw2 * tt[ip1];
  double tindex1 = (temperature_fl - temperature1) / m.d_temperature;
  tindex1 = 1.0 + dmax(0.0, dmin(tindex1, m.ntemp - 1.0001));
  const int it1 = (int)tindex1;
  L.corner[l] = make_int2(ip1, it1);
  L.pw2[l] = pw2; L.tw2[l] = tindex1 - it1;
  const double simple_multiplier = global_multiplier * (p2 - p1);
  L.smult[l] = simple_multiplier;
  for (int j = 0; j < m.ngas; ++j) {
    const CkdGas& G = m.gas[j];
    const double mf = G.slot >= 0 ? LD_IN(in.gas[G.slot], c, l) : 0.0;
    const double scaling = gas_mmr ? G.mmr_scaling : 1.0;   // local_concentration_scaling(igascode), radiation_ecckd.F90:518-519
    double mult;
    if (G.dep == CKD_CONC_LINEAR) mult = simple_multiplier * mf * scaling;
    else if (G.dep == CKD_CONC_RELATIVE_LINEAR) mult = simple_multiplier * (mf * scaling - G.reference_mole_frac);
    else if (G.dep == CKD_CONC_LUT) {
      const double log_conc = log(dmax(mf * scaling, G.mole_frac1));
      double cindex1 = (log_conc - G.log_mole_frac1) / G.d_log_mole_frac;
      cindex1 = 1.0 + dmax(0.0, dmin(cindex1, G.n_mole_frac - 1.0001));
      const int ic1 = (int)cindex1;
      L.ic1[(size_t)l * m.nlut + G.lut] = ic1;
      L.cw2[(size_t)l * m.nlut + G.lut] = cindex1 - ic1;
      mult = simple_multiplier * mf * scaling;
    } else mult = simple_multiplier;
    L.mult[(size_t)l * m.ngas + j] = mult;
  }
}

// absorption optical depth of one (layer, g-point): sum over the gases, clamped at zero (:558-639)
__device__ __forceinline__ double ckd_optical_depth(const CkdModel& m, const double* __restrict__ tab, const CkdLayers& L, int l, int g) {
  const size_t sp = (size_t)m.ng, st = (size_t)m.ng * m.npress, sc = st * m.ntemp;
  const int2 cr = L.corner[l];
  const double pw2 = L.pw2[l], pw1 = 1.0 - pw2, tw2 = L.tw2[l], tw1 = 1.0 - tw2;
  const size_t corner = (size_t)(cr.x - 1) * sp + (size_t)(cr.y - 1) * st + g;
  const double* mult = L.mult + (size_t)l * m.ngas;
  double od = 0.0;
#pragma unroll 8   // (the table loads of several gases in flight: the loop is a chain of L2 hits otherwise)
  for (int j = 0; j < m.ngas; ++j) {
    const CkdGas& G = m.gas[j];
    const double* k00 = tab + G.off + corner;
    if (G.dep == CKD_CONC_LUT) {
      const int ic1 = L.ic1[(size_t)l * m.nlut + G.lut];
      const double cw2 = L.cw2[(size_t)l * m.nlut + G.lut], cw1 = 1.0 - cw2;
      const size_t c0 = (size_t)(ic1 - 1) * sc, c1 = c0 + sc;
      od = od + mult[j] * ((cw1 * tw1 * pw1) * __ldg(k00 + c0) + (cw1 * tw1 * pw2) * __ldg(k00 + c0 + sp) +
                           (cw1 * tw2 * pw1) * __ldg(k00 + c0 + st) + (cw1 * tw2 * pw2) * __ldg(k00 + c0 + sp + st) +
                           (cw2 * tw1 * pw1) * __ldg(k00 + c1) + (cw2 * tw1 * pw2) * __ldg(k00 + c1 + sp) +
                           (cw2 * tw2 * pw1) * __ldg(k00 + c1 + st) + (cw2 * tw2 * pw2) * __ldg(k00 + c1 + sp + st));
    } else {
      od = od + mult[j] * (tw1 * (pw1 * __ldg(k00) + pw2 * __ldg(k00 + sp)) + tw2 * (pw1 * __ldg(k00 + st) + pw2 * __ldg(k00 + sp + st)));
    }
  }
  return dmax(0.0, od);
}

// aerosol state of one layer: humidity bin and (layer mass) x (mixing ratio) of every type (add_aerosol_optics :623-700)
// `row`: table row of each type in this layer (bit 15: hydrophilic table), resolved once per layer so that the per-g-point merge is a
// branch-free chain of independent loads; ignored types get mass 0 and row 0 (adds exact zeros).
struct AerLayers { int* irh; double* fm; unsigned short* row; int ntype; };   // [nlev], [nlev][ntype], [nlev][ntype]
__host__ __device__ inline size_t aer_layers_bytes(int nlev, int ntype) { return (size_t)nlev * (sizeof(int) + ntype * (sizeof(double) + sizeof(unsigned short))) + 16; }
__device__ __forceinline__ void aer_layers_carve(unsigned char* base, int nlev, int ntype, AerLayers& a) {
  a.ntype = ntype;
  a.fm = reinterpret_cast<double*>(base);
  a.irh = reinterpret_cast<int*>(a.fm + (size_t)nlev * ntype);
  a.row = reinterpret_cast<unsigned short*>(a.irh + nlev);
}
__device__ __forceinline__ void aer_layer_state(const AerMeta& A, const DevIn& in, int c, int l, int nlev, const AerLayers& a, bool gas_mmr) {
  // gas%mixing_ratio(:,:,IH2O) is a mole fraction under ecCKD: gas%get(IH2O, IMassMixingRatio) (:611, radiation_gas.F90:603-616)
  const double h2o_mmr = gas_mmr ? LD_IN(in.gas[0], c, l) : LD_IN(in.gas[0], c, l) * (18.0152833 / 28.970);
  const double rh = h2o_mmr / LD_IN(in.h2o_sat_liq, c, l);
  int irh;
  if (rh > A.rh_lower[A.nrh - 1]) irh = A.nrh;
  else { irh = 1; while (rh > A.rh_lower[irh]) ++irh; }
  a.irh[l] = irh;
  const double factor = (LD_IN(in.p_hl, c, l + 1) - LD_IN(in.p_hl, c, l)) * (1.0 / 9.80665);
  for (int jt = 0; jt < A.ntype; ++jt) {
    const int iclass = A.iclass[jt];
    a.fm[(size_t)l * A.ntype + jt] = iclass ? factor * in.aerosol_mmr[((size_t)jt * nlev + l) * in.ld + c] : 0.0;
    a.row[(size_t)l * A.ntype + jt] = iclass == 2 ? (unsigned short)(0x8000u | (unsigned)((A.itype[jt] - 1) * A.nrh + (irh - 1)))
                                    : iclass == 1 ? (unsigned short)(A.itype[jt] - 1) : (unsigned short)0;
  }
}

// =========================================================================================================
// LW: optical depth, Planck function at half-levels, surface emission and albedo per g-point
// =========================================================================================================
template <class SD>
__global__ void __launch_bounds__(CKD_THREADS, 10)
ckd_lw_kernel(DevTables T, DevCfg cfg, DevIn in, Work w, int nlev) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int c = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const CkdModel& m = T.ckd->lw;
  const double* tab = T.ckdtab;
  CkdLayers lay;
  double* ptw = reinterpret_cast<double*>(ckd_layers_carve(smem_raw, nlev, m, lay));   // [nlev+2][2]: Planck interpolation (tw2, it1 or -1)
  const bool do_aer = cfg.use_aerosols && T.aer;
  AerLayers aer;
  if (do_aer) aer_layers_carve(reinterpret_cast<unsigned char*>(ptw + 2 * (nlev + 2)), nlev, T.aer->ntype, aer);
  for (int l = tid; l < nlev; l += CKD_THREADS) {
    ckd_layer_state(m, tab, in, c, l, lay, cfg.gas_mmr != 0);
    if (do_aer) aer_layer_state(*T.aer, in, c, l, nlev, aer, cfg.gas_mmr != 0);
  }
  for (int k = tid; k < nlev + 2; k += CKD_THREADS) {   // half-levels 0..nlev, then the skin temperature
    const double temperature = k <= nlev ? LD_IN(in.t_hl, c, k) : in.skin_t[c];
    double tindex1 = (temperature - m.temperature1_planck) * (1.0 / m.d_temperature_planck);
    if (tindex1 >= 0) {
      tindex1 = 1.0 + tindex1;
      int it1 = (int)tindex1; if (it1 > m.nplanck - 1) it1 = m.nplanck - 1;
      ptw[2 * k] = tindex1 - it1; ptw[2 * k + 1] = (double)it1;
    } else {   // below the table: linear to zero
      ptw[2 * k] = temperature / m.temperature1_planck; ptw[2 * k + 1] = -1.0;
    }
  }
  __syncthreads();
  const double* planck_tab = tab + m.off_planck;
  auto planck = [&](int k, int g) {
    const double a = ptw[2 * k]; const int it1 = (int)ptw[2 * k + 1];
    if (it1 < 0) return __ldg(planck_tab + g) * a;
    const double tw2 = a, tw1 = 1.0 - tw2;
    return tw1 * __ldg(planck_tab + (size_t)(it1 - 1) * SD::NG + g) + tw2 * __ldg(planck_tab + (size_t)it1 * SD::NG + g);
  };
  double* od_out = w.od_lw + (size_t)c * nlev * SD::NG;
  double* pl_out = w.planck + (size_t)c * (nlev + 1) * SD::NG;
  const int me_pb = do_aer ? T.aer->me_lw_phobic : 0, me_pl = do_aer ? T.aer->me_lw_philic : 0;
  const int ss_pb = do_aer ? T.aer->ssa_lw_phobic : 0, ss_pl = do_aer ? T.aer->ssa_lw_philic : 0;
  for (int l = warp; l <= nlev; l += CKD_WARPS) {
    for (int g = lane; g < SD::NG; g += 32) {
      pl_out[(size_t)l * SD::NG + g] = planck(l, g);
      if (l < nlev) {
        double od = ckd_optical_depth(m, tab, lay, l, g);
        if (do_aer) {   // absorption optical depth of the aerosol mixture in this g-point (:700-722, no LW aerosol scattering)
          double od_aer = 0.0;
          const double* fm = aer.fm + (size_t)l * aer.ntype;
          const unsigned short* rw = aer.row + (size_t)l * aer.ntype;
#pragma unroll 4
          for (int jt = 0; jt < aer.ntype; ++jt) {
            const unsigned r = rw[jt];
            const size_t off = (size_t)(r & 0x7FFFu) * SD::NB + g;
            const double me = __ldg(T.aertab + ((r & 0x8000u) ? me_pl : me_pb) + off);
            const double ss = __ldg(T.aertab + ((r & 0x8000u) ? ss_pl : ss_pb) + off);
            od_aer = od_aer + fm[jt] * me * (1.0 - ss);
          }
          od = od + od_aer;
        }
        od_out[(size_t)l * SD::NG + g] = od;
      }
    }
  }
  // get_albedos (radiation_single_level.F90:310-355) and lw_emission (radiation_ecckd_interface.F90:309-315)
  for (int g = tid; g < SD::NG; g += CKD_THREADS) {
    double lw_albedo;
    if (cfg.do_nearest_spectral_lw_emiss) {
      lw_albedo = 1.0 - LD_IN(in.lw_emissivity, c, T.i_emiss_from_band_lw[g] - 1);
    } else {
      lw_albedo = 0.0;
      for (int ja = 0; ja < cfg.n_emiss_lw; ++ja) {
        const double wgt = T.lw_emiss_weights[g * cfg.n_emiss_lw + ja];
        if (wgt != 0.0) lw_albedo = lw_albedo + wgt * (1.0 - LD_IN(in.lw_emissivity, c, ja));
      }
    }
    w.lw_albedo[(size_t)c * SD::NG + g] = lw_albedo;
    w.emission[(size_t)c * SD::NG + g] = planck(nlev + 1, g) * (1.0 - lw_albedo);
  }
}

// =========================================================================================================
// SW: absorption + Rayleigh -> od / ssa, incoming flux per g-point, aerosol merge (od, ssa, g)
// =========================================================================================================
template <class SD>
__global__ void __launch_bounds__(CKD_THREADS, 10)
ckd_sw_kernel(DevTables T, DevCfg cfg, DevIn in, Work w, int nlev) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int c = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (!(in.cos_sza[c] > 0.0)) return;   // night column: the solvers do not read the optical properties
  const CkdModel& m = T.ckd->sw;
  const double* tab = T.ckdtab;
  CkdLayers lay;
  unsigned char* rest = ckd_layers_carve(smem_raw, nlev, m, lay);
  const bool do_aer = cfg.use_aerosols && T.aer;
  AerLayers aer;
  if (do_aer) aer_layers_carve(rest, nlev, T.aer->ntype, aer);
  for (int l = tid; l < nlev; l += CKD_THREADS) {
    ckd_layer_state(m, tab, in, c, l, lay, cfg.gas_mmr != 0);
    if (do_aer) aer_layer_state(*T.aer, in, c, l, nlev, aer, cfg.gas_mmr != 0);
  }
  __syncthreads();
  const size_t n = (size_t)nlev * SD::NG;
  double* od_out = w.od_sw + (size_t)c * n;
  double* ssa_out = w.ssa_sw + (size_t)c * n;
  double* g_out = w.g_sw ? w.g_sw + (size_t)c * n : nullptr;
  const double* rayl = tab + m.off_rayleigh;
  const int me_pb = do_aer ? T.aer->me_sw_phobic : 0, me_pl = do_aer ? T.aer->me_sw_philic : 0;
  const int ss_pb = do_aer ? T.aer->ssa_sw_phobic : 0, ss_pl = do_aer ? T.aer->ssa_sw_philic : 0;
  const int ga_pb = do_aer ? T.aer->g_sw_phobic : 0, ga_pl = do_aer ? T.aer->g_sw_philic : 0;
  for (int l = warp; l < nlev; l += CKD_WARPS) {
    for (int g = lane; g < SD::NG; g += 32) {
      double od = ckd_optical_depth(m, tab, lay, l, g);
      const double rayleigh = lay.smult[l] * __ldg(rayl + g);      // :642-647
      od = od + rayleigh;                                                       // radiation_ecckd_interface.F90:272-279
      double ssa = rayleigh / od;
      double gg = 0.0;
      if (do_aer) {
        double od_aer = 0.0, scat = 0.0, scat_g = 0.0;
        const double* fm = aer.fm + (size_t)l * aer.ntype;
        const unsigned short* rw = aer.row + (size_t)l * aer.ntype;
#pragma unroll 4
        for (int jt = 0; jt < aer.ntype; ++jt) {
          const unsigned r = rw[jt];
          const bool pl = (r & 0x8000u) != 0;
          const size_t off = (size_t)(r & 0x7FFFu) * SD::NB + g;
          const double me = __ldg(T.aertab + (pl ? me_pl : me_pb) + off);
          const double ss = __ldg(T.aertab + (pl ? ss_pl : ss_pb) + off);
          const double ga = __ldg(T.aertab + (pl ? ga_pl : ga_pb) + off);
          const double local_od = fm[jt] * me;
          od_aer = od_aer + local_od;
          scat = scat + local_od * ss;
          scat_g = scat_g + local_od * ss * ga;
        }
        if (!cfg.do_sw_delta_scaling_with_gases) {   // delta_eddington_extensive_vec, radiation_delta_eddington.h:74-96
          const double ge = scat_g / dmax(scat, (double)1.0e-24f);
          const double f = ge * ge;
          od_aer = od_aer - scat * f;
          scat = scat * (1.0 - f);
          scat_g = scat * ge / (1.0 + ge);
        }
        // combine with the gas (:753-790)
        const double local_od = od + od_aer;
        if (local_od > 0.0 && od_aer > 0.0) {
          const double local_scat = ssa * od + scat;
          if (local_scat > 0.0) gg = scat_g / local_scat;
          ssa = local_scat / local_od;
          od = local_od;
        }
      }
      od_out[(size_t)l * SD::NG + g] = od;
      ssa_out[(size_t)l * SD::NG + g] = ssa;
      if (g_out) g_out[(size_t)l * SD::NG + g] = gg;
    }
  }
  // calc_incoming_sw, radiation_ecckd.F90:935-964 (the multiplier is zero unless use_spectral_solar_cycle)
  const double* solar = tab + m.off_solar;
  if (cfg.solar_cycle_multiplier != 0.0 && m.off_solar_amp >= 0) {
    const double* amp = tab + m.off_solar_amp;
    for (int g = tid; g < SD::NG; g += CKD_THREADS)
      w.incoming[(size_t)c * SD::NG + g] = in.solar_irradiance * (__ldg(solar + g) + cfg.solar_cycle_multiplier * __ldg(amp + g));
  } else {
    for (int g = tid; g < SD::NG; g += CKD_THREADS) w.incoming[(size_t)c * SD::NG + g] = in.solar_irradiance * __ldg(solar + g);
  }
}

// =========================================================================================================
// generalised cloud optics: cloud types 1 (liquid) and 2 (ice) from look-up tables in effective radius
// =========================================================================================================
struct GcoAcc { double od, scat, scat_g; };
__device__ __forceinline__ void gco_add(const GcoType& ct, const double* __restrict__ tab, int ng, int g, double water_path, double re, GcoAcc& a) {
  const double re_index = dmax(1.0, dmin(1.0 + (re - ct.re0) / ct.dre, ct.nre - 0.0001));
  const int ire = (int)re_index;
  const double weight2 = re_index - ire, weight1 = 1.0 - weight2;
  const size_t i1 = (size_t)(ire - 1) * ng + g, i2 = i1 + ng;
  double od_local = water_path * (weight1 * __ldg(tab + ct.off_me + i1) + weight2 * __ldg(tab + ct.off_me + i2));
  a.od = a.od + od_local;
  od_local = od_local * (weight1 * __ldg(tab + ct.off_ssa + i1) + weight2 * __ldg(tab + ct.off_ssa + i2));
  a.scat = a.scat + od_local;
  a.scat_g = a.scat_g + od_local * (weight1 * __ldg(tab + ct.off_g + i1) + weight2 * __ldg(tab + ct.off_g + i2));
}
// no-scattering form (:316-326): absorption optical depth, only where there is condensate
__device__ __forceinline__ void gco_add_absorption(const GcoType& ct, const double* __restrict__ tab, int ng, int g, double water_path, double re, double& od) {
  if (!(water_path > 0.0)) return;
  const double re_index = dmax(1.0, dmin(1.0 + (re - ct.re0) / ct.dre, ct.nre - 0.0001));
  const int ire = (int)re_index;
  const double weight2 = re_index - ire, weight1 = 1.0 - weight2;
  const size_t i1 = (size_t)(ire - 1) * ng + g, i2 = i1 + ng;
  od = od + water_path * (weight1 * __ldg(tab + ct.off_me + i1) + weight2 * __ldg(tab + ct.off_me + i2)) *
                (1.0 - (weight1 * __ldg(tab + ct.off_ssa + i1) + weight2 * __ldg(tab + ct.off_ssa + i2)));
}
__device__ __forceinline__ void gco_finish(GcoAcc& a, bool delta_scale) {
  if (delta_scale) {   // delta_eddington_extensive, radiation_delta_eddington.h:46-69
    const double g = a.scat > 0.0 ? a.scat_g / a.scat : 0.0;
    const double f = g * g;
    a.od = a.od - a.scat * f;
    a.scat = a.scat * (1.0 - f);
    a.scat_g = a.scat * g / (1.0 + g);
  }
  a.scat_g = a.scat_g / dmax(a.scat, 1.0e-15);   // asymmetry factor
  a.scat = a.scat / dmax(a.od, 1.0e-15);         // single-scattering albedo
}

// warp = (column, layer), lane = g-point
__global__ void __launch_bounds__(CKD_THREADS)
general_cloud_optics_kernel(DevTables T, DevCfg cfg, DevIn in, Work w, int nc, int nlev) {
  const int c = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int l = blockIdx.y * CKD_WARPS + warp;
  if (l >= nlev) return;
  const double frac = LD_IN(in.frac, c, l);
  // The no-scattering longwave form adds absorption wherever there is condensate, cropped (frac = 0) layers included
  // (radiation_general_cloud_optics_data.F90:311-325); no solver reads those layers, save_radiative_properties shows them.
  const bool cloudy = frac > 0.0;
  if (!cloudy && !(cfg.do_lw && !cfg.do_lw_cloud_scattering)) return;
  const CkdMeta& M = *T.ckd;
  const double* tab = T.ckdtab;
  const double dp = LD_IN(in.p_hl, c, l + 1) - LD_IN(in.p_hl, c, l);
  const double inv = cfg.is_homogeneous ? 1.0 / 9.80665 : 1.0 / (9.80665 * dmax(cfg.cloud_fraction_threshold, frac));   // radiation_general_cloud_optics.F90:191-205
  const double wp_liq = LD_IN(in.q_liq, c, l) * dp * inv, wp_ice = LD_IN(in.q_ice, c, l) * dp * inv;
  const double rel = LD_IN(in.re_liq, c, l), rei = LD_IN(in.re_ice, c, l);
  if (cfg.do_lw) {
    const int ng = cfg.nb_lw;
    double* o = w.cl_lw + ((size_t)c * nlev + l) * 3 * ng;
    for (int g = lane; g < ng; g += 32) {
      GcoAcc a = {0.0, 0.0, 0.0};
      if (cfg.do_lw_cloud_scattering) {
        gco_add(M.gco_lw[0], tab, ng, g, wp_liq, rel, a);
        gco_add(M.gco_lw[1], tab, ng, g, wp_ice, rei, a);
        gco_finish(a, true);
      } else {
        gco_add_absorption(M.gco_lw[0], tab, ng, g, wp_liq, rel, a.od);
        gco_add_absorption(M.gco_lw[1], tab, ng, g, wp_ice, rei, a.od);
      }
      o[g] = a.od; o[ng + g] = a.scat; o[2 * ng + g] = a.scat_g;
    }
  }
  if (cfg.do_sw && cloudy) {
    const int ng = cfg.nb_sw;
    double* o = w.cl_sw + ((size_t)c * nlev + l) * 3 * ng;
    for (int g = lane; g < ng; g += 32) {
      GcoAcc a = {0.0, 0.0, 0.0};
      gco_add(M.gco_sw[0], tab, ng, g, wp_liq, rel, a);
      gco_add(M.gco_sw[1], tab, ng, g, wp_ice, rei, a);
      gco_finish(a, !cfg.do_sw_delta_scaling_with_gases);
      o[g] = a.od; o[ng + g] = a.scat; o[2 * ng + g] = a.scat_g;
    }
  }
}

// =========================================================================================================
// launchers
// =========================================================================================================
static size_t ckd_smem(int nlev, int ngas, int nlut, int n_aerosol_types, bool lw) {
  return ckd_layers_bytes(nlev, ngas, nlut) + (lw ? sizeof(double) * 2 * (nlev + 2) : 0) + (n_aerosol_types ? aer_layers_bytes(nlev, n_aerosol_types) : 0) + 16;
}
template <class SD>
static int launch_ckd_lw_t(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st) {
  const size_t sm = ckd_smem(nlev, cfg.ckd_ngas_lw, cfg.ckd_nlut_lw, (cfg.use_aerosols && T.aer) ? cfg.n_aerosol_types : 0, true);
  cudaFuncSetAttribute(ckd_lw_kernel<SD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  ckd_lw_kernel<SD><<<nc, CKD_THREADS, sm, st>>>(T, cfg, in, w, nlev);
  return 1;
}
template <class SD>
static int launch_ckd_sw_t(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st) {
  const size_t sm = ckd_smem(nlev, cfg.ckd_ngas_sw, cfg.ckd_nlut_sw, (cfg.use_aerosols && T.aer) ? cfg.n_aerosol_types : 0, false);
  cudaFuncSetAttribute(ckd_sw_kernel<SD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  ckd_sw_kernel<SD><<<nc, CKD_THREADS, sm, st>>>(T, cfg, in, w, nlev);
  return 1;
}
int launch_ckd_lw(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st) {
  switch (cfg.ng_lw) {
    case 32: return launch_ckd_lw_t<Ckd32>(T, cfg, in, w, nc, nlev, st);
    case 64: return launch_ckd_lw_t<Ckd64>(T, cfg, in, w, nc, nlev, st);
    case 96: return launch_ckd_lw_t<Ckd96>(T, cfg, in, w, nc, nlev, st);
  }
  return -1;
}
int launch_ckd_sw(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st) {
  switch (cfg.ng_sw) {
    case 32: return launch_ckd_sw_t<Ckd32>(T, cfg, in, w, nc, nlev, st);
    case 64: return launch_ckd_sw_t<Ckd64>(T, cfg, in, w, nc, nlev, st);
    case 96: return launch_ckd_sw_t<Ckd96>(T, cfg, in, w, nc, nlev, st);
  }
  return -1;
}
int launch_general_cloud_optics(const DevTables& T, const DevCfg& cfg, const DevIn& in, const Work& w, int nc, int nlev, cudaStream_t st) {
  dim3 grid(nc, (nlev + CKD_WARPS - 1) / CKD_WARPS);
  general_cloud_optics_kernel<<<grid, CKD_THREADS, 0, st>>>(T, cfg, in, w, nc, nlev);
  return 1;
}

}  // namespace ecb
