// gas_core.h -- RRTMG gas optics recast as a table-driven sparse stencil.
//
// The reference evaluates 16 LW + 14 SW hand-written band routines (ifsrrtm/rrtm_taumol1..16.F90,
// srtm_taumol16..29.F90), each a different mix of 2-D/3-D interpolations in k-tables.  Every one of them has the
// algebraic form
//        tau(ig) = [post(ig)] * sum_k  coef_k * TAB_band[row_k][ig]          (k <= 21 terms)
// where coef_k and row_k depend only on (column, layer, band) and the table row is contiguous in the g-point.
// So the GPU path splits gas optics in two stages:
//   stage A  (one thread per layer x band):  build the list {coef_k, row offset_k}   -- this file
//   stage B  (one lane per g-point):         tau = sum_k coef_k * tab[off_k + ig]    -- coalesced row reads
// Stage A carries all the band-specific physics; stage B is one uniform loop.  Results agree with the reference
// to rounding (sums are re-associated: a + f*(b-a) becomes (1-f)*a + f*b); no discrete decision is changed.
//
// Reference map:  lev_prepare   <- ifsrrtm/rrtm_prepare_gases.F90:150-229
//                 lw_setcoef    <- ifsrrtm/rrtm_setcoef_140gp.F90:84-276
//                 lw_build_list <- ifsrrtm/rrtm_taumol1..16.F90
//                 sw_setcoef    <- ifsrrtm/srtm_setcoef.F90:78-220
//                 sw_build_list <- ifsrrtm/srtm_taumol16..29.F90 (+ srtm_gas_optical_depth.F90:305-321)
//                 planck_band   <- radiation/radiation_ifs_rrtm.F90:676-699
#pragma once
#include "hd.h"

namespace ecb {

enum { NG_LW = 140, NG_SW = 112, NB_LW = 16, NB_SW = 14 };   // RRTMG

// Spectral sizes of a solver instantiation: g-points, bands (cloud/aerosol optics resolution), threads of the one-CTA-per-
// column kernels (one thread per g-point), row stride of the shared-memory reduction tiles.
template <int NG_, int NB_> struct SpecDims { enum { NG = NG_, NB = NB_, THREADS = (NG_ + 31) / 32 * 32, RS = NG_ + 1 }; };
typedef SpecDims<NG_LW, NB_LW> LwRrtmg;
typedef SpecDims<NG_SW, NB_SW> SwRrtmg;
// ecCKD models: cloud and aerosol optics per g-point, i.e. bands == g-points (radiation_ecckd_interface.F90:46-76)
typedef SpecDims<32, 32> Ckd32;
typedef SpecDims<64, 64> Ckd64;
typedef SpecDims<96, 96> Ckd96;
// minimum resident CTAs per SM that keeps the register budget of a kernel tuned as __launch_bounds__(t0, b0)
constexpr int scaled_min_blocks(int threads, int t0, int b0) { return (t0 * b0 / threads) > 32 ? 32 : (t0 * b0 / threads); }
enum { LW_KMAX = 22, SW_KMAX = 14 };

// Sections of a band's packed table (rows of ng doubles, g-point fastest).
enum LwSec { L_ABSA, L_ABSB, L_SELF, L_FOR, L_FRACA, L_FRACB, L_M0, L_M1, L_M2, L_M3, L_M4, L_C0, L_C1, L_POST, L_NSEC };
enum SwSec { S_ABSA, S_ABSB, S_SELF, S_FOR, S_SFLUX, S_RAYA, S_RAYB, S_X0, S_X1, S_ONES, S_NSEC };
// The packer lays a band out as [ABSA][ABSB][everything else]: the two big major-species tables, which the band-wise kernels
// stage by pressure window, then the small sections, staged whole.  sec[SEC_SMALL] = first element of the small sections,
// sec[SEC_END] = one past the band's last element.
enum { SEC_SMALL = 14, SEC_END = 15 };

struct BandMeta {
  int ng;        // g-points in this band
  int g0;        // first g-point (0-based) in the 140 / 112 vector
  int sec[16];   // element offset of row 0 of each section in the packed table, -1 if the band has none
};

// Small read-only tables and per-band scalars (device global memory; ~36 KB).
struct GasMeta {
  BandMeta lw[NG_LW], sw[NG_SW];   // RRTMG: 16 / 14 bands; ecCKD: one single-g-point band per g-point
  double preflog_lw[59], tref_lw[59], chi_mls[7 * 59];
  double preflog_sw[59], tref_sw[59];
  double totplnk[181 * 16], delwave[16];
  // CHI_MLS(i,jp)/CHI_MLS(j,jp) for the species pairs the binary-species bands use (rrtm_setcoef_140gp.F90:169-182),
  // tabulated once on the host with the same IEEE division: [pair][jp-1], pairs: h2o/co2 h2o/o3 h2o/n2o h2o/ch4 n2o/co2 o3/co2
  double chi_rat[6][59];
  double strrat_sw[NB_SW], rayl_sw[NB_SW], givfac_23, scalekur_27;
  int layreffr_sw[NB_SW], nfor_sw[NB_SW];
  int band_of_g_lw[NG_LW], band_of_g_sw[NG_SW];  // 0-based band of each g-point
  // SPARTACUS on RRTMG: 0-based position of each g-point in the order of approximately increasing gas optical depth
  // (inverse of config%i_g_from_reordered_g_{lw,sw}, radiation_ifs_rrtm.F90:50-68, :122-130, :167-174); identity otherwise
  short rank_lw[NG_LW], rank_sw[NG_SW];
  int lw_rows[NB_LW][2], sw_rows[NB_SW][2];      // rows of ABSA / ABSB per band (0: none); 13 (ABSA) or 47 (ABSB) reference pressures each
  short lw_sec_rows[NB_LW][16], sw_sec_rows[NB_SW][16];   // rows of every section of a band's packed table
  unsigned short lw_sec_low[NB_LW], lw_sec_high[NB_LW];   // small sections the band routine reads below / above LAYTROP (bit = section)
};
// g-points per band of the RRTMG-IFS reduction (ifsrrtm/rrtm_init_140gp.F90 NGC, srtm_init.F90 NGC): compile-time sizes of the
// band-wise kernels' register accumulators; ecrad_b200_setup checks the tables against them.
constexpr int kNgLwBand[NB_LW] = {10, 12, 16, 14, 16, 8, 12, 8, 12, 6, 8, 8, 4, 2, 2, 2};
constexpr int kNgSwBand[NB_SW] = {6, 12, 8, 8, 10, 10, 2, 10, 8, 6, 6, 8, 6, 12};

// ---------------------------------------------------------------------------------------------------------
// Per-layer state shared by LW and SW (rrtm_prepare_gases): hPa, K, molecules cm-2
// ---------------------------------------------------------------------------------------------------------
struct LevGas {
  double pavel, tavel, coldry, wbroad;
  double wkl1, wkl2, wkl3, wkl4, wkl6, wkl7;  // h2o co2 o3 n2o ch4 o2 column amounts
  double wx1, wx2, wx3, wx4;                  // ccl4 cfc11 cfc12 cfc22  (x 1e-20)
};

// p_top/p_bot: half-level pressures (Pa) bounding the layer (top = smaller index in ecRad order); mass mixing ratios.
HD void lev_prepare(double p_top, double p_bot, double t_top, double t_bot, double q, double co2, double ch4,
                    double n2o, double cfc11, double cfc12, double hcfc22, double ccl4, double o3, LevGas& L) {
  const double ZAMD = 28.970, ZAMW = 18.0154, ZAMCO2 = 44.011, ZAMO = 47.9982, ZAMCH4 = 16.043, ZAMN2O = 44.013,
               ZAMC11 = 137.3686, ZAMC12 = 120.9140, ZAMC22 = 86.4690, ZAMCL4 = 153.8230, ZAVGDRO = 6.02214E23;
  const double ZGRAVIT = (9.80665 / 1.0) * 1.E2;
  L.pavel = (0.5 * (p_top + p_bot)) / 100.0;   // radiation_ifs_rrtm.F90:388-397 then rrtm_prepare_gases
  L.tavel = 0.5 * (t_top + t_bot);
  double w1 = dmax(q, (double)1.0E-15f) * ZAMD / ZAMW;
  double w2 = co2 * ZAMD / ZAMCO2, w3 = o3 * ZAMD / ZAMO, w4 = n2o * ZAMD / ZAMN2O, w6 = ch4 * ZAMD / ZAMCH4;
  double w7 = 0.209488;
  double zamm = (1.0 - w1) * ZAMD + w1 * ZAMW;
  L.coldry = (p_bot / 100.0 - p_top / 100.0) * 1.E3 * ZAVGDRO / (ZGRAVIT * zamm * (1.0 + w1));
  L.wx1 = L.coldry * (ccl4 * ZAMD / ZAMCL4) * 1.E-20;
  L.wx2 = L.coldry * (cfc11 * ZAMD / ZAMC11) * 1.E-20;
  L.wx3 = L.coldry * (cfc12 * ZAMD / ZAMC12) * 1.E-20;
  L.wx4 = L.coldry * (hcfc22 * ZAMD / ZAMC22) * 1.E-20;
  double summol = 0.0;
  summol = summol + w2; summol = summol + w3; summol = summol + w4; summol = summol + 0.0; summol = summol + w6; summol = summol + w7;
  L.wbroad = L.coldry * (1.0 - summol);
  L.wkl1 = L.coldry * w1; L.wkl2 = L.coldry * w2; L.wkl3 = L.coldry * w3; L.wkl4 = L.coldry * w4;
  L.wkl6 = L.coldry * w6; L.wkl7 = L.coldry * w7;
}

// ---------------------------------------------------------------------------------------------------------
// LW set-coefficients
// ---------------------------------------------------------------------------------------------------------
struct LwLev {
  int jp, jt, jt1, indself, indfor, indminor, tropo;  // tropo: plog > 4.56 (counted to get LAYTROP)
  double fac00, fac01, fac10, fac11, forfac, forfrac, selffac, selffrac, scaleminor, scaleminorn2, minorfrac;
  double colh2o, colco2, colo3, coln2o, colch4, colo2, colbrd, coldry, pavel;
  double wx1, wx2, wx3, wx4;
  double t_top, t_bot;   // half-level temperatures bounding the layer (Planck function of the band-wise kernel)
  double pad_;   // sizeof = 31 doubles: odd stride => conflict-free shared-memory access across layers
};
static_assert(sizeof(LwLev) % 16 == 8, "LwLev stride must be an odd number of doubles");

#define ECB_CHI(i, j) (M.chi_mls[((j) - 1) * 7 + ((i) - 1)])
enum { RAT_H2OCO2 = 0, RAT_H2OO3 = 1, RAT_H2ON2O = 2, RAT_H2OCH4 = 3, RAT_N2OCO2 = 4, RAT_O3CO2 = 5 };
#define ECB_RAT(pair, j) (M.chi_rat[pair][(j) - 1])

HD void lw_setcoef(const GasMeta& M, const LevGas& G, LwLev& L) {
  const double stpfac = 296.0 / 1013.0;
  double plog = log(G.pavel);
  int jp = (int)(36.0 - 5 * (plog + 0.04));
  if (jp < 1) jp = 1; else if (jp > 58) jp = 58;
  int jp1 = jp + 1;
  double fp = 5.0 * (M.preflog_lw[jp - 1] - plog);
  fp = dmax(-1.0, dmin(1.0, fp));
  int jt = (int)(3.0 + (G.tavel - M.tref_lw[jp - 1]) / 15.0);
  if (jt < 1) jt = 1; else if (jt > 4) jt = 4;
  double ft = ((G.tavel - M.tref_lw[jp - 1]) / 15.0) - (double)(jt - 3);
  int jt1 = (int)(3.0 + (G.tavel - M.tref_lw[jp1 - 1]) / 15.0);
  if (jt1 < 1) jt1 = 1; else if (jt1 > 4) jt1 = 4;
  double ft1 = ((G.tavel - M.tref_lw[jp1 - 1]) / 15.0) - (double)(jt1 - 3);
  double water = G.wkl1 / G.coldry;
  double scalefac = G.pavel * stpfac / G.tavel;
  L.jp = jp; L.jt = jt; L.jt1 = jt1;
  L.tropo = plog > 4.56;
  double factor;
  L.forfac = scalefac / (1.0 + water);
  L.selffac = water * L.forfac;
  L.indself = 0; L.selffrac = 0.0;
  if (L.tropo) {
    factor = (332.0 - G.tavel) / 36.0;
    L.indfor = imin(2, imax(1, (int)factor));
    L.forfrac = factor - (double)L.indfor;
    factor = (G.tavel - 188.0) / 7.2;
    L.indself = imin(9, imax(1, (int)factor - 7));
    L.selffrac = factor - (double)(L.indself + 7);
  } else {
    factor = (G.tavel - 188.0) / 36.0;
    L.indfor = 3;
    L.forfrac = factor - 1.0;
  }
  L.scaleminor = G.pavel / G.tavel;
  L.scaleminorn2 = (G.pavel / G.tavel) * (G.wbroad / (G.coldry + G.wkl1));
  factor = (G.tavel - 180.8) / 7.2;
  L.indminor = imin(18, imax(1, (int)factor));
  L.minorfrac = factor - (double)L.indminor;
  L.colh2o = 1.E-20 * G.wkl1; L.colco2 = 1.E-20 * G.wkl2; L.colo3 = 1.E-20 * G.wkl3;
  L.coln2o = 1.E-20 * G.wkl4; L.colch4 = 1.E-20 * G.wkl6; L.colo2 = 1.E-20 * G.wkl7;
  L.colbrd = 1.E-20 * G.wbroad;
  if (L.colco2 == 0.0) L.colco2 = 1.E-32 * G.coldry;
  if (L.coln2o == 0.0) L.coln2o = 1.E-32 * G.coldry;
  if (L.colch4 == 0.0) L.colch4 = 1.E-32 * G.coldry;
  double compfp = 1.0 - fp;
  L.fac10 = compfp * ft; L.fac00 = compfp * (1.0 - ft);
  L.fac11 = fp * ft1;    L.fac01 = fp * (1.0 - ft1);
  L.selffac = L.colh2o * L.selffac;
  L.forfac = L.colh2o * L.forfac;
  L.coldry = G.coldry; L.pavel = G.pavel;
  L.wx1 = G.wx1; L.wx2 = G.wx2; L.wx3 = G.wx3; L.wx4 = G.wx4;
}

// Storage of the per-layer state between gas_prep_kernel and the gas-optics kernels: per column a block [field][layer] of doubles
// (integers stored exactly), layer fastest -- whichever way a kernel maps threads to layers, a field of consecutive layers is one
// contiguous run (a struct-per-layer array made every field load touch one cache line per lane).
enum { LWLEV_NF = 33 };
HD void lwlev_store(double* col, int stride, int l, const LwLev& L) {
  double* p = col + l;
  p[0 * stride] = L.jp; p[1 * stride] = L.jt; p[2 * stride] = L.jt1; p[3 * stride] = L.indself; p[4 * stride] = L.indfor;
  p[5 * stride] = L.indminor; p[6 * stride] = L.tropo;
  p[7 * stride] = L.fac00; p[8 * stride] = L.fac01; p[9 * stride] = L.fac10; p[10 * stride] = L.fac11; p[11 * stride] = L.forfac;
  p[12 * stride] = L.forfrac; p[13 * stride] = L.selffac; p[14 * stride] = L.selffrac; p[15 * stride] = L.scaleminor;
  p[16 * stride] = L.scaleminorn2; p[17 * stride] = L.minorfrac; p[18 * stride] = L.colh2o; p[19 * stride] = L.colco2;
  p[20 * stride] = L.colo3; p[21 * stride] = L.coln2o; p[22 * stride] = L.colch4; p[23 * stride] = L.colo2; p[24 * stride] = L.colbrd;
  p[25 * stride] = L.coldry; p[26 * stride] = L.pavel; p[27 * stride] = L.wx1; p[28 * stride] = L.wx2; p[29 * stride] = L.wx3;
  p[30 * stride] = L.wx4; p[31 * stride] = L.t_top; p[32 * stride] = L.t_bot;
}
HD LwLev lwlev_load(const double* col, int stride, int l) {
  const double* p = col + l;
  LwLev L;
  L.jp = (int)p[0 * stride]; L.jt = (int)p[1 * stride]; L.jt1 = (int)p[2 * stride]; L.indself = (int)p[3 * stride];
  L.indfor = (int)p[4 * stride]; L.indminor = (int)p[5 * stride]; L.tropo = (int)p[6 * stride];
  L.fac00 = p[7 * stride]; L.fac01 = p[8 * stride]; L.fac10 = p[9 * stride]; L.fac11 = p[10 * stride]; L.forfac = p[11 * stride];
  L.forfrac = p[12 * stride]; L.selffac = p[13 * stride]; L.selffrac = p[14 * stride]; L.scaleminor = p[15 * stride];
  L.scaleminorn2 = p[16 * stride]; L.minorfrac = p[17 * stride]; L.colh2o = p[18 * stride]; L.colco2 = p[19 * stride];
  L.colo3 = p[20 * stride]; L.coln2o = p[21 * stride]; L.colch4 = p[22 * stride]; L.colo2 = p[23 * stride]; L.colbrd = p[24 * stride];
  L.coldry = p[25 * stride]; L.pavel = p[26 * stride]; L.wx1 = p[27 * stride]; L.wx2 = p[28 * stride]; L.wx3 = p[29 * stride];
  L.wx4 = p[30 * stride]; L.t_top = p[31 * stride]; L.t_bot = p[32 * stride]; L.pad_ = 0.0;
  return L;
}

// ---------------------------------------------------------------------------------------------------------
// list emitters
// ---------------------------------------------------------------------------------------------------------
struct alignas(16) Term { double c; int o; int pad; };   // 16 bytes: coefficient, element offset of the table row (add the in-band g index)
struct ListOut {
  Term* t;      // term k lives at t[k * stride] (the kernels interleave the lists of 16 layers: stride = 16)
  int n;
  int stride;
  HD void add(double coef, int off) { Term x; x.c = coef; x.o = off; x.pad = 0; t[n * stride] = x; ++n; }   // one 128-bit store
  HD void pad4() { while (n & 3) add(0.0, 0); }   // zero terms so that consumers can unroll by 4
  HD void scale_all(double f) { for (int k = 0; k < n; ++k) t[k * stride].c = f * t[k * stride].c; }   // multiply what has been emitted so far
};

struct Spec { double speccomb, specparm, fs; int js; };

HD Spec mkspec(double cola, double rat, double colb, double mult) {
  const double oneminus = 1.0 - 1.0e-6;
  Spec s;
  s.speccomb = cola + rat * colb;
  s.specparm = cola / s.speccomb;
  if (s.specparm >= oneminus) s.specparm = oneminus;
  double specmult = mult * s.specparm;
  s.js = 1 + (int)specmult;
  s.fs = specmult - (double)(int)specmult;  // MOD(specmult, 1.0), specmult >= 0
  return s;
}
// rrtm_taumol7.F90:126-150 writes SPECPARM as 1/(1 + rat/colh2o*colo3)
HD Spec mkspec7(double cola, double rat, double colb, double mult) {
  const double oneminus = 1.0 - 1.0e-6;
  Spec s;
  s.speccomb = cola + rat * colb;
  s.specparm = 1.0 / (1.0 + rat / cola * colb);
  if (s.specparm >= oneminus) s.specparm = oneminus;
  double specmult = mult * s.specparm;
  s.js = 1 + (int)specmult;
  s.fs = specmult - (double)(int)specmult;
  return s;
}

// col * (fac00*T[i0] + fac10*T[i0+1] + fac01*T[i1] + fac11*T[i1+1]);  i0,i1 are the reference's 1-based rows
template <class Out> HD void emit_major1(Out& out, int sec, int ng, int i0, int i1, double col, double f00, double f10, double f01, double f11) {
  out.add(col * f00, sec + (i0 - 1) * ng);
  out.add(col * f10, sec + i0 * ng);
  out.add(col * f01, sec + (i1 - 1) * ng);
  out.add(col * f11, sec + i1 * ng);
}
// lower atmosphere, NSPA=9 (e.g. rrtm_taumol3.F90:170-229): three-point stencils at the ends of the eta range
template <class Out> HD void emit_low(Out& out, int sec, int ng, int ind, const Spec& s, double fa, double fb) {
  const double sc = s.speccomb;
  if (s.specparm < 0.125) {
    double p = s.fs - 1, p4 = (p * p) * (p * p);
    double fk0 = p4, fk1 = 1 - p - 2.0 * p4, fk2 = p + p4;
    out.add(sc * fk0 * fa, sec + (ind - 1) * ng);  out.add(sc * fk1 * fa, sec + ind * ng);        out.add(sc * fk2 * fa, sec + (ind + 1) * ng);
    out.add(sc * fk0 * fb, sec + (ind + 8) * ng);  out.add(sc * fk1 * fb, sec + (ind + 9) * ng);  out.add(sc * fk2 * fb, sec + (ind + 10) * ng);
  } else if (s.specparm > 0.875) {
    double p = -s.fs, p4 = (p * p) * (p * p);
    double fk0 = p4, fk1 = 1 - p - 2.0 * p4, fk2 = p + p4;
    out.add(sc * fk2 * fa, sec + (ind - 2) * ng);  out.add(sc * fk1 * fa, sec + (ind - 1) * ng);  out.add(sc * fk0 * fa, sec + ind * ng);
    out.add(sc * fk2 * fb, sec + (ind + 7) * ng);  out.add(sc * fk1 * fb, sec + (ind + 8) * ng);  out.add(sc * fk0 * fb, sec + (ind + 9) * ng);
  } else {
    out.add(sc * (1.0 - s.fs) * fa, sec + (ind - 1) * ng);  out.add(sc * s.fs * fa, sec + ind * ng);
    out.add(sc * (1.0 - s.fs) * fb, sec + (ind + 8) * ng);  out.add(sc * s.fs * fb, sec + (ind + 9) * ng);
  }
}
// upper atmosphere, NSPB=5 (e.g. rrtm_taumol3.F90:301-308)
template <class Out> HD void emit_upp(Out& out, int sec, int ng, int ind, const Spec& s, double fa, double fb) {
  const double sc = s.speccomb;
  out.add(sc * (1.0 - s.fs) * fa, sec + (ind - 1) * ng);  out.add(sc * s.fs * fa, sec + ind * ng);
  out.add(sc * (1.0 - s.fs) * fb, sec + (ind + 4) * ng);  out.add(sc * s.fs * fb, sec + (ind + 5) * ng);
}
// scale * (T[i] + f*(T[i+1]-T[i])), i 1-based
template <class Out> HD void emit_lin(Out& out, int sec, int ng, int i, double f, double scale) {
  out.add(scale * (1.0 - f), sec + (i - 1) * ng);
  out.add(scale * f, sec + i * ng);
}
// minor species on a (nj, 19) grid: rows (indm-1)*nj + (j-1)   (e.g. rrtm_taumol3.F90:235-239)
template <class Out> HD void emit_minor2(Out& out, int sec, int ng, int nj, int j, double fj, int indm, double mf, double scale) {
  int r = (indm - 1) * nj + (j - 1);
  out.add(scale * (1.0 - mf) * (1.0 - fj), sec + r * ng);
  out.add(scale * (1.0 - mf) * fj, sec + (r + 1) * ng);
  out.add(scale * mf * (1.0 - fj), sec + (r + nj) * ng);
  out.add(scale * mf * fj, sec + (r + nj + 1) * ng);
}
HD double adjcol(double col, double coldry, double chiref, double thresh, double base, double expo) {
  double chi = col / coldry;
  double rat = 1.E20 * chi / chiref;
  if (rat > thresh) {
    double adjfac = base + pow(rat - base, expo);
    return adjfac * chiref * coldry * 1.E-20;
  }
  return col;
}

struct PlanckFrac { double c0, c1; int o0, o1; };
HD PlanckFrac pf_const(int sec) { PlanckFrac p; p.c0 = 1.0; p.c1 = 0.0; p.o0 = sec; p.o1 = sec; return p; }
HD PlanckFrac pf_interp(int sec, int ng, const Spec& s) {
  PlanckFrac p; p.c0 = 1.0 - s.fs; p.c1 = s.fs; p.o0 = sec + (s.js - 1) * ng; p.o1 = sec + s.js * ng; return p;
}
HD PlanckFrac pf_zero(int sec) { PlanckFrac p; p.c0 = 0.0; p.c1 = 0.0; p.o0 = sec; p.o1 = sec; return p; }

// Build the stencil of LW band `ib` (0-based: 0 = RRTMG band 1) for one layer.  `low` = layer index <= LAYTROP.
// Returns the Planck-fraction stencil; *post = element offset of a per-g multiplier row or -1.
// `B` = the band's section offsets and row stride (B.ng) in the address space `out` reads from: M.lw[ib] for the packed table in
// global memory, a remapped copy for a shared-memory image.  `Out` = ListOut (materialise the list) or an evaluating sink.
template <class Out>
HD PlanckFrac lw_build_list(const GasMeta& M, const BandMeta& B, const LwLev& L, int ib, bool low, Out& out, int* post) {
  const int ng = B.ng;
  const int A = B.sec[L_ABSA], Bb = B.sec[L_ABSB], SF = B.sec[L_SELF], FR = B.sec[L_FOR];
  const int FA = B.sec[L_FRACA], FB = B.sec[L_FRACB];
  const int jp = L.jp;
  const int i0a1 = ((jp - 1) * 5 + (L.jt - 1)) + 1, i1a1 = (jp * 5 + (L.jt1 - 1)) + 1;          // NSPA = 1
  const int i0b1 = ((jp - 13) * 5 + (L.jt - 1)) + 1, i1b1 = ((jp - 12) * 5 + (L.jt1 - 1)) + 1;  // NSPB = 1
  const int i0a9 = ((jp - 1) * 5 + (L.jt - 1)) * 9, i1a9 = (jp * 5 + (L.jt1 - 1)) * 9;          // NSPA = 9 (+js)
  const int i0b5 = ((jp - 13) * 5 + (L.jt - 1)) * 5, i1b5 = ((jp - 12) * 5 + (L.jt1 - 1)) * 5;  // NSPB = 5 (+js)
  const int indm = L.indminor;
  const double mf = L.minorfrac;
  out.n = 0;
  *post = -1;
  PlanckFrac pf = pf_zero(FA);
#define SELF_() emit_lin(out, SF, ng, L.indself, L.selffrac, L.selffac)
#define FOR_() emit_lin(out, FR, ng, L.indfor, L.forfrac, L.forfac)
#define MAJ1A(col) emit_major1(out, A, ng, i0a1, i1a1, col, L.fac00, L.fac10, L.fac01, L.fac11)
#define MAJ1B(col) emit_major1(out, Bb, ng, i0b1, i1b1, col, L.fac00, L.fac10, L.fac01, L.fac11)
#define MAJ9(s, s1) do { emit_low(out, A, ng, i0a9 + (s).js, (s), L.fac00, L.fac10); emit_low(out, A, ng, i1a9 + (s1).js, (s1), L.fac01, L.fac11); } while (0)
#define MAJ5(s, s1) do { emit_upp(out, Bb, ng, i0b5 + (s).js, (s), L.fac00, L.fac10); emit_upp(out, Bb, ng, i1b5 + (s1).js, (s1), L.fac01, L.fac11); } while (0)
  switch (ib + 1) {
    case 1: {  // rrtm_taumol1.F90: H2O / H2O, minor N2
      double scalen2 = L.colbrd * L.scaleminorn2, corradj;
      if (low) {
        corradj = 1.;
        if (L.pavel < 250.0) corradj = 1.0 - 0.15 * (250.0 - L.pavel) / 154.4;
        MAJ1A(L.colh2o); SELF_(); FOR_();
        emit_lin(out, B.sec[L_M0], ng, indm, mf, scalen2);
        pf = pf_const(FA);
      } else {
        corradj = 1.0 - 0.15 * (L.pavel / 95.6);
        MAJ1B(L.colh2o); FOR_();
        emit_lin(out, B.sec[L_M1], ng, indm, mf, scalen2);
        pf = pf_const(FB);
      }
      out.scale_all(corradj);
    } break;
    case 2: {  // rrtm_taumol2.F90: H2O / H2O
      if (low) {
        double corradj = 1.0 - .05 * (L.pavel - 100.0) / 900.0;
        MAJ1A(L.colh2o); SELF_(); FOR_();
        out.scale_all(corradj);
        pf = pf_const(FA);
      } else {
        MAJ1B(L.colh2o); FOR_();
        pf = pf_const(FB);
      }
    } break;
    case 3: {  // rrtm_taumol3.F90: H2O,CO2 / H2O,CO2; minor N2O
      double adjcoln2o = adjcol(L.coln2o, L.coldry, ECB_CHI(4, jp + 1), 1.5, 0.5, 0.65);
      double rat = ECB_RAT(RAT_H2OCO2, jp), rat1 = ECB_RAT(RAT_H2OCO2, jp + 1);
      if (low) {
        Spec s = mkspec(L.colh2o, rat, L.colco2, 8.0), s1 = mkspec(L.colh2o, rat1, L.colco2, 8.0);
        Spec sm = mkspec(L.colh2o, ECB_RAT(RAT_H2OCO2, 3), L.colco2, 8.0);
        Spec sp = mkspec(L.colh2o, ECB_RAT(RAT_H2OCO2, 9), L.colco2, 8.0);
        MAJ9(s, s1); SELF_(); FOR_();
        emit_minor2(out, B.sec[L_M0], ng, 9, sm.js, sm.fs, indm, mf, adjcoln2o);
        pf = pf_interp(FA, ng, sp);
      } else {
        Spec s = mkspec(L.colh2o, rat, L.colco2, 4.0), s1 = mkspec(L.colh2o, rat1, L.colco2, 4.0);
        Spec sm = mkspec(L.colh2o, ECB_RAT(RAT_H2OCO2, 13), L.colco2, 4.0);
        MAJ5(s, s1); FOR_();
        emit_minor2(out, B.sec[L_M1], ng, 5, sm.js, sm.fs, indm, mf, adjcoln2o);
        pf = pf_interp(FB, ng, sm);
      }
    } break;
    case 4: {  // rrtm_taumol4.F90: H2O,CO2 / O3,CO2
      if (low) {
        double rat = ECB_RAT(RAT_H2OCO2, jp), rat1 = ECB_RAT(RAT_H2OCO2, jp + 1);
        Spec s = mkspec(L.colh2o, rat, L.colco2, 8.0), s1 = mkspec(L.colh2o, rat1, L.colco2, 8.0);
        Spec sp = mkspec(L.colh2o, ECB_RAT(RAT_H2OCO2, 11), L.colco2, 8.0);
        MAJ9(s, s1); SELF_(); FOR_();
        pf = pf_interp(FA, ng, sp);
      } else {
        double rat = ECB_RAT(RAT_O3CO2, jp), rat1 = ECB_RAT(RAT_O3CO2, jp + 1);
        Spec s = mkspec(L.colo3, rat, L.colco2, 4.0), s1 = mkspec(L.colo3, rat1, L.colco2, 4.0);
        Spec sp = mkspec(L.colo3, ECB_RAT(RAT_O3CO2, 13), L.colco2, 4.0);
        MAJ5(s, s1);
        pf = pf_interp(FB, ng, sp);
        *post = B.sec[L_POST];  // empirical stratospheric multipliers, rrtm_taumol4.F90:283-289
      }
    } break;
    case 5: {  // rrtm_taumol5.F90: H2O,CO2 / O3,CO2; minor O3, CCl4
      if (low) {
        double rat = ECB_RAT(RAT_H2OCO2, jp), rat1 = ECB_RAT(RAT_H2OCO2, jp + 1);
        Spec s = mkspec(L.colh2o, rat, L.colco2, 8.0), s1 = mkspec(L.colh2o, rat1, L.colco2, 8.0);
        Spec sm = mkspec(L.colh2o, ECB_RAT(RAT_H2OCO2, 7), L.colco2, 8.0);
        Spec sp = mkspec(L.colh2o, ECB_RAT(RAT_H2OCO2, 5), L.colco2, 8.0);
        MAJ9(s, s1); SELF_(); FOR_();
        emit_minor2(out, B.sec[L_M0], ng, 9, sm.js, sm.fs, indm, mf, L.colo3);
        out.add(L.wx1, B.sec[L_C0]);
        pf = pf_interp(FA, ng, sp);
      } else {
        double rat = ECB_RAT(RAT_O3CO2, jp), rat1 = ECB_RAT(RAT_O3CO2, jp + 1);
        Spec s = mkspec(L.colo3, rat, L.colco2, 4.0), s1 = mkspec(L.colo3, rat1, L.colco2, 4.0);
        Spec sp = mkspec(L.colo3, ECB_RAT(RAT_O3CO2, 43), L.colco2, 4.0);
        MAJ5(s, s1);
        out.add(L.wx1, B.sec[L_C0]);
        pf = pf_interp(FB, ng, sp);
      }
    } break;
    case 6: {  // rrtm_taumol6.F90: H2O / -; minor CO2, CFC11, CFC12
      if (low) {
        double adjcolco2 = adjcol(L.colco2, L.coldry, ECB_CHI(2, jp + 1), 3.0, 2.0, 0.77);
        MAJ1A(L.colh2o); SELF_(); FOR_();
        out.add(L.wx2, B.sec[L_C0]); out.add(L.wx3, B.sec[L_C1]);
        emit_lin(out, B.sec[L_M0], ng, indm, mf, adjcolco2);
      } else {
        out.add(L.wx2, B.sec[L_C0]); out.add(L.wx3, B.sec[L_C1]);
      }
      pf = pf_const(FA);
    } break;
    case 7: {  // rrtm_taumol7.F90: H2O,O3 / O3; minor CO2
      if (low) {
        double rat = ECB_RAT(RAT_H2OO3, jp), rat1 = ECB_RAT(RAT_H2OO3, jp + 1);
        Spec s = mkspec7(L.colh2o, rat, L.colo3, 8.0), s1 = mkspec7(L.colh2o, rat1, L.colo3, 8.0);
        Spec sm = mkspec7(L.colh2o, ECB_RAT(RAT_H2OO3, 3), L.colo3, 8.0);
        double adjcolco2 = adjcol(L.colco2, L.coldry, ECB_CHI(2, jp + 1), 3.0, 3.0, 0.79);
        MAJ9(s, s1); SELF_(); FOR_();
        emit_minor2(out, B.sec[L_M0], ng, 9, sm.js, sm.fs, indm, mf, adjcolco2);
        pf = pf_interp(FA, ng, sm);
      } else {
        double adjcolco2 = adjcol(L.colco2, L.coldry, ECB_CHI(2, jp + 1), 3.0, 2.0, 0.79);
        MAJ1B(L.colo3);
        emit_lin(out, B.sec[L_M1], ng, indm, mf, adjcolco2);
        pf = pf_const(FB);
        *post = B.sec[L_POST];
      }
    } break;
    case 8: {  // rrtm_taumol8.F90: H2O / O3; minor CO2, O3, N2O, CFC12, CFC22
      double adjcolco2 = adjcol(L.colco2, L.coldry, ECB_CHI(2, jp + 1), 3.0, 2.0, 0.65);
      if (low) {
        MAJ1A(L.colh2o); SELF_(); FOR_();
        emit_lin(out, B.sec[L_M0], ng, indm, mf, adjcolco2);
        emit_lin(out, B.sec[L_M1], ng, indm, mf, L.colo3);
        emit_lin(out, B.sec[L_M2], ng, indm, mf, L.coln2o);
        out.add(L.wx3, B.sec[L_C0]); out.add(L.wx4, B.sec[L_C1]);
        pf = pf_const(FA);
      } else {
        MAJ1B(L.colo3);
        emit_lin(out, B.sec[L_M3], ng, indm, mf, adjcolco2);
        emit_lin(out, B.sec[L_M4], ng, indm, mf, L.coln2o);
        out.add(L.wx3, B.sec[L_C0]); out.add(L.wx4, B.sec[L_C1]);
        pf = pf_const(FB);
      }
    } break;
    case 9: {  // rrtm_taumol9.F90: H2O,CH4 / CH4; minor N2O
      double adjcoln2o = adjcol(L.coln2o, L.coldry, ECB_CHI(4, jp + 1), 1.5, 0.5, 0.65);
      if (low) {
        double rat = ECB_RAT(RAT_H2OCH4, jp), rat1 = ECB_RAT(RAT_H2OCH4, jp + 1);
        Spec s = mkspec(L.colh2o, rat, L.colch4, 8.0), s1 = mkspec(L.colh2o, rat1, L.colch4, 8.0);
        Spec sm = mkspec(L.colh2o, ECB_RAT(RAT_H2OCH4, 3), L.colch4, 8.0);
        Spec sp = mkspec(L.colh2o, ECB_RAT(RAT_H2OCH4, 9), L.colch4, 8.0);
        MAJ9(s, s1); SELF_(); FOR_();
        emit_minor2(out, B.sec[L_M0], ng, 9, sm.js, sm.fs, indm, mf, adjcoln2o);
        pf = pf_interp(FA, ng, sp);
      } else {
        MAJ1B(L.colch4);
        emit_lin(out, B.sec[L_M1], ng, indm, mf, adjcoln2o);
        pf = pf_const(FB);
      }
    } break;
    case 10: {  // rrtm_taumol10.F90: H2O / H2O
      if (low) { MAJ1A(L.colh2o); SELF_(); FOR_(); pf = pf_const(FA); }
      else { MAJ1B(L.colh2o); FOR_(); pf = pf_const(FB); }
    } break;
    case 11: {  // rrtm_taumol11.F90: H2O / H2O; minor O2
      double scaleo2 = L.colo2 * L.scaleminor;
      if (low) { MAJ1A(L.colh2o); SELF_(); FOR_(); emit_lin(out, B.sec[L_M0], ng, indm, mf, scaleo2); pf = pf_const(FA); }
      else { MAJ1B(L.colh2o); FOR_(); emit_lin(out, B.sec[L_M1], ng, indm, mf, scaleo2); pf = pf_const(FB); }
    } break;
    case 12: {  // rrtm_taumol12.F90: H2O,CO2 / -
      if (low) {
        double rat = ECB_RAT(RAT_H2OCO2, jp), rat1 = ECB_RAT(RAT_H2OCO2, jp + 1);
        Spec s = mkspec(L.colh2o, rat, L.colco2, 8.0), s1 = mkspec(L.colh2o, rat1, L.colco2, 8.0);
        Spec sp = mkspec(L.colh2o, ECB_RAT(RAT_H2OCO2, 10), L.colco2, 8.0);
        MAJ9(s, s1); SELF_(); FOR_();
        pf = pf_interp(FA, ng, sp);
      }
    } break;
    case 13: {  // rrtm_taumol13.F90: H2O,N2O / -; minor CO2, CO (column amount 0 in the IFS), O3
      if (low) {
        double rat = ECB_RAT(RAT_H2ON2O, jp), rat1 = ECB_RAT(RAT_H2ON2O, jp + 1);
        Spec s = mkspec(L.colh2o, rat, L.coln2o, 8.0), s1 = mkspec(L.colh2o, rat1, L.coln2o, 8.0);
        Spec smco2 = mkspec(L.colh2o, ECB_RAT(RAT_H2ON2O, 1), L.coln2o, 8.0);
        Spec sp = mkspec(L.colh2o, ECB_RAT(RAT_H2ON2O, 5), L.coln2o, 8.0);
        double adjcolco2;
        {  // reference CO2 mixing ratio is the constant 3.55e-4 (second occurrence single precision in the source)
          double chi_co2 = L.colco2 / L.coldry;
          double ratco2 = 1.E20 * chi_co2 / 3.55E-4;
          if (ratco2 > 3.0) {
            double adjfac = 2.0 + pow(ratco2 - 2.0, 0.68);
            adjcolco2 = adjfac * (double)3.55E-4f * L.coldry * 1.E-20;
          } else adjcolco2 = L.colco2;
        }
        MAJ9(s, s1); SELF_(); FOR_();
        emit_minor2(out, B.sec[L_M0], ng, 9, smco2.js, smco2.fs, indm, mf, adjcolco2);
        pf = pf_interp(FA, ng, sp);
      } else {
        emit_lin(out, B.sec[L_M2], ng, indm, mf, L.colo3);
        pf = pf_const(FB);
      }
    } break;
    case 14: {  // rrtm_taumol14.F90: CO2 / CO2
      if (low) { MAJ1A(L.colco2); SELF_(); FOR_(); pf = pf_const(FA); }
      else { MAJ1B(L.colco2); pf = pf_const(FB); }
    } break;
    case 15: {  // rrtm_taumol15.F90: N2O,CO2 / -; minor N2
      if (low) {
        double rat = ECB_RAT(RAT_N2OCO2, jp), rat1 = ECB_RAT(RAT_N2OCO2, jp + 1);
        Spec s = mkspec(L.coln2o, rat, L.colco2, 8.0), s1 = mkspec(L.coln2o, rat1, L.colco2, 8.0);
        Spec sm = mkspec(L.coln2o, ECB_RAT(RAT_N2OCO2, 1), L.colco2, 8.0);
        double scalen2 = L.colbrd * L.scaleminor;
        MAJ9(s, s1); SELF_(); FOR_();
        emit_minor2(out, B.sec[L_M0], ng, 9, sm.js, sm.fs, indm, mf, scalen2);
        pf = pf_interp(FA, ng, sm);
      }
    } break;
    case 16: {  // rrtm_taumol16.F90: H2O,CH4 / CH4
      if (low) {
        double rat = ECB_RAT(RAT_H2OCH4, jp), rat1 = ECB_RAT(RAT_H2OCH4, jp + 1);
        Spec s = mkspec(L.colh2o, rat, L.colch4, 8.0), s1 = mkspec(L.colh2o, rat1, L.colch4, 8.0);
        Spec sp = mkspec(L.colh2o, ECB_RAT(RAT_H2OCH4, 6), L.colch4, 8.0);
        MAJ9(s, s1); SELF_(); FOR_();
        pf = pf_interp(FA, ng, sp);
      } else {
        // NSPB(16) = 0 (ifsrrtm/surrtpk.F90:23): the reference always reads rows 1 and 2 of ABSB here
        emit_major1(out, Bb, ng, 1, 1, L.colch4, L.fac00, L.fac10, L.fac01, L.fac11);
        pf = pf_const(FB);
      }
    } break;
  }
#undef SELF_
#undef FOR_
#undef MAJ1A
#undef MAJ1B
#undef MAJ9
#undef MAJ5
  return pf;
}

// Planck function integrated over LW band jb (0-based) at `temperature`: radiation_ifs_rrtm.F90:676-699
HD double planck_band(const GasMeta& M, double temperature, int jb) {
  const double zfluxfac = 2.0 * asin(1.0) * 1.0e4;
  int ind; double frac;
  if (temperature < 339.0 && temperature >= 160.0) {
    ind = (int)(temperature - 159.0);
    frac = temperature - (int)temperature;
  } else if (temperature >= 339.0) {
    ind = 180; frac = temperature - 339.0;
  } else {
    ind = 1; frac = 0.0;
  }
  double factor = zfluxfac * M.delwave[jb];
  const double* tp = M.totplnk + jb * 181;
  return factor * (tp[ind - 1] + frac * (tp[ind] - tp[ind - 1]));
}

// ---------------------------------------------------------------------------------------------------------
// SW
// ---------------------------------------------------------------------------------------------------------
struct SwLev {
  int jp, jt, jt1, indself, indfor, tropo;  // tropo: jp < 13 (counted to get LAYTROP)
  double fac00, fac01, fac10, fac11, forfac, forfrac, selffac, selffrac;
  double colh2o, colco2, colo3, colch4, colo2, colmol;
};
static_assert(sizeof(SwLev) % 16 == 8, "SwLev stride must be an odd number of doubles");

HD void sw_setcoef(const GasMeta& M, const LevGas& G, SwLev& L) {
  const double stpfac = 296.0 / 1013.0;
  const double e20 = (double)1.E-20f, e32 = (double)1.E-32f;
  double plog = log(G.pavel);
  int jp = (int)(36.0 - 5.0 * (plog + 0.04));
  if (jp < 1) jp = 1; else if (jp > 58) jp = 58;
  int jp1 = jp + 1;
  double fp = 5. * (M.preflog_sw[jp - 1] - plog);
  int jt = (int)(3. + (G.tavel - M.tref_sw[jp - 1]) / 15.);
  if (jt < 1) jt = 1; else if (jt > 4) jt = 4;
  double ft = ((G.tavel - M.tref_sw[jp - 1]) / 15.) - (double)(jt - 3);
  int jt1 = (int)(3. + (G.tavel - M.tref_sw[jp1 - 1]) / 15.);
  if (jt1 < 1) jt1 = 1; else if (jt1 > 4) jt1 = 4;
  double ft1 = ((G.tavel - M.tref_sw[jp1 - 1]) / 15.) - (double)(jt1 - 3);
  double water = G.wkl1 / G.coldry;
  double scalefac = G.pavel * stpfac / G.tavel;
  L.jp = jp; L.jt = jt; L.jt1 = jt1;
  L.tropo = jp < 13;
  L.forfac = scalefac / (1. + water);
  if (L.tropo) {
    double factor = (332.0 - G.tavel) / 36.0;
    L.indfor = imin(2, imax(1, (int)factor));
    L.forfrac = factor - (double)L.indfor;
    L.selffac = water * L.forfac;
    factor = (G.tavel - 188.0) / (double)7.2f;
    L.indself = imin(9, imax(1, (int)factor - 7));
    L.selffrac = factor - (double)(L.indself + 7);
  } else {
    double factor = (G.tavel - 188.0) / 36.0;
    L.indfor = 3;
    L.forfrac = factor - 1.0;
    L.selffac = 0.0; L.selffrac = 0.0; L.indself = 1;
  }
  L.colh2o = e20 * G.wkl1; L.colco2 = e20 * G.wkl2; L.colo3 = e20 * G.wkl3;
  L.colch4 = e20 * G.wkl6; L.colo2 = e20 * G.wkl7;
  L.colmol = e20 * G.coldry + L.colh2o;
  if (L.colco2 == 0.) L.colco2 = e32 * G.coldry;
  if (L.colch4 == 0.) L.colch4 = e32 * G.coldry;
  if (L.colo2 == 0.) L.colo2 = e32 * G.coldry;
  double compfp = 1. - fp;
  L.fac10 = compfp * ft; L.fac00 = compfp * (1. - ft);
  L.fac11 = fp * ft1;    L.fac01 = fp * (1. - ft1);
}

enum { SWLEV_NF = 20 };
HD void swlev_store(double* col, int stride, int l, const SwLev& L) {
  double* p = col + l;
  p[0 * stride] = L.jp; p[1 * stride] = L.jt; p[2 * stride] = L.jt1; p[3 * stride] = L.indself; p[4 * stride] = L.indfor; p[5 * stride] = L.tropo;
  p[6 * stride] = L.fac00; p[7 * stride] = L.fac01; p[8 * stride] = L.fac10; p[9 * stride] = L.fac11; p[10 * stride] = L.forfac;
  p[11 * stride] = L.forfrac; p[12 * stride] = L.selffac; p[13 * stride] = L.selffrac; p[14 * stride] = L.colh2o; p[15 * stride] = L.colco2;
  p[16 * stride] = L.colo3; p[17 * stride] = L.colch4; p[18 * stride] = L.colo2; p[19 * stride] = L.colmol;
}
HD SwLev swlev_load(const double* col, int stride, int l) {
  const double* p = col + l;
  SwLev L;
  L.jp = (int)p[0 * stride]; L.jt = (int)p[1 * stride]; L.jt1 = (int)p[2 * stride]; L.indself = (int)p[3 * stride]; L.indfor = (int)p[4 * stride];
  L.tropo = (int)p[5 * stride];
  L.fac00 = p[6 * stride]; L.fac01 = p[7 * stride]; L.fac10 = p[8 * stride]; L.fac11 = p[9 * stride]; L.forfac = p[10 * stride];
  L.forfrac = p[11 * stride]; L.selffac = p[12 * stride]; L.selffrac = p[13 * stride]; L.colh2o = p[14 * stride]; L.colco2 = p[15 * stride];
  L.colo3 = p[16 * stride]; L.colch4 = p[17 * stride]; L.colo2 = p[18 * stride]; L.colmol = p[19 * stride];
  return L;
}

// speccomb*((1-fs)*(T[i0]f00 + T[i0+d]f10 + T[i1]f01 + T[i1+d]f11) + fs*(same rows + 1))
template <class Out> HD void emit_major2(Out& out, int sec, int ng, int i0, int i1, int d, const Spec& s, const SwLev& L) {
  double a = s.speccomb * (1. - s.fs), b = s.speccomb * s.fs;
  out.add(a * L.fac00, sec + (i0 - 1) * ng);  out.add(a * L.fac10, sec + (i0 + d - 1) * ng);
  out.add(a * L.fac01, sec + (i1 - 1) * ng);  out.add(a * L.fac11, sec + (i1 + d - 1) * ng);
  out.add(b * L.fac00, sec + i0 * ng);        out.add(b * L.fac10, sec + (i0 + d) * ng);
  out.add(b * L.fac01, sec + i1 * ng);        out.add(b * L.fac11, sec + (i1 + d) * ng);
}

struct SwAux {     // Rayleigh stencil and (if this layer sets it) the solar-source stencil
  double rc0, rc1; int ro0, ro1;
  double sc0, sc1; int so0, so1;
};

// Build the stencil of SW band `ib` (0-based: 0 = band 16) for one layer.  `low` = layer index <= LAYTROP.
template <class Out>
HD void sw_build_list(const GasMeta& M, const BandMeta& B, const SwLev& L, int ib, bool low, Out& out, SwAux& aux) {
  const int jb = ib + 16, ng = B.ng;
  const int A = B.sec[S_ABSA], Bb = B.sec[S_ABSB], SR = B.sec[S_SELF], FR = B.sec[S_FOR], SFX = B.sec[S_SFLUX];
  const int ONES = B.sec[S_ONES];
  const int jp = L.jp;
  const int i0a1 = ((jp - 1) * 5 + (L.jt - 1)) + 1, i1a1 = (jp * 5 + (L.jt1 - 1)) + 1;
  const int i0b1 = ((jp - 13) * 5 + (L.jt - 1)) + 1, i1b1 = ((jp - 12) * 5 + (L.jt1 - 1)) + 1;
  const int i0a9 = ((jp - 1) * 5 + (L.jt - 1)) * 9, i1a9 = (jp * 5 + (L.jt1 - 1)) * 9;
  const int i0b5 = ((jp - 13) * 5 + (L.jt - 1)) * 5, i1b5 = ((jp - 12) * 5 + (L.jt1 - 1)) * 5;
  const double strrat = M.strrat_sw[ib], rayl = M.rayl_sw[ib];
  out.n = 0;
  // defaults: Rayleigh = colmol * rayl (constant over the band), solar source = SFLUXREFC(:,1)
  aux.rc0 = L.colmol * rayl; aux.rc1 = 0.0; aux.ro0 = ONES; aux.ro1 = ONES;
  aux.sc0 = 1.0; aux.sc1 = 0.0; aux.so0 = SFX; aux.so1 = SFX;
#define SELFFOR(scale) do { emit_lin(out, SR, ng, L.indself, L.selffrac, (scale) * L.selffac); emit_lin(out, FR, ng, L.indfor, L.forfrac, (scale) * L.forfac); } while (0)
#define FORONLY(scale) emit_lin(out, FR, ng, L.indfor, L.forfrac, (scale) * L.forfac)
#define MAJ1(sec, i0, i1, col) emit_major1(out, sec, ng, i0, i1, col, L.fac00, L.fac10, L.fac01, L.fac11)
#define SFLUX_INTERP(s) do { aux.sc0 = 1.0 - (s).fs; aux.sc1 = (s).fs; aux.so0 = SFX + ((s).js - 1) * ng; aux.so1 = SFX + (s).js * ng; } while (0)
#define RAYLC() do { aux.rc0 = L.colmol; aux.ro0 = B.sec[S_RAYA]; } while (0)
  switch (jb) {
    case 16:
      if (low) { Spec s = mkspec(L.colh2o, strrat, L.colch4, 8.0); emit_major2(out, A, ng, i0a9 + s.js, i1a9 + s.js, 9, s, L); SELFFOR(L.colh2o); }
      else { MAJ1(Bb, i0b1, i1b1, L.colch4); }
      break;
    case 17:
      if (low) { Spec s = mkspec(L.colh2o, strrat, L.colco2, 8.); emit_major2(out, A, ng, i0a9 + s.js, i1a9 + s.js, 9, s, L); SELFFOR(L.colh2o); }
      else { Spec s = mkspec(L.colh2o, strrat, L.colco2, 4.); emit_major2(out, Bb, ng, i0b5 + s.js, i1b5 + s.js, 5, s, L); FORONLY(L.colh2o); SFLUX_INTERP(s); }
      break;
    case 18: case 19: case 21: case 22: case 24: {
      double colb = (jb == 18) ? L.colch4 : (jb == 19 || jb == 21) ? L.colco2 : L.colo2;
      const double o2adj = 1.6;
      double o2cont = (double)4.35e-4f * L.colo2 / (double)(350.0f * 2.0f);
      if (low) {
        Spec s = (jb == 22) ? mkspec(L.colh2o, o2adj * strrat, colb, 8.) : mkspec(L.colh2o, strrat, colb, 8.);
        emit_major2(out, A, ng, i0a9 + s.js, i1a9 + s.js, 9, s, L);
        if (jb == 24) out.add(L.colo3, B.sec[S_X0]);
        SELFFOR(L.colh2o);
        if (jb == 22) out.add(o2cont, ONES);
        SFLUX_INTERP(s);
        if (jb == 24) {  // Rayleigh interpolated in RAYLAC(ig, js)
          aux.rc0 = L.colmol * (1.0 - s.fs); aux.rc1 = L.colmol * s.fs;
          aux.ro0 = B.sec[S_RAYA] + (s.js - 1) * ng; aux.ro1 = B.sec[S_RAYA] + s.js * ng;
        }
      } else if (jb == 21) {
        Spec s = mkspec(L.colh2o, strrat, L.colco2, 4.);
        emit_major2(out, Bb, ng, i0b5 + s.js, i1b5 + s.js, 5, s, L); FORONLY(L.colh2o);
      } else {
        if (jb == 22) { MAJ1(Bb, i0b1, i1b1, L.colo2 * o2adj); out.add(o2cont, ONES); }
        else if (jb == 24) { MAJ1(Bb, i0b1, i1b1, L.colo2); out.add(L.colo3, B.sec[S_X1]); aux.rc0 = L.colmol; aux.ro0 = B.sec[S_RAYB]; }
        else { MAJ1(Bb, i0b1, i1b1, colb); }
      }
    } break;
    case 20:
      if (low) { MAJ1(A, i0a1, i1a1, L.colh2o); SELFFOR(L.colh2o); out.add(L.colch4, B.sec[S_X0]); }
      else { MAJ1(Bb, i0b1, i1b1, L.colh2o); FORONLY(L.colh2o); out.add(L.colch4, B.sec[S_X0]); }
      break;
    case 23:
      if (low) { MAJ1(A, i0a1, i1a1, L.colh2o * M.givfac_23); SELFFOR(L.colh2o); }
      RAYLC();
      break;
    case 25:
      if (low) { MAJ1(A, i0a1, i1a1, L.colh2o); out.add(L.colo3, B.sec[S_X0]); }
      else { out.add(L.colo3, B.sec[S_X1]); }
      RAYLC();
      break;
    case 26:
      RAYLC();
      break;
    case 27:
      if (low) { MAJ1(A, i0a1, i1a1, L.colo3); }
      else { MAJ1(Bb, i0b1, i1b1, L.colo3); aux.sc0 = M.scalekur_27; }
      RAYLC();
      break;
    case 28:
      if (low) { Spec s = mkspec(L.colo3, strrat, L.colo2, 8.); emit_major2(out, A, ng, i0a9 + s.js, i1a9 + s.js, 9, s, L); }
      else { Spec s = mkspec(L.colo3, strrat, L.colo2, 4.); emit_major2(out, Bb, ng, i0b5 + s.js, i1b5 + s.js, 5, s, L); SFLUX_INTERP(s); }
      break;
    case 29:
      if (low) { MAJ1(A, i0a1, i1a1, L.colh2o); SELFFOR(L.colh2o); out.add(L.colco2, B.sec[S_X0]); }
      else { MAJ1(Bb, i0b1, i1b1, L.colco2); out.add(L.colh2o, B.sec[S_X1]); }
      break;
  }
#undef SELFFOR
#undef FORONLY
#undef MAJ1
#undef SFLUX_INTERP
#undef RAYLC
}

// Which RRTMG layer (1-based, 1 = bottom) supplies the solar source function of SW band ib, following the
// LAYSOLFR logic of srtm_taumol16..29.F90; 0 if none.  jp_of(il) returns JP of RRTMG layer il (1-based).
template <class JpOf>
HD int sw_solar_layer(const GasMeta& M, int ib, int nlev, int laytrop, JpOf jp_of) {
  const int jb = ib + 16;
  const int layreffr = M.layreffr_sw[ib];
  const bool upper_src = (jb == 16 || jb == 17 || jb == 27 || jb == 28 || jb == 29);
  int laysolfr = upper_src ? nlev : laytrop;
  int last = 0;
  for (int il = 1; il <= nlev; ++il) {
    const bool low = il <= laytrop;
    if (low && !upper_src && jb != 26) {
      int inext = imin(nlev, il + 1);
      if (jp_of(il) < layreffr && jp_of(inext) >= layreffr) laysolfr = imin(il + 1, laytrop);
    }
    if (!low && upper_src) {
      if (il >= 2 && jp_of(il - 1) < layreffr && jp_of(il) >= layreffr) laysolfr = il;
    }
    if (il == laysolfr && (low ? !upper_src : upper_src)) last = il;
  }
  return last;
}

}  // namespace ecb
