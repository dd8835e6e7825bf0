/* rrtmg_lw.c -- oracle restatement of the RRTMG longwave gas optics as wrapped by the IFS/ecRad.
 * TEST INFRASTRUCTURE (see oracle.h).  One column at a time; layers in RRTMG order (1 = bottom).
 *
 * Follows: ifsrrtm/rrtm_prepare_gases.F90, ifsrrtm/rrtm_setcoef_140gp.F90, ifsrrtm/rrtm_taumol1..16.F90
 * (called from ifsrrtm/rrtm_gas_optical_depth.F90:97-183).  Literals written without a kind suffix in the
 * Fortran are single precision there and are written here as (double)<x>f on purpose.
 */
#include <math.h>
#include <string.h>
#include "oracle.h"

#define CHI(i, j) (t->chi_mls[((j) - 1) * 7 + ((i) - 1)])
static inline double dmin(double a, double b) { return a < b ? a : b; }
static inline double dmax(double a, double b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* ifsrrtm/rrtm_prepare_gases.F90:150-229.  Inputs are ecRad-ordered (index 0 = top layer). */
void orc_prepare_gases(int nlev, const double* p_hl, const double* t_hl, const double* p_fl, const double* t_fl,
                       const double* q, const double* co2, const double* ch4, const double* n2o,
                       const double* cfc11, const double* cfc12, const double* hcfc22, const double* ccl4,
                       const double* o3, orc_lay_lw* lay) {
  const double ZAMD = 28.970, ZAMW = 18.0154, ZAMCO2 = 44.011, ZAMO = 47.9982, ZAMCH4 = 16.043, ZAMN2O = 44.013,
               ZAMC11 = 137.3686, ZAMC12 = 120.9140, ZAMC22 = 86.4690, ZAMCL4 = 153.8230, ZAVGDRO = 6.02214E23;
  const double RG = 9.80665, RPLRG = 1.0;              /* ifsaux/yomcst_ecrad.F90:33, yomdyncore.F90:23 */
  const double ZGRAVIT = (RG / RPLRG) * 1.E2;
  (void)t_hl;
  double pz_prev = p_hl[nlev] / 100.0;                 /* PZ(JL,0) = PAPH(JL,KLEV+1)/100 */
  for (int jk = 1; jk <= nlev; ++jk) {
    orc_lay_lw* L = &lay[jk - 1];
    int ie = nlev - jk;                                /* ecRad 0-based layer index = KLEV-JK+1 (1-based) */
    memset(L, 0, sizeof(*L));
    L->pavel = p_fl[ie] / 100.0;
    L->tavel = t_fl[ie];
    double pz = p_hl[ie] / 100.0;                      /* PZ(JL,JK) = PAPH(JL,KLEV-JK+1)/100 */
    L->wkl[1] = dmax(q[ie], (double)1.0E-15f) * ZAMD / ZAMW;
    L->wkl[2] = co2[ie] * ZAMD / ZAMCO2;
    L->wkl[3] = o3[ie] * ZAMD / ZAMO;
    L->wkl[4] = n2o[ie] * ZAMD / ZAMN2O;
    L->wkl[6] = ch4[ie] * ZAMD / ZAMCH4;
    L->wkl[7] = 0.209488;
    double zamm = (1.0 - L->wkl[1]) * ZAMD + L->wkl[1] * ZAMW;
    L->coldry = (pz_prev - pz) * 1.E3 * ZAVGDRO / (ZGRAVIT * zamm * (1.0 + L->wkl[1]));
    pz_prev = pz;
    L->wx[1] = ccl4[ie] * ZAMD / ZAMCL4;
    L->wx[2] = cfc11[ie] * ZAMD / ZAMC11;
    L->wx[3] = cfc12[ie] * ZAMD / ZAMC12;
    L->wx[4] = hcfc22[ie] * ZAMD / ZAMC22;
    for (int i = 1; i <= 4; ++i) L->wx[i] = L->coldry * L->wx[i] * 1.E-20;
    double summol = 0.0;
    for (int m = 2; m <= 7; ++m) summol = summol + L->wkl[m];
    L->wbroad = L->coldry * (1.0 - summol);
    for (int m = 1; m <= 7; ++m) L->wkl[m] = L->coldry * L->wkl[m];
  }
}

/* ifsrrtm/rrtm_setcoef_140gp.F90:84-276 */
void orc_setcoef_lw(const orc_tables* t, int nlev, orc_lay_lw* lay, int* laytrop_out) {
  const double stpfac = 296.0 / 1013.0;
  int laytrop = 0;
  for (int jl = 0; jl < nlev; ++jl) {
    orc_lay_lw* L = &lay[jl];
    double plog = log(L->pavel);
    int jp = (int)(36.0 - 5 * (plog + 0.04));
    if (jp < 1) jp = 1; else if (jp > 58) jp = 58;
    int jp1 = jp + 1;
    double fp = 5.0 * (t->preflog_lw[jp - 1] - plog);
    fp = dmax(-1.0, dmin(1.0, fp));
    int jt = (int)(3.0 + (L->tavel - t->tref_lw[jp - 1]) / 15.0);
    if (jt < 1) jt = 1; else if (jt > 4) jt = 4;
    double ft = ((L->tavel - t->tref_lw[jp - 1]) / 15.0) - (double)(jt - 3);
    int jt1 = (int)(3.0 + (L->tavel - t->tref_lw[jp1 - 1]) / 15.0);
    if (jt1 < 1) jt1 = 1; else if (jt1 > 4) jt1 = 4;
    double ft1 = ((L->tavel - t->tref_lw[jp1 - 1]) / 15.0) - (double)(jt1 - 3);
    double water = L->wkl[1] / L->coldry;
    double scalefac = L->pavel * stpfac / L->tavel;
    L->jp = jp; L->jt = jt; L->jt1 = jt1;
    double factor;
    if (plog > 4.56) {
      laytrop++;
      L->forfac = scalefac / (1.0 + water);
      factor = (332.0 - L->tavel) / 36.0;
      L->indfor = imin(2, imax(1, (int)factor));
      L->forfrac = factor - (double)L->indfor;
      L->selffac = water * L->forfac;
      factor = (L->tavel - 188.0) / 7.2;
      L->indself = imin(9, imax(1, (int)factor - 7));
      L->selffrac = factor - (double)(L->indself + 7);
      L->scaleminor = L->pavel / L->tavel;
      L->scaleminorn2 = (L->pavel / L->tavel) * (L->wbroad / (L->coldry + L->wkl[1]));
      factor = (L->tavel - 180.8) / 7.2;
      L->indminor = imin(18, imax(1, (int)factor));
      L->minorfrac = factor - (double)L->indminor;
      L->rat_h2oco2 = CHI(1, jp) / CHI(2, jp);     L->rat_h2oco2_1 = CHI(1, jp + 1) / CHI(2, jp + 1);
      L->rat_h2oo3 = CHI(1, jp) / CHI(3, jp);      L->rat_h2oo3_1 = CHI(1, jp + 1) / CHI(3, jp + 1);
      L->rat_h2on2o = CHI(1, jp) / CHI(4, jp);     L->rat_h2on2o_1 = CHI(1, jp + 1) / CHI(4, jp + 1);
      L->rat_h2och4 = CHI(1, jp) / CHI(6, jp);     L->rat_h2och4_1 = CHI(1, jp + 1) / CHI(6, jp + 1);
      L->rat_n2oco2 = CHI(4, jp) / CHI(2, jp);     L->rat_n2oco2_1 = CHI(4, jp + 1) / CHI(2, jp + 1);
    } else {
      L->forfac = scalefac / (1.0 + water);
      factor = (L->tavel - 188.0) / 36.0;
      L->indfor = 3;
      L->forfrac = factor - 1.0;
      L->selffac = water * L->forfac;
      L->scaleminor = L->pavel / L->tavel;
      L->scaleminorn2 = (L->pavel / L->tavel) * (L->wbroad / (L->coldry + L->wkl[1]));
      factor = (L->tavel - 180.8) / 7.2;
      L->indminor = imin(18, imax(1, (int)factor));
      L->minorfrac = factor - (double)L->indminor;
      L->rat_h2oco2 = CHI(1, jp) / CHI(2, jp);     L->rat_h2oco2_1 = CHI(1, jp + 1) / CHI(2, jp + 1);
      L->rat_o3co2 = CHI(3, jp) / CHI(2, jp);      L->rat_o3co2_1 = CHI(3, jp + 1) / CHI(2, jp + 1);
    }
    L->colh2o = 1.E-20 * L->wkl[1];
    L->colco2 = 1.E-20 * L->wkl[2];
    L->colo3 = 1.E-20 * L->wkl[3];
    L->coln2o = 1.E-20 * L->wkl[4];
    L->colch4 = 1.E-20 * L->wkl[6];
    L->colo2 = 1.E-20 * L->wkl[7];
    L->colbrd = 1.E-20 * L->wbroad;
    if (L->colco2 == 0.0) L->colco2 = 1.E-32 * L->coldry;
    if (L->coln2o == 0.0) L->coln2o = 1.E-32 * L->coldry;
    if (L->colch4 == 0.0) L->colch4 = 1.E-32 * L->coldry;
    double co2reg = 3.55E-24 * L->coldry;
    L->co2mult = (L->colco2 - co2reg) * 272.63 * exp(-1919.4 / L->tavel) / (8.7604E-4 * L->tavel);
    double compfp = 1.0 - fp;
    L->fac10 = compfp * ft;
    L->fac00 = compfp * (1.0 - ft);
    L->fac11 = fp * ft1;
    L->fac01 = fp * (1.0 - ft1);
    L->selffac = L->colh2o * L->selffac;
    L->forfac = L->colh2o * L->forfac;
  }
  *laytrop_out = laytrop;
}

/* ---------------------------------------------------------------------------------------------------
 * helpers shared by the taumol bands
 * ------------------------------------------------------------------------------------------------- */
typedef struct { double speccomb, specparm, fs; int js; } spec_t;

/* "SPECCOMB = COLA + RAT*COLB; SPECPARM = COLA/SPECCOMB; SPECPARM=MIN(ONEMINUS,.); SPECMULT = n*SPECPARM;
 *  JS = 1 + INT(SPECMULT); FS = MOD(SPECMULT,1.0)"   e.g. rrtm_taumol3.F90:128-133 */
static spec_t mkspec(double cola, double rat, double colb, double mult, double oneminus) {
  spec_t s;
  s.speccomb = cola + rat * colb;
  s.specparm = cola / s.speccomb;
  if (s.specparm >= oneminus) s.specparm = oneminus;
  double specmult = mult * s.specparm;
  s.js = 1 + (int)specmult;
  s.fs = fmod(specmult, 1.0);
  return s;
}
/* band 7 writes specparm as 1/(1+rat/colh2o*colo3)   rrtm_taumol7.F90:126-150 */
static spec_t mkspec7(double cola, double rat, double colb, double mult, double oneminus) {
  spec_t s;
  s.speccomb = cola + rat * colb;
  s.specparm = 1.0 / (1.0 + rat / cola * colb);
  if (s.specparm >= oneminus) s.specparm = oneminus;
  double specmult = mult * s.specparm;
  s.js = 1 + (int)specmult;
  s.fs = fmod(specmult, 1.0);
  return s;
}

typedef struct { int n; double c[6]; int off[6]; } majfac;

/* lower-atmosphere (NSPA=9) interpolation weights, e.g. rrtm_taumol3.F90:170-229 */
static majfac lowfac(const spec_t* s, double fa, double fb) {
  majfac m;
  if (s->specparm < 0.125) {
    double p = s->fs - 1; double p4 = (p * p) * (p * p);
    double fk0 = p4, fk1 = 1 - p - 2.0 * p4, fk2 = p + p4;
    m.n = 6;
    m.c[0] = fk0 * fa; m.off[0] = 0;  m.c[1] = fk1 * fa; m.off[1] = 1;  m.c[2] = fk2 * fa; m.off[2] = 2;
    m.c[3] = fk0 * fb; m.off[3] = 9;  m.c[4] = fk1 * fb; m.off[4] = 10; m.c[5] = fk2 * fb; m.off[5] = 11;
  } else if (s->specparm > 0.875) {
    double p = -s->fs; double p4 = (p * p) * (p * p);
    double fk0 = p4, fk1 = 1 - p - 2.0 * p4, fk2 = p + p4;
    m.n = 6;
    m.c[0] = fk2 * fa; m.off[0] = -1; m.c[1] = fk1 * fa; m.off[1] = 0;  m.c[2] = fk0 * fa; m.off[2] = 1;
    m.c[3] = fk2 * fb; m.off[3] = 8;  m.c[4] = fk1 * fb; m.off[4] = 9;  m.c[5] = fk0 * fb; m.off[5] = 10;
  } else {
    m.n = 4;
    m.c[0] = (1.0 - s->fs) * fa; m.off[0] = 0;  m.c[1] = s->fs * fa; m.off[1] = 1;
    m.c[2] = (1.0 - s->fs) * fb; m.off[2] = 9;  m.c[3] = s->fs * fb; m.off[3] = 10;
  }
  return m;
}
/* upper-atmosphere (NSPB=5) weights, e.g. rrtm_taumol3.F90:301-308 */
static majfac uppfac(const spec_t* s, double fa, double fb) {
  majfac m; m.n = 4;
  m.c[0] = (1.0 - s->fs) * fa; m.off[0] = 0;  m.c[1] = s->fs * fa; m.off[1] = 1;
  m.c[2] = (1.0 - s->fs) * fb; m.off[2] = 5;  m.c[3] = s->fs * fb; m.off[3] = 6;
  return m;
}
/* SPECCOMB * (c0*ABS(ind+off0) + c1*ABS(ind+off1) + ...), summed left to right as in the source */
static inline double majsum(const double* col, int ind, const majfac* m, double speccomb) {
  double s = m->c[0] * col[ind + m->off[0] - 1];
  for (int i = 1; i < m->n; ++i) s = s + m->c[i] * col[ind + m->off[i] - 1];
  return speccomb * s;
}
/* COL * (FAC00*ABS(IND0) + FAC10*ABS(IND0+1) + FAC01*ABS(IND1) + FAC11*ABS(IND1+1)) */
static inline double major1(const double* col, int ind0, int ind1, const orc_lay_lw* L) {
  return L->fac00 * col[ind0 - 1] + L->fac10 * col[ind0] + L->fac01 * col[ind1 - 1] + L->fac11 * col[ind1];
}
/* TAB(i,ig) + f*(TAB(i+1,ig)-TAB(i,ig)) for a (n1, ng) table; `col` points at TAB(1,ig) */
static inline double lin(const double* col, int i, double f) { return col[i - 1] + f * (col[i] - col[i - 1]); }
/* minor species with binary-parameter dependence: K(j, indm, ig), shape (nj, 19, ng)  e.g. rrtm_taumol3.F90:235-239 */
static inline double minor2(const double* k, int nj, int ig, int j, double fj, int indm, double minorfrac) {
  const double* a = k + ((size_t)ig * 19 + (indm - 1)) * nj;
  const double* b = a + nj;
  double m1 = a[j - 1] + fj * (a[j] - a[j - 1]);
  double m2 = b[j - 1] + fj * (b[j] - b[j - 1]);
  return m1 + minorfrac * (m2 - m1);
}
/* adjusted column of a gas whose abundance departs from the reference: e.g. rrtm_taumol3.F90:150-156 */
static inline double adjcol(double col, double coldry, double chiref, double thresh, double base, double expo) {
  double chi = col / coldry;
  double rat = 1.E20 * chi / chiref;
  if (rat > thresh) {
    double adjfac = base + pow(rat - base, expo);
    return adjfac * chiref * coldry * 1.E-20;
  }
  return col;
}

/* FRACREF(IG,JPL) + FPL*(FRACREF(IG,JPL+1)-FRACREF(IG,JPL)) for a (ng, n) table */
#define PLANCK_INTERP(fr, sp, ng, ig) ((fr)[((sp).js - 1) * (ng) + (ig)] + (sp).fs * ((fr)[(sp).js * (ng) + (ig)] - (fr)[((sp).js - 1) * (ng) + (ig)]))
#define IND0A(nspa) (((L->jp - 1) * 5 + (L->jt - 1)) * (nspa))
#define IND1A(nspa) ((L->jp * 5 + (L->jt1 - 1)) * (nspa))
#define IND0B(nspb) (((L->jp - 13) * 5 + (L->jt - 1)) * (nspb))
#define IND1B(nspb) (((L->jp - 12) * 5 + (L->jt1 - 1)) * (nspb))

/* ifsrrtm/rrtm_taumol1..16.F90.  tau/pfrac: [lay][140], RRTMG layer order. */
void orc_taumol_lw(const orc_tables* t, int nlev, const orc_lay_lw* lay, int laytrop, double* tau, double* pfrac) {
  const double oneminus = 1.0 - 1.0e-6;   /* radiation_ifs_rrtm.F90:384 */
  static const int ngs[17] = {0, 10, 22, 38, 52, 68, 76, 88, 96, 108, 114, 122, 130, 134, 136, 138, 140};
  for (int jl = 1; jl <= nlev; ++jl) {
    const orc_lay_lw* L = &lay[jl - 1];
    double* T = tau + (size_t)(jl - 1) * NG_LW;
    double* P = pfrac + (size_t)(jl - 1) * NG_LW;
    const int low = (jl <= laytrop);
    const int inds = L->indself, indf = L->indfor, indm = L->indminor;
#define SELF(b, ig) (L->selffac * lin(t->selfref_lw[b] + (ig) * 10, inds, L->selffrac))
#define FOR(b, ig) (L->forfac * lin(t->forref_lw[b] + (ig) * 4, indf, L->forfrac))
    /* ---- band 1: rrtm_taumol1.F90 ---- */
    {
      const int ng = 10, o = 0; const double *A = t->absa_lw[1], *B = t->absb_lw[1];
      double pp = L->pavel;
      double scalen2 = L->colbrd * L->scaleminorn2;
      if (low) {
        int ind0 = IND0A(1) + 1, ind1 = IND1A(1) + 1;
        double corradj = 1.;
        if (pp < 250.0) corradj = 1.0 - 0.15 * (250.0 - pp) / 154.4;
        for (int ig = 0; ig < ng; ++ig) {
          double tauself = SELF(1, ig), taufor = FOR(1, ig);
          double taun2 = scalen2 * lin(t->ka_mn2_1 + ig * 19, indm, L->minorfrac);
          T[o + ig] = corradj * (L->colh2o * major1(A + ig * 65, ind0, ind1, L) + tauself + taufor + taun2);
          P[o + ig] = t->fracrefa_lw[1][ig];
        }
      } else {
        int ind0 = IND0B(1) + 1, ind1 = IND1B(1) + 1;
        double corradj = 1.0 - 0.15 * (pp / 95.6);
        for (int ig = 0; ig < ng; ++ig) {
          double taufor = FOR(1, ig);
          double taun2 = scalen2 * lin(t->kb_mn2_1 + ig * 19, indm, L->minorfrac);
          T[o + ig] = corradj * (L->colh2o * major1(B + ig * 235, ind0, ind1, L) + taufor + taun2);
          P[o + ig] = t->fracrefb_lw[1][ig];
        }
      }
    }
    /* ---- band 2: rrtm_taumol2.F90 ---- */
    {
      const int ng = 12, o = ngs[1]; const double *A = t->absa_lw[2], *B = t->absb_lw[2];
      if (low) {
        int ind0 = IND0A(1) + 1, ind1 = IND1A(1) + 1;
        double corradj = 1.0 - .05 * (L->pavel - 100.0) / 900.0;
        for (int ig = 0; ig < ng; ++ig) {
          double tauself = SELF(2, ig), taufor = FOR(2, ig);
          T[o + ig] = corradj * (L->colh2o * major1(A + ig * 65, ind0, ind1, L) + tauself + taufor);
          P[o + ig] = t->fracrefa_lw[2][ig];
        }
      } else {
        int ind0 = IND0B(1) + 1, ind1 = IND1B(1) + 1;
        for (int ig = 0; ig < ng; ++ig) {
          double taufor = FOR(2, ig);
          T[o + ig] = L->colh2o * major1(B + ig * 235, ind0, ind1, L) + taufor;
          P[o + ig] = t->fracrefb_lw[2][ig];
        }
      }
    }
    /* ---- band 3: rrtm_taumol3.F90 (H2O,CO2 / H2O,CO2; minor N2O) ---- */
    {
      const int ng = 16, o = ngs[2]; const double *A = t->absa_lw[3], *B = t->absb_lw[3];
      double adjcoln2o = adjcol(L->coln2o, L->coldry, CHI(4, L->jp + 1), 1.5, 0.5, 0.65);
      if (low) {
        spec_t s = mkspec(L->colh2o, L->rat_h2oco2, L->colco2, 8.0, oneminus);
        spec_t s1 = mkspec(L->colh2o, L->rat_h2oco2_1, L->colco2, 8.0, oneminus);
        spec_t sm = mkspec(L->colh2o, CHI(1, 3) / CHI(2, 3), L->colco2, 8.0, oneminus);
        spec_t sp = mkspec(L->colh2o, CHI(1, 9) / CHI(2, 9), L->colco2, 8.0, oneminus);
        int ind0 = IND0A(9) + s.js, ind1 = IND1A(9) + s1.js;
        majfac m0 = lowfac(&s, L->fac00, L->fac10), m1 = lowfac(&s1, L->fac01, L->fac11);
        for (int ig = 0; ig < ng; ++ig) {
          double tauself = SELF(3, ig), taufor = FOR(3, ig);
          double absn2o = minor2(t->ka_mn2o_3, 9, ig, sm.js, sm.fs, indm, L->minorfrac);
          double maj = majsum(A + ig * 585, ind0, &m0, s.speccomb), maj1 = majsum(A + ig * 585, ind1, &m1, s1.speccomb);
          T[o + ig] = maj + maj1 + tauself + taufor + adjcoln2o * absn2o;
          P[o + ig] = PLANCK_INTERP(t->fracrefa_lw[3], sp, ng, ig);
        }
      } else {
        spec_t s = mkspec(L->colh2o, L->rat_h2oco2, L->colco2, 4.0, oneminus);
        spec_t s1 = mkspec(L->colh2o, L->rat_h2oco2_1, L->colco2, 4.0, oneminus);
        spec_t sm = mkspec(L->colh2o, CHI(1, 13) / CHI(2, 13), L->colco2, 4.0, oneminus);
        spec_t sp = mkspec(L->colh2o, CHI(1, 13) / CHI(2, 13), L->colco2, 4.0, oneminus);
        int ind0 = IND0B(5) + s.js, ind1 = IND1B(5) + s1.js;
        majfac m0 = uppfac(&s, L->fac00, L->fac10), m1 = uppfac(&s1, L->fac01, L->fac11);
        for (int ig = 0; ig < ng; ++ig) {
          double taufor = FOR(3, ig);
          double absn2o = minor2(t->kb_mn2o_3, 5, ig, sm.js, sm.fs, indm, L->minorfrac);
          T[o + ig] = majsum(B + ig * 1175, ind0, &m0, s.speccomb) + majsum(B + ig * 1175, ind1, &m1, s1.speccomb) +
                      taufor + adjcoln2o * absn2o;
          P[o + ig] = PLANCK_INTERP(t->fracrefb_lw[3], sp, ng, ig);
        }
      }
    }
    /* ---- band 4: rrtm_taumol4.F90 (H2O,CO2 / O3,CO2) ---- */
    {
      const int ng = 14, o = ngs[3]; const double *A = t->absa_lw[4], *B = t->absb_lw[4];
      if (low) {
        spec_t s = mkspec(L->colh2o, L->rat_h2oco2, L->colco2, 8.0, oneminus);
        spec_t s1 = mkspec(L->colh2o, L->rat_h2oco2_1, L->colco2, 8.0, oneminus);
        spec_t sp = mkspec(L->colh2o, CHI(1, 11) / CHI(2, 11), L->colco2, 8.0, oneminus);
        int ind0 = IND0A(9) + s.js, ind1 = IND1A(9) + s1.js;
        majfac m0 = lowfac(&s, L->fac00, L->fac10), m1 = lowfac(&s1, L->fac01, L->fac11);
        for (int ig = 0; ig < ng; ++ig) {
          double tauself = SELF(4, ig), taufor = FOR(4, ig);
          double maj = majsum(A + ig * 585, ind0, &m0, s.speccomb), maj1 = majsum(A + ig * 585, ind1, &m1, s1.speccomb);
          T[o + ig] = maj + maj1 + tauself + taufor;
          P[o + ig] = PLANCK_INTERP(t->fracrefa_lw[4], sp, ng, ig);
        }
      } else {
        spec_t s = mkspec(L->colo3, L->rat_o3co2, L->colco2, 4.0, oneminus);
        spec_t s1 = mkspec(L->colo3, L->rat_o3co2_1, L->colco2, 4.0, oneminus);
        spec_t sp = mkspec(L->colo3, CHI(3, 13) / CHI(2, 13), L->colco2, 4.0, oneminus);
        int ind0 = IND0B(5) + s.js, ind1 = IND1B(5) + s1.js;
        majfac m0 = uppfac(&s, L->fac00, L->fac10), m1 = uppfac(&s1, L->fac01, L->fac11);
        for (int ig = 0; ig < ng; ++ig) {
          T[o + ig] = majsum(B + ig * 1175, ind0, &m0, s.speccomb) + majsum(B + ig * 1175, ind1, &m1, s1.speccomb);
          P[o + ig] = PLANCK_INTERP(t->fracrefb_lw[4], sp, ng, ig);
        }
        /* empirical stratospheric adjustment, single-precision literals in the source (rrtm_taumol4.F90:283-289) */
        T[o + 7] = T[o + 7] * (double)0.92f;  T[o + 8] = T[o + 8] * (double)0.88f;  T[o + 9] = T[o + 9] * (double)1.07f;
        T[o + 10] = T[o + 10] * (double)1.1f; T[o + 11] = T[o + 11] * (double)0.99f; T[o + 12] = T[o + 12] * (double)0.88f;
        T[o + 13] = T[o + 13] * (double)0.943f;
      }
    }
    /* ---- band 5: rrtm_taumol5.F90 (H2O,CO2 / O3,CO2; minor O3, CCl4) ---- */
    {
      const int ng = 16, o = ngs[4]; const double *A = t->absa_lw[5], *B = t->absb_lw[5];
      if (low) {
        spec_t s = mkspec(L->colh2o, L->rat_h2oco2, L->colco2, 8.0, oneminus);
        spec_t s1 = mkspec(L->colh2o, L->rat_h2oco2_1, L->colco2, 8.0, oneminus);
        spec_t sm = mkspec(L->colh2o, CHI(1, 7) / CHI(2, 7), L->colco2, 8.0, oneminus);
        spec_t sp = mkspec(L->colh2o, CHI(1, 5) / CHI(2, 5), L->colco2, 8.0, oneminus);
        int ind0 = IND0A(9) + s.js, ind1 = IND1A(9) + s1.js;
        majfac m0 = lowfac(&s, L->fac00, L->fac10), m1 = lowfac(&s1, L->fac01, L->fac11);
        for (int ig = 0; ig < ng; ++ig) {
          double tauself = SELF(5, ig), taufor = FOR(5, ig);
          double abso3 = minor2(t->ka_mo3_5, 9, ig, sm.js, sm.fs, indm, L->minorfrac);
          double maj = majsum(A + ig * 585, ind0, &m0, s.speccomb), maj1 = majsum(A + ig * 585, ind1, &m1, s1.speccomb);
          T[o + ig] = maj + maj1 + tauself + taufor + abso3 * L->colo3 + L->wx[1] * t->ccl4_5[ig];
          P[o + ig] = PLANCK_INTERP(t->fracrefa_lw[5], sp, ng, ig);
        }
      } else {
        spec_t s = mkspec(L->colo3, L->rat_o3co2, L->colco2, 4.0, oneminus);
        spec_t s1 = mkspec(L->colo3, L->rat_o3co2_1, L->colco2, 4.0, oneminus);
        spec_t sp = mkspec(L->colo3, CHI(3, 43) / CHI(2, 43), L->colco2, 4.0, oneminus);
        int ind0 = IND0B(5) + s.js, ind1 = IND1B(5) + s1.js;
        majfac m0 = uppfac(&s, L->fac00, L->fac10), m1 = uppfac(&s1, L->fac01, L->fac11);
        for (int ig = 0; ig < ng; ++ig) {
          T[o + ig] = majsum(B + ig * 1175, ind0, &m0, s.speccomb) + majsum(B + ig * 1175, ind1, &m1, s1.speccomb) +
                      L->wx[1] * t->ccl4_5[ig];
          P[o + ig] = PLANCK_INTERP(t->fracrefb_lw[5], sp, ng, ig);
        }
      }
    }
    /* ---- band 6: rrtm_taumol6.F90 (H2O / -; minor CO2, CFC11, CFC12) ---- */
    {
      const int ng = 8, o = ngs[5]; const double* A = t->absa_lw[6];
      if (low) {
        double adjcolco2 = adjcol(L->colco2, L->coldry, CHI(2, L->jp + 1), 3.0, 2.0, 0.77);
        int ind0 = IND0A(1) + 1, ind1 = IND1A(1) + 1;
        for (int ig = 0; ig < ng; ++ig) {
          double tauself = SELF(6, ig), taufor = FOR(6, ig);
          double absco2 = lin(t->ka_mco2_6 + ig * 19, indm, L->minorfrac);
          T[o + ig] = L->colh2o * major1(A + ig * 65, ind0, ind1, L) + tauself + taufor + L->wx[2] * t->cfc11adj_6[ig] +
                      L->wx[3] * t->cfc12_6[ig] + adjcolco2 * absco2;
          P[o + ig] = t->fracrefa_lw[6][ig];
        }
      } else {
        for (int ig = 0; ig < ng; ++ig) {
          T[o + ig] = 0.0 + L->wx[2] * t->cfc11adj_6[ig] + L->wx[3] * t->cfc12_6[ig];
          P[o + ig] = t->fracrefa_lw[6][ig];
        }
      }
    }
    /* ---- band 7: rrtm_taumol7.F90 (H2O,O3 / O3; minor CO2) ---- */
    {
      const int ng = 12, o = ngs[6]; const double *A = t->absa_lw[7], *B = t->absb_lw[7];
      if (low) {
        spec_t s = mkspec7(L->colh2o, L->rat_h2oo3, L->colo3, 8.0, oneminus);
        spec_t s1 = mkspec7(L->colh2o, L->rat_h2oo3_1, L->colo3, 8.0, oneminus);
        spec_t sm = mkspec7(L->colh2o, CHI(1, 3) / CHI(3, 3), L->colo3, 8.0, oneminus);
        spec_t sp = mkspec7(L->colh2o, CHI(1, 3) / CHI(3, 3), L->colo3, 8.0, oneminus);
        double adjcolco2 = adjcol(L->colco2, L->coldry, CHI(2, L->jp + 1), 3.0, 3.0, 0.79);
        int ind0 = IND0A(9) + s.js, ind1 = IND1A(9) + s1.js;
        majfac m0 = lowfac(&s, L->fac00, L->fac10), m1 = lowfac(&s1, L->fac01, L->fac11);
        for (int ig = 0; ig < ng; ++ig) {
          double tauself = SELF(7, ig), taufor = FOR(7, ig);
          double absco2 = minor2(t->ka_mco2_7, 9, ig, sm.js, sm.fs, indm, L->minorfrac);
          double maj = majsum(A + ig * 585, ind0, &m0, s.speccomb), maj1 = majsum(A + ig * 585, ind1, &m1, s1.speccomb);
          T[o + ig] = maj + maj1 + tauself + taufor + adjcolco2 * absco2;
          P[o + ig] = PLANCK_INTERP(t->fracrefa_lw[7], sp, ng, ig);
        }
      } else {
        double adjcolco2 = adjcol(L->colco2, L->coldry, CHI(2, L->jp + 1), 3.0, 2.0, 0.79);
        int ind0 = IND0B(1) + 1, ind1 = IND1B(1) + 1;
        for (int ig = 0; ig < ng; ++ig) {
          double absco2 = lin(t->kb_mco2_7 + ig * 19, indm, L->minorfrac);
          T[o + ig] = L->colo3 * major1(B + ig * 235, ind0, ind1, L) + adjcolco2 * absco2;
          P[o + ig] = t->fracrefb_lw[7][ig];
        }
        /* rrtm_taumol7.F90: empirical adjustment (double-precision literals here) */
        T[o + 5] = T[o + 5] * 0.92;  T[o + 6] = T[o + 6] * 0.88;  T[o + 7] = T[o + 7] * 1.07;
        T[o + 8] = T[o + 8] * 1.1;   T[o + 9] = T[o + 9] * 0.99;  T[o + 10] = T[o + 10] * 0.855;
      }
    }
    /* ---- band 8: rrtm_taumol8.F90 (H2O / O3; minor CO2, O3, N2O, CFC12, CFC22) ---- */
    {
      const int ng = 8, o = ngs[7]; const double *A = t->absa_lw[8], *B = t->absb_lw[8];
      double adjcolco2 = adjcol(L->colco2, L->coldry, CHI(2, L->jp + 1), 3.0, 2.0, 0.65);
      if (low) {
        int ind0 = IND0A(1) + 1, ind1 = IND1A(1) + 1;
        for (int ig = 0; ig < ng; ++ig) {
          double tauself = SELF(8, ig), taufor = FOR(8, ig);
          double absco2 = lin(t->ka_mco2_8 + ig * 19, indm, L->minorfrac);
          double abso3 = lin(t->ka_mo3_8 + ig * 19, indm, L->minorfrac);
          double absn2o = lin(t->ka_mn2o_8 + ig * 19, indm, L->minorfrac);
          T[o + ig] = L->colh2o * major1(A + ig * 65, ind0, ind1, L) + tauself + taufor + adjcolco2 * absco2 +
                      L->colo3 * abso3 + L->coln2o * absn2o + L->wx[3] * t->cfc12_8[ig] + L->wx[4] * t->cfc22adj_8[ig];
          P[o + ig] = t->fracrefa_lw[8][ig];
        }
      } else {
        int ind0 = IND0B(1) + 1, ind1 = IND1B(1) + 1;
        for (int ig = 0; ig < ng; ++ig) {
          double absco2 = lin(t->kb_mco2_8 + ig * 19, indm, L->minorfrac);
          double absn2o = lin(t->kb_mn2o_8 + ig * 19, indm, L->minorfrac);
          T[o + ig] = L->colo3 * major1(B + ig * 235, ind0, ind1, L) + adjcolco2 * absco2 + L->coln2o * absn2o +
                      L->wx[3] * t->cfc12_8[ig] + L->wx[4] * t->cfc22adj_8[ig];
          P[o + ig] = t->fracrefb_lw[8][ig];
        }
      }
    }
    /* ---- band 9: rrtm_taumol9.F90 (H2O,CH4 / CH4; minor N2O) ---- */
    {
      const int ng = 12, o = ngs[8]; const double *A = t->absa_lw[9], *B = t->absb_lw[9];
      double adjcoln2o = adjcol(L->coln2o, L->coldry, CHI(4, L->jp + 1), 1.5, 0.5, 0.65);
      if (low) {
        spec_t s = mkspec(L->colh2o, L->rat_h2och4, L->colch4, 8.0, oneminus);
        spec_t s1 = mkspec(L->colh2o, L->rat_h2och4_1, L->colch4, 8.0, oneminus);
        spec_t sm = mkspec(L->colh2o, CHI(1, 3) / CHI(6, 3), L->colch4, 8.0, oneminus);
        spec_t sp = mkspec(L->colh2o, CHI(1, 9) / CHI(6, 9), L->colch4, 8.0, oneminus);
        int ind0 = IND0A(9) + s.js, ind1 = IND1A(9) + s1.js;
        majfac m0 = lowfac(&s, L->fac00, L->fac10), m1 = lowfac(&s1, L->fac01, L->fac11);
        for (int ig = 0; ig < ng; ++ig) {
          double tauself = SELF(9, ig), taufor = FOR(9, ig);
          double absn2o = minor2(t->ka_mn2o_9, 9, ig, sm.js, sm.fs, indm, L->minorfrac);
          double maj = majsum(A + ig * 585, ind0, &m0, s.speccomb), maj1 = majsum(A + ig * 585, ind1, &m1, s1.speccomb);
          T[o + ig] = maj + maj1 + tauself + taufor + adjcoln2o * absn2o;
          P[o + ig] = PLANCK_INTERP(t->fracrefa_lw[9], sp, ng, ig);
        }
      } else {
        int ind0 = IND0B(1) + 1, ind1 = IND1B(1) + 1;
        for (int ig = 0; ig < ng; ++ig) {
          double absn2o = lin(t->kb_mn2o_9 + ig * 19, indm, L->minorfrac);
          T[o + ig] = L->colch4 * major1(B + ig * 235, ind0, ind1, L) + adjcoln2o * absn2o;
          P[o + ig] = t->fracrefb_lw[9][ig];
        }
      }
    }
    /* ---- band 10: rrtm_taumol10.F90 (H2O / H2O) ---- */
    {
      const int ng = 6, o = ngs[9]; const double *A = t->absa_lw[10], *B = t->absb_lw[10];
      if (low) {
        int ind0 = IND0A(1) + 1, ind1 = IND1A(1) + 1;
        for (int ig = 0; ig < ng; ++ig) {
          double tauself = SELF(10, ig), taufor = FOR(10, ig);
          T[o + ig] = L->colh2o * major1(A + ig * 65, ind0, ind1, L) + tauself + taufor;
          P[o + ig] = t->fracrefa_lw[10][ig];
        }
      } else {
        int ind0 = IND0B(1) + 1, ind1 = IND1B(1) + 1;
        for (int ig = 0; ig < ng; ++ig) {
          double taufor = FOR(10, ig);
          T[o + ig] = L->colh2o * major1(B + ig * 235, ind0, ind1, L) + taufor;
          P[o + ig] = t->fracrefb_lw[10][ig];
        }
      }
    }
    /* ---- band 11: rrtm_taumol11.F90 (H2O / H2O; minor O2) ---- */
    {
      const int ng = 8, o = ngs[10]; const double *A = t->absa_lw[11], *B = t->absb_lw[11];
      double scaleo2 = L->colo2 * L->scaleminor;
      if (low) {
        int ind0 = IND0A(1) + 1, ind1 = IND1A(1) + 1;
        for (int ig = 0; ig < ng; ++ig) {
          double tauself = SELF(11, ig), taufor = FOR(11, ig);
          double tauo2 = scaleo2 * lin(t->ka_mo2_11 + ig * 19, indm, L->minorfrac);
          T[o + ig] = L->colh2o * major1(A + ig * 65, ind0, ind1, L) + tauself + taufor + tauo2;
          P[o + ig] = t->fracrefa_lw[11][ig];
        }
      } else {
        int ind0 = IND0B(1) + 1, ind1 = IND1B(1) + 1;
        for (int ig = 0; ig < ng; ++ig) {
          double taufor = FOR(11, ig);
          double tauo2 = scaleo2 * lin(t->kb_mo2_11 + ig * 19, indm, L->minorfrac);
          T[o + ig] = L->colh2o * major1(B + ig * 235, ind0, ind1, L) + taufor + tauo2;
          P[o + ig] = t->fracrefb_lw[11][ig];
        }
      }
    }
    /* ---- band 12: rrtm_taumol12.F90 (H2O,CO2 / -) ---- */
    {
      const int ng = 8, o = ngs[11]; const double* A = t->absa_lw[12];
      if (low) {
        spec_t s = mkspec(L->colh2o, L->rat_h2oco2, L->colco2, 8.0, oneminus);
        spec_t s1 = mkspec(L->colh2o, L->rat_h2oco2_1, L->colco2, 8.0, oneminus);
        spec_t sp = mkspec(L->colh2o, CHI(1, 10) / CHI(2, 10), L->colco2, 8.0, oneminus);
        int ind0 = IND0A(9) + s.js, ind1 = IND1A(9) + s1.js;
        majfac m0 = lowfac(&s, L->fac00, L->fac10), m1 = lowfac(&s1, L->fac01, L->fac11);
        for (int ig = 0; ig < ng; ++ig) {
          double tauself = SELF(12, ig), taufor = FOR(12, ig);
          double maj = majsum(A + ig * 585, ind0, &m0, s.speccomb), maj1 = majsum(A + ig * 585, ind1, &m1, s1.speccomb);
          T[o + ig] = maj + maj1 + tauself + taufor;
          P[o + ig] = PLANCK_INTERP(t->fracrefa_lw[12], sp, ng, ig);
        }
      } else {
        for (int ig = 0; ig < ng; ++ig) { T[o + ig] = 0.0; P[o + ig] = 0.0; }
      }
    }
    /* ---- band 13: rrtm_taumol13.F90 (H2O,N2O / -; minor CO2, CO(=0), O3) ---- */
    {
      const int ng = 4, o = ngs[12]; const double* A = t->absa_lw[13];
      const double colco = 0.0;
      if (low) {
        spec_t s = mkspec(L->colh2o, L->rat_h2on2o, L->coln2o, 8.0, oneminus);
        spec_t s1 = mkspec(L->colh2o, L->rat_h2on2o_1, L->coln2o, 8.0, oneminus);
        spec_t smco2 = mkspec(L->colh2o, CHI(1, 1) / CHI(4, 1), L->coln2o, 8.0, oneminus);
        spec_t smco = mkspec(L->colh2o, CHI(1, 3) / CHI(4, 3), L->coln2o, 8.0, oneminus);
        spec_t sp = mkspec(L->colh2o, CHI(1, 5) / CHI(4, 5), L->coln2o, 8.0, oneminus);
        /* rrtm_taumol13.F90: reference CO2 mixing ratio is the constant 3.55e-4; the second occurrence has no
         * kind suffix in the source (single precision) */
        double adjcolco2;
        {
          double chi_co2 = L->colco2 / L->coldry;
          double ratco2 = 1.E20 * chi_co2 / 3.55E-4;
          if (ratco2 > 3.0) {
            double adjfac = 2.0 + pow(ratco2 - 2.0, 0.68);
            adjcolco2 = adjfac * (double)3.55E-4f * L->coldry * 1.E-20;
          } else adjcolco2 = L->colco2;
        }
        int ind0 = IND0A(9) + s.js, ind1 = IND1A(9) + s1.js;
        majfac m0 = lowfac(&s, L->fac00, L->fac10), m1 = lowfac(&s1, L->fac01, L->fac11);
        for (int ig = 0; ig < ng; ++ig) {
          double tauself = SELF(13, ig), taufor = FOR(13, ig);
          double absco2 = minor2(t->ka_mco2_13, 9, ig, smco2.js, smco2.fs, indm, L->minorfrac);
          double absco = minor2(t->ka_mco_13, 9, ig, smco.js, smco.fs, indm, L->minorfrac);
          double maj = majsum(A + ig * 585, ind0, &m0, s.speccomb), maj1 = majsum(A + ig * 585, ind1, &m1, s1.speccomb);
          T[o + ig] = maj + maj1 + tauself + taufor + adjcolco2 * absco2 + colco * absco;
          P[o + ig] = PLANCK_INTERP(t->fracrefa_lw[13], sp, ng, ig);
        }
      } else {
        for (int ig = 0; ig < ng; ++ig) {
          double abso3 = lin(t->kb_mo3_13 + ig * 19, indm, L->minorfrac);
          T[o + ig] = L->colo3 * abso3;
          P[o + ig] = t->fracrefb_lw[13][ig];
        }
      }
    }
    /* ---- band 14: rrtm_taumol14.F90 (CO2 / CO2) ---- */
    {
      const int ng = 2, o = ngs[13]; const double *A = t->absa_lw[14], *B = t->absb_lw[14];
      if (low) {
        int ind0 = IND0A(1) + 1, ind1 = IND1A(1) + 1;
        for (int ig = 0; ig < ng; ++ig) {
          double tauself = SELF(14, ig), taufor = FOR(14, ig);
          T[o + ig] = L->colco2 * major1(A + ig * 65, ind0, ind1, L) + tauself + taufor;
          P[o + ig] = t->fracrefa_lw[14][ig];
        }
      } else {
        int ind0 = IND0B(1) + 1, ind1 = IND1B(1) + 1;
        for (int ig = 0; ig < ng; ++ig) {
          T[o + ig] = L->colco2 * major1(B + ig * 235, ind0, ind1, L);
          P[o + ig] = t->fracrefb_lw[14][ig];
        }
      }
    }
    /* ---- band 15: rrtm_taumol15.F90 (N2O,CO2 / -; minor N2) ---- */
    {
      const int ng = 2, o = ngs[14]; const double* A = t->absa_lw[15];
      if (low) {
        spec_t s = mkspec(L->coln2o, L->rat_n2oco2, L->colco2, 8.0, oneminus);
        spec_t s1 = mkspec(L->coln2o, L->rat_n2oco2_1, L->colco2, 8.0, oneminus);
        spec_t sm = mkspec(L->coln2o, CHI(4, 1) / CHI(2, 1), L->colco2, 8.0, oneminus);
        spec_t sp = mkspec(L->coln2o, CHI(4, 1) / CHI(2, 1), L->colco2, 8.0, oneminus);
        int ind0 = IND0A(9) + s.js, ind1 = IND1A(9) + s1.js;
        double scalen2 = L->colbrd * L->scaleminor;
        majfac m0 = lowfac(&s, L->fac00, L->fac10), m1 = lowfac(&s1, L->fac01, L->fac11);
        for (int ig = 0; ig < ng; ++ig) {
          double tauself = SELF(15, ig), taufor = FOR(15, ig);
          double taun2 = scalen2 * minor2(t->ka_mn2_15, 9, ig, sm.js, sm.fs, indm, L->minorfrac);
          double maj = majsum(A + ig * 585, ind0, &m0, s.speccomb), maj1 = majsum(A + ig * 585, ind1, &m1, s1.speccomb);
          T[o + ig] = maj + maj1 + tauself + taufor + taun2;
          P[o + ig] = PLANCK_INTERP(t->fracrefa_lw[15], sp, ng, ig);
        }
      } else {
        for (int ig = 0; ig < ng; ++ig) { T[o + ig] = 0.0; P[o + ig] = 0.0; }
      }
    }
    /* ---- band 16: rrtm_taumol16.F90 (H2O,CH4 / CH4) ---- */
    {
      const int ng = 2, o = ngs[15]; const double *A = t->absa_lw[16], *B = t->absb_lw[16];
      if (low) {
        spec_t s = mkspec(L->colh2o, L->rat_h2och4, L->colch4, 8.0, oneminus);
        spec_t s1 = mkspec(L->colh2o, L->rat_h2och4_1, L->colch4, 8.0, oneminus);
        spec_t sp = mkspec(L->colh2o, CHI(1, 6) / CHI(6, 6), L->colch4, 8.0, oneminus);
        int ind0 = IND0A(9) + s.js, ind1 = IND1A(9) + s1.js;
        majfac m0 = lowfac(&s, L->fac00, L->fac10), m1 = lowfac(&s1, L->fac01, L->fac11);
        for (int ig = 0; ig < ng; ++ig) {
          double tauself = SELF(16, ig), taufor = FOR(16, ig);
          double maj = majsum(A + ig * 585, ind0, &m0, s.speccomb), maj1 = majsum(A + ig * 585, ind1, &m1, s1.speccomb);
          T[o + ig] = maj + maj1 + tauself + taufor;
          P[o + ig] = PLANCK_INTERP(t->fracrefa_lw[16], sp, ng, ig);
        }
      } else {
        /* NSPB(16) = 0 in ifsrrtm/surrtpk.F90:23, so the reference always reads rows 1 and 2 of ABSB here
           (rrtm_taumol16.F90: IND0 = (...)*NSPB(16) + 1); reproduced as is */
        int ind0 = IND0B(0) + 1, ind1 = IND1B(0) + 1;
        for (int ig = 0; ig < ng; ++ig) {
          T[o + ig] = L->colch4 * major1(B + ig * 235, ind0, ind1, L);
          P[o + ig] = t->fracrefb_lw[16][ig];
        }
      }
    }
#undef SELF
#undef FOR
  }
}
