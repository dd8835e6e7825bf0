/* rrtmg_sw.c -- oracle restatement of the RRTMG shortwave gas optics (IFS version).  TEST INFRASTRUCTURE.
 *
 * Follows: ifsrrtm/srtm_setcoef.F90:78-220, ifsrrtm/srtm_gas_optical_depth.F90:136-323,
 * ifsrrtm/srtm_taumol16..29.F90.  Layers in RRTMG order (1 = bottom).  The source uses many default-kind
 * (single precision) literals (1.E-20, 7.2, 4.35e-4 ...); they are reproduced as (double)<x>f.
 */
#include <math.h>
#include <string.h>
#include "oracle.h"

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* srtm_setcoef.F90:78-220.  `gas` supplies PAVEL, TAVEL, COLDRY, WKL from rrtm_prepare_gases. */
void orc_setcoef_sw(const orc_tables* t, int nlev, const orc_lay_lw* gas, orc_lay_sw* lay, int* laytrop_out) {
  const double stpfac = 296.0 / 1013.0;
  const double e20 = (double)1.E-20f, e32 = (double)1.E-32f;
  int laytrop = 0;
  for (int jk = 0; jk < nlev; ++jk) {
    const orc_lay_lw* G = &gas[jk];
    orc_lay_sw* L = &lay[jk];
    memset(L, 0, sizeof(*L));
    double plog = log(G->pavel);
    int jp = (int)(36.0 - 5.0 * (plog + 0.04));
    if (jp < 1) jp = 1; else if (jp > 58) jp = 58;
    int jp1 = jp + 1;
    double fp = 5. * (t->preflog_sw[jp - 1] - plog);
    int jt = (int)(3. + (G->tavel - t->tref_sw[jp - 1]) / 15.);
    if (jt < 1) jt = 1; else if (jt > 4) jt = 4;
    double ft = ((G->tavel - t->tref_sw[jp - 1]) / 15.) - (double)(jt - 3);
    int jt1 = (int)(3. + (G->tavel - t->tref_sw[jp1 - 1]) / 15.);
    if (jt1 < 1) jt1 = 1; else if (jt1 > 4) jt1 = 4;
    double ft1 = ((G->tavel - t->tref_sw[jp1 - 1]) / 15.) - (double)(jt1 - 3);
    double water = G->wkl[1] / G->coldry;
    double scalefac = G->pavel * stpfac / G->tavel;
    L->jp = jp; L->jt = jt; L->jt1 = jt1;
    if (jp < 13) {
      laytrop++;
      L->forfac = scalefac / (1. + water);
      double factor = (332.0 - G->tavel) / 36.0;
      L->indfor = imin(2, imax(1, (int)factor));
      L->forfrac = factor - (double)L->indfor;
      L->selffac = water * L->forfac;
      factor = (G->tavel - 188.0) / (double)7.2f;
      L->indself = imin(9, imax(1, (int)factor - 7));
      L->selffrac = factor - (double)(L->indself + 7);
    } else {
      L->forfac = scalefac / (1. + water);
      double factor = (G->tavel - 188.0) / 36.0;
      L->indfor = 3;
      L->forfrac = factor - 1.0;
      L->selffac = 0.0; L->selffrac = 0.0; L->indself = 0;
    }
    L->colh2o = e20 * G->wkl[1];
    L->colco2 = e20 * G->wkl[2];
    L->colo3 = e20 * G->wkl[3];
    L->colch4 = e20 * G->wkl[6];
    L->colo2 = e20 * G->wkl[7];
    L->colmol = e20 * G->coldry + L->colh2o;
    if (L->colco2 == 0.) L->colco2 = e32 * G->coldry;
    if (L->colch4 == 0.) L->colch4 = e32 * G->coldry;
    if (L->colo2 == 0.) L->colo2 = e32 * G->coldry;
    double compfp = 1. - fp;
    L->fac10 = compfp * ft;
    L->fac00 = compfp * (1. - ft);
    L->fac11 = fp * ft1;
    L->fac01 = fp * (1. - ft1);
  }
  *laytrop_out = laytrop;
}

#define IND0A(nspa) (((L->jp - 1) * 5 + (L->jt - 1)) * (nspa))
#define IND1A(nspa) ((L->jp * 5 + (L->jt1 - 1)) * (nspa))
#define IND0B(nspb) (((L->jp - 13) * 5 + (L->jt - 1)) * (nspb))
#define IND1B(nspb) (((L->jp - 12) * 5 + (L->jt1 - 1)) * (nspb))

typedef struct { double speccomb, fs; int js; } sspec;
/* e.g. srtm_taumol16.F90:96-101 */
static sspec mkspec(double cola, double strrat, double colb, double mult, double oneminus) {
  sspec s;
  s.speccomb = cola + strrat * colb;
  double specparm = cola / s.speccomb;
  if (specparm >= oneminus) specparm = oneminus;
  double specmult = mult * specparm;
  s.js = 1 + (int)specmult;
  s.fs = fmod(specmult, 1.0);
  return s;
}
/* SPECCOMB*((1-FS)*(A(i0)*f00 + A(i0+d)*f10 + A(i1)*f01 + A(i1+d)*f11) + FS*(A(i0+1)*f00 + A(i0+d+1)*f10 + ...)) */
static inline double major2(const double* col, int i0, int i1, int d, const sspec* s, const orc_lay_sw* L) {
  double a = col[i0 - 1] * L->fac00 + col[i0 + d - 1] * L->fac10 + col[i1 - 1] * L->fac01 + col[i1 + d - 1] * L->fac11;
  double b = col[i0] * L->fac00 + col[i0 + d] * L->fac10 + col[i1] * L->fac01 + col[i1 + d] * L->fac11;
  return s->speccomb * ((1. - s->fs) * a + s->fs * b);
}
static inline double major1(const double* col, int i0, int i1, const orc_lay_sw* L) {
  return L->fac00 * col[i0 - 1] + L->fac10 * col[i0] + L->fac01 * col[i1 - 1] + L->fac11 * col[i1];
}
static inline double lin(const double* col, int i, double f) { return col[i - 1] + f * (col[i] - col[i - 1]); }

/* Writes od/ssa (gas + Rayleigh) in [lay][112] RRTMG order and incsol[112]; srtm_gas_optical_depth.F90:305-321 */
void orc_taumol_sw(const orc_tables* t, int nlev, const orc_lay_sw* lay, int laytrop, double* od, double* ssa,
                   double* incsol) {
  const double oneminus = 1.0 - 1.0e-6;
  int iw = 0;
  for (int jb = 16; jb <= 29; ++jb) {
    const int ng = t->ngc_sw[jb - 16];
    const double *A = t->absa_sw[jb], *B = t->absb_sw[jb];
    const double *SR = t->selfref_sw[jb], *FR = t->forref_sw[jb], *SF = t->sfluxref_sw[jb];
    const int nfor = t->nfor_sw[jb];
    const double strrat = t->strrat_sw[jb];
    const int layreffr = t->layreffr_sw[jb];
    const double rayl = t->rayl_sw[jb] ? t->rayl_sw[jb][0] : 0.0;
    const double* raylc = t->raylc_sw[jb];
    double taug[16], taur[16], sflux[16];
    for (int ig = 0; ig < 16; ++ig) sflux[ig] = 0.0;
    /* bands whose solar source is taken in the upper atmosphere initialise LAYSOLFR=NLAYERS (16,17,27,28,29);
       the others start from LAYTROP and look for the reference level in the lower atmosphere */
    const int upper_src = (jb == 16 || jb == 17 || jb == 27 || jb == 28 || jb == 29);
    int laysolfr = upper_src ? nlev : laytrop;
    for (int il = 1; il <= nlev; ++il) {
      const orc_lay_sw* L = &lay[il - 1];
      const int low = il <= laytrop;
      int setflux = 0; sspec s = {0, 0, 1};
      if (low && !upper_src && jb != 26) {
        int inext = imin(nlev, il + 1);
        if (L->jp < layreffr && lay[inext - 1].jp >= layreffr) laysolfr = imin(il + 1, laytrop);
      }
      if (!low && upper_src) {
        if (il >= 2 && lay[il - 2].jp < layreffr && L->jp >= layreffr) laysolfr = il;
      }
      if (il == laysolfr && (low ? !upper_src : upper_src)) setflux = 1;
#define SELFFOR(ig) (L->colh2o * (L->selffac * lin(SR + (ig) * 10, L->indself, L->selffrac) + \
                                  L->forfac * lin(FR + (ig) * nfor, L->indfor, L->forfrac)))
#define FORONLY(ig) (L->colh2o * L->forfac * lin(FR + (ig) * nfor, L->indfor, L->forfrac))
      switch (jb) {
        case 16:
          if (low) {
            s = mkspec(L->colh2o, strrat, L->colch4, 8.0, oneminus);
            int i0 = IND0A(9) + s.js, i1 = IND1A(9) + s.js;
            for (int ig = 0; ig < ng; ++ig) { taug[ig] = major2(A + ig * 585, i0, i1, 9, &s, L) + SELFFOR(ig); taur[ig] = L->colmol * rayl; }
          } else {
            int i0 = IND0B(1) + 1, i1 = IND1B(1) + 1;
            for (int ig = 0; ig < ng; ++ig) {
              taug[ig] = L->colch4 * major1(B + ig * 235, i0, i1, L);
              if (setflux) sflux[ig] = SF[ig];
              taur[ig] = L->colmol * rayl;
            }
          }
          break;
        case 17:
          if (low) {
            s = mkspec(L->colh2o, strrat, L->colco2, 8., oneminus);
            int i0 = IND0A(9) + s.js, i1 = IND1A(9) + s.js;
            for (int ig = 0; ig < ng; ++ig) { taug[ig] = major2(A + ig * 585, i0, i1, 9, &s, L) + SELFFOR(ig); taur[ig] = L->colmol * rayl; }
          } else {
            s = mkspec(L->colh2o, strrat, L->colco2, 4., oneminus);
            int i0 = IND0B(5) + s.js, i1 = IND1B(5) + s.js;
            for (int ig = 0; ig < ng; ++ig) {
              taug[ig] = major2(B + ig * 1175, i0, i1, 5, &s, L) + FORONLY(ig);
              if (setflux) sflux[ig] = SF[(s.js - 1) * ng + ig] + s.fs * (SF[s.js * ng + ig] - SF[(s.js - 1) * ng + ig]);
              taur[ig] = L->colmol * rayl;
            }
          }
          break;
        case 18: case 19: case 21: case 22: case 24: {
          double colb = (jb == 18) ? L->colch4 : (jb == 19 || jb == 21) ? L->colco2 : L->colo2;
          const double o2adj = 1.6;
          double o2cont = (double)4.35e-4f * L->colo2 / (double)(350.0f * 2.0f);
          if (low) {
            s = (jb == 22) ? mkspec(L->colh2o, o2adj * strrat, colb, 8., oneminus) : mkspec(L->colh2o, strrat, colb, 8., oneminus);
            int i0 = IND0A(9) + s.js, i1 = IND1A(9) + s.js;
            for (int ig = 0; ig < ng; ++ig) {
              double maj = major2(A + ig * 585, i0, i1, 9, &s, L);
              if (jb == 22) taug[ig] = maj + SELFFOR(ig) + o2cont;
              else if (jb == 24) taug[ig] = maj + L->colo3 * t->abso3a_24[ig] + SELFFOR(ig);
              else taug[ig] = maj + SELFFOR(ig);
              if (setflux) sflux[ig] = SF[(s.js - 1) * ng + ig] + s.fs * (SF[s.js * ng + ig] - SF[(s.js - 1) * ng + ig]);
              if (jb == 24) {
                const double* R = t->raylac_24;
                taur[ig] = L->colmol * (R[(s.js - 1) * ng + ig] + s.fs * (R[s.js * ng + ig] - R[(s.js - 1) * ng + ig]));
              } else taur[ig] = L->colmol * rayl;
            }
          } else if (jb == 21) {
            s = mkspec(L->colh2o, strrat, L->colco2, 4., oneminus);
            int i0 = IND0B(5) + s.js, i1 = IND1B(5) + s.js;
            for (int ig = 0; ig < ng; ++ig) { taug[ig] = major2(B + ig * 1175, i0, i1, 5, &s, L) + FORONLY(ig); taur[ig] = L->colmol * rayl; }
          } else {
            int i0 = IND0B(1) + 1, i1 = IND1B(1) + 1;
            for (int ig = 0; ig < ng; ++ig) {
              double m = major1(B + ig * 235, i0, i1, L);
              if (jb == 22) { taug[ig] = L->colo2 * o2adj * m + o2cont; taur[ig] = L->colmol * rayl; }
              else if (jb == 24) { taug[ig] = L->colo2 * m + L->colo3 * t->abso3b_24[ig]; taur[ig] = L->colmol * t->raylbc_24[ig]; }
              else { taug[ig] = colb * m; taur[ig] = L->colmol * rayl; }
            }
          }
        } break;
        case 20:
          if (low) {
            int i0 = IND0A(1) + 1, i1 = IND1A(1) + 1;
            for (int ig = 0; ig < ng; ++ig) {
              taug[ig] = L->colh2o * ((major1(A + ig * 65, i0, i1, L)) + L->selffac * lin(SR + ig * 10, L->indself, L->selffrac) +
                                      L->forfac * lin(FR + ig * nfor, L->indfor, L->forfrac)) + L->colch4 * t->absch4_20[ig];
              taur[ig] = L->colmol * rayl;
              if (setflux) sflux[ig] = SF[ig];
            }
          } else {
            int i0 = IND0B(1) + 1, i1 = IND1B(1) + 1;
            for (int ig = 0; ig < ng; ++ig) {
              taug[ig] = L->colh2o * (major1(B + ig * 235, i0, i1, L) + L->forfac * lin(FR + ig * nfor, L->indfor, L->forfrac)) +
                         L->colch4 * t->absch4_20[ig];
              taur[ig] = L->colmol * rayl;
            }
          }
          break;
        case 23:
          if (low) {
            int i0 = IND0A(1) + 1, i1 = IND1A(1) + 1;
            for (int ig = 0; ig < ng; ++ig) {
              taug[ig] = L->colh2o * (t->givfac_23 * (major1(A + ig * 65, i0, i1, L)) +
                                      L->selffac * lin(SR + ig * 10, L->indself, L->selffrac) +
                                      L->forfac * lin(FR + ig * nfor, L->indfor, L->forfrac));
              if (setflux) sflux[ig] = SF[ig];
              taur[ig] = L->colmol * raylc[ig];
            }
          } else {
            for (int ig = 0; ig < ng; ++ig) { taug[ig] = 0.0; taur[ig] = L->colmol * raylc[ig]; }
          }
          break;
        case 25:
          if (low) {
            int i0 = IND0A(1) + 1, i1 = IND1A(1) + 1;
            for (int ig = 0; ig < ng; ++ig) {
              taug[ig] = L->colh2o * major1(A + ig * 65, i0, i1, L) + L->colo3 * t->abso3a_25[ig];
              if (setflux) sflux[ig] = SF[ig];
              taur[ig] = L->colmol * raylc[ig];
            }
          } else {
            for (int ig = 0; ig < ng; ++ig) { taug[ig] = L->colo3 * t->abso3b_25[ig]; taur[ig] = L->colmol * raylc[ig]; }
          }
          break;
        case 26:
          for (int ig = 0; ig < ng; ++ig) {
            if (low && il == laysolfr) sflux[ig] = SF[ig];
            taug[ig] = 0.0; taur[ig] = L->colmol * raylc[ig];
          }
          break;
        case 27:
          if (low) {
            int i0 = IND0A(1) + 1, i1 = IND1A(1) + 1;
            for (int ig = 0; ig < ng; ++ig) { taug[ig] = L->colo3 * major1(A + ig * 65, i0, i1, L); taur[ig] = L->colmol * raylc[ig]; }
          } else {
            int i0 = IND0B(1) + 1, i1 = IND1B(1) + 1;
            for (int ig = 0; ig < ng; ++ig) {
              taug[ig] = L->colo3 * major1(B + ig * 235, i0, i1, L);
              if (setflux) sflux[ig] = t->scalekur_27 * SF[ig];
              taur[ig] = L->colmol * raylc[ig];
            }
          }
          break;
        case 28:
          if (low) {
            s = mkspec(L->colo3, strrat, L->colo2, 8., oneminus);
            int i0 = IND0A(9) + s.js, i1 = IND1A(9) + s.js;
            for (int ig = 0; ig < ng; ++ig) { taug[ig] = major2(A + ig * 585, i0, i1, 9, &s, L); taur[ig] = L->colmol * rayl; }
          } else {
            s = mkspec(L->colo3, strrat, L->colo2, 4., oneminus);
            int i0 = IND0B(5) + s.js, i1 = IND1B(5) + s.js;
            for (int ig = 0; ig < ng; ++ig) {
              taug[ig] = major2(B + ig * 1175, i0, i1, 5, &s, L);
              if (setflux) sflux[ig] = SF[(s.js - 1) * ng + ig] + s.fs * (SF[s.js * ng + ig] - SF[(s.js - 1) * ng + ig]);
              taur[ig] = L->colmol * rayl;
            }
          }
          break;
        case 29:
          if (low) {
            int i0 = IND0A(1) + 1, i1 = IND1A(1) + 1;
            for (int ig = 0; ig < ng; ++ig) {
              taug[ig] = L->colh2o * ((major1(A + ig * 65, i0, i1, L)) + L->selffac * lin(SR + ig * 10, L->indself, L->selffrac) +
                                      L->forfac * lin(FR + ig * nfor, L->indfor, L->forfrac)) + L->colco2 * t->absco2_29[ig];
              taur[ig] = L->colmol * rayl;
            }
          } else {
            int i0 = IND0B(1) + 1, i1 = IND1B(1) + 1;
            for (int ig = 0; ig < ng; ++ig) {
              taug[ig] = L->colco2 * major1(B + ig * 235, i0, i1, L) + L->colh2o * t->absh2o_29[ig];
              if (setflux) sflux[ig] = SF[ig];
              taur[ig] = L->colmol * rayl;
            }
          }
          break;
      }
      /* srtm_gas_optical_depth.F90:314-320 */
      for (int ig = 0; ig < ng; ++ig) {
        double o = taur[ig] + taug[ig];
        od[(size_t)(il - 1) * NG_SW + iw + ig] = o;
        ssa[(size_t)(il - 1) * NG_SW + iw + ig] = taur[ig] / o;
      }
    }
    for (int ig = 0; ig < ng; ++ig) incsol[iw + ig] = sflux[ig];
    iw += ng;
  }
}
