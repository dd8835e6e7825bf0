/* tripleclouds.c -- oracle restatement of the Tripleclouds solvers (3 regions).  TEST INFRASTRUCTURE.
 * Follows radiation/radiation_regions.F90:35-199 (calc_region_properties, gamma PDF), radiation_overlap.F90:130-209
 * (calc_alpha_overlap_matrix) and :280-457 (calc_overlap_matrices), radiation_matrix.F90:110-136 (singlemat_x_vec),
 * radiation_tripleclouds_sw.F90:42-661, radiation_tripleclouds_lw.F90:38-605, radiation_lw_derivatives.F90:200-290
 * (calc_lw_derivatives_region).  Per-column; arrays [lev][reg][g] with g fastest.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

/* spectral sizes from the configuration (RRTMG or ecCKD) */
#undef NG_LW
#undef NG_SW
#undef NB_LW
#undef NB_SW
#define NG_LW (cfg->n_g_lw)
#define NG_SW (cfg->n_g_sw)
#define NB_LW (cfg->n_bands_lw)
#define NB_SW (cfg->n_bands_sw)

#define NREG 3
static inline double dmin(double a, double b) { return a < b ? a : b; }
static inline double dmax(double a, double b) { return a > b ? a : b; }

/* radiation_regions.F90:35-199, nreg = 3, do_gamma = .true. (config%i_cloud_pdf_shape default) */
void orc_region_properties(int nlev, const double* frac, const double* fsd, double frac_threshold, int lognormal,
                           double (*reg_fracs)[NREG], double (*od_scaling)[NREG]) {
  const double MinGammaODScaling = 0.025, MinLowerFrac = 0.5, MaxLowerFrac = 0.9, FSDAtMinLowerFrac = 1.5,
               FSDAtMaxLowerFrac = 3.725;
  const double LowerFracFSDGradient = (MaxLowerFrac - MinLowerFrac) / (FSDAtMaxLowerFrac - FSDAtMinLowerFrac);
  const double LowerFracFSDIntercept = MinLowerFrac - FSDAtMinLowerFrac * LowerFracFSDGradient;
  for (int jl = 0; jl < nlev; ++jl) {
    if (frac[jl] < frac_threshold) {
      reg_fracs[jl][0] = 1.0; reg_fracs[jl][1] = 0.0; reg_fracs[jl][2] = 0.0;
      od_scaling[jl][1] = 1.0; od_scaling[jl][2] = 1.0;
    } else if (lognormal) {   /* radiation_regions.F90:110-126 */
      reg_fracs[jl][0] = 1.0 - frac[jl];
      reg_fracs[jl][1] = frac[jl] * 0.5; reg_fracs[jl][2] = frac[jl] * 0.5;
      od_scaling[jl][1] = exp(-sqrt(log(fsd[jl] * fsd[jl] + 1.0))) / sqrt(fsd[jl] * fsd[jl] + 1.0);
      od_scaling[jl][2] = 2.0 - od_scaling[jl][1];
    } else {
      reg_fracs[jl][0] = 1.0 - frac[jl];
      reg_fracs[jl][1] = frac[jl] * dmax(MinLowerFrac, dmin(MaxLowerFrac, LowerFracFSDIntercept + fsd[jl] * LowerFracFSDGradient));
      od_scaling[jl][1] = MinGammaODScaling +
                          (1.0 - MinGammaODScaling) * exp(-fsd[jl] * (1.0 + 0.5 * fsd[jl] * (1.0 + 0.5 * fsd[jl])));
      reg_fracs[jl][2] = 1.0 - reg_fracs[jl][0] - reg_fracs[jl][1];
      od_scaling[jl][2] = (frac[jl] - reg_fracs[jl][1] * od_scaling[jl][1]) / reg_fracs[jl][2];
    }
    od_scaling[jl][0] = 0.0;
  }
}

/* radiation_overlap.F90:130-209; M[jupper][jlower] */
static void alpha_overlap_matrix(double op, double op_inhom, const double* fu, const double* fl, double M[NREG][NREG]) {
  double cf_upper = fu[1] + fu[2], cf_lower = fl[1] + fl[2];
  double pair_cloud_cover = op * dmax(cf_upper, cf_lower) + (1.0 - op) * (cf_upper + cf_lower - cf_upper * cf_lower);
  M[0][0] = 1.0 - pair_cloud_cover;
  double one_over_cf = 1.0 / dmax(cf_lower, 1.0e-6);
  M[0][1] = (pair_cloud_cover - cf_upper) * fl[1] * one_over_cf;
  M[0][2] = (pair_cloud_cover - cf_upper) * fl[2] * one_over_cf;
  one_over_cf = 1.0 / dmax(cf_upper, 1.0e-6);
  M[1][0] = (pair_cloud_cover - cf_lower) * fu[1] * one_over_cf;
  M[2][0] = (pair_cloud_cover - cf_lower) * fu[2] * one_over_cf;
  double frac_both = cf_upper + cf_lower - pair_cloud_cover;
  cf_upper = fu[2] / dmax(cf_upper, 1.0e-6);
  cf_lower = fl[2] / dmax(cf_lower, 1.0e-6);
  pair_cloud_cover = op_inhom * dmax(cf_upper, cf_lower) + (1.0 - op_inhom) * (cf_upper + cf_lower - cf_upper * cf_lower);
  M[1][1] = frac_both * (1.0 - pair_cloud_cover);
  M[1][2] = frac_both * (pair_cloud_cover - cf_upper);
  M[2][1] = frac_both * (pair_cloud_cover - cf_lower);
  M[2][2] = frac_both * (cf_upper + cf_lower - pair_cloud_cover);
}

/* calc_beta_overlap_matrix, radiation_overlap.F90:63-122 (Shonk et al. 2010 "beta" overlap parameter per region) */
static void beta_overlap_matrix(const double* op, const double* fu, const double* fl, double frac_threshold, double M[NREG][NREG]) {
  double denominator = 1.0, op_x_frac_min[NREG];
  for (int r = 0; r < NREG; ++r) {
    op_x_frac_min[r] = op[r] * dmin(fu[r], fl[r]);
    denominator = denominator - op_x_frac_min[r];
  }
  if (denominator >= frac_threshold) {
    const double factor = 1.0 / denominator;
    for (int ju = 0; ju < NREG; ++ju)
      for (int jw = 0; jw < NREG; ++jw) M[ju][jw] = factor * (fl[jw] - op_x_frac_min[jw]) * (fu[ju] - op_x_frac_min[ju]);
  } else {
    for (int ju = 0; ju < NREG; ++ju) for (int jw = 0; jw < NREG; ++jw) M[ju][jw] = 0.0;
  }
  for (int r = 0; r < NREG; ++r) M[r][r] = M[r][r] + op_x_frac_min[r];
}

/* radiation_overlap.F90:280-457.  U[jlev][jupper][jlower] = u_matrix(jupper,jlower,jlev); V[jlev][a][b] = v_matrix(a,b,jlev);
 * jlev = 0..nlev (half-levels). */
void orc_overlap_matrices(int nlev, double (*reg_fracs)[NREG], const double* overlap_param, double decorrelation_scaling,
                          double frac_threshold, int use_beta_overlap, double (*U)[NREG][NREG], double (*V)[NREG][NREG], double* cloud_cover) {
  double frac_upper[NREG] = {1.0, 0.0, 0.0}, frac_lower[NREG], M[NREG][NREG];
  for (int jlev = 1; jlev <= nlev + 1; ++jlev) {
    if (jlev > nlev) { frac_lower[0] = 1.0; frac_lower[1] = 0.0; frac_lower[2] = 0.0; }
    else for (int r = 0; r < NREG; ++r) frac_lower[r] = reg_fracs[jlev - 1][r];
    double op1, op2;
    if (jlev == 1 || jlev > nlev) { op1 = 1.0; op2 = 1.0; }
    else {
      op1 = overlap_param[jlev - 2];
      op2 = op1 >= 0.0 ? pow(op1, 1.0 / decorrelation_scaling) : op1;
    }
    if (use_beta_overlap) { const double op[NREG] = {op1, op2, op2}; beta_overlap_matrix(op, frac_upper, frac_lower, frac_threshold, M); }
    else alpha_overlap_matrix(op1, op2, frac_upper, frac_lower, M);
    for (int ju = 0; ju < NREG; ++ju)
      for (int jw = 0; jw < NREG; ++jw) {
        U[jlev - 1][ju][jw] = frac_lower[jw] >= frac_threshold ? M[ju][jw] / frac_lower[jw] : 0.0;
        V[jlev - 1][jw][ju] = frac_upper[ju] >= frac_threshold ? M[ju][jw] / frac_upper[ju] : 0.0;
      }
    for (int r = 0; r < NREG; ++r) frac_upper[r] = frac_lower[r];
  }
  double prod = 1.0;
  for (int jlev = 0; jlev <= nlev; ++jlev) prod = prod * V[jlev][0][0];
  *cloud_cover = 1.0 - prod;
}

#define A3(p, l, r, g) ((p)[((size_t)(l) * NREG + (r)) * ng + (g)])
#define A2L(p, l, g) ((p)[(size_t)(l) * ng + (g)])

/* radiation_tripleclouds_sw.F90:42-661 for one sunlit column.  Flux sums are returned per half-level; per-g surface/TOA. */
void orc_tripleclouds_sw(const orc_tables* t, const ecrad_b200_config* cfg, int nlev, double mu0, const double* frac,
                         const double* fsd, const double* overlap_param, const double* od, const double* ssa, const double* g,
                         const double* od_cloud, const double* ssa_cloud, const double* g_cloud, const double* incoming,
                         const double* alb_diff, const double* alb_dir, orc_tc_out* o) {
  const int ng = NG_SW;
  double (*reg)[NREG] = malloc(sizeof(double[NREG]) * nlev), (*ods)[NREG] = malloc(sizeof(double[NREG]) * nlev);
  double (*U)[NREG][NREG] = malloc(sizeof(double[NREG][NREG]) * (nlev + 1)), (*V)[NREG][NREG] = malloc(sizeof(double[NREG][NREG]) * (nlev + 1));
  orc_region_properties(nlev, frac, fsd, cfg->cloud_fraction_threshold, cfg->i_cloud_pdf_shape == ECRAD_PDF_LOGNORMAL, reg, ods);
  orc_overlap_matrices(nlev, reg, overlap_param, cfg->cloud_inhom_decorr_scaling, cfg->cloud_fraction_threshold, cfg->use_beta_overlap, U, V, &o->cloud_cover);
  int* clear = calloc(nlev + 2, sizeof(int));   /* is_clear_sky_layer(0:nlev+1) */
  clear[0] = 1; clear[nlev + 1] = 1;
  for (int jl = 1; jl <= nlev; ++jl) clear[jl] = !(frac[jl - 1] > 0.0);
  const size_t nl = (size_t)nlev * ng;
  double* buf = calloc(5 * nl + 5 * nl * NREG + 2 * (size_t)(nlev + 1) * NREG * ng + 2 * (size_t)(nlev + 1) * ng + 16 * (size_t)ng, sizeof(double));
  double *ref_c = buf, *tr_c = ref_c + nl, *rdir_c = tr_c + nl, *tdd_c = rdir_c + nl, *tdir_c = tdd_c + nl;
  double *ref = tdir_c + nl, *tr = ref + nl * NREG, *rdir = tr + nl * NREG, *tdd = rdir + nl * NREG, *tdir = tdd + nl * NREG;
  double *talb = tdir + nl * NREG, *talb_dir = talb + (size_t)(nlev + 1) * NREG * ng;
  double *talb_clear = talb_dir + (size_t)(nlev + 1) * NREG * ng, *talb_clear_dir = talb_clear + (size_t)(nlev + 1) * ng;
  double *below = talb_clear_dir + (size_t)(nlev + 1) * ng, *below_dir = below + 3 * ng, *inv_denom = below_dir + 3 * ng;
  double *od_total = inv_denom + 3 * ng, *ssa_total = od_total + ng, *g_total = ssa_total + ng;
  orc_calc_ref_trans_sw(ng * nlev, mu0, od, ssa, g, ref_c, tr_c, rdir_c, tdd_c, tdir_c);
  for (int jl = 1; jl <= nlev; ++jl) {
    if (clear[jl]) continue;
    for (int jr = 1; jr < NREG; ++jr) {
      for (int jg = 0; jg < ng; ++jg) {
        const int ib = t->band_sw[jg];
        const size_t i = (size_t)(jl - 1) * ng + jg;
        double scat_od = od[i] * ssa[i];
        double scat_od_cloud = od_cloud[(jl - 1) * NB_SW + ib] * ssa_cloud[(jl - 1) * NB_SW + ib] * ods[jl - 1][jr];
        od_total[jg] = od[i] + od_cloud[(jl - 1) * NB_SW + ib] * ods[jl - 1][jr];
        ssa_total[jg] = (scat_od + scat_od_cloud) / od_total[jg];
        g_total[jg] = (scat_od * g[i] + scat_od_cloud * g_cloud[(jl - 1) * NB_SW + ib]) / (scat_od + scat_od_cloud);
      }
      if (cfg->do_sw_delta_scaling_with_gases)   /* radiation_tripleclouds_sw.F90:298-302: cloudy regions only, the clear region is not scaled */
        for (int jg = 0; jg < ng; ++jg) {
          const double f = g_total[jg] * g_total[jg];
          od_total[jg] = od_total[jg] * (1.0 - ssa_total[jg] * f);
          ssa_total[jg] = ssa_total[jg] * (1.0 - f) / (1.0 - ssa_total[jg] * f);
          g_total[jg] = g_total[jg] / (1.0 + g_total[jg]);
        }
      orc_calc_ref_trans_sw(ng, mu0, od_total, ssa_total, g_total, &A3(ref, jl - 1, jr, 0), &A3(tr, jl - 1, jr, 0),
                            &A3(rdir, jl - 1, jr, 0), &A3(tdd, jl - 1, jr, 0), &A3(tdir, jl - 1, jr, 0));
    }
  }
  /* half-level index hl = 0..nlev  (Fortran jlev+1 -> hl = jlev) */
  for (int jg = 0; jg < ng; ++jg) {
    A3(talb, nlev, 0, jg) = alb_diff[jg];
    A3(talb_dir, nlev, 0, jg) = mu0 * alb_dir[jg];
  }
  if (!clear[nlev])
    for (int jr = 1; jr < NREG; ++jr)
      for (int jg = 0; jg < ng; ++jg) { A3(talb, nlev, jr, jg) = A3(talb, nlev, 0, jg); A3(talb_dir, nlev, jr, jg) = A3(talb_dir, nlev, 0, jg); }
  for (int jg = 0; jg < ng; ++jg) { A2L(talb_clear, nlev, jg) = A3(talb, nlev, 0, jg); A2L(talb_clear_dir, nlev, jg) = A3(talb_dir, nlev, 0, jg); }
  for (int jl = nlev; jl >= 1; --jl) {
    const int l = jl - 1;   /* layer index 0-based; half-level above = l, below = l+1 = jl */
    memset(below, 0, sizeof(double) * 6 * ng);
    for (int jg = 0; jg < ng; ++jg) {
      double id = 1.0 / (1.0 - A2L(talb_clear, jl, jg) * A2L(ref_c, l, jg));
      A2L(talb_clear, l, jg) = A2L(ref_c, l, jg) + A2L(tr_c, l, jg) * A2L(tr_c, l, jg) * A2L(talb_clear, jl, jg) * id;
      A2L(talb_clear_dir, l, jg) = A2L(rdir_c, l, jg) +
          (A2L(tdir_c, l, jg) * A2L(talb_clear_dir, jl, jg) + A2L(tdd_c, l, jg) * A2L(talb_clear, jl, jg)) * A2L(tr_c, l, jg) * id;
    }
    for (int jg = 0; jg < ng; ++jg) {
      double id = 1.0 / (1.0 - A3(talb, jl, 0, jg) * A2L(ref_c, l, jg));
      inv_denom[jg] = id;
      below[jg] = A2L(ref_c, l, jg) + A2L(tr_c, l, jg) * A2L(tr_c, l, jg) * A3(talb, jl, 0, jg) * id;
      below_dir[jg] = A2L(rdir_c, l, jg) +
          (A2L(tdir_c, l, jg) * A3(talb_dir, jl, 0, jg) + A2L(tdd_c, l, jg) * A3(talb, jl, 0, jg)) * A2L(tr_c, l, jg) * id;
    }
    if (!clear[jl])
      for (int jr = 1; jr < NREG; ++jr)
        for (int jg = 0; jg < ng; ++jg) {
          double id = 1.0 / (1.0 - A3(talb, jl, jr, jg) * A3(ref, l, jr, jg));
          below[jr * ng + jg] = A3(ref, l, jr, jg) + A3(tr, l, jr, jg) * A3(tr, l, jr, jg) * A3(talb, jl, jr, jg) * id;
          below_dir[jr * ng + jg] = A3(rdir, l, jr, jg) +
              (A3(tdir, l, jr, jg) * A3(talb_dir, jl, jr, jg) + A3(tdd, l, jr, jg) * A3(talb, jl, jr, jg)) * A3(tr, l, jr, jg) * id;
        }
    if (clear[jl] && clear[jl - 1]) {
      for (int jr = 0; jr < NREG; ++jr)
        for (int jg = 0; jg < ng; ++jg) { A3(talb, l, jr, jg) = below[jr * ng + jg]; A3(talb_dir, l, jr, jg) = below_dir[jr * ng + jg]; }
    } else {
      for (int jr = 0; jr < NREG; ++jr)
        for (int jr2 = 0; jr2 < NREG; ++jr2)
          for (int jg = 0; jg < ng; ++jg) {
            A3(talb, l, jr, jg) = A3(talb, l, jr, jg) + below[jr2 * ng + jg] * V[l][jr2][jr];
            A3(talb_dir, l, jr, jg) = A3(talb_dir, l, jr, jg) + below_dir[jr2 * ng + jg] * V[l][jr2][jr];
          }
    }
  }
  /* fluxes */
  double* fl = calloc((size_t)ng * (3 * NREG + 3 + 3 * NREG), sizeof(double));
  double *flux_dn = fl, *direct_dn = flux_dn + NREG * ng, *flux_up = direct_dn + NREG * ng;
  double *flux_dn_clear = flux_up + NREG * ng, *direct_dn_clear = flux_dn_clear + ng, *flux_up_clear = direct_dn_clear + ng;
  double* tmp = flux_up_clear + ng;
  for (int jr = 0; jr < NREG; ++jr)
    for (int jg = 0; jg < ng; ++jg) {
      direct_dn[jr * ng + jg] = incoming[jg] * reg[0][jr];
      flux_up[jr * ng + jg] = direct_dn[jr * ng + jg] * A3(talb_dir, 0, jr, jg);
    }
  for (int jg = 0; jg < ng; ++jg) { direct_dn_clear[jg] = incoming[jg]; flux_up_clear[jg] = direct_dn_clear[jg] * A2L(talb_clear_dir, 0, jg); }
  for (int jg = 0; jg < ng; ++jg) {
    o->up_toa_g[jg] = flux_up[jg] + flux_up[ng + jg] + flux_up[2 * ng + jg];
    o->up_toa_clear_g[jg] = flux_up_clear[jg];
  }
  for (int hl = 0; hl <= nlev; ++hl) {
    if (hl > 0) {
      const int l = hl - 1, jl = hl;
      for (int jg = 0; jg < ng; ++jg) {
        flux_dn_clear[jg] = (A2L(tr_c, l, jg) * flux_dn_clear[jg] + direct_dn_clear[jg] *
                             (A2L(tdir_c, l, jg) * A2L(talb_clear_dir, jl, jg) * A2L(ref_c, l, jg) + A2L(tdd_c, l, jg))) /
                            (1.0 - A2L(ref_c, l, jg) * A2L(talb_clear, jl, jg));
        direct_dn_clear[jg] = A2L(tdir_c, l, jg) * direct_dn_clear[jg];
        flux_up_clear[jg] = direct_dn_clear[jg] * A2L(talb_clear_dir, jl, jg) + flux_dn_clear[jg] * A2L(talb_clear, jl, jg);
      }
      for (int jg = 0; jg < ng; ++jg) {
        flux_dn[jg] = (A2L(tr_c, l, jg) * flux_dn[jg] + direct_dn[jg] *
                       (A2L(tdir_c, l, jg) * A3(talb_dir, jl, 0, jg) * A2L(ref_c, l, jg) + A2L(tdd_c, l, jg))) /
                      (1.0 - A2L(ref_c, l, jg) * A3(talb, jl, 0, jg));
        direct_dn[jg] = A2L(tdir_c, l, jg) * direct_dn[jg];
        flux_up[jg] = direct_dn[jg] * A3(talb_dir, jl, 0, jg) + flux_dn[jg] * A3(talb, jl, 0, jg);
      }
      if (clear[jl]) {
        for (int jr = 1; jr < NREG; ++jr)
          for (int jg = 0; jg < ng; ++jg) { flux_dn[jr * ng + jg] = 0.0; flux_up[jr * ng + jg] = 0.0; direct_dn[jr * ng + jg] = 0.0; }
      } else {
        for (int jr = 1; jr < NREG; ++jr)
          for (int jg = 0; jg < ng; ++jg) {
            const size_t k = (size_t)jr * ng + jg;
            flux_dn[k] = (A3(tr, l, jr, jg) * flux_dn[k] + direct_dn[k] *
                          (A3(tdir, l, jr, jg) * A3(talb_dir, jl, jr, jg) * A3(ref, l, jr, jg) + A3(tdd, l, jr, jg))) /
                         (1.0 - A3(ref, l, jr, jg) * A3(talb, jl, jr, jg));
            direct_dn[k] = A3(tdir, l, jr, jg) * direct_dn[k];
            flux_up[k] = direct_dn[k] * A3(talb_dir, jl, jr, jg) + flux_dn[k] * A3(talb, jl, jr, jg);
          }
      }
      if (!(clear[jl] && clear[jl + 1])) {
        /* singlemat_x_vec(v_matrix(:,:,jlev+1), x): out(:,j1) = sum_j2 A(j1,j2) x(:,j2) */
        for (int pass = 0; pass < 2; ++pass) {
          double* x = pass ? direct_dn : flux_dn;
          memset(tmp, 0, sizeof(double) * NREG * ng);
          for (int j1 = 0; j1 < NREG; ++j1)
            for (int j2 = 0; j2 < NREG; ++j2)
              for (int jg = 0; jg < ng; ++jg) tmp[j1 * ng + jg] = tmp[j1 * ng + jg] + V[jl][j1][j2] * x[j2 * ng + jg];
          memcpy(x, tmp, sizeof(double) * NREG * ng);
        }
      }
    }
    double sum_up = 0.0, sum_dn_dir = 0.0, sum_dn_diff = 0.0;
    for (int jr = 0; jr < NREG; ++jr)
      for (int jg = 0; jg < ng; ++jg) { sum_up = sum_up + flux_up[jr * ng + jg]; sum_dn_diff = sum_dn_diff + flux_dn[jr * ng + jg]; sum_dn_dir = sum_dn_dir + direct_dn[jr * ng + jg]; }
    o->up[hl] = sum_up;
    o->dn[hl] = hl == 0 ? mu0 * sum_dn_dir : mu0 * sum_dn_dir + sum_dn_diff;
    o->dn_direct[hl] = mu0 * sum_dn_dir;
    sum_up = 0.0; sum_dn_dir = 0.0; sum_dn_diff = 0.0;
    for (int jg = 0; jg < ng; ++jg) { sum_up = sum_up + flux_up_clear[jg]; sum_dn_diff = sum_dn_diff + flux_dn_clear[jg]; sum_dn_dir = sum_dn_dir + direct_dn_clear[jg]; }
    o->up_clear[hl] = sum_up;
    o->dn_clear[hl] = hl == 0 ? mu0 * sum_dn_dir : mu0 * sum_dn_dir + sum_dn_diff;
    o->dn_direct_clear[hl] = mu0 * sum_dn_dir;
    if (o->up_g_prof) {   /* per-g totals over regions for the band profiles (do_save_spectral_flux) */
      for (int jg = 0; jg < ng; ++jg) {
        o->up_g_prof[(size_t)hl * ng + jg] = flux_up[jg] + flux_up[ng + jg] + flux_up[2 * ng + jg];
        o->dn_dir_g_prof[(size_t)hl * ng + jg] = direct_dn[jg] + direct_dn[ng + jg] + direct_dn[2 * ng + jg];
        o->dn_dif_g_prof[(size_t)hl * ng + jg] = flux_dn[jg] + flux_dn[ng + jg] + flux_dn[2 * ng + jg];
      }
    }
  }
  for (int jg = 0; jg < ng; ++jg) {
    o->dn_diffuse_surf_g[jg] = flux_dn[jg] + flux_dn[ng + jg] + flux_dn[2 * ng + jg];
    o->dn_direct_surf_g[jg] = mu0 * (direct_dn[jg] + direct_dn[ng + jg] + direct_dn[2 * ng + jg]);
    o->dn_diffuse_surf_clear_g[jg] = flux_dn_clear[jg];
    o->dn_direct_surf_clear_g[jg] = mu0 * direct_dn_clear[jg];
  }
  free(fl); free(buf); free(clear); free(reg); free(ods); free(U); free(V);
}

/* radiation_tripleclouds_lw.F90:38-605 for one column (no LW aerosol scattering) */
void orc_tripleclouds_lw(const orc_tables* t, const ecrad_b200_config* cfg, int nlev, const double* frac, const double* fsd,
                         const double* overlap_param, const double* od, const double* planck_hl, const double* od_cloud,
                         const double* ssa_cloud, const double* g_cloud, const double* emission, const double* albedo,
                         orc_tc_out* o) {
  const int ng = NG_LW;
  double (*reg)[NREG] = malloc(sizeof(double[NREG]) * nlev), (*ods)[NREG] = malloc(sizeof(double[NREG]) * nlev);
  double (*U)[NREG][NREG] = malloc(sizeof(double[NREG][NREG]) * (nlev + 1)), (*V)[NREG][NREG] = malloc(sizeof(double[NREG][NREG]) * (nlev + 1));
  orc_region_properties(nlev, frac, fsd, cfg->cloud_fraction_threshold, cfg->i_cloud_pdf_shape == ECRAD_PDF_LOGNORMAL, reg, ods);
  orc_overlap_matrices(nlev, reg, overlap_param, cfg->cloud_inhom_decorr_scaling, cfg->cloud_fraction_threshold, cfg->use_beta_overlap, U, V, &o->cloud_cover);
  int* clear = calloc(nlev + 2, sizeof(int));
  clear[0] = 1; clear[nlev + 1] = 1;
  int i_cloud_top = nlev + 1;
  for (int jl = 1; jl <= nlev; ++jl) {
    clear[jl] = !(frac[jl - 1] > 0.0);
    if (!clear[jl] && i_cloud_top > jl) i_cloud_top = jl;
  }
  const size_t nl = (size_t)nlev * ng, nl1 = (size_t)(nlev + 1) * ng;
  double* buf = calloc(3 * nl + 2 * nl1 + 4 * nl * NREG + 2 * nl1 * NREG + 12 * (size_t)ng * NREG, sizeof(double));
  double *tr_c = buf, *su_c = tr_c + nl, *sd_c = su_c + nl, *fu_c = sd_c + nl, *fd_c = fu_c + nl1;
  double *ref = fd_c + nl1, *tr = ref + nl * NREG, *su = tr + nl * NREG, *sd = su + nl * NREG;
  double *talb = sd + nl * NREG, *tsrc = talb + nl1 * NREG;
  double *below = tsrc + nl1 * NREG, *src_below = below + NREG * ng, *flux_up = src_below + NREG * ng, *flux_dn = flux_up + NREG * ng;
  double *tmp = flux_dn + NREG * ng, *od_total = tmp + NREG * ng, *ssa_total = od_total + ng, *g_total = ssa_total + ng;
  double *lw_deriv = g_total + ng, *lw_deriv_below = lw_deriv + NREG * ng;
  orc_calc_no_scattering_transmittance_lw(ng * nlev, od, planck_hl, planck_hl + ng, tr_c, su_c, sd_c);
  orc_calc_fluxes_no_scattering_lw(ng, nlev, tr_c, su_c, sd_c, emission, albedo, fu_c, fd_c);
  for (int hl = 0; hl <= nlev; ++hl) {
    double s_up = 0.0, s_dn = 0.0;
    for (int jg = 0; jg < ng; ++jg) { s_up = s_up + fu_c[(size_t)hl * ng + jg]; s_dn = s_dn + fd_c[(size_t)hl * ng + jg]; }
    o->up_clear[hl] = s_up; o->dn_clear[hl] = s_dn;
  }
  for (int jg = 0; jg < ng; ++jg) { o->dn_diffuse_surf_clear_g[jg] = fd_c[nl1 - ng + jg]; o->up_toa_clear_g[jg] = fu_c[jg]; }
  /* transmittance(:,1,:) = trans_clear; regions 2: = 1 down to cloud top */
  for (int l = 0; l < nlev; ++l)
    for (int jg = 0; jg < ng; ++jg) { A3(tr, l, 0, jg) = A2L(tr_c, l, jg); A3(tr, l, 1, jg) = 1.0; A3(tr, l, 2, jg) = 1.0; }
  for (int jl = i_cloud_top; jl <= nlev; ++jl) {
    const int l = jl - 1;
    for (int jg = 0; jg < ng; ++jg) { A3(ref, l, 0, jg) = 0.0; A3(su, l, 0, jg) = A2L(su_c, l, jg); A3(sd, l, 0, jg) = A2L(sd_c, l, jg); }
    if (clear[jl]) {
      for (int jr = 1; jr < NREG; ++jr)
        for (int jg = 0; jg < ng; ++jg) { A3(ref, l, jr, jg) = 0.0; A3(tr, l, jr, jg) = 1.0; A3(su, l, jr, jg) = 0.0; A3(sd, l, jr, jg) = 0.0; }
    } else {
      for (int jr = 1; jr < NREG; ++jr) {
        for (int jg = 0; jg < ng; ++jg) {
          const int ib = t->band_lw[jg];
          double od_cloud_new = od_cloud[l * NB_LW + ib] * ods[l][jr];
          od_total[jg] = A2L(od, l, jg) + od_cloud_new;
          ssa_total[jg] = 0.0; g_total[jg] = 0.0;
          if (cfg->do_lw_cloud_scattering) {
            if (od_total[jg] > 0.0) ssa_total[jg] = ssa_cloud[l * NB_LW + ib] * od_cloud_new / od_total[jg];
            if (ssa_total[jg] > 0.0 && od_total[jg] > 0.0)
              g_total[jg] = g_cloud[l * NB_LW + ib] * ssa_cloud[l * NB_LW + ib] * od_cloud_new / (ssa_total[jg] * od_total[jg]);
          }
        }
        if (cfg->do_lw_cloud_scattering)
          orc_calc_ref_trans_lw(ng, od_total, ssa_total, g_total, planck_hl + (size_t)l * ng, planck_hl + (size_t)(l + 1) * ng,
                                &A3(ref, l, jr, 0), &A3(tr, l, jr, 0), &A3(su, l, jr, 0), &A3(sd, l, jr, 0));
        else {
          orc_calc_no_scattering_transmittance_lw(ng, od_total, planck_hl + (size_t)l * ng, planck_hl + (size_t)(l + 1) * ng,
                                                  &A3(tr, l, jr, 0), &A3(su, l, jr, 0), &A3(sd, l, jr, 0));
          for (int jg = 0; jg < ng; ++jg) A3(ref, l, jr, jg) = 0.0;
        }
      }
      for (int jr = 0; jr < NREG; ++jr)
        for (int jg = 0; jg < ng; ++jg) { A3(su, l, jr, jg) = reg[l][jr] * A3(su, l, jr, jg); A3(sd, l, jr, jg) = reg[l][jr] * A3(sd, l, jr, jg); }
    }
  }
  for (int jr = 0; jr < NREG; ++jr)
    for (int jg = 0; jg < ng; ++jg) { A3(tsrc, nlev, jr, jg) = reg[nlev - 1][jr] * emission[jg]; A3(talb, nlev, jr, jg) = albedo[jg]; }
  for (int jl = nlev; jl >= i_cloud_top; --jl) {
    const int l = jl - 1;
    memset(below, 0, sizeof(double) * 2 * NREG * ng);
    const int nr = clear[jl] ? 1 : NREG;
    for (int jr = 0; jr < nr; ++jr)
      for (int jg = 0; jg < ng; ++jg) {
        double id = 1.0 / (1.0 - A3(talb, jl, jr, jg) * A3(ref, l, jr, jg));
        below[jr * ng + jg] = A3(ref, l, jr, jg) + A3(tr, l, jr, jg) * A3(tr, l, jr, jg) * A3(talb, jl, jr, jg) * id;
        src_below[jr * ng + jg] = A3(su, l, jr, jg) + A3(tr, l, jr, jg) * (A3(tsrc, jl, jr, jg) + A3(talb, jl, jr, jg) * A3(sd, l, jr, jg)) * id;
      }
    if (clear[jl] && clear[jl - 1]) {
      for (int jr = 0; jr < NREG; ++jr)
        for (int jg = 0; jg < ng; ++jg) { A3(talb, l, jr, jg) = below[jr * ng + jg]; A3(tsrc, l, jr, jg) = src_below[jr * ng + jg]; }
    } else {
      for (int j1 = 0; j1 < NREG; ++j1) {
        for (int jg = 0; jg < ng; ++jg) A3(tsrc, l, j1, jg) = 0.0;
        for (int j2 = 0; j2 < NREG; ++j2)
          for (int jg = 0; jg < ng; ++jg) A3(tsrc, l, j1, jg) = A3(tsrc, l, j1, jg) + U[l][j1][j2] * src_below[j2 * ng + jg];
      }
      for (int jr = 0; jr < NREG; ++jr)
        for (int jr2 = 0; jr2 < NREG; ++jr2)
          for (int jg = 0; jg < ng; ++jg) A3(talb, l, jr, jg) = A3(talb, l, jr, jg) + below[jr2 * ng + jg] * V[l][jr2][jr];
    }
  }
  const int ict = i_cloud_top - 1;   /* 0-based half-level of cloud top */
  for (int hl = 0; hl <= ict && hl <= nlev; ++hl) o->dn[hl] = o->dn_clear[hl];
  for (int jg = 0; jg < ng; ++jg) {
    flux_up[jg] = A3(tsrc, ict, 0, jg) + A3(talb, ict, 0, jg) * fd_c[(size_t)ict * ng + jg];
    flux_up[ng + jg] = 0.0; flux_up[2 * ng + jg] = 0.0;
  }
  {
    double s = 0.0;
    for (int jg = 0; jg < ng; ++jg) s = s + flux_up[jg];
    o->up[ict] = s;
    if (o->up_g_prof) for (int jg = 0; jg < ng; ++jg) o->up_g_prof[(size_t)ict * ng + jg] = flux_up[jg];
  }
  for (int hl = ict - 1; hl >= 0; --hl) {
    double s = 0.0;
    for (int jg = 0; jg < ng; ++jg) { flux_up[jg] = A2L(tr_c, hl, jg) * flux_up[jg] + A2L(su_c, hl, jg); s = s + flux_up[jg]; }
    o->up[hl] = s;
    if (o->up_g_prof) for (int jg = 0; jg < ng; ++jg) o->up_g_prof[(size_t)hl * ng + jg] = flux_up[jg];
  }
  for (int jg = 0; jg < ng; ++jg) o->up_toa_g[jg] = flux_up[jg] + flux_up[ng + jg] + flux_up[2 * ng + jg];
  if (o->dn_dif_g_prof) for (int hl = 0; hl <= ict && hl <= nlev; ++hl) for (int jg = 0; jg < ng; ++jg) o->dn_dif_g_prof[(size_t)hl * ng + jg] = fd_c[(size_t)hl * ng + jg];
  for (int jr = 0; jr < NREG; ++jr)
    for (int jg = 0; jg < ng; ++jg) flux_dn[jr * ng + jg] = V[ict][jr][0] * fd_c[(size_t)ict * ng + jg];
  for (int jl = i_cloud_top; jl <= nlev; ++jl) {
    const int l = jl - 1;
    const int nr = clear[jl] ? 1 : NREG;
    for (int jr = 0; jr < nr; ++jr)
      for (int jg = 0; jg < ng; ++jg) {
        const size_t k = (size_t)jr * ng + jg;
        flux_dn[k] = (A3(tr, l, jr, jg) * flux_dn[k] + A3(ref, l, jr, jg) * A3(tsrc, jl, jr, jg) + A3(sd, l, jr, jg)) /
                     (1.0 - A3(ref, l, jr, jg) * A3(talb, jl, jr, jg));
        flux_up[k] = A3(tsrc, jl, jr, jg) + flux_dn[k] * A3(talb, jl, jr, jg);
      }
    if (clear[jl])
      for (int jr = 1; jr < NREG; ++jr)
        for (int jg = 0; jg < ng; ++jg) { flux_dn[jr * ng + jg] = 0.0; flux_up[jr * ng + jg] = 0.0; }
    if (!(clear[jl] && clear[jl + 1])) {
      memset(tmp, 0, sizeof(double) * NREG * ng);
      for (int j1 = 0; j1 < NREG; ++j1)
        for (int j2 = 0; j2 < NREG; ++j2)
          for (int jg = 0; jg < ng; ++jg) tmp[j1 * ng + jg] = tmp[j1 * ng + jg] + V[jl][j1][j2] * flux_dn[j2 * ng + jg];
      memcpy(flux_dn, tmp, sizeof(double) * NREG * ng);
    }
    double s_up = 0.0, s_dn = 0.0;
    for (int jr = 0; jr < NREG; ++jr)
      for (int jg = 0; jg < ng; ++jg) { s_up = s_up + flux_up[jr * ng + jg]; s_dn = s_dn + flux_dn[jr * ng + jg]; }
    o->up[jl] = s_up; o->dn[jl] = s_dn;
    if (o->up_g_prof)
      for (int jg = 0; jg < ng; ++jg) {
        o->up_g_prof[(size_t)jl * ng + jg] = flux_up[jg] + flux_up[ng + jg] + flux_up[2 * ng + jg];
        o->dn_dif_g_prof[(size_t)jl * ng + jg] = flux_dn[jg] + flux_dn[ng + jg] + flux_dn[2 * ng + jg];
      }
  }
  for (int jg = 0; jg < ng; ++jg) o->dn_diffuse_surf_g[jg] = flux_dn[jg] + flux_dn[ng + jg] + flux_dn[2 * ng + jg];
  /* calc_lw_derivatives_region, radiation_lw_derivatives.F90:200-290 (nreg = 3) */
  if (o->lw_deriv) {
    double s = 0.0;
    for (int jg = 0; jg < ng; ++jg) s = s + (flux_up[jg] + flux_up[ng + jg] + flux_up[2 * ng + jg]);
    for (int jg = 0; jg < ng; ++jg) {
      lw_deriv[jg] = (flux_up[jg] + flux_up[ng + jg] + flux_up[2 * ng + jg]) / s;
      lw_deriv[ng + jg] = 0.0; lw_deriv[2 * ng + jg] = 0.0;
    }
    o->lw_deriv[nlev] = 1.0;
    for (int jl = nlev; jl >= 1; --jl) {
      const int l = jl - 1;
      memcpy(lw_deriv_below, lw_deriv, sizeof(double) * NREG * ng);
      double tot = 0.0;
      for (int jg = 0; jg < ng; ++jg) {
        for (int r = 0; r < NREG; ++r)
          lw_deriv[r * ng + jg] = U[jl][r][0] * lw_deriv_below[jg] + U[jl][r][1] * lw_deriv_below[ng + jg] + U[jl][r][2] * lw_deriv_below[2 * ng + jg];
        for (int r = 0; r < NREG; ++r) lw_deriv[r * ng + jg] = lw_deriv[r * ng + jg] * A3(tr, l, r, jg);
        tmp[jg] = lw_deriv[jg] + lw_deriv[ng + jg] + lw_deriv[2 * ng + jg];
      }
      for (int jg = 0; jg < ng; ++jg) tot = tot + tmp[jg];
      o->lw_deriv[l] = tot;
    }
  }
  free(buf); free(clear); free(reg); free(ods); free(U); free(V);
}
