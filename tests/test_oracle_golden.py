"""Pins the CPU oracle (oracle/*.c) to the reference's own golden outputs.

Golden files: test/ifs/ecrad_meridian_{noaer,cloudless}_out_REFERENCE.nc of the reference (float32, written by
the unmodified Fortran driver; copied by tools/make_golden_fixtures.py).  The reference's own acceptance
thresholds are LW 1e-3 / SW 1e-1 W m-2 (test/ifs/CMakeLists.txt:18-19); the oracle is held to *float32 rounding*
of the golden values, i.e. it reproduces every stored number to within float32 rounding (tolerance 0.51 ulp_f32: exact rounding except for near-ties).
"""
import numpy as np
import pytest

from ecrad_b200 import inputs as I
from ecrad_b200.config import RadiationConfig
from oracle_lib import Oracle

# oracle output name -> golden variable
PROFILES = {
    "lw_up": "flux_up_lw", "lw_dn": "flux_dn_lw", "lw_up_clear": "flux_up_lw_clear", "lw_dn_clear": "flux_dn_lw_clear",
    "sw_up": "flux_up_sw", "sw_dn": "flux_dn_sw", "sw_dn_direct": "flux_dn_direct_sw",
    "sw_up_clear": "flux_up_sw_clear", "sw_dn_clear": "flux_dn_sw_clear", "sw_dn_direct_clear": "flux_dn_direct_sw_clear",
}


def f32_ulp_err(a, g):
    """|a - g| in units of the float32 spacing at g."""
    g64 = g.astype(np.float64)
    return np.abs(a - g64) / np.maximum(np.spacing(np.abs(g).astype(np.float32)).astype(np.float64), 1e-30)


def _run(raw, spectral=False, **cfgkw):
    cfg = RadiationConfig(**cfgkw).consolidate()
    inp = I.to_radiation_inputs(raw)
    ncol, nlev = inp["pressure_hl"].shape[0], inp["pressure_hl"].shape[1] - 1
    out = Oracle(cfg).radiation(inp, ncol, nlev, spectral_profiles=spectral or cfg.sw_solver_name == "Cloudless")
    return cfg, out


def test_oracle_mcica_noaer_matches_reference_golden(meridian_raw, golden_noaer):
    _, out = _run(meridian_raw)
    for nm, gname in PROFILES.items():
        err = f32_ulp_err(out[nm], golden_noaer[gname])
        assert err.max() <= 0.51, (nm, err.max())
    # McICA RNG / cloud generator pin: total cloud cover per column
    for nm in ("cloud_cover_lw", "cloud_cover_sw"):
        assert f32_ulp_err(out[nm], golden_noaer[nm]).max() <= 0.51, nm
    assert f32_ulp_err(out["lw_derivatives"], golden_noaer["lw_derivative"]).max() <= 0.51
    # surface spectral / canopy fluxes (radiation_flux.F90 calc_surface_spectral); golden is (col, band)
    for nm, gname in (("sw_dn_surf_band", "spectral_flux_dn_sw_surf"), ("sw_dn_direct_surf_band", "spectral_flux_dn_direct_sw_surf"),
                      ("sw_dn_surf_clear_band", "spectral_flux_dn_sw_surf_clear"),
                      ("sw_dn_direct_surf_clear_band", "spectral_flux_dn_direct_sw_surf_clear"),
                      ("sw_dn_diffuse_surf_canopy", "canopy_flux_dn_diffuse_sw_surf"),
                      ("sw_dn_direct_surf_canopy", "canopy_flux_dn_direct_sw_surf"), ("lw_dn_surf_canopy", "canopy_flux_dn_lw_surf")):
        g = golden_noaer[gname]
        a = out[nm].T
        # canopy diffuse is a difference of two O(100) numbers: allow 2 ulp of the larger operand
        tol = 1.0e-4 if "canopy" in nm else None
        if tol is None:
            assert f32_ulp_err(a, g).max() <= 0.51, nm
        else:
            assert np.abs(a - g).max() <= tol, (nm, np.abs(a - g).max())


def test_oracle_mcica_with_aerosols_matches_reference_default_golden(meridian_raw, golden_default):
    """test/ifs `default` ctest: McICA + RRTMG + general aerosol optics (12 IFS aerosol types); also pins the setup-time
    spectral averaging restated in tools/extract_rrtmg_tables.py (aerosol_tables)."""
    _, out = _run(meridian_raw, use_aerosols=True)
    for nm, gname in PROFILES.items():
        err = f32_ulp_err(out[nm], golden_default[gname])
        assert err.max() <= 0.51, (nm, err.max())
    for nm in ("cloud_cover_lw", "cloud_cover_sw"):
        assert f32_ulp_err(out[nm], golden_default[nm]).max() <= 0.51, nm
    assert f32_ulp_err(out["lw_derivatives"], golden_default["lw_derivative"]).max() <= 0.51
    for nm, gname in (("sw_dn_surf_band", "spectral_flux_dn_sw_surf"), ("sw_dn_direct_surf_band", "spectral_flux_dn_direct_sw_surf")):
        assert f32_ulp_err(out[nm].T, golden_default[gname]).max() <= 0.51, nm


def test_oracle_expexp_matches_reference_golden(meridian_raw, golden_expexp):
    """test/ifs `expexp` ctest: as `default` with overlap_scheme_name='Exp-Exp' (cum_cloud_cover_exp_exp + generate_column_exp_exp)."""
    _, out = _run(meridian_raw, use_aerosols=True, overlap_scheme_name="Exp-Exp")
    for nm, gname in PROFILES.items():
        err = f32_ulp_err(out[nm], golden_expexp[gname])
        assert err.max() <= 0.51, (nm, err.max())
    for nm in ("cloud_cover_lw", "cloud_cover_sw"):
        assert f32_ulp_err(out[nm], golden_expexp[nm]).max() <= 0.51, nm


def test_oracle_tripleclouds_matches_reference_golden(meridian_raw, golden_tripleclouds):
    """test/ifs `tripleclouds` ctest: Tripleclouds LW+SW (3 regions, gamma PDF) + RRTMG + aerosols, incl. per-band profiles."""
    _, out = _run(meridian_raw, use_aerosols=True, sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds", spectral=True)
    for nm, gname in PROFILES.items():
        err = f32_ulp_err(out[nm], golden_tripleclouds[gname])
        assert err.max() <= 0.51, (nm, err.max())
    for nm in ("cloud_cover_lw", "cloud_cover_sw"):
        assert f32_ulp_err(out[nm], golden_tripleclouds[nm]).max() <= 0.51, nm
    assert f32_ulp_err(out["lw_derivatives"], golden_tripleclouds["lw_derivative"]).max() <= 0.51
    lev = golden_tripleclouds["band_levels"]
    for nm, gname in (("lw_up_band", "spectral_flux_up_lw"), ("lw_dn_band", "spectral_flux_dn_lw"), ("sw_up_band", "spectral_flux_up_sw"),
                      ("sw_dn_band", "spectral_flux_dn_sw"), ("sw_dn_direct_band", "spectral_flux_dn_direct_sw")):
        a = np.transpose(out[nm], (1, 2, 0))[:, lev, :]
        assert f32_ulp_err(a, golden_tripleclouds[gname]).max() <= 0.51, nm


def test_oracle_cloudless_matches_reference_golden(meridian_raw, golden_cloudless):
    _, out = _run(meridian_raw, sw_solver_name="Cloudless", lw_solver_name="Cloudless")
    for nm, gname in PROFILES.items():
        err = f32_ulp_err(out[nm], golden_cloudless[gname])
        assert err.max() <= 0.51, (nm, err.max())
    lev = golden_cloudless["band_levels"]
    # per-band profiles pin each taumol band separately: oracle (nband, ncol, nlev+1) vs golden (ncol, levels, nband)
    for nm, gname in (("lw_up_band", "spectral_flux_up_lw"), ("lw_dn_band", "spectral_flux_dn_lw"),
                      ("sw_up_band", "spectral_flux_up_sw"), ("sw_dn_band", "spectral_flux_dn_sw"),
                      ("sw_dn_direct_band", "spectral_flux_dn_direct_sw")):
        a = np.transpose(out[nm], (1, 2, 0))[:, lev, :]
        err = f32_ulp_err(a, golden_cloudless[gname])
        assert err.max() <= 0.51, (nm, err.max())


ECCKD = dict(gas_model_name="ECCKD", use_aerosols=True, do_nearest_spectral_lw_emiss=False)   # test/ifs/configCY49R1_ecckd.nam
CANOPY = (("lw_dn_surf_canopy", "canopy_flux_dn_lw_surf"), ("sw_dn_diffuse_surf_canopy", "canopy_flux_dn_diffuse_sw_surf"),
          ("sw_dn_direct_surf_canopy", "canopy_flux_dn_direct_sw_surf"))


def _run_ecckd(raw, spectral=False, **kw):
    cfg = RadiationConfig(**ECCKD, **kw).consolidate()
    return cfg, Oracle(cfg).radiation(I.to_radiation_inputs(raw, cfg), 32, 137, spectral_profiles=spectral)


def test_oracle_ecckd_tripleclouds_matches_reference_golden(meridian_raw, golden_ecckd_tc):
    """test/ifs `ecckd_tc` ctest: ecCKD 32-term gas optics (LW fsck-32b, SW rgb-32b) + generalised cloud optics (thick
    averaging) + generalised aerosol optics, all per g-point, weighted albedo/emissivity intervals, Tripleclouds."""
    cfg, out = _run_ecckd(meridian_raw, spectral=True, sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds")
    assert (cfg.n_g, cfg.n_bands) == ((32, 32), (32, 32))
    for nm, gname in PROFILES.items():
        assert f32_ulp_err(out[nm], golden_ecckd_tc[gname]).max() <= 0.51, nm
    for nm in ("cloud_cover_lw", "cloud_cover_sw"):
        assert f32_ulp_err(out[nm], golden_ecckd_tc[nm]).max() <= 0.51, nm
    assert f32_ulp_err(out["lw_derivatives"], golden_ecckd_tc["lw_derivative"]).max() <= 0.51
    for nm, gname in CANOPY:
        assert f32_ulp_err(out[nm].T, golden_ecckd_tc[gname]).max() <= 0.51, nm
    lev = golden_ecckd_tc["band_levels"]   # per-g-point profiles pin every k-term separately
    for nm, gname in (("lw_up_band", "spectral_flux_up_lw"), ("lw_dn_band", "spectral_flux_dn_lw"), ("sw_up_band", "spectral_flux_up_sw"),
                      ("sw_dn_band", "spectral_flux_dn_sw"), ("sw_dn_direct_band", "spectral_flux_dn_direct_sw")):
        a = np.transpose(out[nm], (1, 2, 0))[:, lev, :]
        assert f32_ulp_err(a, golden_ecckd_tc[gname]).max() <= 0.51, nm


def test_oracle_ecckd_mcica_matches_reference_golden(meridian_raw, golden_ecckd_mcica):
    """test/ifs `ecckd_mcica` ctest: same optics, McICA solvers (32 g-point streams per spectrum)."""
    _, out = _run_ecckd(meridian_raw, sw_solver_name="McICA", lw_solver_name="McICA")
    for nm, gname in PROFILES.items():
        assert f32_ulp_err(out[nm], golden_ecckd_mcica[gname]).max() <= 0.51, nm
    for nm in ("cloud_cover_lw", "cloud_cover_sw"):
        assert f32_ulp_err(out[nm], golden_ecckd_mcica[nm]).max() <= 0.51, nm
    assert f32_ulp_err(out["lw_derivatives"], golden_ecckd_mcica["lw_derivative"]).max() <= 0.51
    for nm, gname in CANOPY + (("sw_dn_surf_band", "spectral_flux_dn_sw_surf"), ("sw_dn_direct_surf_band", "spectral_flux_dn_direct_sw_surf"),
                               ("sw_dn_surf_clear_band", "spectral_flux_dn_sw_surf_clear")):
        assert f32_ulp_err(out[nm].T, golden_ecckd_mcica[gname]).max() <= 0.51, nm


def test_oracle_unpinned_ecckd_models_agree_with_the_pinned_one(meridian_raw):
    """The 64-term pair (BASELINE configs[2]) and the 96-term SW model have no golden file (SURVEY section 8c gaps): same code,
    other tables.  Cross-model check: their clear-sky fluxes agree with the pinned 32-term models within 1.5 W m-2 (and all of
    them with RRTMG within 6 W m-2, the known inter-model spread)."""
    out = {}
    for tabs in ("ecckd_tables_32b.bin", "ecckd_tables_64b.bin", "ecckd_tables_lw32_sw96.bin"):
        cfg = RadiationConfig(gas_model_name="ECCKD", do_nearest_spectral_lw_emiss=False, ecckd_tables=tabs,
                              sw_solver_name="Cloudless", lw_solver_name="Cloudless").consolidate()
        out[tabs] = Oracle(cfg).radiation(I.to_radiation_inputs(meridian_raw, cfg), 32, 137)
    cfg = RadiationConfig(sw_solver_name="Cloudless", lw_solver_name="Cloudless").consolidate()
    rrtmg = Oracle(cfg).radiation(I.to_radiation_inputs(meridian_raw), 32, 137)
    ref = out["ecckd_tables_32b.bin"]
    for tabs, o in out.items():
        for nm, lev in (("sw_dn", -1), ("sw_up", 0), ("lw_up", 0), ("lw_dn", -1)):
            assert np.abs(o[nm][:, lev] - ref[nm][:, lev]).max() < 1.5, (tabs, nm)
            assert np.abs(o[nm][:, lev] - rrtmg[nm][:, lev]).max() < 6.0, (tabs, nm)


def test_oracle_vectorizable_generator_is_statistically_consistent(meridian_raw):
    """use_vectorizable_generator (ctest `vec`) has no reference output: it draws other random numbers (vector MINSTD streams,
    radiation_random_numbers.F90), so it can only agree with the default generator in the mean.  Same cloud cover and clear-sky
    fluxes exactly; all-sky fluxes differ by McICA noise with a mean well below 1 W m-2 over 600 columns."""
    n = 600
    raw = I.synthetic_columns(meridian_raw, n)
    out = {}
    for vec in (False, True):
        cfg = RadiationConfig(use_vectorizable_generator=vec).consolidate()
        out[vec] = Oracle(cfg).radiation(I.to_radiation_inputs(raw), n, 137)
    assert np.array_equal(out[0]["cloud_cover_sw"], out[1]["cloud_cover_sw"])
    assert np.array_equal(out[0]["sw_up_clear"], out[1]["sw_up_clear"]) and np.array_equal(out[0]["lw_dn_clear"], out[1]["lw_dn_clear"])
    for nm in ("sw_up", "sw_dn", "lw_up", "lw_dn"):
        d = out[1][nm] - out[0][nm]
        assert np.abs(d).max() > 1.0, nm                      # it really is another sample of sub-columns
        assert abs(d.mean()) < 1.0 and np.sqrt((d ** 2).mean()) < 30.0, (nm, d.mean())


def test_oracle_crop_cloud_fraction_side_effect(meridian_raw):
    """radiation_cloud.F90:700-740: fraction below threshold (or with negligible water) is zeroed in the caller's array."""
    cfg = RadiationConfig().consolidate()
    inp = I.to_radiation_inputs(meridian_raw)
    before = inp["cloud_fraction"].copy()
    out = Oracle(cfg).radiation(inp, 32, 137)
    after = out["cloud_fraction"]
    qsum = inp["q_liq"] + inp["q_ice"]
    expect = np.where((before < cfg.cloud_fraction_threshold) | (qsum < cfg.cloud_mixing_ratio_threshold), 0.0, before)
    assert np.array_equal(after, expect)


@pytest.mark.parametrize("first,n", [(0, 40), (4090, 12)])
def test_synthetic_columns_are_shard_invariant(meridian_raw, first, n):
    """bench inputs: any shard [first, first+n) of the global synthetic problem sees identical columns."""
    a = I.synthetic_columns(meridian_raw, n, first=first)
    b = I.synthetic_columns(meridian_raw, n + 7, first=first - min(first, 3))
    off = min(first, 3)
    for k in ("temperature_hl", "q", "cloud_fraction", "cos_solar_zenith_angle", "iseed"):
        assert np.array_equal(a[k], b[k][off:off + n]), k
