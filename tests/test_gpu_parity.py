"""GPU parity tests: the CUDA path, called through the C-ABI (libecrad_b200.so), against the CPU oracle on the same
seeded inputs, against the reference's golden files, and at BASELINE size through size-independent properties.

Tolerance (BASELINE.json north_star): |flux - reference| <= 1e-6 W m-2 in double precision, on every flux component;
integer / decision work (McICA cloud masks -> total cloud cover, cropped cloud fraction) must be bit-exact.
"""
import os

import numpy as np
import pytest

from ecrad_b200 import inputs as I
from ecrad_b200.config import RadiationConfig

pytestmark = pytest.mark.gpu

TOL = 1.0e-6  # W m-2
NLEV = 137
FLUXES = ["lw_up", "lw_dn", "lw_up_clear", "lw_dn_clear", "sw_up", "sw_dn", "sw_dn_direct", "sw_up_clear", "sw_dn_clear",
          "sw_dn_direct_clear"]
OTHERS = ["lw_derivatives", "lw_dn_surf_g", "lw_dn_surf_clear_g", "lw_up_toa_g", "lw_up_toa_clear_g", "sw_dn_diffuse_surf_g",
          "sw_dn_direct_surf_g", "sw_dn_diffuse_surf_clear_g", "sw_dn_direct_surf_clear_g", "sw_up_toa_g", "sw_up_toa_clear_g",
          "sw_dn_surf_band", "sw_dn_direct_surf_band", "sw_dn_surf_clear_band", "sw_dn_direct_surf_clear_band",
          "sw_dn_diffuse_surf_canopy", "sw_dn_direct_surf_canopy", "lw_dn_surf_canopy"]
BANDS = ["lw_up_band", "lw_dn_band", "sw_up_band", "sw_dn_band", "sw_dn_direct_band"]   # flux%*_band(nband, ncol, nlev+1)
BANDS_GOLDEN = {"lw_up_band": "spectral_flux_up_lw", "lw_dn_band": "spectral_flux_dn_lw", "sw_up_band": "spectral_flux_up_sw",
                "sw_dn_band": "spectral_flux_dn_sw", "sw_dn_direct_band": "spectral_flux_dn_direct_sw"}


def check_band_profiles(out, ref, golden):
    """Per-band profiles (do_save_spectral_flux): vs the oracle, vs the golden file's stored half-levels, and band sums = broadband."""
    compare(out, ref, BANDS)
    lev = golden["band_levels"]
    for nm, gname in BANDS_GOLDEN.items():
        a = np.transpose(out[nm], (1, 2, 0))[:, lev, :]   # (nband, ncol, nlev+1) -> (ncol, levels, nband)
        assert f32_ulp_err(a, golden[gname]).max() <= 1.0, nm
    for nm in BANDS:
        assert np.abs(out[nm].sum(axis=0) - out[nm[:-5]]).max() <= 1e-9, nm


@pytest.fixture(scope="module")
def handles():
    from ecrad_b200.radiation_interface import setup_radiation
    from oracle_lib import Oracle

    made = {}

    def get(**kw):
        key = tuple(sorted(kw.items()))
        if key not in made:
            cfg = RadiationConfig(**kw).consolidate()
            made[key] = (setup_radiation(cfg), Oracle(cfg), cfg)
        return made[key]

    yield get
    for h, _, _ in made.values():
        h.finalize()


def compare(out, ref, names, tol=TOL):
    for nm in names:
        a, b = out[nm], ref[nm]
        assert a.shape == b.shape, nm
        m = np.isfinite(b)
        assert np.isfinite(a[m]).all(), f"{nm}: non-finite values"
        err = np.abs(a[m] - b[m]).max() if m.any() else 0.0
        assert err <= tol, f"{nm}: max |gpu - oracle| = {err:.3e} W m-2 > {tol}"


def f32_ulp_err(a, g):
    return np.abs(a - g.astype(np.float64)) / np.maximum(np.spacing(np.abs(g).astype(np.float32)).astype(np.float64), 1e-30)


def test_mcica_meridian_vs_oracle(handles, meridian_raw):
    h, orc, _ = handles()
    out = h.radiation(I.to_radiation_inputs(meridian_raw), 32, NLEV)
    ref = orc.radiation(I.to_radiation_inputs(meridian_raw), 32, NLEV)
    compare(out, ref, FLUXES + OTHERS)
    for nm in ("cloud_cover_lw", "cloud_cover_sw"):  # night columns keep the caller's value (-1) in both
        assert np.array_equal(out[nm], ref[nm]), nm
    assert np.array_equal(out["cloud_fraction"], ref["cloud_fraction"])  # crop_cloud_fraction side effect


def test_mcica_meridian_vs_reference_golden(handles, meridian_raw, golden_noaer):
    """Direct check against the reference's own golden file (float32): within 1 float32 ulp of every stored value."""
    h, _, _ = handles()
    out = h.radiation(I.to_radiation_inputs(meridian_raw), 32, NLEV)
    gmap = {"lw_up": "flux_up_lw", "lw_dn": "flux_dn_lw", "lw_up_clear": "flux_up_lw_clear", "lw_dn_clear": "flux_dn_lw_clear",
            "sw_up": "flux_up_sw", "sw_dn": "flux_dn_sw", "sw_dn_direct": "flux_dn_direct_sw", "sw_up_clear": "flux_up_sw_clear",
            "sw_dn_clear": "flux_dn_sw_clear", "sw_dn_direct_clear": "flux_dn_direct_sw_clear", "lw_derivatives": "lw_derivative",
            "cloud_cover_lw": "cloud_cover_lw", "cloud_cover_sw": "cloud_cover_sw"}
    for nm, gname in gmap.items():
        assert f32_ulp_err(out[nm], golden_noaer[gname]).max() <= 1.0, nm
    # and well inside the reference's own acceptance thresholds (test/ifs/CMakeLists.txt:18-19: LW 1e-3, SW 1e-1)
    for nm, gname in gmap.items():
        assert np.abs(out[nm] - golden_noaer[gname]).max() <= 1.0e-3, nm


def test_mcica_with_aerosols_vs_reference_default_golden(handles, meridian_raw, golden_default):
    """The reference's `default` ctest (McICA + RRTMG + 12 IFS aerosol types): within 1 float32 ulp of its golden file."""
    h, orc, _ = handles(use_aerosols=True)
    out = h.radiation(I.to_radiation_inputs(meridian_raw), 32, NLEV)
    ref = orc.radiation(I.to_radiation_inputs(meridian_raw), 32, NLEV)
    compare(out, ref, FLUXES + OTHERS)
    gmap = {"lw_up": "flux_up_lw", "lw_dn": "flux_dn_lw", "sw_up": "flux_up_sw", "sw_dn": "flux_dn_sw", "sw_dn_direct": "flux_dn_direct_sw",
            "sw_up_clear": "flux_up_sw_clear", "sw_dn_clear": "flux_dn_sw_clear", "lw_up_clear": "flux_up_lw_clear",
            "lw_derivatives": "lw_derivative", "cloud_cover_sw": "cloud_cover_sw"}
    for nm, gname in gmap.items():
        assert f32_ulp_err(out[nm], golden_default[gname]).max() <= 1.0, nm


def test_expexp_vs_reference_golden(handles, meridian_raw, golden_expexp):
    """The reference's `expexp` ctest (Exp-Exp overlap + aerosols): within 1 float32 ulp of its golden file."""
    h, _, _ = handles(use_aerosols=True, overlap_scheme_name="Exp-Exp")
    out = h.radiation(I.to_radiation_inputs(meridian_raw), 32, NLEV)
    for nm, gname in (("lw_up", "flux_up_lw"), ("lw_dn", "flux_dn_lw"), ("sw_up", "flux_up_sw"), ("sw_dn", "flux_dn_sw"),
                      ("sw_dn_direct", "flux_dn_direct_sw"), ("cloud_cover_sw", "cloud_cover_sw"), ("cloud_cover_lw", "cloud_cover_lw")):
        assert f32_ulp_err(out[nm], golden_expexp[gname]).max() <= 1.0, nm


def test_tripleclouds_vs_reference_golden(handles, meridian_raw, golden_tripleclouds):
    """The reference's `tripleclouds` ctest (Tripleclouds LW+SW + RRTMG + aerosols): within 1 float32 ulp of its golden file."""
    h, orc, _ = handles(use_aerosols=True, sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds")
    out = h.radiation(I.to_radiation_inputs(meridian_raw), 32, NLEV, spectral_profiles=True)
    ref = orc.radiation(I.to_radiation_inputs(meridian_raw), 32, NLEV, spectral_profiles=True)
    compare(out, ref, FLUXES + OTHERS)
    check_band_profiles(out, ref, golden_tripleclouds)
    assert np.array_equal(out["cloud_cover_sw"], ref["cloud_cover_sw"]) and np.array_equal(out["cloud_cover_lw"], ref["cloud_cover_lw"])
    for nm, gname in (("lw_up", "flux_up_lw"), ("lw_dn", "flux_dn_lw"), ("sw_up", "flux_up_sw"), ("sw_dn", "flux_dn_sw"),
                      ("sw_dn_direct", "flux_dn_direct_sw"), ("lw_up_clear", "flux_up_lw_clear"), ("sw_dn_clear", "flux_dn_sw_clear"),
                      ("cloud_cover_sw", "cloud_cover_sw"), ("lw_derivatives", "lw_derivative")):
        assert f32_ulp_err(out[nm], golden_tripleclouds[gname]).max() <= 1.0, nm


def test_cloudless_vs_oracle_and_golden(handles, meridian_raw, golden_cloudless):
    h, orc, _ = handles(sw_solver_name="Cloudless", lw_solver_name="Cloudless")
    out = h.radiation(I.to_radiation_inputs(meridian_raw), 32, NLEV, spectral_profiles=True)
    ref = orc.radiation(I.to_radiation_inputs(meridian_raw), 32, NLEV, spectral_profiles=True)
    compare(out, ref, FLUXES + OTHERS)
    check_band_profiles(out, ref, golden_cloudless)
    for nm, gname in (("lw_up", "flux_up_lw"), ("lw_dn", "flux_dn_lw"), ("sw_up", "flux_up_sw"), ("sw_dn", "flux_dn_sw"),
                      ("sw_dn_direct", "flux_dn_direct_sw")):
        assert f32_ulp_err(out[nm], golden_cloudless[gname]).max() <= 1.0, nm


@pytest.mark.parametrize("kw", [dict(), dict(overlap_scheme_name="Max-Ran"), dict(do_lw_cloud_scattering=False),
                                dict(use_beta_overlap=True), dict(use_aerosols=True), dict(overlap_scheme_name="Exp-Exp"),
                                dict(overlap_scheme_name="Exp-Exp", use_beta_overlap=True),
                                dict(sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds"),
                                dict(sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds", use_aerosols=True, do_lw_cloud_scattering=False),
                                dict(use_aerosols=True, sw_solver_name="Cloudless", lw_solver_name="Cloudless"),
                                dict(sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds", use_beta_overlap=True),
                                dict(use_vectorizable_generator=True), dict(use_vectorizable_generator=True, overlap_scheme_name="Max-Ran", use_aerosols=True),
                                dict(sw_solver_name="Homogeneous", lw_solver_name="Homogeneous"),
                                dict(sw_solver_name="Homogeneous", lw_solver_name="Homogeneous", use_aerosols=True, do_lw_cloud_scattering=False),
                                dict(sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds", cloud_pdf_shape_name="Lognormal"),
                                dict(sw_solver_name="SPARTACUS", lw_solver_name="SPARTACUS", do_3d_effects=True, cloud_pdf_shape_name="Lognormal"),
                                dict(do_nearest_spectral_sw_albedo=True), dict(do_nearest_spectral_sw_albedo=True, sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds"),
                                dict(do_nearest_spectral_sw_albedo=True, gas_model_name="ECCKD", do_nearest_spectral_lw_emiss=False),
                                dict(do_nearest_spectral_lw_emiss=False), dict(do_nearest_spectral_lw_emiss=False, sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds"),
                                dict(do_sw_delta_scaling_with_gases=True), dict(do_sw_delta_scaling_with_gases=True, use_aerosols=True),
                                dict(do_sw_delta_scaling_with_gases=True, use_aerosols=True, sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds"),
                                dict(do_sw_delta_scaling_with_gases=True, use_aerosols=True, sw_solver_name="Cloudless", lw_solver_name="Cloudless"),
                                dict(do_sw_delta_scaling_with_gases=True, use_aerosols=True, sw_solver_name="Homogeneous", lw_solver_name="Homogeneous"),
                                # use_general_cloud_optics with RRTMG-IFS (radiation_config.F90:185, :1078-1090): look-up tables per band
                                dict(use_general_cloud_optics=True), dict(use_general_cloud_optics=True, use_aerosols=True, do_lw_cloud_scattering=False),
                                dict(use_general_cloud_optics=True, sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds"),
                                dict(use_general_cloud_optics=True, sw_solver_name="Homogeneous", lw_solver_name="Homogeneous", do_sw_delta_scaling_with_gases=True)])
def test_synthetic_columns_vs_oracle(handles, meridian_raw, kw):
    """600 perturbed columns (BASELINE.md section 4 generator): different cloud profiles, seeds, sun angles."""
    n = 600
    h, orc, cfg = handles(**kw)
    raw = I.synthetic_columns(meridian_raw, n)
    out = h.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)
    ref = orc.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)
    compare(out, ref, FLUXES + OTHERS)
    assert np.array_equal(out["cloud_cover_lw"], ref["cloud_cover_lw"])
    assert np.array_equal(out["cloud_cover_sw"], ref["cloud_cover_sw"])
    assert np.array_equal(out["cloud_fraction"], ref["cloud_fraction"])


MIXED = [dict(), dict(use_aerosols=True), dict(sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds", use_aerosols=True),
         dict(sw_solver_name="SPARTACUS", lw_solver_name="SPARTACUS", do_3d_effects=True),
         dict(sw_solver_name="Homogeneous", lw_solver_name="Homogeneous"), dict(sw_solver_name="Cloudless", lw_solver_name="Cloudless", use_aerosols=True),
         dict(ecckd_tables="ecckd_tables_64b.bin", use_aerosols=True, overlap_scheme_name="Exp-Exp", do_lw_cloud_scattering=False)]


@pytest.mark.parametrize("ckd_spectrum", ["sw", "lw"])
@pytest.mark.parametrize("kw", MIXED)
def test_mixed_gas_models_vs_oracle(handles, meridian_raw, kw, ckd_spectrum):
    """RRTMG-IFS in one spectrum, ecCKD in the other (radiation_interface.F90:333-355; test/ifs `test_mixed_gas` with
    configCY49R1_mixed.nam): mass mixing ratios in, ecCKD scales them itself, cloud optics from the generalised look-up tables per
    band (RRTMG spectrum) and per g-point (ecCKD spectrum).  Each spectrum must also reproduce the run that uses its model twice."""
    n = 300
    E = dict(do_nearest_spectral_lw_emiss=False, **kw)
    h, orc, cfg = handles(**{ckd_spectrum + "_gas_model_name": "ECCKD"}, **E)
    raw = I.synthetic_columns(meridian_raw, n)
    out = h.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)
    ref = orc.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)
    compare(out, ref, FLUXES + OTHERS)
    for nm in ("cloud_cover_lw", "cloud_cover_sw", "cloud_fraction"):
        assert np.array_equal(out[nm], ref[nm]), nm
    nck = 64 if "64b" in str(kw) else 32
    assert (h.cfg.n_g_sw, h.cfg.n_g_lw) == ((nck, 140) if ckd_spectrum == "sw" else (112, nck))
    # the RRTMG spectrum of the mixed run == the same spectrum of the all-RRTMG run with the generalised cloud optics (bit for bit:
    # the same kernels on the same inputs); the ecCKD spectrum == the all-ecCKD run up to the rounding of the unit conversion
    hr, _, cr = handles(use_general_cloud_optics=True, **E)
    hc, _, cc = handles(gas_model_name="ECCKD", **E)
    rr = hr.radiation(I.to_radiation_inputs(raw, cr), n, NLEV)
    ck = hc.radiation(I.to_radiation_inputs(raw, cc), n, NLEV)
    lw, sw = ("lw_up", "lw_dn", "lw_up_clear", "lw_dn_clear"), ("sw_up", "sw_dn", "sw_dn_direct", "sw_up_clear", "sw_dn_clear")
    for nm in (lw if ckd_spectrum == "sw" else sw):
        assert np.array_equal(out[nm], rr[nm]), nm
    for nm in (sw if ckd_spectrum == "sw" else lw):
        m = np.isfinite(ck[nm])
        assert np.abs(out[nm][m] - ck[nm][m]).max() < TOL, nm   # (SPARTACUS amplifies the conversion's rounding to 2e-7)


@pytest.mark.parametrize("kw", [dict(gas_model_name="ECCKD"), dict(sw_gas_model_name="ECCKD", sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds", use_aerosols=True),
                                dict(gas_model_name="ECCKD", ecckd_tables="ecckd_tables_lw32_sw96.bin", sw_solver_name="Cloudless", lw_solver_name="Cloudless")])
def test_spectral_solar_cycle(meridian_raw, kw):
    """single_level%spectral_solar_cycle_multiplier (use_spectral_solar_cycle; calc_incoming_sw, radiation_ecckd.F90:935-964)."""
    from ecrad_b200.radiation_interface import RadiationError, setup_radiation
    from ecrad_b200.tables import read_blob
    from oracle_lib import Oracle

    n = 200
    cfg = RadiationConfig(do_nearest_spectral_lw_emiss=False, **kw).consolidate()
    raw = I.synthetic_columns(meridian_raw, n)
    h, orc = setup_radiation(cfg), Oracle(cfg)
    try:
        base = h.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)
        for mult in (1.0, -0.37):
            h.set_solar_cycle_multiplier(mult); orc.set_solar_cycle_multiplier(mult)
            out = h.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)
            ref = orc.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)
            compare(out, ref, FLUXES + OTHERS)
            assert np.abs(out["sw_dn"] - base["sw_dn"]).max() > 1e-4
            inc = h.radiative_properties(I.to_radiation_inputs(raw, cfg), n, NLEV)["incoming_sw"]
            assert np.abs(inc.sum(axis=0) - float(raw["solar_irradiance"])).max() < 1e-9
        h.set_solar_cycle_multiplier(0.0)
        again = h.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)
        assert np.array_equal(again["sw_dn"], base["sw_dn"]) and np.array_equal(again["sw_up"], base["sw_up"])
    finally:
        h.finalize()
    # no solar-cycle information: RRTMG-IFS in the shortwave, or an ecCKD model registered without the amplitude
    hr = setup_radiation(RadiationConfig().consolidate())
    with pytest.raises(RadiationError, match="solar cycle only available with ecCKD"):
        hr.set_solar_cycle_multiplier(1.0)
    hr.set_solar_cycle_multiplier(0.0)
    hr.finalize()
    arrays = {k: v for k, v in read_blob(cfg.tables_path()).items() if k != "ckd_sw_norm_amplitude_solar_irradiance"}
    hn = setup_radiation(cfg, tables_arrays=arrays)
    with pytest.raises(RadiationError, match="no information present on solar cycle"):
        hn.set_solar_cycle_multiplier(-1.0)
    hn.finalize()


@pytest.mark.parametrize("kw", [dict(), dict(sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds", do_save_spectral_flux=True),
                                dict(sw_solver_name="SPARTACUS", lw_solver_name="SPARTACUS", do_3d_effects=True)])
def test_do_sw_direct_false(handles, meridian_raw, kw):
    """config%do_sw_direct = false (radiation_flux.F90:208-240): sw_dn_direct, sw_dn_direct_clear and sw_dn_direct_band do not exist for the
    solvers (they test allocated()); every other output is what it is with the direct arrays."""
    n = 100
    raw = I.synthetic_columns(meridian_raw, n)
    h1, _, cfg1 = handles(**kw)
    h0, _, cfg0 = handles(do_sw_direct=False, **kw)
    prof = bool(kw.get("do_save_spectral_flux"))
    a = h1.radiation(I.to_radiation_inputs(raw, cfg1), n, NLEV, spectral_profiles=prof)
    b = h0.radiation(I.to_radiation_inputs(raw, cfg0), n, NLEV, spectral_profiles=prof)
    for nm in a:
        if a[nm] is None:
            continue
        if nm in ("sw_dn_direct", "sw_dn_direct_clear", "sw_dn_direct_band"):
            assert np.isnan(b[nm]).all(), nm           # alloc_outputs fills with NaN: left untouched
            assert np.isfinite(a[nm]).any(), nm
        else:
            assert np.array_equal(a[nm], b[nm], equal_nan=True), nm


def test_column_range_and_untouched_columns(handles, meridian_raw):
    """istartcol/iendcol semantics of radiation(): only that range is written (1-based inclusive)."""
    h, orc, _ = handles()
    n = 64
    raw = I.synthetic_columns(meridian_raw, n)
    out = h.radiation(I.to_radiation_inputs(raw), n, NLEV, istartcol=11, iendcol=40)
    ref = orc.radiation(I.to_radiation_inputs(raw), n, NLEV, istartcol=11, iendcol=40)
    for nm in FLUXES:
        assert np.isnan(out[nm][:10]).all() and np.isnan(out[nm][40:]).all(), nm  # alloc_outputs fills with NaN
        assert np.abs(out[nm][10:40] - ref[nm][10:40]).max() <= TOL, nm
    assert np.isnan(out["sw_up_toa_g"][:, :10]).all() and np.isnan(out["sw_up_toa_g"][:, 40:]).all()
    # single column, first and last
    for j in (1, n):
        o1 = h.radiation(I.to_radiation_inputs(raw), n, NLEV, istartcol=j, iendcol=j)
        r1 = orc.radiation(I.to_radiation_inputs(raw), n, NLEV, istartcol=j, iendcol=j)
        for nm in FLUXES:
            assert np.abs(o1[nm][j - 1] - r1[nm][j - 1]).max() <= TOL


def test_tiling_is_invisible(meridian_raw):
    """Results do not depend on the internal column tile (ragged last tile included)."""
    from ecrad_b200.radiation_interface import setup_radiation

    n = 333
    raw = I.synthetic_columns(meridian_raw, n)
    cfg = RadiationConfig().consolidate()
    outs = []
    for tile in ("4096", "100", "7"):
        os.environ["ECRAD_B200_TILE"] = tile
        try:
            h = setup_radiation(cfg)
        finally:
            del os.environ["ECRAD_B200_TILE"]
        outs.append(h.radiation(I.to_radiation_inputs(raw), n, NLEV))
        h.finalize()
    # the host entry's schedule options (first / last tile, several short tiles at the end, doubling tile sizes)
    h = setup_radiation(cfg)
    for opts in (dict(tile_cols=128, edge_cols=64, tail_tiles=3), dict(tile_cols=256, edge_cols=64, tile_ramp=1), dict(tile_cols=64, edge_cols=64)):
        for k, v in opts.items():
            h.set_option(k, v)
        outs.append(h.radiation(I.to_radiation_inputs(raw), n, NLEV))
    h.finalize()
    for nm in FLUXES + OTHERS + ["cloud_cover_sw", "cloud_cover_lw", "cloud_fraction"]:
        for o in outs[1:]:
            assert np.array_equal(outs[0][nm], o[nm], equal_nan=True), nm


ECCKD = dict(gas_model_name="ECCKD", do_nearest_spectral_lw_emiss=False)   # test/ifs/configCY49R1_ecckd.nam
CANOPY = (("lw_dn_surf_canopy", "canopy_flux_dn_lw_surf"), ("sw_dn_diffuse_surf_canopy", "canopy_flux_dn_diffuse_sw_surf"),
          ("sw_dn_direct_surf_canopy", "canopy_flux_dn_direct_sw_surf"))
GOLDEN_PROFILES = {"lw_up": "flux_up_lw", "lw_dn": "flux_dn_lw", "lw_up_clear": "flux_up_lw_clear", "lw_dn_clear": "flux_dn_lw_clear",
                   "sw_up": "flux_up_sw", "sw_dn": "flux_dn_sw", "sw_dn_direct": "flux_dn_direct_sw", "sw_up_clear": "flux_up_sw_clear",
                   "sw_dn_clear": "flux_dn_sw_clear", "sw_dn_direct_clear": "flux_dn_direct_sw_clear", "lw_derivatives": "lw_derivative",
                   "cloud_cover_lw": "cloud_cover_lw", "cloud_cover_sw": "cloud_cover_sw"}


def test_ecckd_tripleclouds_vs_reference_golden(handles, meridian_raw, golden_ecckd_tc):
    """The reference's `ecckd_tc` ctest: ecCKD 32-term gas optics, generalised cloud + aerosol optics per g-point,
    Tripleclouds; per-g-point flux profiles included.  Within 1 float32 ulp of the golden file, <= 1e-6 W m-2 of the oracle."""
    h, orc, cfg = handles(**ECCKD, use_aerosols=True, sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds")
    out = h.radiation(I.to_radiation_inputs(meridian_raw, cfg), 32, NLEV, spectral_profiles=True)
    ref = orc.radiation(I.to_radiation_inputs(meridian_raw, cfg), 32, NLEV, spectral_profiles=True)
    compare(out, ref, FLUXES + OTHERS)
    check_band_profiles(out, ref, golden_ecckd_tc)
    for nm, gname in GOLDEN_PROFILES.items():
        assert f32_ulp_err(out[nm], golden_ecckd_tc[gname]).max() <= 1.0, nm
    for nm, gname in CANOPY:
        assert f32_ulp_err(out[nm].T, golden_ecckd_tc[gname]).max() <= 1.0, nm


def test_ecckd_mcica_vs_reference_golden(handles, meridian_raw, golden_ecckd_mcica):
    """The reference's `ecckd_mcica` ctest (McICA over 32 + 32 g-points)."""
    h, orc, cfg = handles(**ECCKD, use_aerosols=True)
    out = h.radiation(I.to_radiation_inputs(meridian_raw, cfg), 32, NLEV)
    ref = orc.radiation(I.to_radiation_inputs(meridian_raw, cfg), 32, NLEV)
    compare(out, ref, FLUXES + OTHERS)
    assert np.array_equal(out["cloud_cover_sw"], ref["cloud_cover_sw"]) and np.array_equal(out["cloud_cover_lw"], ref["cloud_cover_lw"])
    for nm, gname in GOLDEN_PROFILES.items():
        assert f32_ulp_err(out[nm], golden_ecckd_mcica[gname]).max() <= 1.0, nm
    for nm, gname in CANOPY + (("sw_dn_surf_band", "spectral_flux_dn_sw_surf"), ("sw_dn_direct_surf_band", "spectral_flux_dn_direct_sw_surf")):
        assert f32_ulp_err(out[nm].T, golden_ecckd_mcica[gname]).max() <= 1.0, nm


@pytest.mark.parametrize("kw", [dict(), dict(use_aerosols=True), dict(sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds"),
                                dict(sw_solver_name="Cloudless", lw_solver_name="Cloudless", use_aerosols=True),
                                dict(do_lw_cloud_scattering=False, overlap_scheme_name="Exp-Exp"),
                                dict(do_nearest_spectral_lw_emiss=True, overlap_scheme_name="Max-Ran"),
                                dict(ecckd_tables="ecckd_tables_64b.bin", use_aerosols=True), dict(use_vectorizable_generator=True),
                                dict(ecckd_tables="ecckd_tables_64b.bin", sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds"),
                                dict(ecckd_tables="ecckd_tables_lw32_sw96.bin", use_aerosols=True),
                                dict(ecckd_tables="ecckd_tables_lw32_sw96.bin", sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds"),
                                # do_sw_delta_scaling_with_gases on the ecCKD spectra (radiation_mcica_sw.F90:156-180, :274-278 and the other solvers)
                                dict(do_sw_delta_scaling_with_gases=True, use_aerosols=True),
                                dict(do_sw_delta_scaling_with_gases=True, use_aerosols=True, sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds"),
                                dict(do_sw_delta_scaling_with_gases=True, use_aerosols=True, sw_solver_name="Cloudless", lw_solver_name="Cloudless"),
                                dict(do_sw_delta_scaling_with_gases=True, sw_solver_name="Homogeneous", lw_solver_name="Homogeneous", ecckd_tables="ecckd_tables_64b.bin")])
def test_ecckd_synthetic_columns_vs_oracle(handles, meridian_raw, kw):
    """ecCKD configurations (32- and 64-term models; BASELINE configs 1 and 3 have no golden file) on 300 perturbed columns."""
    n = 300
    cfgkw = dict(ECCKD); cfgkw.update(kw)
    h, orc, cfg = handles(**cfgkw)
    raw = I.synthetic_columns(meridian_raw, n)
    out = h.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV, spectral_profiles=True)
    ref = orc.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV, spectral_profiles=True)
    compare(out, ref, FLUXES + OTHERS)
    if cfg.sw_solver_name != "McICA":
        compare(out, ref, BANDS)
    for nm in ("cloud_cover_lw", "cloud_cover_sw", "cloud_fraction"):
        assert np.array_equal(out[nm], ref[nm]), nm


def test_ecckd_tiling_is_invisible(meridian_raw):
    from ecrad_b200.radiation_interface import setup_radiation

    n = 150
    raw = I.synthetic_columns(meridian_raw, n)
    cfg = RadiationConfig(**ECCKD, use_aerosols=True).consolidate()
    outs = []
    for tile in ("4096", "37"):
        os.environ["ECRAD_B200_TILE"] = tile
        try:
            h = setup_radiation(cfg)
        finally:
            del os.environ["ECRAD_B200_TILE"]
        outs.append(h.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV))
        h.finalize()
    for nm in FLUXES + OTHERS + ["cloud_cover_sw", "cloud_cover_lw"]:
        assert np.array_equal(outs[0][nm], outs[1][nm], equal_nan=True), nm


TOA = ["sw_dn_toa_g", "sw_dn_toa_band", "sw_up_toa_band", "sw_up_toa_clear_band", "lw_up_toa_band", "lw_up_toa_clear_band"]


@pytest.mark.parametrize("kw", [dict(), dict(sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds"),
                                dict(sw_solver_name="SPARTACUS", lw_solver_name="SPARTACUS", do_3d_effects=True),
                                dict(gas_model_name="ECCKD", do_nearest_spectral_lw_emiss=False, sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds")])
def test_toa_spectral_fluxes(handles, meridian_raw, kw):
    """flux%calc_toa_spectral (do_toa_spectral_flux): band sums of the per-g-point TOA fluxes; sw_dn_toa_g / sw_dn_toa_band exist for
    the Tripleclouds solver only (the reference's other solvers never set sw_dn_toa_g) and stay untouched otherwise."""
    n = 200
    h, orc, cfg = handles(do_toa_spectral_flux=True, **kw)
    raw = I.synthetic_columns(meridian_raw, n)
    out = h.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)
    ref = orc.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)
    compare(out, ref, FLUXES + OTHERS + TOA)
    tc = cfg.sw_solver_name == "Tripleclouds"
    sun = raw["cos_solar_zenith_angle"] >= 1e-10
    for nm in ("sw_dn_toa_g", "sw_dn_toa_band"):
        assert np.array_equal(np.isnan(out[nm]), np.isnan(ref[nm])), nm
        assert np.isnan(out[nm]).all() != tc, nm
    if tc:
        assert np.abs(out["sw_dn_toa_band"][:, sun].sum(axis=0) - out["sw_dn"][sun, 0]).max() <= 1e-9
    assert np.abs(out["sw_up_toa_band"].sum(axis=0) - out["sw_up"][:, 0]).max() <= 1e-9
    assert np.abs(out["lw_up_toa_band"].sum(axis=0) - out["lw_up"][:, 0]).max() <= 1e-9
    assert np.abs(out["lw_up_toa_clear_band"].sum(axis=0) - out["lw_up_clear"][:, 0]).max() <= 1e-9


@pytest.mark.parametrize("solver", ["Tripleclouds", "Cloudless", "Homogeneous"])
def test_band_profiles_synthetic_and_tiled(handles, meridian_raw, solver):
    """Per-band profiles on perturbed columns (night columns included), and the same through ragged column tiles."""
    from ecrad_b200.radiation_interface import setup_radiation

    n = 150
    raw = I.synthetic_columns(meridian_raw, n)
    h, orc, cfg = handles(sw_solver_name=solver, lw_solver_name=solver, use_aerosols=True)
    out = h.radiation(I.to_radiation_inputs(raw), n, NLEV, spectral_profiles=True)
    ref = orc.radiation(I.to_radiation_inputs(raw), n, NLEV, spectral_profiles=True)
    compare(out, ref, FLUXES + BANDS)
    os.environ["ECRAD_B200_TILE"] = "64"
    try:
        h2 = setup_radiation(cfg)
    finally:
        del os.environ["ECRAD_B200_TILE"]
    out2 = h2.radiation(I.to_radiation_inputs(raw), n, NLEV, spectral_profiles=True)
    h2.finalize()
    for nm in FLUXES + BANDS:
        assert np.array_equal(out[nm], out2[nm]), nm
    # do_save_spectral_flux = false: the band arrays are left untouched
    h3 = setup_radiation(RadiationConfig(sw_solver_name=solver, lw_solver_name=solver, do_save_spectral_flux=False).consolidate())
    out3 = h3.radiation(I.to_radiation_inputs(raw), n, NLEV, spectral_profiles=True)
    h3.finalize()
    for nm in BANDS:
        assert np.isnan(out3[nm]).all(), nm


def coarsen_levels(raw, nlev_new):
    """Column set on fewer model levels: keep nlev_new+1 of the 138 half-levels (TOA and surface included); each new layer takes
    the composition/cloud of the first old layer it spans.  Only the array shapes matter for the test."""
    hl = np.unique(np.round(np.linspace(0, 137, nlev_new + 1)).astype(int))
    assert len(hl) == nlev_new + 1
    lay = hl[:-1]
    out = {}
    for k, v in raw.items():
        if np.ndim(v) < 2:
            out[k] = v
        elif k == "aerosol_mmr":
            out[k] = v[:, :, lay]
        elif v.shape[1] == 138:
            out[k] = v[:, hl]
        elif v.shape[1] == 137:
            out[k] = v[:, lay]
        elif v.shape[1] == 136:
            out[k] = v[:, lay[:-1]]
        else:
            out[k] = v
    return out


@pytest.mark.parametrize("nlev,kw", [(60, dict()), (91, dict(use_aerosols=True, overlap_scheme_name="Exp-Exp")),
                                     (19, dict(sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds")),
                                     (50, dict(gas_model_name="ECCKD", do_nearest_spectral_lw_emiss=False, use_aerosols=True)),
                                     (33, dict(gas_model_name="ECCKD", do_nearest_spectral_lw_emiss=False, sw_solver_name="Tripleclouds",
                                               lw_solver_name="Tripleclouds")),
                                     (45, dict(sw_solver_name="SPARTACUS", lw_solver_name="SPARTACUS", do_3d_effects=True)),
                                     (27, dict(sw_solver_name="Homogeneous", lw_solver_name="Homogeneous"))])
def test_other_level_counts(handles, meridian_raw, nlev, kw):
    """The reference takes any nlev; the kernels stage per-layer state in shared memory sized by nlev (odd counts, not a multiple of 4)."""
    n = 96
    h, orc, cfg = handles(**kw)
    raw = coarsen_levels(I.synthetic_columns(meridian_raw, n), nlev)
    out = h.radiation(I.to_radiation_inputs(raw, cfg), n, nlev)
    ref = orc.radiation(I.to_radiation_inputs(raw, cfg), n, nlev)
    compare(out, ref, FLUXES + OTHERS)
    for nm in ("cloud_cover_lw", "cloud_cover_sw", "cloud_fraction"):
        assert np.array_equal(out[nm], ref[nm]), nm


@pytest.mark.parametrize("kw", [dict(do_sw=False), dict(do_lw=False), dict(do_sw=False, gas_model_name="ECCKD", do_nearest_spectral_lw_emiss=False),
                                dict(do_lw=False, sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds"),
                                dict(do_sw=False, sw_solver_name="SPARTACUS", lw_solver_name="SPARTACUS", do_3d_effects=True),
                                dict(do_lw=False, sw_solver_name="SPARTACUS", lw_solver_name="SPARTACUS", do_3d_effects=True),
                                dict(do_lw=False, sw_solver_name="Homogeneous", lw_solver_name="Homogeneous")])
def test_one_spectrum_only_and_single_column(handles, meridian_raw, kw):
    """do_sw = false / do_lw = false: the other spectrum's outputs stay untouched; ncol = 1 and a 1-column range work."""
    n = 40
    h, orc, cfg = handles(**kw)
    raw = I.synthetic_columns(meridian_raw, n)
    out = h.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)
    ref = orc.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)
    on = [nm for nm in FLUXES + OTHERS if nm.startswith("lw" if cfg.do_lw else "sw")]
    off = [nm for nm in FLUXES if nm.startswith("sw" if cfg.do_lw else "lw")]
    compare(out, ref, on)
    for nm in off:
        assert np.isnan(out[nm]).all(), nm
    one = {k: (v if np.ndim(v) == 0 else v[7:8]) for k, v in raw.items()}
    out1 = h.radiation(I.to_radiation_inputs(one, cfg), 1, NLEV)
    for nm in on:
        a = out1[nm]
        b = out[nm][7:8] if a.shape[0] == 1 else out[nm][:, 7:8]
        assert np.array_equal(a, b, equal_nan=True), nm


def test_all_night_and_all_clear_edge_cases(handles, meridian_raw):
    h, orc, _ = handles()
    n = 40
    raw = I.synthetic_columns(meridian_raw, n)
    raw["cos_solar_zenith_angle"][:] = -0.1           # no sunlit column at all
    raw["cloud_fraction"][: n // 2] = 0.0              # half the columns cloud free
    raw["cloud_fraction"][n // 2:, 100:] = 1.0         # the rest overcast in the lowest layers (MaxCloudFrac branch)
    out = h.radiation(I.to_radiation_inputs(raw), n, NLEV)
    ref = orc.radiation(I.to_radiation_inputs(raw), n, NLEV)
    compare(out, ref, FLUXES + OTHERS)
    assert (out["sw_dn"] == 0.0).all() and (out["cloud_cover_sw"] == -1.0).all()
    assert np.array_equal(out["cloud_cover_lw"], ref["cloud_cover_lw"])


def test_full_size_properties_10000_columns(handles, meridian_raw, golden_noaer):
    """BASELINE config 2 size (10 000 columns): column independence (permutation invariance, bit-exact), the first 32
    columns reproduce the golden file, physical bounds, and an oracle spot check on a random subset."""
    h, orc, _ = handles()
    n = 10000
    raw = I.synthetic_columns(meridian_raw, n)
    out = h.radiation(I.to_radiation_inputs(raw), n, NLEV)
    for nm in FLUXES:
        assert np.isfinite(out[nm]).all(), nm
    # (a) first 32 columns are the unperturbed test slice
    assert f32_ulp_err(out["sw_dn"][:32], golden_noaer["flux_dn_sw"]).max() <= 1.0
    assert f32_ulp_err(out["lw_up"][:32], golden_noaer["flux_up_lw"]).max() <= 1.0
    # (b) permutation invariance: columns are independent, so shuffling them only shuffles the results
    rng = np.random.default_rng(7)
    perm = rng.permutation(n)
    rawp = {k: (v if np.ndim(v) == 0 else v[perm]) for k, v in raw.items()}
    outp = h.radiation(I.to_radiation_inputs(rawp), n, NLEV)
    for nm in FLUXES + ["cloud_cover_sw", "cloud_cover_lw", "lw_derivatives"]:
        assert np.array_equal(outp[nm], out[nm][perm]), nm
    # (c) physics: fluxes non-negative, direct <= total, TOA incoming = S0*mu0 for sunlit columns
    assert (out["sw_dn"] >= -1e-9).all() and (out["sw_up"] >= -1e-9).all() and (out["lw_up"] > 0).all()
    assert (out["sw_dn_direct"] <= out["sw_dn"] + 1e-9).all()
    mu0 = raw["cos_solar_zenith_angle"]
    sun = mu0 > 0
    assert np.abs(out["sw_dn"][sun, 0] - raw["solar_irradiance"] * mu0[sun]).max() <= 1e-9 * 1400
    assert (out["sw_dn"][~sun] == 0).all()
    # (d) oracle on a random subset of 256 columns
    idx = np.sort(rng.choice(n, 256, replace=False))
    sub = {k: (v if np.ndim(v) == 0 else v[idx]) for k, v in raw.items()}
    ref = orc.radiation(I.to_radiation_inputs(sub), len(idx), NLEV)
    for nm in FLUXES:
        assert np.abs(out[nm][idx] - ref[nm]).max() <= TOL, nm
    assert np.array_equal(out["cloud_cover_sw"][idx], ref["cloud_cover_sw"])


@pytest.mark.parametrize("kw", [dict(), dict(gas_model_name="ECCKD", do_nearest_spectral_lw_emiss=False, use_aerosols=True)])
def test_repeated_runs_are_bit_identical(handles, meridian_raw, kw):
    """Race detector for the asynchronous pieces (TMA bulk-copy rings in the flux kernels, overlapping column tiles, three
    concurrent kernel chains): the same 10 000 columns, run repeatedly, must reproduce every output bit for bit."""
    h, _, cfg = handles(**kw)
    n = 10000
    raw = I.synthetic_columns(meridian_raw, n)
    ref = h.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)
    for _ in range(8):
        out = h.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)
        for nm in FLUXES + OTHERS + ["cloud_cover_sw", "cloud_cover_lw"]:
            assert np.array_equal(out[nm], ref[nm], equal_nan=True), nm


@pytest.mark.parametrize("kw", [dict(), dict(sw_solver_name="SPARTACUS", lw_solver_name="SPARTACUS", do_3d_effects=True)])
def test_device_resident_entry_matches_host_entry(handles, meridian_raw, kw):
    import ctypes as C

    import torch

    from ecrad_b200 import abi

    h, _, cfg = handles(**kw)
    n = 257
    raw = I.synthetic_columns(meridian_raw, n)
    inp = I.to_radiation_inputs(raw, cfg)
    host = h.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)
    dev = torch.device("cuda:0")
    ist = abi.Inputs(); ist.struct_bytes = C.sizeof(abi.Inputs); ist.solar_irradiance = inp["solar_irradiance"]
    keep = {}
    for nm, dt, _ in abi.INPUT_ARRAYS:
        if nm not in inp:   # optional inputs (cloud effective sizes: SPARTACUS only)
            continue
        a = np.asfortranarray(inp[nm], dtype=np.int32 if dt == "i4" else np.float64)
        t = torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)   # (rows, ncol) C-order == (ncol, rows) Fortran order
        keep[nm] = t
        setattr(ist, nm, C.cast(t.data_ptr(), abi.c_ip if dt == "i4" else abi.c_dp))
    ost = abi.Outputs(); ost.struct_bytes = C.sizeof(abi.Outputs)
    outs = {}
    for nm in FLUXES:
        t = torch.full((NLEV + 1, n), float("nan"), dtype=torch.float64, device=dev)
        outs[nm] = t
        setattr(ost, nm, C.cast(t.data_ptr(), abi.c_dp))
    h.radiation_device(n, NLEV, ist, ost, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for nm in FLUXES:
        assert np.array_equal(outs[nm].cpu().numpy().T, host[nm]), nm


def test_device_entry_with_separate_leading_dimensions(handles, meridian_raw):
    """ecrad_b200_radiation_device_ld: two column ranges write their slices of ONE set of output arrays (ld_out = total columns),
    as the ranks of a multi-GPU run do with peer-mapped arrays; result = the host entry on all columns, bit for bit."""
    import ctypes as C

    import torch

    from ecrad_b200 import abi

    h, _, cfg = handles()
    n, n1 = 300, 130   # ranges [0, 130) and [130, 300)
    raw = I.synthetic_columns(meridian_raw, n)
    inp = I.to_radiation_inputs(raw)
    host = h.radiation(I.to_radiation_inputs(raw), n, NLEV)
    dev = torch.device("cuda:0")
    keep = {}
    for nm, dt, _ in abi.INPUT_ARRAYS:
        if nm not in inp:   # optional inputs (cloud effective sizes: SPARTACUS only)
            continue
        a = np.asfortranarray(inp[nm], dtype=np.int32 if dt == "i4" else np.float64)
        keep[nm] = torch.from_numpy(np.ascontiguousarray(a.T)).to(dev)   # (rows, n): leading dimension n
    outs = {nm: torch.full((NLEV + 1, n), float("nan"), dtype=torch.float64, device=dev) for nm in FLUXES}
    for c0, nc in ((0, n1), (n1, n - n1)):
        ist = abi.Inputs(); ist.struct_bytes = C.sizeof(abi.Inputs); ist.solar_irradiance = inp["solar_irradiance"]
        for nm, dt, _ in abi.INPUT_ARRAYS:
            if nm not in keep:
                continue
            esz = 4 if dt == "i4" else 8
            setattr(ist, nm, C.cast(keep[nm].data_ptr() + esz * c0, abi.c_ip if dt == "i4" else abi.c_dp))
        ost = abi.Outputs(); ost.struct_bytes = C.sizeof(abi.Outputs)
        for nm in FLUXES:
            setattr(ost, nm, C.cast(outs[nm].data_ptr() + 8 * c0, abi.c_dp))
        h.radiation_device_ld(nc, NLEV, n, n, ist, ost, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    for nm in FLUXES:
        assert np.array_equal(outs[nm].cpu().numpy().T, host[nm]), nm


def test_error_behaviour(meridian_raw):
    """Non-zero status + message instead of the reference's radiation_abort."""
    from ecrad_b200.radiation_interface import RadiationError, setup_radiation

    with pytest.raises(RadiationError, match="do_lw_aerosol_scattering"):   # built for McICA / Cloudless on RRTMG only
        setup_radiation(RadiationConfig(do_lw_aerosol_scattering=True, sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds").consolidate())
    with pytest.raises(RadiationError, match="requires longwave cloud scattering"):   # radiation_interface.F90:84-88
        setup_radiation(RadiationConfig(do_lw_aerosol_scattering=True, do_lw_cloud_scattering=False).consolidate())
    h = setup_radiation(RadiationConfig().consolidate())
    with pytest.raises(RadiationError, match="bad dimensions"):
        h.radiation(I.to_radiation_inputs(meridian_raw), 32, NLEV, istartcol=5, iendcol=40)
    h.finalize()


@pytest.mark.parametrize("opts", [dict(gas_variant=0), dict(gas_variant=3), dict(scan_solvers=1), dict(scan_solvers=1, gas_variant=3)])
@pytest.mark.parametrize("kw", [dict(use_aerosols=True), dict(sw_solver_name="Cloudless", lw_solver_name="Cloudless"),
                                dict(overlap_scheme_name="Exp-Exp", do_lw_cloud_scattering=False, do_sw_delta_scaling_with_gases=True, use_aerosols=True)])
def test_selectable_kernel_variants_vs_oracle(meridian_raw, kw, opts):
    """Every kernel the library can be told to use (set_option): the per-column and the band-wise RRTMG gas optics, the lanes-are-g-points
    and the warp-scan McICA / Cloudless solvers -- same parity bar against the oracle, with 137 and with 60 levels."""
    from ecrad_b200.radiation_interface import setup_radiation
    from oracle_lib import Oracle

    cfg = RadiationConfig(**kw).consolidate()
    h, orc = setup_radiation(cfg), Oracle(cfg)
    for k, v in opts.items():
        h.set_option(k, v)
    try:
        for n, nlev in ((300, NLEV), (40, 60)):
            raw = I.synthetic_columns(meridian_raw, n)
            if nlev != NLEV:
                raw = coarsen_levels(raw, nlev)
            out = h.radiation(I.to_radiation_inputs(raw, cfg), n, nlev)
            ref = orc.radiation(I.to_radiation_inputs(raw, cfg), n, nlev)
            compare(out, ref, FLUXES + OTHERS)
            assert np.array_equal(out["cloud_cover_sw"], ref["cloud_cover_sw"])
    finally:
        h.finalize()


@pytest.mark.parametrize("kw", [dict(), dict(sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds"), dict(scan_solvers=1),
                                dict(sw_solver_name="SPARTACUS", lw_solver_name="SPARTACUS", do_3d_effects=True, do_3d_lw_multilayer_effects=True,
                                     sw_entrapment_name="Maximum", overhang_factor=1.0, overhead_sun_factor=0.06),
                                dict(sw_solver_name="SPARTACUS", lw_solver_name="SPARTACUS", do_3d_effects=True, sw_entrapment_name="Explicit")])
def test_i3rc_cumulus_profile_164_layers(kw):
    """More than 160 layers: the reference's I3RC cumulus case (test/i3rc/i3rc_mls_cumulus.nc, 164 layers, eight solar zenith angles),
    with the options of test/i3rc/configI3RC.nam for SPARTACUS."""
    from ecrad_b200.radiation_interface import setup_radiation
    from oracle_lib import Oracle

    kw = dict(kw)
    scan = kw.pop("scan_solvers", 0)
    raw = {k: np.array(v, dtype=np.float64) for k, v in np.load(os.path.join(os.path.dirname(__file__), "golden", "i3rc_mls_cumulus_inputs.npz")).items()}
    cfg = RadiationConfig(**kw).consolidate()
    h, orc = setup_radiation(cfg), Oracle(cfg)
    h.set_option("scan_solvers", scan)
    try:
        n, nlev = 8, 164
        out = h.radiation(I.to_radiation_inputs(raw, cfg), n, nlev)
        ref = orc.radiation(I.to_radiation_inputs(raw, cfg), n, nlev)
        compare(out, ref, FLUXES + OTHERS)
        for nm in ("cloud_cover_lw", "cloud_cover_sw", "cloud_fraction"):
            assert np.array_equal(out[nm], ref[nm]), nm
        assert out["sw_dn"][0, 0] > 1300.0 and np.all(out["lw_up"] > 0.0)   # overhead sun: ~1366 W m-2 at the top
    finally:
        h.finalize()


@pytest.mark.parametrize("kw,nproma", [(dict(use_aerosols=True), 32), (dict(), 13),
                                       (dict(sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds", do_save_spectral_flux=True), 16),
                                       (dict(gas_model_name="ECCKD", do_nearest_spectral_lw_emiss=False), 64)])
def test_blocked_nproma_entry_is_bit_identical(handles, meridian_raw, kw, nproma):
    """the IFS-style blocked layout zrgp(nproma, nfields, nblocks) (driver/ifs_blocking.F90) through
    ecrad_b200_radiation_blocked gives the bits of the column-layout entry, ragged last block included (the reference's own cross-driver
    check: test/ifs/CMakeLists.txt:139-200 compares the blocked driver with the plain one)."""
    h, _, cfg = handles(**kw)
    n = 150
    raw = I.synthetic_columns(meridian_raw, n)
    sp = bool(kw.get("do_save_spectral_flux"))
    a = h.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV, spectral_profiles=sp)
    b = h.radiation_blocked(I.to_radiation_inputs(raw, cfg), n, NLEV, nproma, spectral_profiles=sp)
    for nm in a:
        if nm == "cloud_fraction" or a[nm] is None:
            continue
        assert np.array_equal(a[nm], b[nm], equal_nan=True), nm


@pytest.mark.parametrize("kw", [dict(use_aerosols=True, do_lw_aerosol_scattering=True), dict(do_lw_aerosol_scattering=True),
                                dict(use_aerosols=True, do_lw_aerosol_scattering=True, sw_solver_name="Cloudless", lw_solver_name="Cloudless",
                                     do_save_spectral_flux=False),
                                dict(use_aerosols=True, do_lw_aerosol_scattering=True, overlap_scheme_name="Exp-Exp", do_nearest_spectral_lw_emiss=False)])
def test_lw_aerosol_scattering(handles, meridian_raw, kw):
    """do_lw_aerosol_scattering (the default of config_type, radiation_config.F90:260): aerosols scatter in the longwave, so the clear-sky
    sub-column needs the full adding method as well (radiation_mcica_lw.F90:160-173, :324-329; radiation_aerosol_optics.F90:657-801).
    McICA and Cloudless on RRTMG; the oracle restates the same branches."""
    n = 300
    h, orc, cfg = handles(**kw)
    raw = I.synthetic_columns(meridian_raw, n)
    out = h.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)
    ref = orc.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)
    compare(out, ref, FLUXES + OTHERS)
    assert np.array_equal(out["cloud_cover_lw"], ref["cloud_cover_lw"])
    if kw.get("use_aerosols"):
        # scattering aerosols change the longwave by a few tenths of a W m-2 against absorption-only aerosols
        h0, _, cfg0 = handles(**dict(kw, do_lw_aerosol_scattering=False))
        out0 = h0.radiation(I.to_radiation_inputs(raw, cfg0), n, NLEV)
        d = np.abs(out["lw_up"] - out0["lw_up"]).max()
        assert 1e-3 < d < 5.0, d


def test_ckdmip_profiles_vs_line_by_line(handles):
    """The reference's CKDMIP clear-sky test (test/ckdmip: 50 profiles x 54 layers, line-by-line fluxes) on the GPU: within 1e-6 W m-2
    of the oracle and within the published accuracy of each gas-optics model of the line-by-line truth (tests/test_ckdmip_lbl.py)."""
    import test_ckdmip_lbl as K
    fix = K.load_fixture()
    for name, kw, lw_bound, sw_bound in K.MODELS:
        h, orc, cfg = handles(**K.CLOUDLESS, **kw)
        K.check_against_lbl(lambda raw: h.radiation(I.to_radiation_inputs(raw, cfg), 50, 54), fix, lw_bound, sw_bound)
        raw = I.ckdmip_raw(fix, 0.5)
        out = h.radiation(I.to_radiation_inputs(raw, cfg), 50, 54)
        ref = orc.radiation(I.to_radiation_inputs(raw, cfg), 50, 54)
        compare(out, ref, ["lw_up", "lw_dn", "sw_up", "sw_dn", "sw_dn_direct"])


@pytest.mark.parametrize("kw", [dict(liquid_model_name="Slingo"), dict(ice_model_name="Baran-EXPERIMENTAL"), dict(ice_model_name="Baran2016", use_aerosols=True),
                                dict(ice_model_name="Baran2017-EXPERIMENTAL", do_lw_cloud_scattering=False),
                                dict(ice_model_name="Yi", sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds"),
                                dict(liquid_model_name="Slingo", ice_model_name="Yi", do_sw_delta_scaling_with_gases=True),
                                dict(ice_model_name="Baran2016", sw_solver_name="SPARTACUS", lw_solver_name="SPARTACUS", do_3d_effects=True)])
def test_other_liquid_and_ice_optics_models(handles, meridian_raw, kw):
    """config%i_liq_model / i_ice_model other than SOCRATES + Fu (radiation_cloud_optics.F90:345-447): Slingo / Lindner-Li droplets, Baran,
    Baran-2016, Baran-2017 and Yi ice, through every solver family.  Oracle: tests/test_cloud_optics_models.py pins its formulas."""
    n = 200 if "SPARTACUS" in kw.values() else 400
    h, orc, cfg = handles(**kw)
    raw = I.synthetic_columns(meridian_raw, n)
    out = h.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)
    ref = orc.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)
    compare(out, ref, FLUXES + OTHERS)
    assert np.array_equal(out["cloud_cover_sw"], ref["cloud_cover_sw"])
    # and it is not the default pair's answer
    h0, _, cfg0 = handles(**{k: v for k, v in kw.items() if k not in ("liquid_model_name", "ice_model_name")})
    out0 = h0.radiation(I.to_radiation_inputs(raw, cfg0), n, NLEV)
    assert np.abs(out["sw_up"] - out0["sw_up"]).max() > 1.0


@pytest.mark.parametrize("kw", [dict(), dict(use_aerosols=True, sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds", do_save_spectral_flux=True)])
def test_single_precision_boundary(handles, meridian_raw, golden_noaer, kw):
    """ecrad_b200_radiation_sp: the call of a host built with JPRB = JPRM -- float arrays in, float arrays out, double-precision kernels
    in between.  Gate of BASELINE.json's north_star for single precision: 1e-3 W m-2 against the double-precision result (the
    differences are the rounding of the derived inputs -- mass mixing ratios, saturation -- to float on the way in and of the fluxes on
    the way out)."""
    h, _, cfg = handles(**kw)
    sp = bool(kw.get("do_save_spectral_flux"))
    n = 32
    d = h.radiation(I.to_radiation_inputs(meridian_raw, cfg), n, NLEV, spectral_profiles=sp)
    s = h.radiation_sp(I.to_radiation_inputs(meridian_raw, cfg), n, NLEV, spectral_profiles=sp)
    for nm in FLUXES + OTHERS + (BANDS if sp else []):
        assert s[nm].dtype == np.float32
        m = np.isfinite(d[nm])
        assert np.isfinite(s[nm][m]).all(), nm
        assert np.abs(s[nm][m].astype(np.float64) - d[nm][m]).max() <= 1e-3, nm
    assert np.array_equal(s["cloud_fraction"], d["cloud_fraction"].astype(np.float32))
    assert np.array_equal(s["cloud_cover_sw"], d["cloud_cover_sw"].astype(np.float32))
    if not kw:   # and the reference's own (float32) golden output
        assert np.abs(s["sw_dn"].astype(np.float64) - golden_noaer["flux_dn_sw"]).max() <= 1e-3
    # synthetic columns (inputs rounded to float on the way in) and tiling: still within the single-precision gate of the fp64 run
    n = 700
    raw = I.synthetic_columns(meridian_raw, n)
    d = h.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)
    s = h.radiation_sp(I.to_radiation_inputs(raw, cfg), n, NLEV, istartcol=3, iendcol=690)
    for nm in FLUXES:
        assert np.isnan(s[nm][:2]).all() and np.isnan(s[nm][690:]).all(), nm
        err = np.abs(s[nm][2:690].astype(np.float64) - d[nm][2:690]).max()
        assert err <= 2e-3, (nm, err)   # input rounding (temperature to 2e-5 K) + output rounding (6e-5 W m-2 at 1000 W m-2)
