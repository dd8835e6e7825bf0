"""GPU parity of the SPARTACUS solvers (solver_sp.cu, through the C-ABI) against the CPU oracle (oracle/spartacus.c).

The reference ships no SPARTACUS output (its ctest targets `spartacus*` are XFAIL_VALIDATION without a reference file), so the
oracle is pinned indirectly (tests/test_oracle_spartacus.py) and the bar here is the fp64 tolerance of BASELINE.json's north_star:
|flux - oracle| <= 1e-6 W m-2 on every flux component, cloud cover and cropped fraction bit-exact.
"""
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
BANDS = ["lw_up_band", "lw_dn_band", "sw_up_band", "sw_dn_band", "sw_dn_direct_band"]
SP = dict(sw_solver_name="SPARTACUS", lw_solver_name="SPARTACUS", do_3d_effects=True)


def run_pair(kw, raw, n, nlev=NLEV, spectral_profiles=False):
    from ecrad_b200.radiation_interface import setup_radiation
    from oracle_lib import Oracle

    cfg = RadiationConfig(**kw).consolidate()
    h = setup_radiation(cfg)
    try:
        out = h.radiation(I.to_radiation_inputs(raw, cfg), n, nlev, spectral_profiles=spectral_profiles)
    finally:
        h.finalize()
    ref = Oracle(cfg).radiation(I.to_radiation_inputs(raw, cfg), n, nlev, spectral_profiles=spectral_profiles)
    return out, ref


def compare(out, ref, names, tol=TOL):
    worst = {}
    for nm in names:
        a, b = out[nm], ref[nm]
        assert a.shape == b.shape, nm
        m = np.isfinite(b)
        assert np.isfinite(a[m]).all(), f"{nm}: non-finite values"
        worst[nm] = np.abs(a[m] - b[m]).max() if m.any() else 0.0
    bad = {k: v for k, v in worst.items() if v > tol}
    assert not bad, f"max |gpu - oracle| (W m-2) above {tol}: {bad}"
    return worst


def check_exact(out, ref):
    assert np.array_equal(out["cloud_cover_lw"], ref["cloud_cover_lw"])
    assert np.array_equal(out["cloud_cover_sw"], ref["cloud_cover_sw"])
    assert np.array_equal(out["cloud_fraction"], ref["cloud_fraction"])


@pytest.mark.parametrize("kw", [dict(), dict(sw_entrapment_name="Maximum"), dict(sw_entrapment_name="Zero"),
                                dict(sw_entrapment_name="Edge-only"), dict(sw_entrapment_name="Non-fractal"),
                                dict(do_3d_effects=False), dict(do_3d_lw_multilayer_effects=True),
                                dict(do_lw_side_emissivity=False, clear_to_thick_fraction=0.3),
                                dict(use_aerosols=True, do_lw_cloud_scattering=False),
                                dict(use_expm_everywhere=True), dict(use_expm_everywhere=True, do_3d_effects=False),
                                dict(n_regions=2), dict(n_regions=2, sw_entrapment_name="Maximum", do_3d_lw_multilayer_effects=True),
                                dict(n_regions=2, do_3d_effects=False),
                                dict(use_aerosols=True, do_lw_aerosol_scattering=True), dict(do_lw_aerosol_scattering=True, n_regions=2),
                                dict(use_general_cloud_optics=True)])
def test_spartacus_meridian_vs_oracle(meridian_raw, kw):
    """The reference's own 32-column slice (the input of its `spartacus` / `spartacus_maxentr` ctest targets)."""
    out, ref = run_pair({**SP, **kw}, meridian_raw, 32, spectral_profiles=True)
    compare(out, ref, FLUXES + OTHERS + BANDS)
    check_exact(out, ref)
    for nm in BANDS:   # band sums = broadband
        assert np.abs(out[nm].sum(axis=0) - out[nm[:-5]]).max() <= 1e-9, nm


@pytest.mark.parametrize("kw", [dict(), dict(sw_entrapment_name="Maximum", use_beta_overlap=True)])
def test_spartacus_synthetic_columns_vs_oracle(meridian_raw, kw):
    """300 perturbed columns: other cloud profiles and sun angles (incl. night columns and very low sun)."""
    n = 300
    raw = I.synthetic_columns(meridian_raw, n)
    out, ref = run_pair({**SP, **kw}, raw, n)
    compare(out, ref, FLUXES + OTHERS)
    check_exact(out, ref)


def test_spartacus_ecckd_vs_oracle(meridian_raw):
    """SPARTACUS on the 32-term ecCKD spectra (per-g-point cloud optics): the same kernels, other spectral sizes."""
    kw = dict(gas_model_name="ECCKD", do_nearest_spectral_lw_emiss=False, **SP)
    out, ref = run_pair(kw, meridian_raw, 32)
    compare(out, ref, FLUXES + OTHERS)
    check_exact(out, ref)


def test_spartacus_without_3d_is_tripleclouds(meridian_raw):
    """Physical cross-check on the GPU alone: with 3D effects off SPARTACUS and Tripleclouds solve the same SW equations."""
    from ecrad_b200.radiation_interface import setup_radiation

    res = {}
    for name in ("SPARTACUS", "Tripleclouds"):
        cfg = RadiationConfig(sw_solver_name=name, lw_solver_name=name, do_3d_effects=False).consolidate()
        h = setup_radiation(cfg)
        res[name] = h.radiation(I.to_radiation_inputs(meridian_raw, cfg), 32, NLEV)
        h.finalize()
    for nm in ("sw_up", "sw_dn", "sw_dn_direct"):
        assert np.abs(res["SPARTACUS"][nm] - res["Tripleclouds"][nm]).max() < 1e-5, nm


def test_spartacus_tiling_and_column_range(meridian_raw):
    """Host-entry tiles must not show: 700 columns in one tile == tiles of 128; istartcol:iendcol leaves the rest untouched."""
    from ecrad_b200.radiation_interface import setup_radiation

    n = 700
    cfg = RadiationConfig(**SP).consolidate()
    raw = I.synthetic_columns(meridian_raw, n)
    h = setup_radiation(cfg)
    h.set_option("tile_cols", 1024)
    a = h.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)
    h.set_option("tile_cols", 128)
    b = h.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)
    for nm in FLUXES + OTHERS:
        assert np.array_equal(a[nm], b[nm], equal_nan=True), nm
    c = h.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV, istartcol=101, iendcol=400)
    h.finalize()
    for nm in FLUXES:
        assert np.array_equal(c[nm][100:400], a[nm][100:400]), nm
        assert np.isnan(c[nm][:100]).all() and np.isnan(c[nm][400:]).all(), nm


def test_spartacus_error_behaviour():
    from ecrad_b200.radiation_interface import RadiationError, setup_radiation

    with pytest.raises(RadiationError, match="Exponential-Random"):
        setup_radiation(RadiationConfig(overlap_scheme_name="Max-Ran", **SP).consolidate())
    with pytest.raises(RadiationError, match="Exponential-Random"):
        setup_radiation(RadiationConfig(overlap_scheme_name="Exp-Exp", sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds").consolidate())
    with pytest.raises(RadiationError, match="Homogeneous"):
        setup_radiation(RadiationConfig(sw_solver_name="Homogeneous").consolidate())
    with pytest.raises(RadiationError, match="delta-Eddington scaling with gases"):
        setup_radiation(RadiationConfig(do_sw_delta_scaling_with_gases=True, **SP).consolidate())
    with pytest.raises(RadiationError, match="n_regions must be 2 or 3"):
        setup_radiation(RadiationConfig(n_regions=4, **SP).consolidate())
    with pytest.raises(RadiationError, match="n_regions = 2 needs SPARTACUS in both spectra"):
        setup_radiation(RadiationConfig(n_regions=2, sw_solver_name="SPARTACUS", lw_solver_name="Tripleclouds").consolidate())


def test_spartacus_full_size_properties(meridian_raw):
    """5000 columns through the tiled host entry (10 tiles, two overlapping compute sets): column independence (permutation
    invariance, bit-exact), bit-identical repeated runs (race detector), physical bounds, oracle spot check on a random subset."""
    from ecrad_b200.radiation_interface import setup_radiation
    from oracle_lib import Oracle

    n = 5000
    cfg = RadiationConfig(**SP).consolidate()
    raw = I.synthetic_columns(meridian_raw, n)
    h = setup_radiation(cfg)
    out = h.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)
    for nm in FLUXES:
        assert np.isfinite(out[nm]).all(), nm
    again = h.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)
    for nm in FLUXES + OTHERS:
        assert np.array_equal(again[nm], out[nm], equal_nan=True), nm
    rng = np.random.default_rng(11)
    perm = rng.permutation(n)
    rawp = {k: (v if np.ndim(v) == 0 else v[perm]) for k, v in raw.items()}
    outp = h.radiation(I.to_radiation_inputs(rawp, cfg), n, NLEV)
    h.finalize()
    for nm in FLUXES + ["cloud_cover_sw", "cloud_cover_lw", "lw_derivatives"]:
        assert np.array_equal(outp[nm], out[nm][perm]), nm
    mu0 = raw["cos_solar_zenith_angle"]
    sun = mu0 >= 1e-10
    assert (out["sw_dn"] >= -1e-9).all() and (out["sw_up"] >= -1e-9).all() and (out["lw_up"] > 0).all()
    assert (out["sw_dn_direct"] <= out["sw_dn"] + 1e-9).all()
    assert np.abs(out["sw_dn"][sun, 0] - raw["solar_irradiance"] * mu0[sun]).max() <= 1e-9 * 1400
    assert (out["sw_dn"][~sun] == 0).all()
    net = out["sw_dn"] - out["sw_up"]
    assert (np.diff(net[sun], axis=1) <= 1e-6).all()          # absorption only: the SW net flux decreases downwards
    idx = np.sort(rng.choice(n, 128, replace=False))
    sub = {k: (v if np.ndim(v) == 0 else v[idx]) for k, v in raw.items()}
    ref = Oracle(cfg).radiation(I.to_radiation_inputs(sub, cfg), len(idx), NLEV)
    for nm in FLUXES:
        assert np.abs(out[nm][idx] - ref[nm]).max() <= TOL, nm
    assert np.array_equal(out["cloud_cover_sw"][idx], ref["cloud_cover_sw"])


def test_spartacus_i3rc_vs_mystic_and_oracle():
    """The reference's I3RC cumulus test (test/i3rc) on the GPU: within the bounds of tests/test_i3rc_libradtran.py of libRadtran's
    MYSTIC 3D Monte-Carlo fluxes, and within 1e-6 W m-2 of the oracle, per-g-point outputs (in the reordered SPARTACUS sequence of
    radiation_ifs_rrtm.F90:122-130) included."""
    import test_i3rc_libradtran as L
    from ecrad_b200.radiation_interface import setup_radiation
    from oracle_lib import Oracle
    fix, lib = L.load()
    pairs = []

    def run(raw, **kw):
        cfg = RadiationConfig(**kw).consolidate()
        h = setup_radiation(cfg)
        n = len(raw["cos_solar_zenith_angle"])
        out = h.radiation(I.to_radiation_inputs(raw, cfg), n, 164)
        h.finalize()
        pairs.append((out, Oracle(cfg).radiation(I.to_radiation_inputs(raw, cfg), n, 164)))
        return out

    L.check_against_libradtran(run, fix, lib)
    L.check_two_regions(run, fix, lib)
    for out, ref in pairs:
        compare(out, ref, FLUXES + OTHERS)
