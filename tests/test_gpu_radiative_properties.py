"""save_radiative_properties (radiation_interface.F90:405-425, radiation_save.F90:716-726): the optical properties the gas, aerosol and
cloud optics hand to the solvers, read back from the GPU through `ecrad_b200_save_radiative_properties` and compared array by array
with the oracle at the same point of `radiation()` -- the per-g-point / per-band parity of every optics kernel, not only of the fluxes.
Tolerance: 1e-9 relative (+1e-13 absolute) on every element; the fluxes these properties produce agree to 1e-6 W m-2 elsewhere."""
import numpy as np
import pytest

from ecrad_b200 import inputs as I
from ecrad_b200.config import RadiationConfig

pytestmark = pytest.mark.gpu
NLEV = 137
TC = dict(sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds")
SP = dict(sw_solver_name="SPARTACUS", lw_solver_name="SPARTACUS", do_3d_effects=True)
CKD = dict(gas_model_name="ECCKD", do_nearest_spectral_lw_emiss=False)


def run_pair(kw, raw, n, opts=None, **rng):
    from ecrad_b200.radiation_interface import setup_radiation
    from oracle_lib import Oracle

    cfg = RadiationConfig(**kw).consolidate()
    h = setup_radiation(cfg)
    for k, v in (opts or {}).items():
        assert h.lib.ecrad_b200_set_option(h.h, k.encode(), v) == 0
    try:
        out = h.radiative_properties(I.to_radiation_inputs(raw, cfg), n, NLEV, **rng)
        flux = h.radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)   # (the handle still works, and on the same scratch)
    finally:
        h.finalize()
    ref = Oracle(cfg).radiative_properties(I.to_radiation_inputs(raw, cfg), n, NLEV, **rng)
    return out, ref, flux, cfg


def check(out, ref, cos_sza, cfg, cols=slice(None)):
    """Every element, night columns included (RRTMG-IFS: aerosols only there; ecCKD: computed like any column)."""
    for nm, a in out.items():
        b = ref[nm]
        assert a.shape == b.shape, nm
        a, b = a[..., cols], b[..., cols]
        if nm in ("ssa_lw", "g_lw") and not cfg.do_lw_aerosol_scattering:
            assert not a.any(), nm   # not defined by the reference without longwave aerosol scattering: zero here
            continue
        assert np.isfinite(a).all(), nm
        err = np.abs(a - b) / (1e-9 * np.abs(b) + 1e-13)
        assert err.max() <= 1.0, f"{nm}: max |gpu - oracle| = {np.abs(a - b).max():.3e} at value {b.flat[err.argmax()]:.6e}"


@pytest.mark.parametrize("kw,opts", [
    (dict(use_aerosols=True), None), (dict(use_aerosols=True), dict(gas_variant=0)), (dict(use_aerosols=True), dict(scan_solvers=1)),
    (dict(use_aerosols=True, do_lw_aerosol_scattering=True), None), (dict(use_aerosols=True, do_lw_aerosol_scattering=True, **SP), None),
    (dict(**SP), None), (dict(use_general_cloud_optics=True, do_lw_cloud_scattering=False, **TC), None),
    (dict(liquid_model_name="Slingo", ice_model_name="Yi", do_sw_delta_scaling_with_gases=True, use_aerosols=True), None),
    (dict(use_aerosols=True, **CKD, **TC), None), (dict(ecckd_tables="ecckd_tables_64b.bin", **CKD), None),
    (dict(sw_gas_model_name="ECCKD", do_nearest_spectral_lw_emiss=False, use_aerosols=True), None),
    (dict(lw_gas_model_name="ECCKD", do_nearest_spectral_lw_emiss=False, use_aerosols=True, **SP), None),
    (dict(sw_solver_name="Homogeneous", lw_solver_name="Homogeneous"), None), (dict(do_sw_delta_scaling_with_gases=True, use_aerosols=True, **CKD), None), (dict(do_nearest_spectral_sw_albedo=True), None)])
def test_radiative_properties_vs_oracle(meridian_raw, kw, opts):
    n = 96
    raw = I.synthetic_columns(meridian_raw, n)
    out, ref, flux, cfg = run_pair(kw, raw, n, opts)
    check(out, ref, raw["cos_solar_zenith_angle"], cfg)
    assert np.isfinite(flux["lw_up"]).all()
    if cfg.do_sw and not cfg.use_aerosols:   # gases alone scatter isotropically
        assert not out["g_sw"].any()


def test_column_range_and_tiles(meridian_raw):
    """istartcol..iendcol only (the rest stays untouched), and more columns than one internal tile (2048)."""
    n = 2300
    raw = I.synthetic_columns(meridian_raw, n)
    out, ref, _, cfg = run_pair(dict(use_aerosols=True), raw, n, istartcol=7, iendcol=2290)
    for nm, a in out.items():
        assert np.isnan(a[..., :6]).all() and np.isnan(a[..., 2290:]).all(), nm
    check(out, ref, raw["cos_solar_zenith_angle"], cfg, cols=slice(6, 2290))


def test_band_sums_are_consistent(meridian_raw):
    """Physical sanity of what comes back: incoming_sw sums to the solar irradiance in sunlit columns, the Planck function at the
    surface half-level times (1 - albedo) is not the emission (skin temperature), cloud properties vanish outside clouds."""
    n = 64
    raw = I.synthetic_columns(meridian_raw, n)
    out, _, flux, cfg = run_pair(dict(), raw, n)
    sun = raw["cos_solar_zenith_angle"] > 0.0
    assert np.abs(out["incoming_sw"][:, sun].sum(axis=0) - float(raw["solar_irradiance"])).max() < 1e-9
    frac = np.asarray(flux["cloud_fraction"])   # cropped by radiation(): (ncol, nlev)
    clear = (frac == 0.0).T                      # (nlev, ncol)
    for nm in ("od_lw_cloud", "od_sw_cloud", "ssa_sw_cloud", "g_sw_cloud"):
        assert not out[nm][:, clear].any(), nm
    assert (out["od_sw_cloud"][:, ~clear] > 0.0).all()
    assert (out["od_lw"] > 0.0).all() and (out["planck_hl"] >= 0.0).all()
