"""CPU checks of the oracle's option branches added late in round 1 (no reference output exists for any of them: invariants only).

  do_sw_delta_scaling_with_gases   radiation_mcica_sw.F90:156-180, :274-278 (+ cloudless / tripleclouds / homogeneous)
  do_toa_spectral_flux             radiation_flux.F90:579-660 calc_toa_spectral
  do_nearest_spectral_sw_albedo    radiation_single_level.F90:266-285, radiation_flux.F90:479-497
  do_nearest_spectral_lw_emiss     radiation_single_level.F90:310-355 (weighted intervals with RRTMG)
  cloud_pdf_shape_name             radiation_regions.F90:110-126 (lognormal regions)
"""
import numpy as np
import pytest

from ecrad_b200 import inputs as I
from ecrad_b200.config import RadiationConfig
from oracle_lib import Oracle

NLEV = 137
TC = dict(sw_solver_name="Tripleclouds", lw_solver_name="Tripleclouds")


def run(raw, n=32, **kw):
    cfg = RadiationConfig(**kw).consolidate()
    return Oracle(cfg).radiation(I.to_radiation_inputs(raw, cfg), n, NLEV)


def test_delta_scaling_with_gases(meridian_raw):
    base = run(meridian_raw)
    dsg = run(meridian_raw, do_sw_delta_scaling_with_gases=True)
    # without aerosols the clear-sky mixture has g = 0: the scaling is the identity there, and it never touches the longwave
    for nm in ("sw_up_clear", "sw_dn_clear", "sw_dn_direct_clear", "lw_up", "lw_dn"):
        assert np.array_equal(dsg[nm], base[nm]), nm
    # scaling the gas-cloud mixture instead of the cloud alone is a small change of the cloudy fluxes, not a different answer
    d = np.abs(dsg["sw_up"] - base["sw_up"]).max()
    assert 0.0 < d < 10.0
    aer = run(meridian_raw, use_aerosols=True, do_sw_delta_scaling_with_gases=True)
    aer0 = run(meridian_raw, use_aerosols=True)
    assert 0.0 < np.abs(aer["sw_up_clear"] - aer0["sw_up_clear"]).max() < 5.0
    # Tripleclouds does not scale its clear-sky region (radiation_tripleclouds_sw.F90:269)
    t1, t0 = run(meridian_raw, use_aerosols=True, do_sw_delta_scaling_with_gases=True, **TC), run(meridian_raw, use_aerosols=True, **TC)
    assert np.abs(t1["sw_up"] - t0["sw_up"]).max() > 0.0


@pytest.mark.parametrize("kw", [dict(), TC])
def test_toa_spectral_flux(meridian_raw, kw):
    out = run(meridian_raw, do_toa_spectral_flux=True, **kw)
    assert np.abs(out["sw_up_toa_band"].sum(axis=0) - out["sw_up"][:, 0]).max() <= 1e-9
    assert np.abs(out["sw_up_toa_clear_band"].sum(axis=0) - out["sw_up_clear"][:, 0]).max() <= 1e-9
    assert np.abs(out["lw_up_toa_band"].sum(axis=0) - out["lw_up"][:, 0]).max() <= 1e-9
    assert np.abs(out["lw_up_toa_clear_band"].sum(axis=0) - out["lw_up_clear"][:, 0]).max() <= 1e-9
    sun = np.asarray(meridian_raw["cos_solar_zenith_angle"]) >= 1e-10
    if kw:   # only the Tripleclouds solver sets sw_dn_toa_g (sunlit columns)
        assert np.abs(out["sw_dn_toa_band"][:, sun].sum(axis=0) - out["sw_dn"][sun, 0]).max() <= 1e-9
        assert np.isnan(out["sw_dn_toa_g"][:, ~sun]).all()
    else:
        assert np.isnan(out["sw_dn_toa_g"]).all() and np.isnan(out["sw_dn_toa_band"]).all()
    off = run(meridian_raw, **kw)
    assert np.isnan(off["sw_up_toa_band"]).all() and np.isnan(off["lw_up_toa_band"]).all()


def test_nearest_albedo_and_weighted_emissivity(meridian_raw):
    near = run(meridian_raw, do_nearest_spectral_sw_albedo=True)
    wgt = run(meridian_raw)
    # canopy fluxes partition the surface downwelling flux whichever mapping is used
    for o in (near, wgt):
        tot = o["sw_dn_diffuse_surf_canopy"].sum(axis=0) + o["sw_dn_direct_surf_canopy"].sum(axis=0)
        assert np.abs(tot - o["sw_dn"][:, -1]).max() <= 1e-9
    assert np.isfinite(near["sw_up"]).all() and np.abs(near["sw_up"] - wgt["sw_up"]).max() < 40.0
    assert np.array_equal(near["lw_up"], wgt["lw_up"])
    we = run(meridian_raw, do_nearest_spectral_lw_emiss=False)
    assert np.array_equal(we["sw_up"], wgt["sw_up"])
    assert np.abs(we["lw_dn_surf_canopy"].sum(axis=0) - we["lw_dn"][:, -1]).max() <= 1e-9
    assert 0.0 < np.abs(we["lw_up"] - wgt["lw_up"]).max() < 3.0


def test_lognormal_regions(meridian_raw):
    gam = run(meridian_raw, **TC)
    logn = run(meridian_raw, cloud_pdf_shape_name="Lognormal", **TC)
    # the overlap of the cloud boundaries does not depend on how the cloudy part is split: same cloud cover, same clear-sky fluxes
    assert np.array_equal(logn["cloud_cover_sw"], gam["cloud_cover_sw"])
    for nm in ("sw_up_clear", "lw_up_clear"):
        assert np.array_equal(logn[nm], gam[nm]), nm
    assert np.isfinite(logn["sw_up"]).all() and 0.0 < np.abs(logn["sw_up"] - gam["sw_up"]).max() < 60.0
    with pytest.raises(ValueError, match="gamma PDF"):
        RadiationConfig(cloud_pdf_shape_name="Lognormal").consolidate().to_struct()


@pytest.mark.parametrize("kw", [dict(), TC, dict(sw_solver_name="SPARTACUS", lw_solver_name="SPARTACUS", do_3d_effects=True)])
def test_general_cloud_optics_with_rrtmg(meridian_raw, kw):
    """config%use_general_cloud_optics = true with RRTMG-IFS (radiation_interface.F90:378-392, radiation_config.F90:1078-1090: cloud
    properties per band): Mie droplets + Baum ice from the look-up tables instead of SOCRATES + Fu-IFS.  No reference output exists
    for the combination: the clear-sky fluxes cannot change, the cloudy ones move by what two cloud-optics models differ by."""
    base = run(meridian_raw, **kw)
    gen = run(meridian_raw, use_general_cloud_optics=True, **kw)
    for nm in ("sw_up_clear", "sw_dn_clear", "lw_up_clear", "lw_dn_clear"):
        assert np.array_equal(gen[nm], base[nm]), nm
    for nm, bound in (("sw_up", 60.0), ("sw_dn", 80.0), ("lw_up", 15.0), ("lw_dn", 15.0)):
        d = np.abs(gen[nm] - base[nm])
        assert 0.0 < d.max() < bound and d.mean() < 0.1 * bound, (nm, d.max(), d.mean())
    assert np.array_equal(gen["cloud_cover_sw"], base["cloud_cover_sw"])


def test_general_cloud_optics_flag_is_required_by_ecckd(meridian_raw):
    """ecCKD has no band parameterisations to fall back to (the reference warns, radiation_config.F90:1155-1157, then stops on the band
    count in radiation_cloud_optics.F90:67-79)."""
    cfg = RadiationConfig(gas_model_name="ECCKD", do_nearest_spectral_lw_emiss=False, use_general_cloud_optics=False)
    with pytest.raises(ValueError, match="use_general_cloud_optics"):
        cfg.consolidate()


@pytest.mark.parametrize("kw", [dict(), dict(use_aerosols=True, **TC), dict(sw_solver_name="SPARTACUS", lw_solver_name="SPARTACUS", do_3d_effects=True)])
def test_mixed_gas_models(meridian_raw, kw):
    """One gas model per spectrum (radiation_interface.F90:333-355; the four runs of test/ifs `test_mixed_gas`).  The gas arrays are
    mass mixing ratios as soon as one spectrum uses RRTMG (set_gas_units, :164-186) and ecCKD scales them itself
    (radiation_ecckd.F90:518-623), so each spectrum of a mixed run reproduces the same spectrum of the run with that model in both:
    exactly for RRTMG, to the rounding of the unit conversion for ecCKD."""
    E = dict(do_nearest_spectral_lw_emiss=False, **kw)
    rrtmg = run(meridian_raw, use_general_cloud_optics=True, **E)
    ecckd = run(meridian_raw, gas_model_name="ECCKD", **E)
    sw_ckd = run(meridian_raw, sw_gas_model_name="ECCKD", **E)
    lw_ckd = run(meridian_raw, lw_gas_model_name="ECCKD", **E)
    for nm in ("lw_up", "lw_dn", "lw_up_clear", "lw_dn_clear"):
        assert np.array_equal(sw_ckd[nm], rrtmg[nm]), nm
        assert np.abs(lw_ckd[nm] - ecckd[nm]).max() < 1e-7, nm   # (SPARTACUS amplifies the rounding to a few 1e-9)
    for nm in ("sw_up", "sw_dn", "sw_dn_direct", "sw_up_clear", "sw_dn_clear"):
        assert np.array_equal(lw_ckd[nm], rrtmg[nm]), nm
        assert np.abs(sw_ckd[nm] - ecckd[nm]).max() < 1e-7, nm
    assert sw_ckd["sw_up_toa_g"].shape[0] == 32 and sw_ckd["lw_up_toa_g"].shape[0] == 140


def test_radiative_properties(meridian_raw):
    """orc_radiative_properties = the arrays radiation() hands to save_radiative_properties (radiation_interface.F90:405-425): without
    aerosols the gas part is the gas-optics stage dump, the SPARTACUS order is the permutation of radiation_ifs_rrtm.F90:50-68, a
    column range leaves the other columns alone, and the incoming flux sums to the solar irradiance."""
    cfg = RadiationConfig().consolidate()
    o = Oracle(cfg)
    p = o.radiative_properties(I.to_radiation_inputs(meridian_raw, cfg), 32, NLEV, istartcol=3, iendcol=30)
    for nm, a in p.items():
        assert np.isnan(a[..., :2]).all() and np.isnan(a[..., 30:]).all() and np.isfinite(a[..., 2:30]).all(), nm
    g = o.gas_optics_column(I.to_radiation_inputs(meridian_raw, cfg), 32, NLEV, 11)
    for nm in ("od_lw", "planck_hl", "lw_emission", "od_sw", "ssa_sw", "incoming_sw"):
        assert np.array_equal(p[nm][..., 10], np.asarray(g[nm]).T), nm   # (the stage dump is C-ordered [level][g])
    sun = meridian_raw["cos_solar_zenith_angle"][2:30] > 0.0
    assert np.abs(p["incoming_sw"][:, 2:30][:, sun].sum(axis=0) - float(meridian_raw["solar_irradiance"])).max() < 1e-9
    assert not p["g_sw"][..., 2:30].any() and not p["ssa_lw"][..., 2:30].any()
    # SPARTACUS on RRTMG: the same numbers at the reordered positions
    from ecrad_b200.tables import read_blob
    perm = read_blob(cfg.tables_path())["i_g_from_reordered_g_lw"].astype(int) - 1
    sp = RadiationConfig(sw_solver_name="SPARTACUS", lw_solver_name="SPARTACUS", do_3d_effects=True).consolidate()
    q = Oracle(sp).radiative_properties(I.to_radiation_inputs(meridian_raw, sp), 32, NLEV, istartcol=3, iendcol=30)
    assert np.array_equal(q["od_lw"][..., 2:30], p["od_lw"][perm][..., 2:30])
    assert np.array_equal(q["od_lw_cloud"][..., 2:30], p["od_lw_cloud"][..., 2:30])


def test_spectral_solar_cycle(meridian_raw):
    """use_spectral_solar_cycle (radiation_config.F90:174, read_spectral_solar_cycle radiation_ecckd.F90:295-451, calc_incoming_sw :935-964):
    the multiplier redistributes the incoming flux among the g-points (more ultraviolet at solar maximum) without changing its sum."""
    cfg = RadiationConfig(gas_model_name="ECCKD", do_nearest_spectral_lw_emiss=False).consolidate()
    o = Oracle(cfg)
    base = o.radiation(I.to_radiation_inputs(meridian_raw, cfg), 32, NLEV)
    p0 = o.radiative_properties(I.to_radiation_inputs(meridian_raw, cfg), 32, NLEV)["incoming_sw"]
    o.set_solar_cycle_multiplier(1.0)
    smax = o.radiation(I.to_radiation_inputs(meridian_raw, cfg), 32, NLEV)
    p1 = o.radiative_properties(I.to_radiation_inputs(meridian_raw, cfg), 32, NLEV)["incoming_sw"]
    o.set_solar_cycle_multiplier(-1.0)
    pm = o.radiative_properties(I.to_radiation_inputs(meridian_raw, cfg), 32, NLEV)["incoming_sw"]
    o.set_solar_cycle_multiplier(0.0)
    again = o.radiation(I.to_radiation_inputs(meridian_raw, cfg), 32, NLEV)
    assert np.array_equal(again["sw_dn"], base["sw_dn"])
    tsi = float(meridian_raw["solar_irradiance"])
    assert np.abs(p1.sum(axis=0) - tsi).max() < 1e-9 and np.abs(pm.sum(axis=0) - tsi).max() < 1e-9
    rel = p1[:, 0] / p0[:, 0] - 1.0
    assert 0.002 < rel.max() < 0.05 and -0.001 < rel.min() < 0.0      # a few 0.1 % more in the ultraviolet terms, slightly less elsewhere
    assert np.allclose(p1 + pm, 2.0 * p0, rtol=0, atol=1e-12)         # linear in the multiplier
    sun = meridian_raw["cos_solar_zenith_angle"] > 0.0
    assert np.abs(smax["sw_dn"][sun, 0] - base["sw_dn"][sun, 0]).max() < 1e-9     # the same flux enters ...
    d = np.abs(smax["sw_dn"][sun, -1] - base["sw_dn"][sun, -1])
    assert 0.0 < d.max() < 1.0                                                     # ... a little less of it reaches the surface
    assert np.array_equal(smax["lw_up"], base["lw_up"])
    # RRTMG-IFS has no solar-cycle information (radiation_config.F90:1200-1203)
    with pytest.raises(RuntimeError):
        Oracle(RadiationConfig().consolidate()).set_solar_cycle_multiplier(0.5)
