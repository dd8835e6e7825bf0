"""SPARTACUS against a 3D Monte-Carlo benchmark: the reference's I3RC cumulus test (test/i3rc: Makefile target i3rc_spartacus with
configI3RC.nam, judged by plot_i3rc.m against i3rc_mls_cumulus_LIBRADTRAN.mat -- Hogan et al. 2016, Fig. 4).

libRadtran ran on the full 3D cloud field: DISORT in independent columns ("1D") and MYSTIC ("3D").  SPARTACUS sees the one
164-layer profile of cloud fraction, water content, fractional standard deviation and cloud effective size.  This is the only external
truth for the SPARTACUS solver (no reference output of ecRad itself exists for it), and it is sensitive to what makes SPARTACUS
SPARTACUS: the lateral transfer through cloud sides lowers the direct beam at the surface by 30-60 W m-2 at low sun and changes the
sign of the 3D effect on the reflected flux with solar zenith angle.  It caught the missing g-point reordering of
radiation_ifs_rrtm.F90:122-130 (without it only the first few g-points get 3D transfer and the effect all but vanishes).
Bounds: a little above what the restated solver achieves; the published agreement of SPARTACUS with MYSTIC is of this size
(MYSTIC's own standard error is up to 9.5 W m-2 on the reflected and 28 W m-2 on the direct flux).
"""
import os

import numpy as np
import pytest

from ecrad_b200 import inputs as I
from ecrad_b200.config import RadiationConfig
from oracle_lib import Oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# test/i3rc/configI3RC.nam (+ the Makefile's config_3reg_3d / config_3reg_1d variants)
I3RC = dict(sw_solver_name="SPARTACUS", lw_solver_name="SPARTACUS", do_3d_lw_multilayer_effects=True, min_cloud_effective_size=1.0e-6,
            sw_entrapment_name="Maximum", overhang_factor=1.0, overhead_sun_factor=0.06, cloud_inhom_decorr_scaling=0.5)


def load():
    fix = dict(np.load(os.path.join(GOLDEN, "i3rc_mls_cumulus_inputs.npz")))
    lib = dict(np.load(os.path.join(GOLDEN, "i3rc_libradtran.npz")))
    return fix, lib


def check_against_libradtran(run, fix, lib):
    """run(raw, **config) -> outputs for the eight sunlit zenith angles of the benchmark (0-85 degrees)."""
    sza = lib["sza"][:8]
    raw = I.i3rc_raw(fix, sza)
    s3 = run(raw, do_3d_effects=True, **I3RC)
    s1 = run(raw, do_3d_effects=False, **I3RC)
    up3, dir3, dn3 = s3["sw_up"][:, 0], s3["sw_dn_direct"][:, -1], s3["sw_dn"][:, -1]
    # clear sky: the gas optics and the surface against libRadtran's
    assert np.abs(s3["sw_up_clear"][:, 0] - lib["up_toa_clear"][:8]).max() <= 0.5
    assert (np.abs(s3["sw_dn_clear"][:, -1] / lib["dn_surf_clear"][:8] - 1.0) <= 0.02).all()   # (RRTMG absorbs 1 % more than libRadtran's gas model)
    # SPARTACUS with 3D effects against MYSTIC
    assert np.abs(up3 - lib["up_toa_3D"][:8]).max() <= 8.5, up3 - lib["up_toa_3D"][:8]
    assert np.abs(dir3 - lib["dn_direct_surf_3D"][:8]).max() <= 14.0, dir3 - lib["dn_direct_surf_3D"][:8]
    assert np.abs(dn3 - lib["dn_surf_3D"][:8]).max() <= 10.0, dn3 - lib["dn_surf_3D"][:8]
    # the 3D effect itself (plot_i3rc.m panels c/d): 3D minus the independent-column benchmark; negative for high sun (radiation
    # escapes through cloud sides), positive for low sun (cloud sides intercept the beam)
    eff, eff_lib = up3 - lib["up_toa_1D"][:8], (lib["up_toa_3D"] - lib["up_toa_1D"])[:8]
    assert np.abs(eff - eff_lib).max() <= 8.5, eff - eff_lib
    assert (eff[:4] < -8.0).all() and (eff[5:] > 5.0).all()
    # cloud sides shadow the surface at low sun: far more than any 1D scheme can (MYSTIC: -57 W m-2 at 60 degrees)
    d = dir3 - s1["sw_dn_direct"][:, -1]
    assert d[4] < -35.0 and d[5] < -30.0 and abs(d[0]) < 20.0, d
    # longwave: emission from cloud sides increases the downwelling flux at the surface by a few W m-2
    dl = s3["lw_dn"][0, -1] - s1["lw_dn"][0, -1]
    assert 2.0 < dl < 12.0, dl
    return s3


def test_oracle_spartacus_vs_mystic():
    fix, lib = load()

    def run(raw, **kw):
        cfg = RadiationConfig(**kw).consolidate()
        return Oracle(cfg).radiation(I.to_radiation_inputs(raw, cfg), len(raw["cos_solar_zenith_angle"]), 164)

    check_against_libradtran(run, fix, lib)


def check_two_regions(run, fix, lib):
    """test/i3rc `i3rc_spartacus2` (n_regions = 2): the homogeneous-cloud version of the same case; looser agreement with MYSTIC."""
    raw = I.i3rc_raw(fix, lib["sza"][:8])
    s2 = run(raw, do_3d_effects=True, n_regions=2, **I3RC)
    assert np.abs(s2["sw_up"][:, 0] - lib["up_toa_3D"][:8]).max() <= 8.5
    assert np.abs(s2["sw_dn_direct"][:, -1] - lib["dn_direct_surf_3D"][:8]).max() <= 24.0
    return s2


def test_oracle_two_region_spartacus_vs_mystic():
    fix, lib = load()

    def run(raw, **kw):
        cfg = RadiationConfig(**kw).consolidate()
        return Oracle(cfg).radiation(I.to_radiation_inputs(raw, cfg), len(raw["cos_solar_zenith_angle"]), 164)

    check_two_regions(run, fix, lib)


def test_per_g_outputs_are_in_the_reordered_sequence():
    """flux_type's g-point arrays of a SPARTACUS run are in the order the solver works in (config%i_g_from_reordered_g_*): the same
    numbers as a Tripleclouds run's, permuted, in clear sky (both solvers use the same clear-sky two-stream there)."""
    from ecrad_b200 import tables
    fix, _ = load()
    raw = I.i3rc_raw(fix, [30.0])
    raw["cloud_fraction"][:] = 0.0
    out = {}
    for name in ("SPARTACUS", "Tripleclouds"):
        cfg = RadiationConfig(sw_solver_name=name, lw_solver_name=name).consolidate()
        out[name] = Oracle(cfg).radiation(I.to_radiation_inputs(raw, cfg), 1, 164)
    blob = tables.read_blob(RadiationConfig().consolidate().tables_path())
    ps, pl = blob["i_g_from_reordered_g_sw"] - 1, blob["i_g_from_reordered_g_lw"] - 1
    a, b = out["SPARTACUS"]["sw_dn_direct_surf_g"][:, 0], out["Tripleclouds"]["sw_dn_direct_surf_g"][:, 0]
    assert np.abs(a - b[ps]).max() <= 1e-9 and np.abs(a - b).max() > 1.0
    a, b = out["SPARTACUS"]["lw_dn_surf_g"][:, 0], out["Tripleclouds"]["lw_dn_surf_g"][:, 0]
    assert np.abs(a - b[pl]).max() <= 5e-3 and np.abs(a - b).max() > 0.1
