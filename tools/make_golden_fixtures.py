#!/usr/bin/env python
"""Copy the reference's own regression fixtures into small committed files under tests/golden/.

Run in the authoring container (needs /root/reference); the GPU box only sees the .npz files.
  inputs : test/ifs/ecrad_meridian.nc                          -> tests/golden/ecrad_meridian_inputs.npz
  golden : test/ifs/ecrad_meridian_noaer_out_REFERENCE.nc      -> tests/golden/ecrad_meridian_noaer_ref.npz
           test/ifs/ecrad_meridian_cloudless_out_REFERENCE.nc  -> tests/golden/ecrad_meridian_cloudless_ref.npz
           ... and the default, expexp, tripleclouds, ecckd_mcica, ecckd_tc reference outputs likewise
  i3rc   : test/i3rc/i3rc_mls_cumulus.nc                       -> tests/golden/i3rc_mls_cumulus_inputs.npz
           test/i3rc/i3rc_mls_cumulus_LIBRADTRAN.mat            -> tests/golden/i3rc_libradtran.npz (DISORT-ICA and MYSTIC-3D fluxes)
  ckdmip : test/ckdmip/ckdmip_evaluation1_*_present_reduced.nc -> tests/golden/ckdmip_evaluation1.npz (50 clear-sky profiles + line-by-line fluxes)
Usage: make_golden_fixtures.py [reference root] [section ...]   (sections: meridian i3rc ckdmip; default all)
The golden outputs are float32 as written by the reference driver (do_write_double_precision=false); the
per-band profiles are kept at every half-level.
"""
import os
import sys

import numpy as np
from scipy.io import netcdf_file

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ecrad_b200.inputs import NC_VARS  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
BAND_LEVELS = list(range(138))   # every half-level (the per-band profiles pin each taumol band at every height)



def meridian():
    with netcdf_file(f"{REF}/test/ifs/ecrad_meridian.nc", mmap=False) as f:
        np.savez_compressed(f"{OUT}/ecrad_meridian_inputs.npz", **{k: np.array(f.variables[k][...]) for k in NC_VARS})
    for name in ("noaer", "cloudless", "default", "expexp", "tripleclouds", "ecckd_mcica", "ecckd_tc"):
        with netcdf_file(f"{REF}/test/ifs/ecrad_meridian_{name}_out_REFERENCE.nc", mmap=False) as f:
            d = {}
            for k, v in f.variables.items():
                a = np.array(v[...])
                if a.ndim == 3:  # (column, half_level, band)
                    a = a[:, BAND_LEVELS, :]
                d[k] = a
            d["band_levels"] = np.array(BAND_LEVELS)
            d["history"] = np.array(getattr(f, "history", b"").decode(errors="replace"))
            np.savez_compressed(f"{OUT}/ecrad_meridian_{name}_ref.npz", **d)


# ---- the reference's I3RC cumulus profile (test/i3rc/i3rc_mls_cumulus.nc): 164 layers, the SPARTACUS case of Hogan et al. (2016).
# Stored in the key format of ecrad_meridian_inputs.npz (what ecrad_b200.inputs.to_radiation_inputs reads), the single profile
# duplicated for eight of the solar zenith angles of test/i3rc/duplicate_profiles.sh; surface albedo 0.08 and the solar
# irradiance 1366 W m-2 of test/i3rc/configI3RC.nam in all six albedo intervals of the CY49R1 configuration.
def i3rc():
    with netcdf_file(f"{REF}/test/i3rc/i3rc_mls_cumulus.nc", mmap=False) as f:
        V = {k: np.array(v[...], dtype=np.float32) for k, v in f.variables.items()}
    cos_sza = np.array([1.0, 0.939693, 0.788011, 0.615661, 0.438371, 0.241922, 0.104528, 0.01], dtype=np.float32)
    n, nl = len(cos_sza), 164
    rep = lambda a: np.repeat(np.asarray(a, dtype=np.float32).reshape(1, -1), n, axis=0)  # noqa: E731
    full = lambda x: np.full((n, nl), np.float32(x), dtype=np.float32)  # noqa: E731
    i3 = {"solar_irradiance": np.float32(1366.0), "skin_temperature": np.repeat(V["skin_temperature"], n), "cos_solar_zenith_angle": cos_sza,
          "sw_albedo": np.full((n, 6), np.float32(0.08)), "sw_albedo_direct": np.full((n, 6), np.float32(0.08)),
          "lw_emissivity": np.full((n, 2), V["lw_emissivity"][0], dtype=np.float32), "iseed": np.arange(1, n + 1, dtype=np.float64),
          "pressure_hl": rep(V["pressure_hl"]), "temperature_hl": rep(V["temperature_hl"]), "q": rep(V["q"]), "o3_mmr": rep(V["o3_mmr"]),
          "co2_vmr": full(V["co2_vmr"]), "n2o_vmr": full(V["n2o_vmr"]), "ch4_vmr": full(V["ch4_vmr"]), "cfc11_vmr": full(V["cfc1_vmr"]),
          "cfc12_vmr": full(V["cfc2_vmr"]), "hcfc22_vmr": full(0.0), "ccl4_vmr": full(0.0),
          "cloud_fraction": rep(V["cloud_fraction"]), "q_liquid": rep(V["q_liquid"]), "q_ice": rep(V["q_ice"]), "re_liquid": rep(V["re_liquid"]),
          "re_ice": rep(V["re_ice"]), "overlap_param": rep(V["overlap_param"]), "fractional_std": rep(V["fractional_std"]),
          "inv_cloud_effective_size": rep(V["inv_cloud_effective_size"])}
    np.savez_compressed(f"{OUT}/i3rc_mls_cumulus_inputs.npz", **i3)
    print("i3rc_mls_cumulus_inputs.npz:", n, "columns,", nl, "layers")
    # the benchmark the reference's plot_i3rc.m judges SPARTACUS by (Hogan et al. 2016, Fig. 4): libRadtran on the full 3D cloud field,
    # DISORT in independent columns ("1D") and the MYSTIC Monte-Carlo model ("3D", with its standard error), nine solar zenith angles
    from scipy.io import loadmat
    m = loadmat(f"{REF}/test/i3rc/i3rc_mls_cumulus_LIBRADTRAN.mat")
    keep = ("sza", "up_toa_1D", "up_toa_3D", "up_toa_std_3D", "dn_surf_1D", "dn_surf_3D", "dn_direct_surf_1D", "dn_direct_surf_3D",
            "dn_direct_surf_std_3D")
    lib = {k: np.asarray(m[k], dtype=np.float64).ravel() for k in keep}
    lib["up_toa_clear"] = np.asarray(m["sw_up_clear"], dtype=np.float64)[-1]
    lib["dn_surf_clear"] = np.asarray(m["sw_dn_clear"], dtype=np.float64)[0]
    np.savez_compressed(f"{OUT}/i3rc_libradtran.npz", **lib)
    print("i3rc_libradtran.npz:", {k: v.shape for k, v in lib.items()})


# ---- CKDMIP "evaluation-1" clear-sky data set (test/ckdmip, Hogan & Matricardi 2020): 50 profiles x 54 layers with the line-by-line
# fluxes the reference's `make test` there is judged against (evaluate_ckd_lw_fluxes.m / evaluate_ckd_sw_fluxes.m): longwave, and
# shortwave at five solar zenith angles.  The only line-by-line truth in the reference tree: it pins every gas-optics model
# (RRTMG and the 32-, 64- and 96-term ecCKD models) against something that is not ecRad.
def ckdmip():
    d = {}
    with netcdf_file(f"{REF}/test/ckdmip/ckdmip_evaluation1_concentrations_present_reduced.nc", mmap=False) as f:
        for k in ("pressure_hl", "temperature_hl", "h2o_mole_fraction_fl", "o3_mole_fraction_fl", "co2_mole_fraction_fl", "ch4_mole_fraction_fl",
                  "n2o_mole_fraction_fl", "cfc11_mole_fraction_fl", "cfc12_mole_fraction_fl", "o2_mole_fraction_fl", "n2_mole_fraction_fl"):
            d[k] = np.array(f.variables[k][...])
    with netcdf_file(f"{REF}/test/ckdmip/ckdmip_evaluation1_lw_fluxes_present_reduced.nc", mmap=False) as f:
        for k in ("flux_up_lw", "flux_dn_lw"):
            d["lbl_" + k] = np.array(f.variables[k][...])
        assert np.array_equal(np.array(f.variables["pressure_hl"][...]), d["pressure_hl"])
    with netcdf_file(f"{REF}/test/ckdmip/ckdmip_evaluation1_sw_fluxes_present_reduced.nc", mmap=False) as f:
        for k in ("flux_up_sw", "flux_dn_sw", "flux_dn_direct_sw", "mu0"):
            d["lbl_" + k] = np.array(f.variables[k][...])
    np.savez_compressed(f"{OUT}/ckdmip_evaluation1.npz", **d)
    print("ckdmip_evaluation1.npz:", {k: v.shape for k, v in d.items()})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for section in sys.argv[2:] or ["meridian", "i3rc", "ckdmip"]:
        {"meridian": meridian, "i3rc": i3rc, "ckdmip": ckdmip}[section]()
    print(sorted(os.listdir(OUT)))
