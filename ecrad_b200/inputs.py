"""Column inputs for the hot path: reader for the reference's IFS-style netCDF input and the synthetic
IFS-shaped generator used by bench.py.

The arrays produced here are exactly what the reference's offline driver hands to `radiation()` after
driver/ecrad_driver_read_input.F90:21 and `set_gas_units` (radiation_interface.F90:164 ->
radiation_gas.F90 set_units_gas: vmr * M_gas/M_air), in Fortran memory order (column fastest).
"""
from __future__ import annotations

import numpy as np

SYNTH_BLOCK = 4096
AIR_MOLAR_MASS = 28.970  # radiation_gas_constants.F90
GAS_MOLAR_MASS = {"co2": 44.011, "n2o": 44.013, "ch4": 16.043, "cfc11": 137.3686, "cfc12": 120.914,
                  "hcfc22": 86.469, "ccl4": 153.823}

# variables copied from the reference input file into the small committed fixture
NC_VARS = ["solar_irradiance", "skin_temperature", "cos_solar_zenith_angle", "sw_albedo", "sw_albedo_direct",
           "lw_emissivity", "iseed", "pressure_hl", "temperature_hl", "q", "o3_mmr", "co2_vmr", "n2o_vmr", "ch4_vmr",
           "cfc11_vmr", "cfc12_vmr", "hcfc22_vmr", "ccl4_vmr", "cloud_fraction", "q_liquid", "q_ice", "re_liquid",
           "re_ice", "overlap_param", "fractional_std", "aerosol_mmr"]


def read_ifs_netcdf(path):
    """netCDF-3 classic reader (scipy) -> dict of raw file variables (float32 promoted exactly to float64)."""
    from scipy.io import netcdf_file

    raw = {}
    with netcdf_file(path, mmap=False) as f:
        for nm in NC_VARS:
            raw[nm] = np.array(f.variables[nm][...], dtype=np.float64)
    return raw


def to_radiation_inputs(raw, config=None):
    """File variables (C order: column slowest) -> radiation() inputs (Fortran order: column fastest).

    Gas arrays are what gas%mixing_ratio holds after set_gas_units (radiation_interface.F90:164-193): mass mixing ratios
    for RRTMG (radiation_ifs_rrtm.F90 set_gas_units), volume mixing ratios for ecCKD (radiation_ecckd_interface.F90:148-163);
    pass the RadiationConfig to get the latter.  The dict keys stay the C-ABI field names (`*_mmr`)."""
    F = np.asfortranarray
    vmr = config is not None and getattr(config, "is_ecckd", False)
    d = {
        "cos_sza": raw["cos_solar_zenith_angle"].copy(),
        "skin_temperature": raw["skin_temperature"].copy(),
        "sw_albedo": F(raw["sw_albedo"]), "sw_albedo_direct": F(raw["sw_albedo_direct"]),
        "lw_emissivity": F(raw["lw_emissivity"]),
        "iseed": np.asarray(raw["iseed"]).astype(np.int64).astype(np.int32),
        "pressure_hl": F(raw["pressure_hl"]), "temperature_hl": F(raw["temperature_hl"]),
        "h2o_mmr": F(raw["q"]), "o3_mmr": F(raw["o3_mmr"]),
        "cloud_fraction": F(raw["cloud_fraction"]), "q_liq": F(raw["q_liquid"]), "q_ice": F(raw["q_ice"]),
        "re_liq": F(raw["re_liquid"]), "re_ice": F(raw["re_ice"]),
        "overlap_param": F(raw["overlap_param"]), "fractional_std": F(raw["fractional_std"]),
    }
    d.update(set_gas_units(raw, vmr))
    d["solar_irradiance"] = float(raw["solar_irradiance"])
    if "aerosol_mmr" in raw:
        # file (column, type, level) -> aerosol%mixing_ratio(ncol, nlev, ntype)  (driver/ecrad_driver_read_input.F90:546)
        d["aerosol_mmr"] = F(np.transpose(raw["aerosol_mmr"], (0, 2, 1)))
        d["h2o_sat_liq"] = F(saturation_wrt_liquid(raw["pressure_hl"], raw["temperature_hl"]))
    if config is not None and "spartacus" in (config.sw_solver_name.lower(), config.lw_solver_name.lower()):
        if "inv_cloud_effective_size" in raw:   # given in the input file (driver/ecrad_driver_read_input.F90:334-360), e.g. the I3RC profile
            ic = ii = np.asarray(raw["inv_cloud_effective_size"], dtype=np.float64)
        else:
            ic, ii = cloud_effective_separation_eta(raw["pressure_hl"], raw["cloud_fraction"])
        d["inv_cloud_effective_size"], d["inv_inhom_effective_size"] = F(ic), F(ii)
    return d


def set_gas_units(raw, want_vmr):
    """What `set_gas_units` (radiation_interface.F90:164-193 -> gas%set_units, radiation_gas.F90:392-461) leaves in
    gas%mixing_ratio for the gases the hot path reads: mass mixing ratios for RRTMG-IFS (radiation_ifs_rrtm.F90 set_gas_units),
    volume mixing ratios for ecCKD (radiation_ecckd_interface.F90:148-163).  The input file gives H2O and O3 as mass mixing ratios
    (`q`, `o3_mmr`) and the well-mixed gases as volume mixing ratios (`*_vmr`).  Keys are the C-ABI field names (`*_mmr`)."""
    F = np.asfortranarray
    d = {"h2o_mmr": F(raw["q"]), "o3_mmr": F(raw["o3_mmr"])}
    if want_vmr:
        # radiation_gas.F90:444-451 set_units_gas: sf = 1*AirMolarMass/GasMolarMass for the two gases given as mass mixing ratio
        d["h2o_mmr"] = F(raw["q"] * (1.0 * AIR_MOLAR_MASS / 18.0152833))
        d["o3_mmr"] = F(raw["o3_mmr"] * (1.0 * AIR_MOLAR_MASS / 47.9982))
    for g, m in GAS_MOLAR_MASS.items():
        sf = 1.0 * m / AIR_MOLAR_MASS  # radiation_gas.F90 set_units_gas: sf = sf*GasMolarMass/AirMolarMass
        d[f"{g}_mmr"] = F(raw[f"{g}_vmr"]) if want_vmr else F(raw[f"{g}_vmr"] * sf)
    return d


def i3rc_raw(fix, sza_deg, overlap_decorr_length_scaling=1.13):
    """The reference's I3RC cumulus test (test/i3rc/Makefile + configI3RC.nam) as the raw-variable dict `to_radiation_inputs` reads:
    the single profile of i3rc_mls_cumulus.nc once per solar zenith angle (duplicate_profiles.sh), and the driver's
    overlap_decorr_length_scaling applied as driver/ecrad_driver_read_input.F90:247-255 does (overlap_param ** (1 / scaling) where
    positive).  `fix`: tests/golden/i3rc_mls_cumulus_inputs.npz (surface albedo 0.08 and 1366 W m-2 of the namelist already in it)."""
    sza = np.atleast_1d(np.asarray(sza_deg, dtype=np.float64))
    raw = {}
    for k, v in fix.items():
        v = np.array(v, dtype=np.float64)
        raw[k] = v if v.ndim == 0 else np.repeat(v[:1], len(sza), axis=0)
    raw["cos_solar_zenith_angle"] = np.cos(np.deg2rad(sza))
    raw["iseed"] = np.arange(1, len(sza) + 1, dtype=np.float64)
    op = raw["overlap_param"]
    raw["overlap_param"] = np.where(op > 0.0, np.abs(op) ** (1.0 / overlap_decorr_length_scaling), op)
    return raw


def ckdmip_raw(fix, mu0, sw_albedo=0.15, lw_emissivity=1.0, solar_irradiance=1361.0, n_albedo=6, n_emiss=2):
    """The reference's CKDMIP clear-sky test (test/ckdmip/config-*.nam + ckdmip_evaluation1_concentrations_present_reduced.nc) as the
    raw-variable dict `to_radiation_inputs` reads: what driver/ecrad_driver_read_input.F90 builds from that file -- gases from the
    `*_mole_fraction_fl` variables (vmr_suffix_str, :566-600), skin temperature = temperature of the lowest half-level (:477-478),
    cos_solar_zenith_angle / sw_albedo / lw_emissivity / solar irradiance overridden by the namelist, no clouds, no aerosols.
    `fix`: tests/golden/ckdmip_evaluation1.npz."""
    f64 = lambda a: np.array(a, dtype=np.float64)   # noqa: E731
    p = f64(fix["pressure_hl"])
    ncol, nlev = p.shape[0], p.shape[1] - 1
    zeros = np.zeros((ncol, nlev))
    raw = {
        "solar_irradiance": solar_irradiance, "skin_temperature": f64(fix["temperature_hl"])[:, -1].copy(),
        "cos_solar_zenith_angle": np.full(ncol, float(mu0)), "sw_albedo": np.full((ncol, n_albedo), sw_albedo),
        "sw_albedo_direct": np.full((ncol, n_albedo), sw_albedo), "lw_emissivity": np.full((ncol, n_emiss), lw_emissivity),
        "iseed": np.arange(1, ncol + 1, dtype=np.float64), "pressure_hl": p, "temperature_hl": f64(fix["temperature_hl"]),
        # H2O and O3 travel as mass mixing ratios in this dict (the IFS file convention); set_gas_units turns them back for ecCKD
        "q": f64(fix["h2o_mole_fraction_fl"]) * (18.0152833 / AIR_MOLAR_MASS), "o3_mmr": f64(fix["o3_mole_fraction_fl"]) * (47.9982 / AIR_MOLAR_MASS),
        "hcfc22_vmr": zeros.copy(), "ccl4_vmr": zeros.copy(),
        "cloud_fraction": zeros.copy(), "q_liquid": zeros.copy(), "q_ice": zeros.copy(), "re_liquid": np.full((ncol, nlev), 1.0e-5),
        "re_ice": np.full((ncol, nlev), 5.0e-5), "overlap_param": np.ones((ncol, nlev - 1)), "fractional_std": np.ones((ncol, nlev)),
    }
    for g in ("co2", "ch4", "n2o", "cfc11", "cfc12"):
        raw[f"{g}_vmr"] = f64(fix[f"{g}_mole_fraction_fl"])
    return raw


def cloud_effective_separation_eta(pressure_hl, cloud_fraction, separation_surf=2500.0, separation_toa=14000.0, power=3.5,
                                   inhom_separation_factor=0.75):
    """cloud%param_cloud_effective_separation_eta (radiation_cloud.F90:602-690) with the driver's namelist values of
    test/ifs/configCY49R1.nam (cloud_separation_scale_surface/_toa/_power, cloud_inhom_separation_factor;
    driver/ecrad_driver_read_input.F90:333-353): the inverse cloud and inhomogeneity effective sizes (m-1) the SPARTACUS
    solvers read, (ncol, nlev) each.  Arguments in file order (column slowest), levels top-down."""
    coeff_e = 1.0 - np.exp(-1.0)
    coeff_b = (separation_toa - separation_surf) / coeff_e
    coeff_a = separation_toa - coeff_b
    eta = (pressure_hl[:, :-1] + pressure_hl[:, 1:]) * (0.5 / pressure_hl[:, -1:])
    eff_separation = coeff_a + coeff_b * np.exp(-eta**power)
    f = cloud_fraction
    inv_cloud = 1.0 / (eff_separation * np.sqrt(np.maximum(1.0e-5, f * (1.0 - f))))
    inv_inhom = 1.0 / (eff_separation * inhom_separation_factor * np.sqrt(np.maximum(1.0e-5, 0.5 * f * (1.0 - 0.5 * f))))
    return inv_cloud, inv_inhom


def saturation_wrt_liquid(pressure_hl, temperature_hl):
    """thermodynamics%calc_saturation_wrt_liquid (radiation_thermodynamics.F90:118-158): driver-side preparation of the
    saturation mass mixing ratio used for the aerosol relative-humidity index (driver/ecrad_driver.F90:301)."""
    p = 0.5 * (pressure_hl[:, :-1] + pressure_hl[:, 1:])
    t = 0.5 * (temperature_hl[:, :-1] + temperature_hl[:, 1:])
    e_sat = 6.11e2 * np.exp(17.269 * (t - 273.16) / (t - 35.86))
    return np.minimum(1.0, 0.622 * e_sat / p)


def synthetic_columns(base_raw, ncol, seed=20261017, first=0):
    """IFS-shaped synthetic columns (BASELINE.md section 4 / SURVEY.md section 8d).

    Column i (global index first+i) = column (i mod 32) of the 32-column test slice; columns >= 32 are perturbed:
    T += U(-2,2) K, q *= exp N(0,0.1), o3 *= exp N(0,0.05), cloud fraction *= U(0.5,1.5) clipped to [0,1],
    q_liq,q_ice *= exp N(0,0.3), cos_sza = U(-0.2,1), T_skin += U(-3,3), iseed = i+1.
    Draws are made per block of 4096 global columns from a counter-based stream, so shards agree across ranks.
    """
    nbase = base_raw["pressure_hl"].shape[0]
    idx = (first + np.arange(ncol)) % nbase
    out = {}
    for k, v in base_raw.items():
        out[k] = v if np.ndim(v) == 0 else np.array(v[idx], dtype=np.float64)
    gidx = first + np.arange(ncol)
    pert = gidx >= nbase
    if pert.any():
        # perturbations are drawn per fixed block of global columns (counter-based Philox), so any shard
        # [first, first+ncol) of the same global problem sees identical columns on every rank
        B = SYNTH_BLOCK
        draws = {k: np.empty(ncol) for k in ("dT", "q", "o3", "cf", "ql", "qi", "mu0", "dTs")}
        for blk in range(first // B, (first + ncol - 1) // B + 1):
            rng = np.random.Generator(np.random.Philox(key=seed, counter=[0, 0, 0, blk]))
            b = {"dT": rng.uniform(-2.0, 2.0, B), "q": np.exp(rng.normal(0.0, 0.1, B)),
                 "o3": np.exp(rng.normal(0.0, 0.05, B)), "cf": rng.uniform(0.5, 1.5, B),
                 "ql": np.exp(rng.normal(0.0, 0.3, B)), "qi": np.exp(rng.normal(0.0, 0.3, B)),
                 "mu0": rng.uniform(-0.2, 1.0, B), "dTs": rng.uniform(-3.0, 3.0, B)}
            lo, hi = max(first, blk * B), min(first + ncol, (blk + 1) * B)
            for k in draws:
                draws[k][lo - first:hi - first] = b[k][lo - blk * B:hi - blk * B]
        m = pert
        out["temperature_hl"][m] += draws["dT"][m, None]
        out["q"][m] *= draws["q"][m, None]
        out["o3_mmr"][m] *= draws["o3"][m, None]
        out["cloud_fraction"][m] = np.clip(out["cloud_fraction"][m] * draws["cf"][m, None], 0.0, 1.0)
        out["q_liquid"][m] *= draws["ql"][m, None]
        out["q_ice"][m] *= draws["qi"][m, None]
        out["cos_solar_zenith_angle"][m] = draws["mu0"][m]
        out["skin_temperature"][m] += draws["dTs"][m]
        out["iseed"][m] = (gidx[m] + 1).astype(np.float64)
    return out
