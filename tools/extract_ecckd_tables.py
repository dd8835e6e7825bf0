#!/usr/bin/env python
"""Build the packed table blob for an ecCKD gas-optics configuration from the reference's data files.

Run in the authoring container (needs /root/reference):
    python tools/extract_ecckd_tables.py                       # 32-term defaults -> ecrad_b200/data/ecckd_tables_32b.bin
    python tools/extract_ecckd_tables.py --lw ecckd-1.2_lw_climate_narrow-64b_ckd-definition.nc \
        --sw ecckd-1.2_sw_climate_window-64b_ckd-definition.nc --out ecrad_b200/data/ecckd_tables_64b.bin

What it replaces: the table-filling half of `setup_radiation` for gas_model_name="ECCKD" with generalised cloud and
aerosol optics, everything per g-point (do_cloud_aerosol_per_{sw,lw}_g_point = true, the default):
    radiation_ecckd_interface.F90:26 setup_gas_optics -> radiation_ecckd.F90:128 read_ckd_model, radiation_ecckd_gas.F90:83
    radiation_general_cloud_optics.F90:32 -> radiation_general_cloud_optics_data.F90:71 (mie_droplet, baum-general-habit-mixture_ice)
    radiation_aerosol_optics.F90:96 setup_general_aerosol_optics (aerosol_ifs_49R1_20230119.nc)
    radiation_pdf_sampler.F90:44 (mcica_gamma.nc)
The spectral definition of each model is stored too, so that the host can derive config%sw_albedo_weights and
config%lw_emiss_weights for its albedo / emissivity intervals (ecrad_b200/spectral.py).
Array shapes are the reference's Fortran shapes (first index fastest).
"""
import argparse
import os
import sys

import numpy as np
from scipy.io import netcdf_file

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ecrad_b200.spectral import SpectralDefinition  # noqa: E402
from ecrad_b200.tables import write_blob  # noqa: E402

GAS_CODES = {"h2o": 1, "co2": 2, "o3": 3, "n2o": 4, "co": 5, "ch4": 6, "o2": 7, "cfc11": 8, "cfc12": 9, "hcfc22": 10,
             "ccl4": 11, "no2": 12}                                   # radiation_gas_constants.F90:26-39
CONC_NONE, CONC_LINEAR, CONC_LUT, CONC_RELATIVE_LINEAR = 0, 1, 2, 3   # radiation_ecckd_gas.F90:27-32
MAX_GASES = 12


def f64(v):
    return np.array(v[:], dtype=np.float64)


def read_ckd_model(path, prefix):
    """read_ckd_model (radiation_ecckd.F90:128-226).  Returns (tables, SpectralDefinition)."""
    out = {}
    with netcdf_file(path, mmap=False) as f:
        V = f.variables
        pressure = f64(V["pressure"])
        temperature_full = f64(V["temperature"]).T          # Fortran (npress, ntemp)
        npress, ntemp = temperature_full.shape
        log_pressure1 = np.log(pressure[0])
        d_log_pressure = np.log(pressure[1]) - log_pressure1
        d_temperature = temperature_full[0, 1] - temperature_full[0, 0]
        out[prefix + "temperature1"] = temperature_full[:, 0].copy()
        is_sw = "solar_irradiance" in V
        nplanck, t1_planck, dt_planck = 0, 0.0, 1.0
        if is_sw:
            ssi = f64(V["solar_irradiance"])
            out[prefix + "norm_solar_irradiance"] = ssi / ssi.sum()
            out[prefix + "rayleigh_molar_scat"] = f64(V["rayleigh_molar_scattering_coeff"])
        else:
            tp = f64(V["temperature_planck"])
            nplanck, t1_planck, dt_planck = len(tp), tp[0], tp[1] - tp[0]
            out[prefix + "planck_function"] = f64(V["planck_function"]).T   # (ng, nplanck)
        sd = SpectralDefinition(f64(V["wavenumber1"]), f64(V["wavenumber2"]), f64(V["gpoint_fraction"]).T,
                                f64(V["wavenumber1_band"]), f64(V["wavenumber2_band"]),
                                np.array(V["band_number"][:], dtype=np.int32) + 1,
                                f64(V["solar_spectral_irradiance"]) if "solar_spectral_irradiance" in V else None,
                                f64(V["solar_irradiance"]) if is_sw else None)
        out.update(sd.to_tables(prefix))
        ngas = int(V["n_gases"][()])
        names = f._attributes["constituent_id"].decode().split()
        assert len(names) == ngas
        gas_meta = np.zeros((6, ngas))   # code, conc dependence, reference mole fraction, n_mole_frac, log_mole_frac1, d_log_mole_frac
        for j, nm in enumerate(names):
            code = GAS_CODES.get(nm, 0)   # "composite" -> 0 (well-mixed background gases, no concentration dependence)
            dep = int(V[nm + "_conc_dependence_code"][()])
            k = f64(V[nm + "_molar_absorption_coeff"])
            gas_meta[0, j], gas_meta[1, j] = code, dep
            if dep == CONC_LUT:
                mf = f64(V[nm + "_mole_fraction"])
                gas_meta[3, j] = len(mf)
                gas_meta[4, j] = np.log(mf[0])
                gas_meta[5, j] = (np.log(mf[-1]) - np.log(mf[0])) / (len(mf) - 1)
                out[f"{prefix}gas{j}_molar_abs"] = np.transpose(k, (3, 2, 1, 0)).copy()   # (ng, npress, ntemp, nconc)
            else:
                out[f"{prefix}gas{j}_molar_abs"] = np.transpose(k, (2, 1, 0)).copy()      # (ng, npress, ntemp)
            if dep == CONC_RELATIVE_LINEAR:
                gas_meta[2, j] = float(V[nm + "_reference_mole_fraction"][()])
        out[prefix + "gas_meta"] = gas_meta
        out[prefix + "meta"] = np.array([sd.ng, npress, ntemp, nplanck, ngas, log_pressure1, d_log_pressure, d_temperature,
                                         t1_planck, dt_planck, 1.0 if is_sw else 0.0])
    return out, sd


def delta_eddington(od, ssa, g):
    """radiation_delta_eddington.h:23-37 (elemental, intensive form)."""
    f = g * g
    od2 = od * (1.0 - ssa * f)
    ssa2 = ssa * (1.0 - f) / (1.0 - ssa * f)
    return od2, ssa2, g / (1.0 + g)


def revert_delta_eddington(od, ssa, g):
    """radiation_delta_eddington.h:76-96."""
    g = g / (1.0 - g)
    f = g * g
    ssa = ssa / (1.0 - f + f * ssa)
    od = od / (1.0 - ssa * f)
    return od, ssa, g


def general_cloud_optics(path, sd, thick):
    """setup_general_cloud_optics, radiation_general_cloud_optics_data.F90:71-243, per g-point.
    Returns meta [n_effective_radius, effective_radius_0, d_effective_radius] and mass_ext/ssa/asymmetry (ng, nre)."""
    with netcdf_file(path, mmap=False) as f:
        V = f.variables
        wavenumber = f64(V["wavenumber"])
        re = f64(V["effective_radius"])
        mass_ext = f64(V["mass_extinction_coefficient"]).T     # Fortran (nwav, nre)
        ssa = f64(V["single_scattering_albedo"]).T
        asym = f64(V["asymmetry_factor"]).T
    assert mass_ext.shape == (len(wavenumber), len(re)), mass_ext.shape
    mapping = sd.calc_mapping(wavenumber)
    mass_ext, ssa, asym = delta_eddington(mass_ext, ssa, asym)
    me = mapping @ mass_ext
    ss = (mapping @ (mass_ext * ssa)) / me
    gg = (mapping @ (mass_ext * ssa * asym)) / (me * ss)
    if thick:   # Edwards & Slingo (1996) eqs 17-19
        ref_inf = np.sqrt((1.0 - ssa) / (1.0 - ssa * asym))
        ref_inf = (1.0 - ref_inf) / (1.0 + ref_inf)
        ss = mapping @ ref_inf
        ss = 4.0 * ss / ((1.0 + ss) ** 2 - gg * (1.0 - ss) ** 2)
    me, ss, gg = revert_delta_eddington(me, ss, gg)
    return np.array([len(re), re[0], re[1] - re[0]]), me, ss, gg


def aerosol_tables(path, sd_sw, sd_lw):
    """setup_general_aerosol_optics (radiation_aerosol_optics.F90:96-338) per g-point: *_phobic(ng, ntype), *_philic(ng, nrh, ntype)."""
    out = {}
    with netcdf_file(path, mmap=False) as f:
        g = lambda n: f64(f.variables[n])  # noqa: E731
        wn = g("wavenumber")
        phobic = {k: g(f"{k}_hydrophobic").T for k in ("mass_ext", "ssa", "asymmetry")}
        philic = {k: np.transpose(g(f"{k}_hydrophilic"), (2, 1, 0)) for k in ("mass_ext", "ssa", "asymmetry")}
        out["aer_rh_lower"] = g("relative_humidity1")
    for spec, sd in (("sw", sd_sw), ("lw", sd_lw)):
        m = sd.calc_mapping(wn)
        me = m @ phobic["mass_ext"]
        ssa = (m @ (phobic["mass_ext"] * phobic["ssa"])) / me
        gg = (m @ (phobic["mass_ext"] * phobic["ssa"] * phobic["asymmetry"])) / (me * ssa)
        out[f"aer_mass_ext_{spec}_phobic"], out[f"aer_ssa_{spec}_phobic"], out[f"aer_g_{spec}_phobic"] = me, ssa, gg
        nrh, nty = philic["mass_ext"].shape[1:]
        me3 = np.zeros((sd.ng, nrh, nty)); ssa3 = np.zeros_like(me3); g3 = np.zeros_like(me3)
        for jt in range(nty):
            e, s_, a = (philic[k][:, :, jt] for k in ("mass_ext", "ssa", "asymmetry"))
            me3[:, :, jt] = m @ e
            ssa3[:, :, jt] = (m @ (e * s_)) / me3[:, :, jt]
            g3[:, :, jt] = (m @ (e * s_ * a)) / (me3[:, :, jt] * ssa3[:, :, jt])
        out[f"aer_mass_ext_{spec}_philic"], out[f"aer_ssa_{spec}_philic"], out[f"aer_g_{spec}_philic"] = me3, ssa3, g3
    return out


def solar_cycle_amplitude(path, sd, norm_solar_irradiance):
    """read_spectral_solar_cycle (radiation_ecckd.F90:295-451, use_updated_solar_spectrum = false): the solar-cycle amplitude of the
    spectral solar irradiance (data/ssi_nrl2.nc) interpolated to the model's wavenumber grid, mapped to the g-points, and shifted so that
    it does not change the total (the caller scales by the total solar irradiance).  calc_incoming_sw (:935-964) then gives
    incoming = TSI * (norm_solar_irradiance + spectral_solar_cycle_multiplier * norm_amplitude_solar_irradiance)."""
    with netcdf_file(path, mmap=False) as f:
        wavenumber = f64(f.variables["wavenumber"])
        ssi = f64(f.variables["mean_solar_spectral_irradiance"])
        ssi_amplitude = f64(f.variables["ssi_solar_cycle_amplitude"])
    grid = 0.5 * (sd.wavenumber1 + sd.wavenumber2)
    dwav = sd.wavenumber2[0] - sd.wavenumber1[0]
    ssi_grid, amp_grid = np.zeros(sd.nwav), np.zeros(sd.nwav)
    for j, wn in enumerate(grid):   # linear interpolation, first bracketing pair (:379-393)
        k = np.nonzero((wavenumber[:-1] < wn) & (wavenumber[1:] >= wn))[0]
        if len(k):
            k = k[0]
            dw = wavenumber[k + 1] - wavenumber[k]
            ssi_grid[j] = (ssi[k] * (wavenumber[k + 1] - wn) + ssi[k + 1] * (wn - wavenumber[k])) * dwav / dw
            amp_grid[j] = (ssi_amplitude[k] * (wavenumber[k + 1] - wn) + ssi_amplitude[k + 1] * (wn - wavenumber[k])) * dwav / dw
    amp = norm_solar_irradiance * (amp_grid @ sd.gpoint_fraction) / (ssi_grid @ sd.gpoint_fraction)   # :421-424
    return (norm_solar_irradiance + amp) / (norm_solar_irradiance + amp).sum() - norm_solar_irradiance   # :428-431


def build(ref, lw_file, sw_file):
    d = os.path.join(ref, "data")
    tabs = {}
    lw, sd_lw = read_ckd_model(os.path.join(d, lw_file), "ckd_lw_")
    sw, sd_sw = read_ckd_model(os.path.join(d, sw_file), "ckd_sw_")
    tabs.update(lw); tabs.update(sw)
    # use_spectral_solar_cycle (radiation_ecckd_interface.F90:79-82): ssi_nrl2.nc, radiation_config.F90:1200-1203
    tabs["ckd_sw_norm_amplitude_solar_irradiance"] = solar_cycle_amplitude(os.path.join(d, "ssi_nrl2.nc"), sd_sw, tabs["ckd_sw_norm_solar_irradiance"])
    # cloud types 1 (liquid) and 2 (ice): radiation_general_cloud_optics.F90:62-71, thick averaging (radiation_config.F90:352)
    for jt, nm in enumerate(("mie_droplet", "baum-general-habit-mixture_ice")):
        for spec, sd in (("sw", sd_sw), ("lw", sd_lw)):
            meta, me, ss, gg = general_cloud_optics(os.path.join(d, nm + "_scattering.nc"), sd, thick=True)
            tabs[f"gco_{spec}_{jt}_meta"] = meta
            tabs[f"gco_{spec}_{jt}_mass_ext"], tabs[f"gco_{spec}_{jt}_ssa"], tabs[f"gco_{spec}_{jt}_asymmetry"] = me, ss, gg
    tabs.update(aerosol_tables(os.path.join(d, "aerosol_ifs_49R1_20230119.nc"), sd_sw, sd_lw))
    with netcdf_file(os.path.join(d, "mcica_gamma.nc"), mmap=False) as f:   # radiation_pdf_sampler.F90:44-107
        tabs["pdf_val"] = f64(f.variables["x"]).T.copy()
        tabs["pdf_fsd"] = f64(f.variables["fsd"])
    return tabs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--lw", default="ecckd-1.0_lw_climate_fsck-32b_ckd-definition.nc")   # radiation_config.F90:1194-1195
    ap.add_argument("--sw", default="ecckd-1.4_sw_climate_rgb-32b_ckd-definition.nc")    # radiation_config.F90:1176-1177
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "ecrad_b200", "data", "ecckd_tables_32b.bin"))
    args = ap.parse_args()
    tabs = build(args.ref, args.lw, args.sw)
    write_blob(args.out, tabs)
    print(f"wrote {args.out}: {len(tabs)} arrays, {sum(np.asarray(v).nbytes for v in tabs.values())/1e6:.2f} MB")


if __name__ == "__main__":
    main()
