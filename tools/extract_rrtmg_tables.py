#!/usr/bin/env python
"""Build the packed RRTMG/McICA/cloud-optics table blob from the reference's data + setup sources.

Run in the authoring container (needs /root/reference):
    python tools/extract_rrtmg_tables.py [--ref /root/reference] [--out ecrad_b200/data/rrtmg_tables.bin]

What it replaces: the table-filling half of `setup_radiation` for the RRTMG configuration
(radiation/radiation_interface.F90:37-156 -> radiation_ifs_rrtm.F90:34 setup_gas_optics ->
SURRTPK/SURRTRF/RRTM_INIT_140GP/SRTM_INIT; radiation_cloud_optics.F90 setup_cloud_optics;
radiation_pdf_sampler.F90:44 setup_pdf_sampler).  In a Fortran host the shim passes the very same arrays by
pointer; the standalone bench/tests load this blob instead.

Blob format ("ETB1"): see ecrad_b200/tables.py.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from minifortran import FArray, Interp  # noqa: E402
from ecrad_b200.tables import write_blob  # noqa: E402

NG_LW = [10, 12, 16, 14, 16, 8, 12, 8, 12, 6, 8, 8, 4, 2, 2, 2]          # ifsrrtm/yoerrtm.F90:61-76
NG_SW = [6, 12, 8, 8, 10, 10, 2, 10, 8, 6, 6, 8, 6, 12]                  # ifsrrtm/yoesrtm.F90:43-56


def rrtmg_tables(ref):
    it = Interp([os.path.join(ref, "ifsrrtm"), os.path.join(ref, "ifsaux")], os.path.join(ref, "data"))
    env = {"MPL_NPROC": lambda: 1, "MPL_MYRANK": lambda: 1, "CDIRECTORY": "."}
    # radiation_ifs_rrtm.F90:91-99
    it.run("SURRTPK")
    it.run("SURRTRF")
    it.run("RRTM_INIT_140GP", **env)
    it.run("SRTM_INIT", **env)

    out = {}

    # ---- LW
    for b in range(1, 17):
        mod = it.module(f"YOERRTA{b}")
        ng = NG_LW[b - 1]
        for name, v in mod.items():
            if name.startswith("__") or not isinstance(v, FArray):
                continue
            if name in ("ABSA", "ABSB"):
                continue  # EQUIVALENCE'd with KA/KB (yoerrta*.F90), rebuilt below
            a = v.a
            if name.startswith("FRACREF"):
                assert a.shape[0] == ng, (b, name, a.shape)   # FRACREFA(ng[,9]) : g-point first
            else:
                assert a.shape[-1] == ng, (b, name, a.shape)
            if name in ("KA", "KB"):
                a = a.reshape((-1, ng), order="F")  # = ABSA(65*nspa, ng) / ABSB(235*nspb, ng)
                name = "ABSA" if name == "KA" else "ABSB"
            out[f"lw{b}_{name}"] = a
    wn = it.module("YOERRTWN")
    out["lw_TOTPLNK"] = wn["TOTPLNK"].a
    out["lw_DELWAVE"] = wn["DELWAVE"].a
    out["lw_NSPA"] = wn["NSPA"].a.astype(np.int32)
    out["lw_NSPB"] = wn["NSPB"].a.astype(np.int32)
    rf = it.module("YOERRTRF")
    out["lw_PREFLOG"] = rf["PREFLOG"].a
    out["lw_TREF"] = rf["TREF"].a
    out["lw_CHI_MLS"] = rf["CHI_MLS"].a
    ftr = it.module("YOERRTFTR")
    out["lw_NGB"] = ftr["NGB"].a[:140].astype(np.int32)
    out["lw_NGC"] = ftr["NGC"].a.astype(np.int32)
    assert list(out["lw_NGC"]) == NG_LW

    # ---- SW
    for b in range(16, 30):
        mod = it.module(f"YOESRTA{b}")
        ng = NG_SW[b - 16]
        for name, v in mod.items():
            if name.startswith("__") or name in ("ABSA", "ABSB", "JPG"):
                continue
            if isinstance(v, FArray):
                # keep only the g-point-reduced ("C") arrays; the unreduced ones are setup intermediates
                if not name.endswith("C"):
                    continue
                a = v.a
                if name in ("KAC", "KBC"):
                    a = a.reshape((-1, a.shape[-1]), order="F")[:, :ng]
                    name = "ABSA" if name == "KAC" else "ABSB"
                elif name in ("SFLUXREFC", "RAYLAC") and a.ndim == 2:
                    a = a[:ng, :]
                else:
                    a = a[..., :ng] if a.ndim > 1 else a[:ng]
                out[f"sw{b}_{name}"] = a
            elif name in ("RAYL", "STRRAT", "STRRAT1", "GIVFAC", "SCALEKUR"):
                out[f"sw{b}_{name}"] = np.array([float(v)])
            elif name == "LAYREFFR":
                out[f"sw{b}_{name}"] = np.array([int(v)], dtype=np.int32)
    sw = it.module("YOESRTWN")
    out["sw_PREFLOG"] = sw["PREFLOG"].a
    out["sw_TREF"] = sw["TREF"].a
    out["sw_NSPA"] = sw["NSPA"].a.astype(np.int32)
    out["sw_NSPB"] = sw["NSPB"].a.astype(np.int32)
    out["sw_NGC"] = sw["NGC"].a.astype(np.int32)
    assert list(out["sw_NGC"]) == NG_SW
    out["sw_NGBSW"] = it.module("YOESRTM")["NGBSW"].a[:112].astype(np.int32)
    return out


def nc_tables(ref):
    """Cloud-optics coefficient files and the McICA PDF look-up table (netCDF-3 classic)."""
    from scipy.io import netcdf_file

    out = {}
    d = os.path.join(ref, "data")
    # radiation_cloud_optics.F90:46-110 (setup_cloud_optics): coeff arrays read as (nband, ncoeff)
    with netcdf_file(os.path.join(d, "socrates_droplet_scattering_rrtm.nc"), mmap=False) as f:
        out["liq_coeff_lw"] = np.array(f.variables["coeff_lw"][:], dtype=np.float64)
        out["liq_coeff_sw"] = np.array(f.variables["coeff_sw"][:], dtype=np.float64)
    with netcdf_file(os.path.join(d, "fu_ice_scattering_rrtm.nc"), mmap=False) as f:
        out["ice_coeff_lw"] = np.array(f.variables["coeff_lw"][:], dtype=np.float64)
        out["ice_coeff_sw"] = np.array(f.variables["coeff_sw"][:], dtype=np.float64)
    # the other parameterisations radiation_cloud_optics.F90 dispatches (file names: radiation_config.F90:1240-1283), stored next to
    # the default pair as "<name>.<model>"; setup picks the pair of config%i_liq_model / i_ice_model
    for tag, fn in (("slingo", "slingo_droplet_scattering_rrtm.nc"),):
        with netcdf_file(os.path.join(d, fn), mmap=False) as f:
            out[f"liq_coeff_lw.{tag}"] = np.array(f.variables["coeff_lw"][:], dtype=np.float64)
            out[f"liq_coeff_sw.{tag}"] = np.array(f.variables["coeff_sw"][:], dtype=np.float64)
    for tag, fn in (("baran", "baran_ice_scattering_rrtm.nc"), ("baran2016", "baran2016_ice_scattering_rrtm.nc"),
                    ("baran2017", "baran2017_ice_scattering_rrtm.nc"), ("yi", "yi_ice_scattering_rrtm.nc")):
        with netcdf_file(os.path.join(d, fn), mmap=False) as f:
            out[f"ice_coeff_lw.{tag}"] = np.array(f.variables["coeff_lw"][:], dtype=np.float64)
            out[f"ice_coeff_sw.{tag}"] = np.array(f.variables["coeff_sw"][:], dtype=np.float64)
            if "coeff_gen" in f.variables:
                out[f"ice_coeff_gen.{tag}"] = np.array(f.variables["coeff_gen"][:], dtype=np.float64)
    # radiation_pdf_sampler.F90:44-107 (setup_pdf_sampler)
    with netcdf_file(os.path.join(d, "mcica_gamma.nc"), mmap=False) as f:
        out["pdf_val"] = np.array(f.variables["x"][:], dtype=np.float64).T.copy()  # val(ncdf, nfsd), radiation_pdf_sampler.F90:83-93
        out["pdf_fsd"] = np.array(f.variables["fsd"][:], dtype=np.float64)
    return out


def calc_mapping_bands(wavenumber, wn1_band, wn2_band, reference_temperature):
    """radiation_spectral_definition.F90:222-338 calc_mapping, use_bands=.true. branch: mapping(nband, nwav)."""
    from ecrad_b200.config import _planck_wavenumber

    nwav, nband = len(wavenumber), len(wn1_band)
    planck_weight = _planck_wavenumber(wavenumber, reference_temperature)
    mapping = np.zeros((nband, nwav))
    for jb in range(nband):
        weight = np.zeros(nwav)
        for jw in range(nwav):
            if wn1_band[jb] <= wavenumber[jw] <= wn2_band[jb]:
                w1 = max(wn1_band[jb], 0.5 * (wavenumber[jw - 1] + wavenumber[jw])) if jw > 0 else wn1_band[jb]
                w2 = min(wn2_band[jb], 0.5 * (wavenumber[jw] + wavenumber[jw + 1])) if jw < nwav - 1 else wn2_band[jb]
                weight[jw] = (w2 - w1) * planck_weight[jw]
        if weight.sum() <= 0.0:
            if wavenumber[0] >= wn2_band[jb]:
                weight[0] = 1.0
            elif wavenumber[-1] <= wn1_band[jb]:
                weight[-1] = 1.0
            else:
                iw = 1
                while wavenumber[iw] < wn2_band[jb]:
                    iw += 1
                mid = 0.5 * (wn2_band[jb] + wn1_band[jb])
                weight[iw - 1] = planck_weight[iw - 1] * (wavenumber[iw] - mid)
                weight[iw] = planck_weight[iw] * (-wavenumber[iw - 1] + mid)
        mapping[jb, :] = weight / weight.sum()
    return mapping


def aerosol_tables(ref):
    """Band-averaged aerosol optical properties for the RRTMG bands: setup_general_aerosol_optics,
    radiation/radiation_aerosol_optics.F90:96-338, applied to data/aerosol_ifs_49R1_20230119.nc (the default of
    use_general_aerosol_optics=true, radiation_config.F90:1221-1240).  Arrays keep the reference's Fortran shapes:
    *_phobic(nband, ntype), *_philic(nband, nrh, ntype)."""
    from scipy.io import netcdf_file

    from ecrad_b200.config import LW_WN1, LW_WN2, SOLAR_REF_T, SW_WN1, SW_WN2, TERRESTRIAL_REF_T

    out = {}
    with netcdf_file(os.path.join(ref, "data", "aerosol_ifs_49R1_20230119.nc"), mmap=False) as f:
        g = lambda n: np.array(f.variables[n][:], dtype=np.float64)  # noqa: E731
        wn = g("wavenumber")
        phobic = {k: g(f"{k}_hydrophobic").T for k in ("mass_ext", "ssa", "asymmetry")}          # (nwav, ntype)
        philic = {k: np.transpose(g(f"{k}_hydrophilic"), (2, 1, 0)) for k in ("mass_ext", "ssa", "asymmetry")}  # (nwav, nrh, ntype)
        out["aer_rh_lower"] = g("relative_humidity1")
    for spec, wn1, wn2, tref in (("sw", SW_WN1, SW_WN2, SOLAR_REF_T), ("lw", LW_WN1, LW_WN2, TERRESTRIAL_REF_T)):
        m = calc_mapping_bands(wn, wn1, wn2, tref)
        me = m @ phobic["mass_ext"]
        ssa = (m @ (phobic["mass_ext"] * phobic["ssa"])) / me
        gg = (m @ (phobic["mass_ext"] * phobic["ssa"] * phobic["asymmetry"])) / (me * ssa)
        out[f"aer_mass_ext_{spec}_phobic"], out[f"aer_ssa_{spec}_phobic"], out[f"aer_g_{spec}_phobic"] = me, ssa, gg
        nrh, nty = philic["mass_ext"].shape[1:]
        me3 = np.zeros((len(wn1), nrh, nty)); ssa3 = np.zeros_like(me3); g3 = np.zeros_like(me3)
        for jt in range(nty):
            e, s_, a = (philic[k][:, :, jt] for k in ("mass_ext", "ssa", "asymmetry"))
            me3[:, :, jt] = m @ e
            ssa3[:, :, jt] = (m @ (e * s_)) / me3[:, :, jt]
            g3[:, :, jt] = (m @ (e * s_ * a)) / (me3[:, :, jt] * ssa3[:, :, jt])
        out[f"aer_mass_ext_{spec}_philic"], out[f"aer_ssa_{spec}_philic"], out[f"aer_g_{spec}_philic"] = me3, ssa3, g3
    return out


def general_cloud_tables(ref):
    """use_general_cloud_optics = true with RRTMG-IFS (the config_type default, radiation_config.F90:185): the look-up tables of
    radiation_general_cloud_optics.F90:39-110 (cloud types 1 = mie_droplet, 2 = baum-general-habit-mixture_ice) averaged to the 14 + 16
    RRTMG bands -- setup_general_cloud_optics with use_bands = .true. (RRTMG supports cloud properties per band only,
    radiation_config.F90:1078-1090) on the bands-only spectral definitions of radiation_ifs_rrtm.F90:108-114, :156-162, "thick"
    averaging (radiation_config.F90:352).  Same names as in the ecCKD blobs: gco_{sw,lw}_{0,1}_{meta,mass_ext,ssa,asymmetry}."""
    from extract_ecckd_tables import general_cloud_optics
    from ecrad_b200.config import LW_WN1, LW_WN2, SOLAR_REF_T, SW_WN1, SW_WN2, TERRESTRIAL_REF_T

    class Bands:
        def __init__(self, wn1, wn2, tref):
            self.wn1, self.wn2, self.tref = wn1, wn2, tref

        def calc_mapping(self, wavenumber):
            return calc_mapping_bands(wavenumber, self.wn1, self.wn2, self.tref)

    out = {}
    d = os.path.join(ref, "data")
    for jt, nm in enumerate(("mie_droplet", "baum-general-habit-mixture_ice")):
        for spec, sd in (("sw", Bands(SW_WN1, SW_WN2, SOLAR_REF_T)), ("lw", Bands(LW_WN1, LW_WN2, TERRESTRIAL_REF_T))):
            meta, me, ss, gg = general_cloud_optics(os.path.join(d, nm + "_scattering.nc"), sd, thick=True)
            out[f"gco_{spec}_{jt}_meta"] = meta
            out[f"gco_{spec}_{jt}_mass_ext"], out[f"gco_{spec}_{jt}_ssa"], out[f"gco_{spec}_{jt}_asymmetry"] = me, ss, gg
    return out


def gpoint_reordering(ref):
    """RRTM_GPOINT_REORDERING_SW / _LW (radiation_ifs_rrtm.F90:50-68): the order SPARTACUS wants the g-points in (approximately
    increasing gas optical depth), i.e. config%i_g_from_reordered_g_{sw,lw} when that spectrum's solver is SPARTACUS
    (radiation_ifs_rrtm.F90:122-130, :167-174); every other solver gets the identity and does not read these arrays."""
    import re

    src = open(os.path.join(ref, "radiation", "radiation_ifs_rrtm.F90")).read()
    out = {}
    for spec, n in (("SW", 112), ("LW", 140)):
        m = re.search(r"RRTM_GPOINT_REORDERING_%s\(%d\)\s*=\s*\(/(.*?)/\)" % (spec, n), src, re.S)
        vals = np.array([int(v) for v in re.findall(r"\d+", m.group(1).replace("&", " "))], dtype=np.int32)
        assert len(vals) == n and sorted(vals) == list(range(1, n + 1)), spec
        out[f"i_g_from_reordered_g_{spec.lower()}"] = vals
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "ecrad_b200", "data", "rrtmg_tables.bin"))
    args = ap.parse_args()
    tabs = rrtmg_tables(args.ref)
    tabs.update(nc_tables(args.ref))
    tabs.update(aerosol_tables(args.ref))
    tabs.update(gpoint_reordering(args.ref))
    tabs.update(general_cloud_tables(args.ref))
    write_blob(args.out, tabs)
    tot = sum(v.nbytes for v in tabs.values())
    print(f"wrote {args.out}: {len(tabs)} arrays, {tot/1e6:.2f} MB")


if __name__ == "__main__":
    main()
