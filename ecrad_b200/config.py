"""Host-side mirror of the part of `config_type` the hot path needs (radiation/radiation_config.F90:163-649).

Key names are the reference's namelist keys (`&radiation`, radiation_config.F90:730-764); defaults are those of
test/ifs/configCY49R1.nam with use_aerosols=false (the `noaer` ctest, BASELINE config 2).  `consolidate()` restates
the two pieces of setup arithmetic whose results cross the C-ABI as tables:
  * config%sw_albedo_weights        radiation_config.F90:1947-2019 -> radiation_spectral_definition.F90 calc_mapping_from_bands
  * config%i_emiss_from_band_lw     radiation_config.F90:2025-2097
In a Fortran host these come straight out of `config_type`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import os

import numpy as np

from . import abi

# radiation_ifs_rrtm.F90:104-111 / :143-150 : RRTMG band limits (cm-1)
SW_WN1 = np.array([2600, 3250, 4000, 4650, 5150, 6150, 7700, 8050, 12850, 16000, 22650, 29000, 38000, 820], dtype=np.float64)
SW_WN2 = np.array([3250, 4000, 4650, 5150, 6150, 7700, 8050, 12850, 16000, 22650, 29000, 38000, 50000, 2600], dtype=np.float64)
LW_WN1 = np.array([10, 350, 500, 630, 700, 820, 980, 1080, 1180, 1390, 1480, 1800, 2080, 2250, 2380, 2600], dtype=np.float64)
LW_WN2 = np.array([350, 500, 630, 700, 820, 980, 1080, 1180, 1390, 1480, 1800, 2080, 2250, 2380, 2600, 3250], dtype=np.float64)
SOLAR_REF_T = 5777.0        # radiation_spectral_definition.F90:27
TERRESTRIAL_REF_T = 273.15  # radiation_spectral_definition.F90:28


def _planck_wavenumber(wn, temperature):
    """radiation_spectral_definition.F90 calc_planck_function_wavenumber (constants radiation_constants.F90)."""
    c, kb, h = 299792458.0, 1.380648813e-23, 6.6260695729e-34
    freq = 100.0 * c * np.asarray(wn, dtype=np.float64)
    pf = 2.0 * h * freq**3 / (c**2 * (np.exp(h * freq / (kb * temperature)) - 1.0))
    return pf * 100.0 * c


def mapping_from_bands(wn1_band, wn2_band, ref_temperature, wavelength_bound, i_intervals):
    """calc_mapping_from_bands(..., use_bands=.true.), radiation_spectral_definition.F90 (band branch).
    Returns mapping(ninput, nband)."""
    ninterval = len(i_intervals)
    ninput = max(i_intervals)
    nband = len(wn1_band)
    mapping = np.zeros((ninput, nband))
    weight = np.array([0.5, 1.0, 1.0, 1.0, 0.5])
    for jb in range(nband):
        for jint in range(1, ninterval + 1):
            wn2 = wn2_band[jb] if jint == 1 else min(wn2_band[jb], 0.01 / wavelength_bound[jint - 2])
            wn1 = wn1_band[jb] if jint == ninterval else max(wn1_band[jb], 0.01 / wavelength_bound[jint - 1])
            if wn2 > wn1:
                samp = wn1 + np.arange(5) * (wn2 - wn1) / 4.0
                pl = _planck_wavenumber(samp, ref_temperature)
                mapping[i_intervals[jint - 1] - 1, jb] += np.sum(pl * weight) * (wn2 - wn1)
    for jb in range(nband):
        mapping[:, jb] = mapping[:, jb] * (1.0 / np.sum(mapping[:, jb]))
    return mapping


@dataclass
class RadiationConfig:
    # names = namelist keys of &radiation (test/ifs/configCY49R1.nam)
    do_sw: bool = True
    do_lw: bool = True
    do_sw_direct: bool = True
    do_clear: bool = True
    sw_solver_name: str = "McICA"
    lw_solver_name: str = "McICA"
    gas_model_name: str = "RRTMG-IFS"
    # one gas model per spectrum (sw_gas_model_name / lw_gas_model_name, radiation_config.F90:712-713); None = gas_model_name.
    # RRTMG-IFS in one spectrum and ECCKD in the other is the mixed configuration of test/ifs/configCY49R1_mixed.nam
    sw_gas_model_name: str | None = None
    lw_gas_model_name: str | None = None
    # cloud optics from the generalised look-up tables (config%use_general_cloud_optics).  None = what the reference's namelists of
    # test/ifs say: false with RRTMG-IFS (configCY49R1.nam:37), true with ecCKD (configCY49R1_ecckd.nam:39)
    use_general_cloud_optics: bool | None = None
    liquid_model_name: str = "SOCRATES"
    ice_model_name: str = "Fu-IFS"
    overlap_scheme_name: str = "Exp-Ran"
    cloud_fraction_threshold: float = 0.001e-3
    cloud_mixing_ratio_threshold: float = 1.0e-9
    do_lw_aerosol_scattering: bool = False
    do_lw_cloud_scattering: bool = True
    cloud_inhom_decorr_scaling: float = 0.5
    cloud_pdf_shape_name: str = "Gamma"   # shape of the sub-grid cloud water PDF: regions of Tripleclouds / SPARTACUS (the McICA
                                          # look-up table of the shipped blob is the gamma one, data/mcica_gamma.nc)
    use_beta_overlap: bool = False
    use_vectorizable_generator: bool = False
    use_aerosols: bool = False
    do_save_spectral_flux: bool = True
    do_lw_derivatives: bool = True
    do_surface_sw_spectral_flux: bool = True
    do_toa_spectral_flux: bool = False
    do_fu_lw_ice_optics_bug: bool = False
    do_sw_delta_scaling_with_gases: bool = False
    do_canopy_fluxes_lw: bool = True
    do_canopy_fluxes_sw: bool = True
    do_nearest_spectral_sw_albedo: bool = False
    sw_albedo_wavelength_bound: tuple = (0.25e-6, 0.44e-6, 0.69e-6, 1.19e-6, 2.38e-6)
    i_sw_albedo_index: tuple = (1, 2, 3, 4, 5, 6)
    do_nearest_spectral_lw_emiss: bool = True
    lw_emiss_wavelength_bound: tuple = (8.0e-6, 13.0e-6)
    i_lw_emiss_index: tuple = (1, 2, 1)
    use_general_aerosol_optics: bool = True
    n_aerosol_types: int = 12
    i_aerosol_type_map: tuple = (-1, -2, -3, 7, 8, 9, -4, 10, 11, 11, -5, 14)
    min_gas_od_lw: float = 1.0e-15
    min_gas_od_sw: float = 0.0
    # SPARTACUS (radiation_config.F90:225-411 defaults; test/ifs `test_spartacus` sets do_3d_effects = true)
    do_3d_effects: bool = False
    n_regions: int = 3               # 2 = one homogeneous cloudy region (SPARTACUS only; test/i3rc `i3rc_spartacus2`)
    sw_entrapment_name: str = "Explicit"
    do_3d_lw_multilayer_effects: bool = False
    do_lw_side_emissivity: bool = True
    use_expm_everywhere: bool = False
    max_gas_od_3d: float = 8.0
    max_cloud_od: float = 16.0
    max_3d_transfer_rate: float = 10.0
    min_cloud_effective_size: float = 100.0
    overhead_sun_factor: float = 0.0
    overhang_factor: float = 0.0
    clear_to_thick_fraction: float = 0.0
    # gas_model_name = "ECCKD": table blob made by tools/extract_ecckd_tables.py (file name inside ecrad_b200/data/, or a
    # path); stands for gas_optics_{sw,lw}_override_file_name + the general cloud / aerosol optics files.  Cloud and
    # aerosol optics are per g-point (do_cloud_aerosol_per_{sw,lw}_g_point = true, radiation_config.F90 defaults for ecCKD).
    ecckd_tables: str = "ecckd_tables_32b.bin"
    derived: dict = field(default_factory=dict)
    n_g: tuple = (112, 140)      # (n_g_sw, n_g_lw), set by consolidate()
    n_bands: tuple = (14, 16)    # (n_bands_sw, n_bands_lw)

    @property
    def is_ecckd_sw(self):
        return (self.sw_gas_model_name or self.gas_model_name).lower() == "ecckd"

    @property
    def is_ecckd_lw(self):
        return (self.lw_gas_model_name or self.gas_model_name).lower() == "ecckd"

    @property
    def is_ecckd(self):
        """Both spectra on ecCKD: the gas arrays are volume mixing ratios then (set_gas_units, radiation_interface.F90:164-186)."""
        return self.is_ecckd_sw and self.is_ecckd_lw

    @property
    def is_mixed(self):
        return self.is_ecckd_sw != self.is_ecckd_lw

    def _data_path(self, name):
        here = os.path.dirname(os.path.abspath(__file__))
        return name if os.path.isabs(name) else os.path.join(here, "data", name)

    def tables_path(self):
        """The table blob `setup_radiation` loads for this configuration."""
        if self.is_ecckd:
            return self._data_path(self.ecckd_tables)
        if not self.is_mixed:
            return self._data_path("rrtmg_tables.bin")
        # mixed gas models: the RRTMG directory plus the ecCKD spectrum's model, cloud and aerosol tables (their names carry the
        # spectrum), what a host that ran both setup_gas_optics would register.  Written once next to the shipped blobs.
        from .tables import read_blob, write_blob
        spec = "sw" if self.is_ecckd_sw else "lw"
        src = self._data_path(self.ecckd_tables)
        out = self._data_path(os.path.join("_mixed", f"rrtmg_{spec}_{os.path.basename(src)}"))
        if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(self._data_path("rrtmg_tables.bin"))):
            tabs = read_blob(self._data_path("rrtmg_tables.bin"))
            for nm, arr in read_blob(src).items():
                if f"_{spec}_" in nm:
                    tabs[nm] = arr
            os.makedirs(os.path.dirname(out), exist_ok=True)
            tmp = f"{out}.{os.getpid()}.tmp"
            write_blob(tmp, tabs)
            os.replace(tmp, out)
        return out

    def consolidate(self):
        """Tables derived from the config (what `setup_radiation` stores in config_type)."""
        if self.do_canopy_fluxes_sw:   # radiation_config.F90:1119-1124
            self.do_surface_sw_spectral_flux = True
        any_ckd = self.is_ecckd_sw or self.is_ecckd_lw
        if any_ckd and self.use_general_cloud_optics is not None and not self.use_general_cloud_optics:
            raise ValueError("ecCKD gas optics needs use_general_cloud_optics = true: the band parameterisations are on the RRTMG "
                             "bands (the reference stops in radiation_cloud_optics.F90:67-79)")
        if any_ckd:
            # consolidate_sw_albedo_intervals / consolidate_lw_emiss_intervals (radiation_config.F90:1947-2100) with the
            # model's own spectral definition, one weight vector per g-point; bands == g-points
            # (radiation_ecckd_interface.F90:46-76)
            from .spectral import SpectralDefinition
            from .tables import read_blob
            tabs = read_blob(self._data_path(self.ecckd_tables))
        if self.is_ecckd_sw:
            sd_sw = SpectralDefinition.from_tables(tabs, "ckd_sw_")
            w = sd_sw.calc_mapping_from_bands(self.sw_albedo_wavelength_bound, self.i_sw_albedo_index)
            ng_sw = nb_sw = sd_sw.ng
        else:
            w = mapping_from_bands(SW_WN1, SW_WN2, SOLAR_REF_T, self.sw_albedo_wavelength_bound, self.i_sw_albedo_index)
            ng_sw, nb_sw = 112, 14
        if self.is_ecckd_lw:
            sd_lw = SpectralDefinition.from_tables(tabs, "ckd_lw_")
            e = sd_lw.calc_mapping_from_bands(self.lw_emiss_wavelength_bound, self.i_lw_emiss_index)
            ng_lw = nb_lw = sd_lw.ng
        else:
            e = mapping_from_bands(LW_WN1, LW_WN2, TERRESTRIAL_REF_T, self.lw_emiss_wavelength_bound, self.i_lw_emiss_index)
            ng_lw, nb_lw = 140, 16
        self.n_g, self.n_bands = (ng_sw, ng_lw), (nb_sw, nb_lw)
        self.derived = {
            "sw_albedo_weights": np.asfortranarray(w),                                   # (n_albedo, 14)
            "i_emiss_from_band_lw": (np.argmax(e, axis=0) + 1).astype(np.int32),         # maxloc(dim=1)
            "lw_emiss_weights": np.asfortranarray(e),
        }
        if self.do_nearest_spectral_sw_albedo:   # radiation_config.F90:1994-1997: maxloc(sw_albedo_weights, dim=1)
            self.derived["i_albedo_from_band_sw"] = (np.argmax(w, axis=0) + 1).astype(np.int32)
        if self.use_aerosols:
            # aerosol_optics%set_types (radiation_aerosol_optics_data.F90:605-633): >0 hydrophobic, <0 hydrophilic, 0 ignored
            m = np.array(self.i_aerosol_type_map[: self.n_aerosol_types], dtype=np.int32)
            self.derived["aerosol_iclass"] = np.where(m > 0, 1, np.where(m < 0, 2, 0)).astype(np.int32)
            self.derived["aerosol_itype"] = np.abs(m).astype(np.int32)
        return self

    def to_struct(self) -> abi.Config:
        if not self.derived:
            self.consolidate()
        c = abi.Config()
        c.struct_bytes = C.sizeof(abi.Config)
        c.i_solver_sw = abi.SOLVER[self.sw_solver_name.lower()]
        c.i_solver_lw = abi.SOLVER[self.lw_solver_name.lower()]
        c.i_gas_model_sw = abi.GAS_MODEL[(self.sw_gas_model_name or self.gas_model_name).lower()]
        c.i_gas_model_lw = abi.GAS_MODEL[(self.lw_gas_model_name or self.gas_model_name).lower()]
        c.i_overlap_scheme = abi.OVERLAP[self.overlap_scheme_name.lower()]
        c.i_liq_model = abi.LIQ_MODEL[self.liquid_model_name.lower()]
        c.i_ice_model = abi.ICE_MODEL[self.ice_model_name.lower()]
        for k in ("do_sw", "do_lw", "do_sw_direct", "do_clear", "use_aerosols", "do_lw_cloud_scattering",
                  "do_lw_aerosol_scattering", "do_lw_derivatives", "do_sw_delta_scaling_with_gases",
                  "do_fu_lw_ice_optics_bug", "use_beta_overlap", "use_vectorizable_generator",
                  "do_surface_sw_spectral_flux", "do_canopy_fluxes_sw", "do_canopy_fluxes_lw", "do_save_spectral_flux",
                  "do_nearest_spectral_sw_albedo", "do_nearest_spectral_lw_emiss",
                  "do_3d_effects", "do_3d_lw_multilayer_effects", "do_lw_side_emissivity", "use_expm_everywhere", "do_toa_spectral_flux"):
            setattr(c, k, int(getattr(self, k)))
        c.i_3d_sw_entrapment = abi.ENTRAPMENT[self.sw_entrapment_name.lower()]
        c.i_cloud_pdf_shape = abi.PDF_SHAPE[self.cloud_pdf_shape_name.lower()]
        c.n_regions = int(self.n_regions)
        c.use_general_cloud_optics = int((self.is_ecckd_sw or self.is_ecckd_lw) if self.use_general_cloud_optics is None else bool(self.use_general_cloud_optics))
        if c.i_cloud_pdf_shape == 0 and "mcica" in (self.sw_solver_name.lower(), self.lw_solver_name.lower()):
            raise ValueError("the shipped table blob holds the gamma PDF look-up table of the McICA generator (mcica_gamma.nc); "
                             "a lognormal McICA run needs 'pdf_val' from mcica_lognormal.nc")
        for k in ("max_gas_od_3d", "max_cloud_od", "max_3d_transfer_rate", "overhead_sun_factor", "overhang_factor",
                  "clear_to_thick_fraction"):
            setattr(c, k, float(getattr(self, k)))
        c.min_cloud_effective_size = max(1.0e-6, self.min_cloud_effective_size)   # radiation_config.F90:970
        # radiation_config.F90:1127-1132 consolidate: clouds matter if an active spectrum has a solver other than Cloudless
        c.do_clouds = int(bool((self.do_sw and c.i_solver_sw != 0) or (self.do_lw and c.i_solver_lw != 0)))
        c.n_g_sw, c.n_g_lw = self.n_g
        c.n_bands_sw, c.n_bands_lw = self.n_bands
        c.n_albedo_sw = self.derived["sw_albedo_weights"].shape[0]
        c.n_emiss_lw = max(self.i_lw_emiss_index)
        c.n_canopy_bands_sw = max(self.i_sw_albedo_index)
        c.n_canopy_bands_lw = max(self.i_lw_emiss_index)
        c.n_aerosol_types = self.n_aerosol_types if self.use_aerosols else 0
        c.cloud_fraction_threshold = self.cloud_fraction_threshold
        c.cloud_mixing_ratio_threshold = self.cloud_mixing_ratio_threshold
        c.min_gas_od_lw, c.min_gas_od_sw = self.min_gas_od_lw, self.min_gas_od_sw
        c.cloud_inhom_decorr_scaling = self.cloud_inhom_decorr_scaling
        return c
