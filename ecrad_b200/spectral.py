"""Spectral definition of an ecCKD gas-optics model and the setup-time mappings built from it.

Host-side mirror of radiation/radiation_spectral_definition.F90 (type spectral_definition_type), g-point branches:
    find                      :198-213
    calc_mapping              :222-486   (cloud / aerosol spectra -> g-points)
    calc_mapping_from_bands   :507-791   (albedo / emissivity intervals -> g-points; non-coarse variant)
Only setup-time work (run once per model, on the host); the arrays it produces travel in the table directory.
"""
import numpy as np

SOLAR_REF_T = 5777.0        # radiation_spectral_definition.F90:27
TERRESTRIAL_REF_T = 273.15  # radiation_spectral_definition.F90:28


def planck_wavenumber(wn, temperature):
    """calc_planck_function_wavenumber, radiation_spectral_definition.F90:1094-1116 (constants radiation_constants.F90)."""
    c, kb, h = 299792458.0, 1.380648813e-23, 6.6260695729e-34
    freq = 100.0 * c * np.asarray(wn, dtype=np.float64)
    pf = 2.0 * h * freq**3 / (c**2 * (np.exp(h * freq / (kb * temperature)) - 1.0))
    return pf * 100.0 * c


class SpectralDefinition:
    def __init__(self, wavenumber1, wavenumber2, gpoint_fraction, wavenumber1_band, wavenumber2_band, i_band_number,
                 solar_spectral_irradiance=None, solar_irradiance=None):
        self.wavenumber1 = np.asarray(wavenumber1, dtype=np.float64)
        self.wavenumber2 = np.asarray(wavenumber2, dtype=np.float64)
        self.gpoint_fraction = np.asarray(gpoint_fraction, dtype=np.float64)   # (nwav, ng)
        self.wavenumber1_band = np.asarray(wavenumber1_band, dtype=np.float64)
        self.wavenumber2_band = np.asarray(wavenumber2_band, dtype=np.float64)
        self.i_band_number = np.asarray(i_band_number, dtype=np.int32)         # 1-based
        self.solar_spectral_irradiance = None if solar_spectral_irradiance is None else np.asarray(solar_spectral_irradiance, np.float64)
        self.solar_irradiance = None if solar_irradiance is None else np.asarray(solar_irradiance, np.float64)
        self.reference_temperature = SOLAR_REF_T if solar_irradiance is not None else TERRESTRIAL_REF_T
        self.nwav, self.ng = self.gpoint_fraction.shape
        self.nband = len(self.wavenumber1_band)

    def find(self, wn):
        """1-based index of the wavenumber interval containing wn, 0 if outside (:198-213)."""
        if wn < self.wavenumber1[0] or wn > self.wavenumber2[-1]:
            return 0
        i = 1
        while wn > self.wavenumber2[i - 1] and i < self.nwav:
            i += 1
        return i

    def _spectral_weight(self):
        if self.solar_spectral_irradiance is not None:
            return self.solar_spectral_irradiance
        return planck_wavenumber(0.5 * (self.wavenumber1 + self.wavenumber2), self.reference_temperature)

    def calc_mapping(self, wavenumber):
        """mapping(ng, nwav_in) such that y = mapping @ x maps a property x sampled at `wavenumber` (cm-1, increasing)
        to the g-points (:356-482).  1-based indices kept in the comments; arrays are 0-based."""
        wavenumber = np.asarray(wavenumber, dtype=np.float64)
        nw = len(wavenumber)
        w1, w2 = self.wavenumber1, self.wavenumber2
        planck_weight = self._spectral_weight()
        mapping = np.zeros((self.ng, nw))
        for jw in range(nw):
            weight = np.zeros(self.nwav)
            wavenum1 = wavenumber[jw]
            isd1 = self.find(wavenum1)
            if isd1 < 1:
                continue
            k1 = isd1 - 1
            if jw > 0:
                wavenum0 = wavenumber[jw - 1]
                isd0 = self.find(wavenum0)
                k0 = isd0 - 1
                if isd0 == isd1:
                    weight[k0] = 0.5 * (wavenum1 - wavenum0) / (w2[k0] - w1[k0])
                else:
                    if isd0 >= 1:
                        weight[k0] = 0.5 * (w2[k0] - wavenum0) ** 2 / ((w2[k0] - w1[k0]) * (wavenum1 - wavenum0))
                    weight[k1] = 0.5 * (1.0 + (w1[k1] - wavenum1) / (wavenum1 - wavenum0)) * (wavenum1 - w1[k1]) / (w2[k1] - w1[k1])
                    if isd1 - isd0 > 1:
                        for isd in range(isd0 + 1, isd1):
                            k = isd - 1
                            weight[k] = 0.5 * (w1[k] + w2[k] - 2.0 * wavenum0) / (wavenum1 - wavenum0)
            else:
                weight[: k1] = 1.0
                weight[k1] = (wavenum1 - w1[k1]) / (w2[k1] - w1[k1])
            if jw < nw - 1:
                wavenum2 = wavenumber[jw + 1]
                isd2 = self.find(wavenum2)
                k2 = isd2 - 1
                if isd1 == isd2:
                    weight[k1] += 0.5 * (wavenum2 - wavenum1) / (w2[k1] - w1[k1])
                else:
                    if 1 <= isd2 <= self.nwav:
                        weight[k2] += 0.5 * (wavenum2 - w1[k2]) ** 2 / ((w2[k2] - w1[k2]) * (wavenum2 - wavenum1))
                    weight[k1] += 0.5 * (1.0 + (wavenum2 - w2[k1]) / (wavenum2 - wavenum1)) * (w2[k1] - wavenum1) / (w2[k1] - w1[k1])
                    if isd2 - isd1 > 1:
                        for isd in range(isd1 + 1, isd2):
                            k = isd - 1
                            weight[k] += 0.5 * (2.0 * wavenum2 - w1[k] - w2[k]) / (wavenum2 - wavenum1)
            else:
                weight[k1 + 1:] = 1.0
                weight[k1] = (w2[k1] - wavenum1) / (w2[k1] - w1[k1])
            weight = weight * planck_weight
            mapping[:, jw] = (weight[:, None] * self.gpoint_fraction).sum(axis=0)
        for jg in range(self.ng):
            mapping[jg, :] = mapping[jg, :] * (1.0 / mapping[jg, :].sum())
        return mapping

    def calc_mapping_from_bands(self, wavelength_bound, i_intervals):
        """mapping(ninput, ng): weights of the ninput albedo/emissivity values in each g-point (:700-790, the variant
        compiled without USE_COARSE_MAPPING), normalised per g-point."""
        ninterval = len(i_intervals)
        ninput = int(max(i_intervals))
        w1, w2 = self.wavenumber1, self.wavenumber2
        planck = self._spectral_weight()
        mapping = np.zeros((ninput, self.ng))
        for jint in range(1, ninterval + 1):
            for jw in range(self.nwav):
                wn2 = w2[jw] if jint == 1 else min(w2[jw], 0.01 / wavelength_bound[jint - 2])
                wn1 = w1[jw] if jint == ninterval else max(w1[jw], 0.01 / wavelength_bound[jint - 1])
                if wn2 > wn1:
                    mapping[i_intervals[jint - 1] - 1, :] += self.gpoint_fraction[jw, :] * (planck[jw] * (wn2 - wn1) / (w2[jw] - w1[jw]))
        for jg in range(self.ng):
            mapping[:, jg] = mapping[:, jg] * (1.0 / mapping[:, jg].sum())
        return mapping

    # ---- table-directory round trip ("ckd_<sw|lw>_*" arrays) ----
    def to_tables(self, prefix):
        out = {f"{prefix}wavenumber1": self.wavenumber1, f"{prefix}wavenumber2": self.wavenumber2,
               f"{prefix}gpoint_fraction": self.gpoint_fraction, f"{prefix}wavenumber1_band": self.wavenumber1_band,
               f"{prefix}wavenumber2_band": self.wavenumber2_band, f"{prefix}band_number": self.i_band_number}
        if self.solar_spectral_irradiance is not None:
            out[f"{prefix}solar_spectral_irradiance"] = self.solar_spectral_irradiance
        if self.solar_irradiance is not None:
            out[f"{prefix}solar_irradiance"] = self.solar_irradiance
        return out

    @classmethod
    def from_tables(cls, tabs, prefix):
        g = lambda n: tabs.get(prefix + n)  # noqa: E731
        return cls(g("wavenumber1"), g("wavenumber2"), g("gpoint_fraction"), g("wavenumber1_band"), g("wavenumber2_band"),
                   g("band_number"), g("solar_spectral_irradiance"), g("solar_irradiance"))
